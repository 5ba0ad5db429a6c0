"""Per-phase CUDA-event timing of the data-parallel optimizer step (run under torchrun):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_phase_timing.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
import bench  # noqa: E402
from ebrec.models.newsrec import _ebk  # noqa: E402
from ebrec.models.newsrec._engine import NRMSEngine, keras_adam_alpha  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
eng = NRMSEngine(V=w["V"], E=w["E"], T=w["T"], H=w["H"], nh=w["nh"], dh=w["dh"], att=w["att"], dropout=0.2, lr=1e-4, seed=1)
eng.params.theta.normal_(0, 0.02)
rng = np.random.default_rng(dist.get_rank())
tok, lab = eng.to_device_batch(*bench.synth_batch(rng, w, w["B"]))
P, lib = eng.params, _ebk.lib()
shard = P.n // eng.world
lo = eng.rank * shard
gsh = torch.empty(shard, device="cuda")
names = ["fwd+bwd", "reduce_scatter", "adam_shard", "all_gather", "zero_grad"]
tot = {k: 0.0 for k in names}
steps = 12
for it in range(steps + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    eng.loss_and_grads_dev(tok, lab, w["B"], w["C"], training=True)
    ev[1].record()
    dist.reduce_scatter_tensor(gsh, P.grad)
    ev[2].record()
    eng.step_count += 1
    alpha = keras_adam_alpha(eng.lr, eng.step_count, eng.beta1, eng.beta2)
    th, m, v = P.theta[lo: lo + shard], P.m[lo: lo + shard], P.v[lo: lo + shard]
    _ebk.check(lib.ebk_adam_keras_step(_ebk.ptr(th), _ebk.ptr(gsh), _ebk.ptr(m), _ebk.ptr(v), shard, alpha, eng.beta1,
                                       eng.beta2, eng.eps, 0, _ebk.stream()))
    ev[3].record()
    dist.all_gather_into_tensor(P.theta, th)
    ev[4].record()
    P.grad.zero_()
    ev[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i, k in enumerate(names):
            tot[k] += ev[i].elapsed_time(ev[i + 1])
if dist.get_rank() == 0:
    print({k: round(v / steps, 3) for k, v in tot.items()}, "world", eng.world, flush=True)
dist.destroy_process_group()
