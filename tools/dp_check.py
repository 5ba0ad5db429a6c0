"""Data-parallel correctness on real GPUs (run under torchrun, world 2+):
every rank trains on its shard of a global batch for a few steps (overlapped reduce-scatter / sharded Adam /
all-gather path of NRMSEngine.apply_adam); rank 0 also trains a world-1 engine on the whole batch.  The
parameters must agree across ranks bit for bit and with the single-process run up to summation order.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
from oracle import nrms_oracle as O  # noqa: E402  (checker only)
from ebrec.models.newsrec._engine import NRMSEngine  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
V, E, nh, dh, att, Bg, H, C, T = 4096, 64, 4, 8, 24, 8 * world, 10, 5, 12
rng = np.random.default_rng(0)
P = O.init_nrms_params(rng, V, E, nh, dh, att)
lr, steps = 1e-3, 4
batches = []
for _ in range(steps):
    his = rng.integers(0, V, (Bg, H, T)).astype(np.int32)
    pred = rng.integers(0, V, (Bg, C, T)).astype(np.int32)
    y = np.zeros((Bg, C), np.float32)
    y[np.arange(Bg), rng.integers(0, C, Bg)] = 1
    batches.append((his, pred, y))


def run(engine, sl):
    for his, pred, y in batches:
        tok, lab = engine.to_device_batch(his[sl], pred[sl], y[sl])
        engine.train_step_dev(tok, lab, his[sl].shape[0], C)
    engine._sync_table()   # rank-sharded table: bring this replica up to date (no-op on one GPU)
    torch.cuda.synchronize()
    return engine.params.theta.clone()


eng = NRMSEngine(V=V, E=E, T=T, H=H, nh=nh, dh=dh, att=att, dropout=0.0, lr=lr, seed=1)
eng.set_weights([P[k] for k in O.NRMS_PARAM_ORDER])
Bl = Bg // world
theta = run(eng, slice(rank * Bl, (rank + 1) * Bl))
others = [torch.empty_like(theta) for _ in range(world)]
dist.all_gather(others, theta)
same = all(bool(torch.equal(others[0], o)) for o in others)
ok = same
if rank == 0:
    solo = NRMSEngine(V=V, E=E, T=T, H=H, nh=nh, dh=dh, att=att, dropout=0.0, lr=lr, seed=1)
    solo.world, solo.rank = 1, 0
    solo.sparse_table_grad = False
    solo.set_weights([P[k] for k in O.NRMS_PARAM_ORDER])
    ref = run(solo, slice(0, Bg))
    dev = float((theta - ref).abs().mean() / (steps * lr))
    moved = float((ref - torch.cat([torch.from_numpy(P[k].ravel()) for k in ["table"]]).cuda().new_zeros(1)).abs().mean())
    ok = ok and dev < 3e-2
    print(f"DP_CHECK world={world} ranks_identical={same} mean|dp-solo|/(steps*lr)={dev:.3e} {'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
