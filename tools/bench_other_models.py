"""Step time of the other two model families at their BASELINE.json shapes (configs[1] and configs[4], per GPU):
  NRMSDocVec bs=512 H=20 C=5 Ddoc=768 units 512x3 ; NAML bs=64 H=50 C=5 title 30 body 40 V=32000 E=300.
  python tools/bench_other_models.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
from ebrec.models.newsrec import _ebk  # noqa: E402
from ebrec.models.newsrec._engine_docvec import DocVecEngine  # noqa: E402
from ebrec.models.newsrec._engine_naml import NAMLEngine  # noqa: E402

_ebk.require_device()
rng = np.random.default_rng(0)


def timeit(step, n=10, warm=3):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# ---- NRMSDocVec (configs[1])
B, H, C, Dd = 512, 20, 5, 768
e = DocVecEngine(Ddoc=Dd, units=[512, 512, 512], H=H, nh=16, dh=16, att=200, dropout=0.2, lr=1e-4, l2=1e-4, seed=1)
for name, shape in e.params.spec:
    if name.endswith(("_W", "Wqkv", "attW", "attq")):
        e.params.p(name).normal_(0, 0.05)
    if name.endswith("_gamma"):
        e.params.p(name).fill_(1.0)
his, pred = rng.standard_normal((B, H, Dd)).astype(np.float32), rng.standard_normal((B, C, Dd)).astype(np.float32)
y = np.zeros((B, C), np.float32)
y[np.arange(B), rng.integers(0, C, B)] = 1
x, lab = e.to_device_batch(his, pred, y)
ms = timeit(lambda: e.train_step_dev(x, lab, B, C))
print(f"NRMSDocVec bs={B}: {ms:.3f} ms/step = {B / ms * 1e3:,.0f} impressions/s", flush=True)

# ---- NAML (configs[4], per-GPU share of bs=512 over 8 GPUs)
B, H, C, T, Tb, V, E = 64, 50, 5, 30, 40, 32000, 300
n = NAMLEngine(V=V, E=E, T=T, Tb=Tb, H=H, F=400, att=200, window=3, vert_num=100, vert_dim=10, subvert_num=100, subvert_dim=10,
               dropout=0.2, lr=1e-4, seed=1)
for name, shape in n.params.spec:
    if len(shape) >= 2 or name.endswith("_q"):
        n.params.p(name).normal_(0, 0.05)
arrays = (rng.integers(0, V, (B, H, T)), rng.integers(0, V, (B, H, Tb)), rng.integers(0, 100, (B, H, 1)), rng.integers(0, 100, (B, H, 1)),
          rng.integers(0, V, (B, C, T)), rng.integers(0, V, (B, C, Tb)), rng.integers(0, 100, (B, C, 1)), rng.integers(0, 100, (B, C, 1)))
y = np.zeros((B, C), np.float32)
y[np.arange(B), rng.integers(0, C, B)] = 1
xn, labn = n.to_device_batch(arrays, y)
ms = timeit(lambda: n.train_step_dev(xn, labn, B, C))
print(f"NAML bs={B}/GPU H=50: {ms:.3f} ms/step = {B / ms * 1e3:,.0f} impressions/s", flush=True)
lib = _ebk.lib()
lib.ebk_prof_enable(1)
for _ in range(5):
    n.train_step_dev(xn, labn, B, C)
torch.cuda.synchronize()
pr = {k: round(v[0] / 5, 4) for k, v in _ebk.prof_collect().items()}
lib.ebk_prof_enable(0)
print("NAML kernel ms/step:", dict(sorted(pr.items(), key=lambda kv: -kv[1])), "sum", round(sum(pr.values()), 3), flush=True)
lib.ebk_prof_enable(1)
for _ in range(5):
    e.train_step_dev(x, lab, 512, 5)
torch.cuda.synchronize()
pr = {k: round(v[0] / 5, 4) for k, v in _ebk.prof_collect().items()}
lib.ebk_prof_enable(0)
print("DocVec kernel ms/step:", dict(sorted(pr.items(), key=lambda kv: -kv[1])), "sum", round(sum(pr.values()), 3), flush=True)
