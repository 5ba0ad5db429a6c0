"""Time the training GEMM shapes of the bench config on the all-TMA kernel (and the register-path kernel).
  python tools/bench_gemm_tma.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
from ebrec.models.newsrec import _ebk  # noqa: E402

_ebk.require_device()
lib = _ebk.lib()
R, E, D3, D, ATT = 192000, 768, 1200, 400, 200


def rnd(*shape):
    x = torch.randn(*shape, device="cuda")
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


cases = [  # name, tA, tB, M, N, K, A shape, B shape
    ("qkv_fwd   X[R,E].W[E,3D]", 0, 0, R, D3, E, (R, E), (E, D3)),
    ("qkv_dgrad dQKV[R,3D].W^T", 0, 1, R, E, D3, (R, D3), (E, D3)),
    ("qkv_wgrad X^T.dQKV", 1, 0, E, D3, R, (R, E), (R, D3)),
    ("att_fwd   Y[R,D].W[D,att]", 0, 0, R, ATT, D, (R, D), (D, ATT)),
    ("att_dgrad dpre[R,att].W^T", 0, 1, R, D, ATT, (R, ATT), (D, ATT)),
    ("att_wgrad Y^T.dpre", 1, 0, D, ATT, R, (R, D), (R, ATT)),
]
for name, tA, tB, M, N, K, sa, sb in cases:
    A, B = rnd(*sa), rnd(*sb)
    C = torch.zeros(M, N, device="cuda")
    fl = 2.0 * M * N * K
    line = f"{name:28s}"
    for tall in (0, 1, 0 | (1 << 4), 1 | (1 << 4)):
        f = lambda: _ebk.check(lib.ebk_gemm_tma(tA, tB, tall, M, N, K, _ebk.ptr(A), sa[1], _ebk.ptr(B), sb[1],
                                                _ebk.ptr(C), N, 0.0, 1.0, _ebk.stream()))
        ms = timeit(f)
        line += f"  t{tall & 15}c{tall >> 4} {ms:6.3f}ms {fl / ms / 1e9:5.0f}TF"
    print(line, flush=True)
