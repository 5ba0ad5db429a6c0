"""Run one ebk_gemm_tma case and print where the result differs (per 128 x BN tile): debugging aid."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
from ebrec.models.newsrec import _ebk as ebk

M, N, K, tA, tB, tall = [int(v) for v in sys.argv[1:7]]
rng = np.random.default_rng(1)
rt = lambda a: ((a.astype(np.float32).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
pad = lambda n: (n + 3) // 4 * 4
A = np.zeros((K, pad(M)) if tA else (M, pad(K)), np.float32)
B = np.zeros((N, pad(K)) if tB else (K, pad(N)), np.float32)
A[:, :(M if tA else K)] = rt(rng.standard_normal((K, M) if tA else (M, K)))
B[:, :(K if tB else N)] = rt(rng.standard_normal((N, K) if tB else (K, N)))
a = A[:, :M].T if tA else A[:, :K]
b = B[:, :K].T if tB else B[:, :N]
want = a.astype(np.float64) @ b.astype(np.float64)
Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
for beta, alpha in ((0.0, 1.0), (1.0, 1.25)):
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    Cd = torch.from_numpy(C0).cuda()
    ebk.check(ebk.lib().ebk_gemm_tma(tA, tB, tall, M, N, K, ebk.ptr(Ad), A.shape[1], ebk.ptr(Bd), B.shape[1], ebk.ptr(Cd), N, beta, alpha, ebk.stream()))
    got = Cd.cpu().numpy().astype(np.float64)
    diff = np.abs(got - (alpha * want + (C0 if beta else 0)))
    print(f"beta={beta} max diff {diff.max():.3e} at {np.unravel_index(diff.argmax(), diff.shape)} norm {diff.max() / (np.sqrt(K) + np.abs(C0).max()):.3e}")
    bad = diff > 1e-3 * np.sqrt(K)
    print("bad entries", int(bad.sum()), "rows", np.unique(np.nonzero(bad)[0] // 128), "col blocks of 16", np.unique(np.nonzero(bad)[1] // 16)[:40])
    if bad.any():
        i, j = np.argwhere(bad)[0]
        print("first bad", i, j, got[i, j], (alpha * want + (C0 if beta else 0))[i, j], "partial k-split check:", alpha * (a[i, :61 * 32].astype(np.float64) @ b[:61 * 32, j].astype(np.float64)), alpha * (a[i, 61 * 32:].astype(np.float64) @ b[61 * 32:, j].astype(np.float64)))
