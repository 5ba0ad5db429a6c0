"""GPU parity of the all-TMA tcgen05 training GEMM (gemm_tma_sm100.cu) through the C-ABI.

Operands are rounded to tf32 on the host with the same bit arithmetic the producing kernels use, so the
tensor core's truncation is exact and the only error left is fp32 accumulation order: the result is
compared with the float64 product of the rounded operands at 2e-5 of sqrt(K).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ebk():
    from ebrec.models.newsrec import _ebk

    _ebk.require_device()
    return _ebk


def round_tf32(a: np.ndarray) -> np.ndarray:
    u = a.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


SHAPES = [(1, 1, 1), (37, 53, 29), (128, 256, 64), (300, 1200, 768), (130, 200, 5000), (257, 72, 132),
          (3840, 1200, 100), (768, 1200, 3841), (400, 200, 6000), (640, 400, 200), (513, 768, 1200)]


# tile/pair mode: bits 0-3 tile rows (0 = 128, 1 = 256, 15 = auto), bits 4-7 CTA pairs / cta_group::2 (0, 1, 15 = auto)
@pytest.mark.parametrize("tall", [0, 1, 0 | (1 << 4), 1 | (1 << 4), 15 | (15 << 4)])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tma(ebk, tall, tA, tB, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K + tA * 2 + tB)
    pad = lambda n: (n + 3) // 4 * 4   # row strides must be multiples of 4 floats (16-byte TMA strides)
    A = np.zeros((K, pad(M)) if tA else (M, pad(K)), np.float32)
    B = np.zeros((N, pad(K)) if tB else (K, pad(N)), np.float32)
    A[:, :(M if tA else K)] = round_tf32(rng.standard_normal((K, M) if tA else (M, K)))
    B[:, :(K if tB else N)] = round_tf32(rng.standard_normal((N, K) if tB else (K, N)))
    # poison the padding columns: they must never leak into the result
    A[:, (M if tA else K):] = np.nan
    B[:, (K if tB else N):] = np.nan
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    a = A[:, :M].T if tA else A[:, :K]
    b = B[:, :K].T if tB else B[:, :N]
    want = a.astype(np.float64) @ b.astype(np.float64)
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    for beta, alpha in ((0.0, 1.0), (1.0, 1.25)):
        Cd = torch.from_numpy(C0).cuda()
        ebk.check(ebk.lib().ebk_gemm_tma(tA, tB, tall, M, N, K, ebk.ptr(Ad), A.shape[1], ebk.ptr(Bd), B.shape[1],
                                         ebk.ptr(Cd), N, beta, alpha, ebk.stream()))
        ref = alpha * want + (C0 if beta else 0)
        err = np.abs(Cd.cpu().numpy().astype(np.float64) - ref).max() / (np.sqrt(K) + np.abs(C0).max())
        # fp32 accumulation in TMEM: the error grows with the length of one accumulation chain (K / split-K parts; the
        # tensor core's adds truncate).  Measured worst case over these shapes: 3.2e-5 (K = 3841 in two chains, alpha 1.25)
        assert err < 5e-5, f"beta={beta} err={err:.3e}"
