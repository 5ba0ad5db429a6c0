"""GPU parity tests of the ebk building blocks against the numpy oracle (float64).

All calls go through the C-ABI (ctypes) exactly as the product does.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import nrms_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ebk():
    from ebrec.models.newsrec import _ebk

    _ebk.require_device()
    return _ebk


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype).cuda()


def relerr(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


MATHS = [0, 1, 2]  # fp32 FMA, tcgen05 tf32, tcgen05 3xTF32


def test_dropout_mask_bit_exact(ebk):
    for seed, p, n in [(1, 0.2, 1000), (2**63 + 12345, 0.5, 4099), (7, 0.0, 64)]:
        out = torch.empty(n, device="cuda")
        ebk.check(ebk.lib().ebk_dropout_mask(seed, p, n, ebk.ptr(out), ebk.stream()))
        want = O.dropout_keep_mask(seed, n, p) if p > 0 else np.ones(n, bool)
        assert np.array_equal(out.cpu().numpy() > 0.5, want)


@pytest.mark.parametrize("math", MATHS)
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 53, 29), (128, 256, 64), (300, 1200, 768), (130, 200, 5000), (257, 72, 132)])
def test_gemm(ebk, math, tA, tB, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K + tA * 2 + tB)
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    want = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    Ad, Bd = dev(A), dev(B)  # keep alive: ptr() of a temporary would dangle
    for beta in (0.0, 1.0):
        Cd = dev(C0)
        ebk.check(ebk.lib().ebk_gemm(math, tA, tB, M, N, K, ebk.ptr(Ad), A.shape[1], ebk.ptr(Bd), B.shape[1],
                                     ebk.ptr(Cd), N, beta, ebk.stream()))
        ref = want + (C0 if beta else 0)
        # tf32: 10-bit mantissa inputs, fp32 accumulate -> ~1e-3 relative to the row/col norms
        tol = {0: 2e-5, 1: 2e-3, 2: 2e-5}[math]
        scale = np.sqrt(K) + np.abs(C0).max()
        err = np.abs(Cd.cpu().numpy() - ref).max() / scale
        assert err < tol, f"beta={beta} err={err:.3e}"


@pytest.mark.parametrize("n_seq,L,nh,dh", [(3, 30, 20, 20), (5, 20, 16, 16), (2, 50, 20, 20), (4, 7, 3, 4), (1, 64, 2, 20), (2, 33, 2, 32)])
def test_attention_core(ebk, n_seq, L, nh, dh):
    rng = np.random.default_rng(L + nh)
    D = nh * dh
    qkv = rng.standard_normal((n_seq, L, 3 * D))
    Q, K, V = (qkv[..., i * D:(i + 1) * D].reshape(n_seq, L, nh, dh).transpose(0, 2, 1, 3) for i in range(3))
    S = np.einsum("nhqd,nhkd->nhqk", Q, K) / np.sqrt(dh)
    A = O.softmax(S)
    Oo = np.einsum("nhqk,nhqd->nhkd", A, V).transpose(0, 2, 1, 3).reshape(n_seq, L, D)
    y = torch.empty(n_seq * L, D, device="cuda")
    qd = dev(qkv.reshape(n_seq * L, 3 * D))
    ebk.check(ebk.lib().ebk_attention_core_fwd(n_seq, L, nh, dh, ebk.ptr(qd), ebk.ptr(y), ebk.stream()))
    assert relerr(y.cpu().numpy().reshape(n_seq, L, D), Oo) < 2e-5
    # the discriminating check: A V (standard attention) must NOT match  (layers.py:249 adjoint_a=True)
    std = np.einsum("nhqk,nhkd->nhqd", A, V).transpose(0, 2, 1, 3).reshape(n_seq, L, D)
    if L > 1:
        assert relerr(y.cpu().numpy().reshape(n_seq, L, D), std) > 1e-2

    # backward with a dropout mask on dy
    dy = rng.standard_normal((n_seq, L, D))
    p, seed = 0.2, 99
    keep = O.dropout_keep_mask(seed, dy.size, p).reshape(dy.shape)
    dO = (dy * keep / (1 - p)).reshape(n_seq, L, nh, dh).transpose(0, 2, 1, 3)
    dV = np.einsum("nhqk,nhkd->nhqd", A, dO)
    dA = np.einsum("nhqd,nhkd->nhqk", V, dO)
    dS = A * (dA - (dA * A).sum(-1, keepdims=True))
    dQ = np.einsum("nhqk,nhkd->nhqd", dS, K) / np.sqrt(dh)
    dK = np.einsum("nhqk,nhqd->nhkd", dS, Q) / np.sqrt(dh)
    want = np.concatenate([x.transpose(0, 2, 1, 3).reshape(n_seq, L, D) for x in (dQ, dK, dV)], axis=-1)
    dqkv = torch.empty(n_seq * L, 3 * D, device="cuda")
    dyd = dev(dy.reshape(n_seq * L, D))
    ebk.check(ebk.lib().ebk_attention_core_bwd(n_seq, L, nh, dh, ebk.ptr(qd), ebk.ptr(dyd),
                                               p, seed, ebk.ptr(dqkv), ebk.stream()))
    assert relerr(dqkv.cpu().numpy().reshape(n_seq, L, 3 * D), want) < 5e-5


def test_score_softmax_ce_and_sigmoid(ebk):
    rng = np.random.default_rng(5)
    for B, Cc, D in [(7, 5, 400), (3, 1, 256), (2, 250, 400)]:
        news, user = rng.standard_normal((B, Cc, D)) * 0.2, rng.standard_normal((B, D)) * 0.2
        y = np.zeros((B, Cc))
        y[np.arange(B), rng.integers(0, Cc, B)] = 1
        z = O.click_logits(news, user)
        loss, p, dz = O.softmax_ce(z, y)
        scale = 1.0 / B
        probs = torch.empty(B, Cc, device="cuda")
        ls = torch.zeros(1, device="cuda")
        dn, du = torch.empty(B, Cc, D, device="cuda"), torch.empty(B, D, device="cuda")
        nd, ud, yd = dev(news), dev(user), dev(y)
        ebk.check(ebk.lib().ebk_score_softmax_ce(B, Cc, D, ebk.ptr(nd), ebk.ptr(ud), ebk.ptr(yd), scale,
                                                 ebk.ptr(probs), ebk.ptr(ls), ebk.ptr(dn), ebk.ptr(du), ebk.stream()))
        assert relerr(probs.cpu().numpy(), p) < 1e-5
        assert abs(float(ls) - loss) < 1e-5 * max(1, abs(loss))
        assert relerr(dn.cpu().numpy(), dz[..., None] * user[:, None, :]) < 1e-5
        assert relerr(du.cpu().numpy(), np.einsum("bc,bcd->bd", dz, news)) < 1e-5
        sg = torch.empty(B, Cc, device="cuda")
        ebk.check(ebk.lib().ebk_score_sigmoid(B, Cc, D, ebk.ptr(nd), ebk.ptr(ud), ebk.ptr(sg), ebk.stream()))
        assert relerr(sg.cpu().numpy(), O.sigmoid(z)) < 1e-5


def test_adam_keras_matches_oracle(ebk):
    rng = np.random.default_rng(9)
    n = 4 * 1000 + 3
    th = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    thd, md, vd = dev(th), dev(m), dev(v)
    from ebrec.models.newsrec._engine import keras_adam_alpha

    for t in range(1, 4):
        g = (rng.standard_normal(n) * 0.01).astype(np.float32)
        g[::3] = 0  # rows without gradient still decay m, v and move (non-lazy Keras Adam)
        gd = dev(g)
        O.keras_adam_step(th, g, m, v, t, 1e-3)
        ebk.check(ebk.lib().ebk_adam_keras_step(ebk.ptr(thd), ebk.ptr(gd), ebk.ptr(md), ebk.ptr(vd), n,
                                                keras_adam_alpha(1e-3, t, 0.9, 0.999), 0.9, 0.999, 1e-7, 1, ebk.stream()))
        assert float(gd.abs().max()) == 0.0
        np.testing.assert_allclose(thd.cpu().numpy(), th, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(md.cpu().numpy(), m, rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(vd.cpu().numpy(), v, rtol=2e-6, atol=1e-12)
