"""CPU tests of the oracle (the checker itself): structure, known answers, analytic backward."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import nrms_oracle as O, torch_port as TP

GOLD = Path(__file__).parent / "golden"


def small(rng=None, V=50, E=16, nh=3, dh=4, att=10, B=3, H=5, C=4, T=6):
    rng = rng or np.random.default_rng(0)
    P = O.init_nrms_params(rng, V, E, nh, dh, att, dtype=np.float64)
    P["news_b"] = rng.standard_normal(att) * 0.1
    his = rng.integers(0, V, (B, H, T))
    pred = rng.integers(0, V, (B, C, T))
    y = np.zeros((B, C), int)
    y[np.arange(B), rng.integers(0, C, B)] = 1
    return P, his, pred, y, (nh, dh)


def test_param_count_matches_keras_summary():
    # nrms_dummy.py: table 1000x100, head 20x20, att 200 -> 860 800 parameters
    P = O.init_nrms_params(np.random.default_rng(0), 1000, 100, 20, 20, 200)
    assert O.count_params(P) == 860_800
    assert [P[k].shape for k in O.NRMS_PARAM_ORDER[:7]] == [(1000, 100), (100, 400), (100, 400), (100, 400), (400, 200), (200,), (200, 1)]


def test_attention_uses_adjoint_product_not_standard_attention():
    """layers.py:249 tf.matmul(A, V, adjoint_a=True): O[k] = sum_q A[q,k] V[q]."""
    rng = np.random.default_rng(1)
    X = rng.standard_normal((2, 5, 8))
    W = [rng.standard_normal((8, 6)) for _ in range(3)]
    out, cache = O.self_attention_fwd(X, *W, 2, 3)
    Q, K, V, A = cache[4], cache[5], cache[6], cache[7]
    adj = np.einsum("nhqk,nhqd->nhkd", A, V).transpose(0, 2, 1, 3).reshape(2, 5, 6)
    std = np.einsum("nhqk,nhkd->nhqd", A, V).transpose(0, 2, 1, 3).reshape(2, 5, 6)
    np.testing.assert_allclose(out, adj, rtol=1e-12)
    assert np.abs(out - std).max() > 1e-2
    np.testing.assert_allclose(A.sum(-1), 1.0, rtol=1e-12)  # softmax over keys


def test_attlayer2_plain_exp_and_epsilon():
    """layers.py:70-77: w = exp(a) / (sum exp(a) + 1e-7), no max subtraction."""
    rng = np.random.default_rng(2)
    X = rng.standard_normal((1, 4, 6))
    W, b, q = rng.standard_normal((6, 5)), rng.standard_normal(5), rng.standard_normal((5, 1))
    y, (_, _, _, h, w) = O.att_layer2_fwd(X, W, b, q)
    a = (np.tanh(X @ W + b) @ q)[..., 0]
    e = np.exp(a)
    np.testing.assert_allclose(w, e / (e.sum(-1, keepdims=True) + 1e-7), rtol=1e-13)
    assert w.sum() < 1.0  # the epsilon keeps the weights from summing to exactly one
    # very negative scores: epsilon dominates and the pooled vector shrinks towards zero
    y2, _ = O.att_layer2_fwd(X, W, b - 100.0, q * 0 + 1.0)
    assert np.abs(y2).max() < np.abs(y).max()


def test_analytic_backward_matches_autograd():
    P, his, pred, y, (nh, dh) = small()
    p = 0.2
    loss, prob, G = O.nrms_loss_and_grads(his, pred, y, P, nh, dh, training=True, p_drop=p, seed1=11, seed2=22)
    N = his.shape[0] * (his.shape[1] + pred.shape[1])
    T, E, D = his.shape[2], P["table"].shape[1], nh * dh
    k1 = O.dropout_keep_mask(11, N * T * E, p).reshape(N, T, E)
    k2 = O.dropout_keep_mask(22, N * T * D, p).reshape(N, T, D)
    Pt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
    l, _ = TP.nrms_loss(torch.tensor(his), torch.tensor(pred), torch.tensor(y), Pt, nh, dh,
                        keep1=torch.tensor(k1, dtype=torch.float64), keep2=torch.tensor(k2, dtype=torch.float64), p_drop=p)
    l.backward()
    assert abs(float(l) - loss) < 1e-12
    for k in P:
        np.testing.assert_allclose(G[k], Pt[k].grad.numpy(), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(prob.sum(-1), 1.0, rtol=1e-12)


def test_dropout_mask_statistics_and_determinism():
    m = O.dropout_keep_mask(123456789, 1 << 18, 0.2)
    assert abs(m.mean() - 0.8) < 3e-3
    assert np.array_equal(m, O.dropout_keep_mask(123456789, 1 << 18, 0.2))
    assert not np.array_equal(m, O.dropout_keep_mask(123456790, 1 << 18, 0.2))
    x = np.ones(1000)
    yv, keep = O.dropout_fwd(x, 5, 0.25)
    assert set(np.unique(yv)) <= {0.0, 1.0 / 0.75}  # inverted dropout, nrms.py:136


def test_keras_adam_first_step_closed_form():
    """t=1: m=(1-b1)g, v=(1-b2)g^2, alpha=lr*sqrt(1-b2)/(1-b1) -> theta -= lr*g/(|g| + eps/sqrt(1-b2))."""
    g = np.array([0.5, -2.0, 0.0, 1e-9])
    th = np.zeros(4)
    m, v = np.zeros(4), np.zeros(4)
    O.keras_adam_step(th, g, m, v, 1, 1e-3)
    want = -1e-3 * g / (np.abs(g) + 1e-7 / np.sqrt(1 - 0.999))
    np.testing.assert_allclose(th, want, rtol=1e-9, atol=1e-18)
    # non-lazy: a zero gradient at step 2 still decays m, v and moves theta
    th2 = th.copy()
    O.keras_adam_step(th2, np.zeros(4), m, v, 2, 1e-3)
    assert np.abs(th2 - th)[:2].min() > 0


def test_out_of_range_token_reads_zero_row_and_gets_no_gradient():
    P, his, pred, y, (nh, dh) = small()
    V = P["table"].shape[0]
    his2 = his.copy()
    his2[0, 0, 0] = V + 5
    P2 = {k: v.copy() for k, v in P.items()}
    P2["table"] = np.concatenate([P["table"], np.zeros((10, P["table"].shape[1]))])  # explicit zero rows
    z1, _ = O.nrms_forward(his2, pred, P, nh, dh)
    z2, _ = O.nrms_forward(his2, pred, P2, nh, dh)
    np.testing.assert_allclose(z1, z2, rtol=1e-12)


def test_oracle_fixture_has_not_drifted():
    """tests/golden/nrms_oracle_case.npz is SELF-generated (the reference has no golden model vectors)."""
    f = np.load(GOLD / "nrms_oracle_case.npz")
    V, E, nh, dh, att, B, H, C, T = (int(x) for x in f["dims"])
    P = {k: f[f"P_{k}"] for k in O.NRMS_PARAM_ORDER}
    np.testing.assert_allclose(O.nrms_predict(f["his"], f["pred"], P, nh, dh), f["probs"], rtol=1e-10)
    np.testing.assert_allclose(O.nrms_score(f["his"], f["pred"], P, nh, dh), f["sigmoid"], rtol=1e-10)
    loss, _, G = O.nrms_loss_and_grads(f["his"], f["pred"], f["y"], P, nh, dh, training=True, p_drop=0.2, seed1=11, seed2=22)
    assert abs(loss - float(f["loss"])) < 1e-10
    np.testing.assert_allclose(G["news_WV"], f["g_news_WV"], rtol=1e-8, atol=1e-14)
    assert np.array_equal(O.dropout_keep_mask(11, 256, 0.2), f["keep_head"])


def test_float32_oracle_close_to_float64():
    P, his, pred, y, (nh, dh) = small()
    P32 = {k: v.astype(np.float32) for k, v in P.items()}
    a = O.nrms_predict(his, pred, P, nh, dh)
    b = O.nrms_predict(his, pred, P32, nh, dh)
    assert b.dtype == np.float32 and np.abs(a - b).max() < 1e-5


def test_dense_stack_oracle_gradients_match_finite_differences():
    """NRMS with the Dense/BN/Dropout stack (nrms.py:142-152): the analytic backward of
    oracle/nrms_dense_oracle.py against central differences of its own float64 loss (training-mode BN,
    dropout masks fixed by the seeds)."""
    from oracle import nrms_dense_oracle as ND

    rng = np.random.default_rng(5)
    V, E, units, nh, dh, att, B, H, C, T = 30, 12, [10, 8], 2, 4, 6, 3, 4, 3, 5
    P = ND.init_params(rng, V, E, units, nh, dh, att, dtype=np.float64)
    for i, u in enumerate(units):
        P[f"d{i}_b"] = rng.standard_normal(u) * 0.3 + 0.2
        P[f"d{i}_gamma"] = 1 + rng.standard_normal(u) * 0.1
    his, pred = rng.integers(0, V, (B, H, T)), rng.integers(0, V, (B, C, T))
    y = np.zeros((B, C))
    y[np.arange(B), rng.integers(0, C, B)] = 1
    kw = dict(p_drop=0.2, seed1=11, seed_h=22, seed_c=33, l2=1e-2)
    loss, _, G = ND.loss_and_grads(his, pred, y, P, len(units), nh, dh, **kw)
    for k in ("table", "news_WQ", "d0_W", "d0_b", "d1_gamma", "d1_beta", "news_W", "news_q", "user_WV"):
        idx = tuple(rng.integers(0, s) for s in P[k].shape)
        if k == "table":
            idx = (int(his[0, 0, 0]), idx[1])
        old = P[k][idx]
        eps = 1e-6
        P[k][idx] = old + eps
        lp = ND.loss_and_grads(his, pred, y, P, len(units), nh, dh, **kw)[0]
        P[k][idx] = old - eps
        lm = ND.loss_and_grads(his, pred, y, P, len(units), nh, dh, **kw)[0]
        P[k][idx] = old
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - G[k][idx]) < 1e-6 * max(1.0, abs(fd)) + 1e-8, (k, fd, G[k][idx])
