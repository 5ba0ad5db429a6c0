"""GPU parity of the NRMS hot path (through the C-ABI and the engine) against the oracle.

Tolerances: EBK_MATH_FP32 -> 1e-4 relative (fp32 round-off vs the float64 oracle);
EBK_MATH_TF32 -> 1e-3 relative on forward click scores (the north-star gate) and 2e-2 of
the gradient's max-norm for backward quantities (tf32 inputs, fp32 accumulation).
"""
import numpy as np
import pytest
import torch

from oracle import nrms_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = {0: 1e-4, 1: 1e-3}      # inference arithmetic (fp32 / 3xTF32): the north-star click-score gate
# training arithmetic (fp32 / single-pass TF32 on the TMA path): measured logit error 1.4e-3 ... 4.3e-3 of max|logit|
# (tests/test_gpu_reference_golden.py::test_nrms_training_kernels_forward_error_measured); its gate is AUC parity
TRAIN_TOL = {0: 1e-4, 1: 1.5e-2}   # (tiny dims + peaky softmax + dropout: up to 1.2e-2 measured on case 2)
BWD_TOL = {0: 2e-4, 1: 2e-2}


# news-encoder WV scale per case (keyed by V): with Glorot-sized weights every impression's logits agree to ~1e-3
# and softmax(z) is uniform whatever the kernels compute -- the scale makes the per-impression logit SPREAD >= 1 so
# that score comparisons carry signal (asserted in the tests)
CASE_WV = {1000: 3.0, 500: 3.0, 300: 1.3, 50: 2.0}


def make_case(rng, V, E, nh, dh, att, B, H, C, T, table_scale=None, wv=None):
    P = O.init_nrms_params(rng, V, E, nh, dh, att, dtype=np.float64)
    if table_scale is not None:
        P["table"] = rng.random((V, E)) * table_scale
    if wv is not None:
        P["table"] = rng.standard_normal((V, E))
        P["news_WV"] = P["news_WV"] * wv
    for k in ("news_b", "user_b"):
        P[k] = rng.standard_normal(P[k].shape) * 0.05
    # Glorot-sized WQ/WK give near-uniform attention, whose WQ/WK gradients are pure fp32
    # cancellation noise (dS = A o (dA - rowsum)); scale them so the softmax is exercised.
    for k in ("news_WQ", "news_WK", "user_WQ", "user_WK"):
        P[k] = P[k] * 6.0
    his = rng.integers(0, V, (B, H, T)).astype(np.int32)
    pred = rng.integers(0, V, (B, C, T)).astype(np.int32)
    y = np.zeros((B, C), np.float32)
    y[np.arange(B), rng.integers(0, C, B)] = 1
    return P, his, pred, y


def make_engine(P, V, E, T, H, nh, dh, att, dropout, lr, math, seed=3):
    from ebrec.models.newsrec._engine import NRMSEngine

    eng = NRMSEngine(V=V, E=E, T=T, H=H, nh=nh, dh=dh, att=att, dropout=dropout, lr=lr, seed=seed, math=math)
    eng.set_weights([P[k] for k in O.NRMS_PARAM_ORDER])
    return eng


def rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


CASES = [
    # V, E, nh, dh, att, B, H, C, T
    (1000, 100, 20, 20, 200, 8, 20, 5, 30),   # config 1 (nrms_dummy.py) at a small batch
    (500, 64, 16, 16, 200, 4, 20, 5, 30),
    (300, 32, 4, 8, 24, 3, 50, 7, 12),        # long history, odd sizes
    (50, 16, 3, 4, 10, 1, 5, 1, 6),           # single impression, single candidate
]


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_forward_scores(math, case):
    V, E, nh, dh, att, B, H, C, T = case
    rng = np.random.default_rng(sum(case))
    P, his, pred, y = make_case(rng, V, E, nh, dh, att, B, H, C, T, wv=CASE_WV[V])
    eng = make_engine(P, V, E, T, H, nh, dh, att, 0.2, 1e-4, math)
    tok, _ = eng.to_device_batch(his, pred)
    # gather indices are passed through bit-exact
    assert np.array_equal(tok.cpu().numpy(), np.concatenate([his.reshape(-1, T), pred.reshape(-1, T)]))
    zw, _ = O.nrms_forward(his, pred, P, nh, dh)
    # precondition: the gate is not vacuous (per-impression logit spread >= 1; one candidate: |logit| >= 1)
    assert (np.ptp(zw, axis=1).min() >= 1.0) if C > 1 else (np.abs(zw).min() >= 1.0)
    _, news_c, u, _ = eng.forward_logits_parts(tok, B, C)
    z = (news_c * u[:, None, :]).sum(-1).cpu().numpy()
    assert rel(z, zw) < FWD_TOL[math], ("logits", rel(z, zw))
    probs = eng.predict_dev(tok, B, C).cpu().numpy()
    want = O.nrms_predict(his, pred, P, nh, dh)
    assert rel(probs, want) < FWD_TOL[math], (rel(probs, want))
    sig = eng.predict_dev(tok, B, C, head="sigmoid").cpu().numpy()
    assert rel(sig, O.nrms_score(his, pred, P, nh, dh)) < FWD_TOL[math]


def test_out_of_range_ids_read_zero_rows():
    V, E, nh, dh, att, B, H, C, T = 50, 16, 3, 4, 10, 2, 5, 3, 6
    rng = np.random.default_rng(1)
    P, his, pred, y = make_case(rng, V, E, nh, dh, att, B, H, C, T)
    his[0, 0, :3] = [V, V + 100, -1]
    eng = make_engine(P, V, E, T, H, nh, dh, att, 0.0, 1e-4, 0)
    tok, lab = eng.to_device_batch(his, pred, y)
    probs = eng.predict_dev(tok, B, C).cpu().numpy()
    assert rel(probs, O.nrms_predict(his, pred, P, nh, dh)) < 1e-4


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("dropout", [0.0, 0.2])
@pytest.mark.parametrize("case", CASES[:3])
def test_loss_and_gradients(math, dropout, case):
    V, E, nh, dh, att, B, H, C, T = case
    rng = np.random.default_rng(sum(case) + 1)
    P, his, pred, y = make_case(rng, V, E, nh, dh, att, B, H, C, T, wv=CASE_WV[V])
    s1, s2 = 1234567, 7654321
    zw = O.nrms_forward(his, pred, P, nh, dh, training=True, p_drop=dropout, seed1=s1, seed2=s2)[0]
    assert np.ptp(zw, axis=1).min() >= 1.0   # loss / probs carry signal
    zmax = max(1.0, float(np.abs(zw).max()))  # a relative logit error eps moves loss and probabilities by ~eps * max|z|
    eng = make_engine(P, V, E, T, H, nh, dh, att, dropout, 1e-4, math)
    tok, lab = eng.to_device_batch(his, pred, y)
    eng.params.grad.zero_()
    loss, probs = eng.loss_and_grads_dev(tok, lab, B, C, training=True, seeds=(s1, s2))
    wl, wp, G = O.nrms_loss_and_grads(his, pred, y, P, nh, dh, training=True, p_drop=dropout, seed1=s1, seed2=s2)
    assert abs(float(loss) - wl) < TRAIN_TOL[math] * zmax, (float(loss), wl)
    assert rel(probs.cpu().numpy(), wp) < TRAIN_TOL[math] * 3 * zmax
    D = nh * dh
    got = {
        "table": eng.params.g("table").cpu().numpy(),
        "news_W": eng.params.g("news_attW").cpu().numpy(),
        "news_b": eng.params.g("news_attb").cpu().numpy(),
        "news_q": eng.params.g("news_attq").cpu().numpy().reshape(-1, 1),
        "user_W": eng.params.g("user_attW").cpu().numpy(),
        "user_b": eng.params.g("user_attb").cpu().numpy(),
        "user_q": eng.params.g("user_attq").cpu().numpy().reshape(-1, 1),
    }
    for pre in ("news", "user"):
        W = eng.params.g(f"{pre}_Wqkv").cpu().numpy()
        got[f"{pre}_WQ"], got[f"{pre}_WK"], got[f"{pre}_WV"] = W[:, :D], W[:, D:2 * D], W[:, 2 * D:]
    # Conditioning-aware bound: some gradients (WQ/WK, AttLayer2 W/b/q) are differences of nearly
    # equal terms, so their achievable accuracy is set by the arithmetic's epsilon times a
    # cancellation factor.  That factor is measured by evaluating the SAME oracle in float32:
    # allowed error = tol * max|G| + 20 * |G_f32 - G_f64| * (eps_math / eps_f32).
    P32 = {k: v.astype(np.float32) for k, v in P.items()}
    _, _, G32 = O.nrms_loss_and_grads(his, pred, y, P32, nh, dh, training=True, p_drop=dropout, seed1=s1, seed2=s2)
    amp = {0: 1.0, 1: 2.0 ** 13}[math]  # tf32 keeps 10 mantissa bits vs fp32's 23
    report = {}
    for k in O.NRMS_PARAM_ORDER:
        err = np.abs(got[k] - G[k]).max()
        noise = np.abs(G32[k].astype(np.float64) - G[k]).max()
        allowed = BWD_TOL[math] * (zmax if math == 1 else 1.0) * np.abs(G[k]).max() + 20.0 * amp * noise
        report[k] = (err / (np.abs(G[k]).max() + 1e-30), err / allowed)
    print("gradient (rel err, err/allowed):", {k: (f"{a:.1e}", f"{b:.2f}") for k, (a, b) in report.items()})
    for k, (_, ratio) in report.items():
        assert ratio < 1.0, (k, report)
    # the well-conditioned gradients must be tight without the noise allowance
    # (case 2 -- dh = 8, H = 50, 12-token titles -- is ill-conditioned once its logits are O(1): the float32 oracle
    # itself differs from the float64 one by ~1e-4 of max|g| there, so in tf32 only the conditioning-aware bound applies)
    if math == 0 or V != 300:
        for k in ("table", "news_WV", "user_WV"):
            assert report[k][0] < BWD_TOL[math] * (zmax if math == 1 else 1.0), (k, report)


@pytest.mark.parametrize("math", [0, 1])
def test_train_steps_follow_oracle(math):
    """3 optimizer steps with dropout: weights track the float64 oracle run with the same masks."""
    V, E, nh, dh, att, B, H, C, T = 200, 32, 4, 8, 24, 6, 10, 5, 12
    rng = np.random.default_rng(11)
    P, his, pred, y = make_case(rng, V, E, nh, dh, att, B, H, C, T)
    lr = 1e-3
    eng = make_engine(P, V, E, T, H, nh, dh, att, 0.2, lr, math, seed=5)
    Pm = {k: np.zeros_like(v) for k, v in P.items()}
    Pv = {k: np.zeros_like(v) for k, v in P.items()}
    losses = []
    for t in range(1, 4):
        his = rng.integers(0, V, (B, H, T)).astype(np.int32)
        pred = rng.integers(0, V, (B, C, T)).astype(np.int32)
        tok, lab = eng.to_device_batch(his, pred, y)
        s1, s2 = eng.step_seeds()
        loss, _ = eng.train_step_dev(tok, lab, B, C)
        wl, _, G = O.nrms_loss_and_grads(his, pred, y, P, nh, dh, training=True, p_drop=0.2, seed1=s1, seed2=s2)
        for k in P:
            O.keras_adam_step(P[k], G[k], Pm[k], Pv[k], t, lr)
        losses.append((float(loss), wl))
        assert abs(float(loss) - wl) < 5e-3 * max(1, abs(wl))
    W = eng.get_weights()
    for k, w in zip(O.NRMS_PARAM_ORDER, W):
        # Adam's first steps move every weight by ~lr whatever the gradient scale (and flip sign
        # with the gradient's sign), so compare the mean deviation with the total travel 3*lr.
        dev = np.abs(w - P[k]).mean() / (3 * lr)
        assert dev < (2e-3 if math == 0 else 3e-2), (k, dev)
    assert float(eng.params.grad.abs().max()) == 0.0


@pytest.mark.parametrize("V", [40, 5000])
def test_fused_embedding_adam_matches_dense_path(V):
    """ebk_embed_adam_step (row-sparse gradient summed inside the table's Adam pass) vs the dense path
    (scatter into a [V, E] gradient + ebk_adam_keras_step): same weights after 3 steps up to fp32 summation
    order.  V = 40 makes every token heavy (> 32 occurrences: pre-reduced with atomics), V = 5000 exercises
    the per-warp gather of short segments and rows without any gradient (still moved by the non-lazy Adam)."""
    E, nh, dh, att, B, H, C, T = 64, 4, 8, 24, 8, 10, 5, 12
    rng = np.random.default_rng(3)
    P, his, pred, y = make_case(rng, V, E, nh, dh, att, B, H, C, T)
    engs = [make_engine(P, V, E, T, H, nh, dh, att, 0.2, 1e-3, 1, seed=9) for _ in range(2)]
    engs[1].sparse_table_grad = False
    for e in engs:
        # Both paths sum gradients with atomics somewhere, so g differs by fp32 order noise between them; with the Keras
        # eps = 1e-7 Adam's first steps are ~lr * sign(g) and amplify that noise without bound wherever a gradient
        # nearly cancels.  A large eps makes the update ~linear in g, so the comparison measures the optimizer
        # arithmetic (gather-sum, m / v / theta update of EVERY row) instead of the noise.
        e.eps = 1e-3
    for t in range(3):
        his = rng.integers(0, V, (B, H, T)).astype(np.int32)
        his[:, 0, :] = -1 if t == 1 else his[:, 0, :]   # out-of-range ids: zero row, no gradient
        pred = rng.integers(0, V, (B, C, T)).astype(np.int32)
        losses = []
        for e in engs:
            tok, lab = e.to_device_batch(his, pred, y)
            loss, _ = e.train_step_dev(tok, lab, B, C)
            losses.append(float(loss))
        assert abs(losses[0] - losses[1]) < 5e-5 * max(1.0, abs(losses[1]))
    # Both paths sum with atomics somewhere (heavy-token pre-reduction / scatter, split-K), so gradients differ by
    # fp32 summation order; Adam turns that into ~lr * dg/(|g| + eps): ~1e-8 typically, up to ~lr for the rare element
    # whose gradient nearly cancels.  Hence a tight bound on the mean and a bound on the outlier FRACTION.
    for k, a, b in zip(O.NRMS_PARAM_ORDER, engs[0].get_weights(), engs[1].get_weights()):
        diff = np.abs(a - b)
        assert diff.mean() < 2e-7, (k, diff.mean())
        assert (diff > 1e-5).mean() < 1e-3, (k, (diff > 1e-5).mean())
        assert diff.max() < 3.5e-3, (k, diff.max())       # never more than the total Adam travel (3 steps * lr)
    for e in engs:
        assert float(e.params.grad.abs().max()) == 0.0
        assert float((e.params.m != 0).float().sum()) > 0


def test_graph_replay_matches_eager_steps(monkeypatch):
    """The single-GPU train step is captured in a CUDA graph (dropout seeds and Adam's alpha come from a
    device-resident ebk_step_params written before every replay): 5 steps with a new batch, new masks and a new
    alpha each -- eager warm-up, capture + replay, replays -- must follow the eager engine (EBK_NO_GRAPH=1)."""
    V, E, nh, dh, att, B, H, C, T = 3000, 64, 4, 8, 24, 16, 10, 5, 12
    rng = np.random.default_rng(5)
    P, _, _, y = make_case(rng, V, E, nh, dh, att, B, H, C, T)
    batches = [(rng.integers(0, V, (B, H, T)).astype(np.int32), rng.integers(0, V, (B, C, T)).astype(np.int32)) for _ in range(5)]
    results = {}
    for mode in ("graph", "eager"):
        if mode == "eager":
            monkeypatch.setenv("EBK_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("EBK_NO_GRAPH", raising=False)
        eng = make_engine(P, V, E, T, H, nh, dh, att, 0.2, 1e-3, 1, seed=9)
        eng.eps = 1e-3    # see test_fused_embedding_adam_matches_dense_path: keeps atomics-order noise un-amplified
        losses = []
        for i, (his, pred) in enumerate(batches):
            if i == 3:
                eng.lr = 5e-4                                   # ReduceLROnPlateau between steps: alpha is not baked in
            loss, _ = eng.train_step_host(his, pred, y) if i % 2 else eng.train_step_dev(*eng.to_device_batch(his, pred, y), B, C)
            losses.append(float(loss))
        results[mode] = (losses, eng.get_weights(), getattr(eng, "graph_steps", 0), eng.step_count)
    (lg, wg, ng, tg), (le, we, ne, te) = results["graph"], results["eager"]
    assert ng == 4 and ne == 0 and tg == te == 5                # step 0 eager (warm-up), steps 1-4 replayed
    assert np.allclose(lg, le, rtol=0, atol=5e-5 * max(1.0, max(abs(x) for x in le)))
    for k, a, b in zip(O.NRMS_PARAM_ORDER, wg, we):
        diff = np.abs(a - b)
        assert diff.mean() < 2e-7, (k, diff.mean())
        assert (diff > 1e-5).mean() < 1e-3, (k, (diff > 1e-5).mean())


@pytest.mark.parametrize("shape", [(20, 20, 30, 20, 100), (16, 16, 30, 7, 64), (4, 32, 17, 5, 48), (6, 24, 32, 3, 40)])
def test_fused_projection_attention_matches_unfused(monkeypatch, shape):
    """North-star kernel (QKV projection with the per-head attention in the tcgen05 GEMM's epilogue, Q|K|V saved as
    per-(sequence, head) tiles) against the unfused path (GEMM -> HBM -> attention kernel): same loss, same gradients
    up to accumulation order, with dropout, in single-CTA and CTA-pair (cta_group::2) form, incl. a ragged last tile."""
    nh, dh, T, H, E = shape
    V, att, B, C = 500, 40, 5, 3            # N = B * (H + C) sequences: not a multiple of 4 or 8 -> ragged tiles
    rng = np.random.default_rng(sum(shape))
    P, his, pred, y = make_case(rng, V, E, nh, dh, att, B, H, C, T)
    outs = {}
    for mode in ("unfused", "fused", "fused_pair"):
        monkeypatch.setenv("EBK_FUSED_ATTN", "0" if mode == "unfused" else "1")
        monkeypatch.setenv("EBK_FUSED_PAIR", "1" if mode == "fused_pair" else "0")
        monkeypatch.setenv("EBK_NO_GRAPH", "1")
        eng = make_engine(P, V, E, T, H, nh, dh, att, 0.2, 1e-3, 1, seed=2)
        tok, lab = eng.to_device_batch(his, pred, y)
        eng.params.grad.zero_()
        loss, probs = eng.loss_and_grads_dev(tok, lab, B, C, training=True, seeds=(21, 22))
        torch.cuda.synchronize()
        outs[mode] = (float(loss), probs.cpu().numpy().copy(), eng.params.grad.clone())
    l0, p0, g0 = outs["unfused"]
    for mode in ("fused", "fused_pair"):
        l, p, g = outs[mode]
        assert abs(l - l0) < 2e-5 * max(1.0, abs(l0)), (mode, l, l0)
        assert np.abs(p - p0).max() < 2e-5, mode
        gmax = float(g0.abs().max())
        assert float((g - g0).abs().max()) < 2e-4 * gmax, (mode, float((g - g0).abs().max()), gmax)


def test_scorer_dedup_matches_plain_predict():
    """Eval-mode batches repeat the history once per candidate (dataloader.py:94-107): the deduplicating
    predict path (distinct articles and histories encoded once) must reproduce the plain path's scores."""
    V, E, nh, dh, att, H, T = 300, 32, 4, 8, 24, 10, 12
    rng = np.random.default_rng(21)
    P, _, _, _ = make_case(rng, V, E, nh, dh, att, 2, H, 5, T)
    eng = make_engine(P, V, E, T, H, nh, dh, att, 0.2, 1e-3, 1, seed=1)
    pool = rng.integers(0, V, (40, T)).astype(np.int32)              # article pool: many repeats
    n_inview = [3, 7, 1, 12, 5]
    his = np.concatenate([np.repeat(pool[rng.integers(0, 40, H)][None], n, axis=0) for n in n_inview])   # [sumN, H, T]
    pred = pool[rng.integers(0, 40, his.shape[0])][:, None, :]                                           # [sumN, 1, T]
    B = his.shape[0]
    tok, _ = eng.to_device_batch(his, pred)
    plain = eng.predict_dev(tok, B, 1, head="sigmoid").cpu().numpy().copy()
    dedup = eng.predict_host_dedup(his, pred, head="sigmoid").cpu().numpy()
    rows, uniq, b, users = eng.last_dedup
    assert uniq <= 40 and users == len(n_inview) and rows == B * (H + 1)
    assert np.abs(plain - dedup).max() < 1e-6
    want = O.nrms_score(his, pred, P, nh, dh)
    assert float(np.abs(dedup - want).max() / np.abs(want).max()) < 1e-3     # and the fp32-oracle gate of the scorer


AUC_CASES = {
    # V, E, nh, dh, att, B, H, C, T, topics, topic-token fraction, lr, steps, held-out impressions
    "small": (400, 32, 4, 8, 24, 64, 8, 5, 10, 8, 0.5, 2e-3, 60, 2048),
    # xlm-roberta-base width, the BASELINE config-3 encoder (20 heads x 20, attention_hidden 200)
    "e768_d400": (20000, 768, 20, 20, 200, 32, 8, 5, 10, 16, 0.35, 1e-3, 24, 1024),
}


@pytest.mark.parametrize("case", list(AUC_CASES))
def test_auc_parity_with_oracle_trained_model(case):
    """North-star gate: AUC within +-0.002 of the reference.  The same NRMS (same initial weights, same batches,
    same dropout masks) is trained by the GPU engine (tcgen05 tf32 TMA path -- the benchmarked arithmetic) and by
    the float64 oracle (pinned to the reference source by tests/test_cpu_reference_golden.py); both models then
    score a held-out set and are compared on the reference's AucScore (mean per-impression roc_auc_score,
    evaluation/metrics_protocols.py:73-86).  The task is learnable: the clicked candidate shares tokens with the
    user's history.  "e768_d400" runs it at the BASELINE encoder dimensions (E=768, D=400, att=200)."""
    from sklearn.metrics import roc_auc_score

    V, E, nh, dh, att, B, H, C, T, n_topics, frac, lr, steps, n_held = AUC_CASES[case]
    rng = np.random.default_rng(2024)
    P, _, _, _ = make_case(rng, V, E, nh, dh, att, 2, H, C, T)
    if E >= 256:
        P["table"] = rng.normal(0, 0.02, (V, E))

    def batch(n):
        topic = rng.integers(0, n_topics, n)                            # each user reads one "topic"

        def toks(shape, tp):                                            # a fraction of the tokens carries the topic
            return np.where(rng.random(shape) < frac, tp * 50 + rng.integers(0, 50, shape),
                            rng.integers(0, V, shape)).astype(np.int32)

        his = toks((n, H, T), topic[:, None, None])
        pred = rng.integers(0, V, (n, C, T)).astype(np.int32)
        pos = rng.integers(0, C, n)
        pred[np.arange(n), pos] = toks((n, T), topic[:, None])
        y = np.zeros((n, C), np.float32)
        y[np.arange(n), pos] = 1
        return his, pred, y

    p_drop = 0.2
    eng = make_engine(P, V, E, T, H, nh, dh, att, p_drop, lr, 1, seed=5)
    Po = {k: v.copy() for k, v in P.items()}
    Pm = {k: np.zeros_like(v) for k, v in P.items()}
    Pv = {k: np.zeros_like(v) for k, v in P.items()}
    for t in range(1, steps + 1):
        his, pred, y = batch(B)
        tok, lab = eng.to_device_batch(his, pred, y)
        s1, s2 = eng.step_seeds()
        eng.train_step_dev(tok, lab, B, C)
        _, _, G = O.nrms_loss_and_grads(his, pred, y, Po, nh, dh, training=True, p_drop=p_drop, seed1=s1, seed2=s2)
        for k in Po:
            O.keras_adam_step(Po[k], G[k], Pm[k], Pv[k], t, lr)
    his, pred, y = batch(n_held)
    tok, _ = eng.to_device_batch(his, pred)
    p_gpu = eng.predict_dev(tok, n_held, C).cpu().numpy()
    p_orc = O.nrms_predict(his, pred, Po, nh, dh)
    auc = lambda p: float(np.mean([roc_auc_score(y[i], p[i]) for i in range(len(y))]))
    a_gpu, a_orc = auc(p_gpu), auc(p_orc)
    print(f"[parity] AUC {case}: GPU-trained {a_gpu:.4f} vs oracle-trained {a_orc:.4f} (|diff| {abs(a_gpu - a_orc):.4f})")
    assert 0.7 < a_orc < 0.99, a_orc               # learned, not saturated
    assert abs(a_gpu - a_orc) <= 0.002, (a_gpu, a_orc)


def test_device_feed_matches_host_feed():
    """Device-resident batch feed (article row indices + token matrix in HBM) against the host-gather feed:
    identical training losses / weights and identical predictions through the Keras-shaped facade."""
    from ebrec.models.newsrec.dataloader import NRMSDataLoader, NRMSDataLoaderDevice
    from ebrec.models.newsrec.model_config import hparams_nrms
    from ebrec.models.newsrec.nrms import NRMSModel

    class hp(hparams_nrms):
        history_size, title_size, head_num, head_dim, attention_hidden_dim, dropout = 6, 10, 4, 8, 24, 0.0

    rng = np.random.default_rng(4)
    n_art, n_imp, C = 120, 96, 5
    articles = {1000 + i: rng.integers(1, 300, 10).tolist() for i in range(n_art)}
    ids = list(articles)
    beh = {"user_id": list(range(n_imp)),
           "hist": [[ids[j] for j in rng.integers(0, n_art, 6)] for _ in range(n_imp)],
           "article_ids_inview": [[ids[j] for j in rng.integers(0, n_art, C)] for _ in range(n_imp)],
           "labels": [np.eye(C, dtype=int)[rng.integers(0, C)].tolist() for _ in range(n_imp)]}
    table = rng.standard_normal((300, 32)).astype(np.float32) * 0.3
    kw = dict(behaviors=beh, article_dict=articles, history_column="hist", unknown_representation="zeros", batch_size=32)
    runs = []
    for cls in (NRMSDataLoader, NRMSDataLoaderDevice):
        m = NRMSModel(hp, word2vec_embedding=table.copy(), seed=3)
        m._engine.eps = 1e-3   # see test_fused_embedding_adam_matches_dense_path: keeps atomics-order noise un-amplified
        h = m.model.fit(cls(**kw), epochs=2, verbose=0, shuffle=False)
        pred = m.model.predict(cls(**kw))
        sc = m.scorer.predict(cls(**dict(kw, eval_mode=True)))
        runs.append((h.history["loss"], m.model.get_weights(), pred, sc))
    (l0, w0, p0, s0), (l1, w1, p1, s1) = runs
    assert np.allclose(l0, l1, rtol=0, atol=2e-5)
    # (not bit-equal: split-K reduce-adds and the heavy-token scatter sum with atomics in arbitrary order)
    assert all(np.abs(a - b).mean() < 2e-7 and (np.abs(a - b) > 1e-5).mean() < 1e-3 for a, b in zip(w0, w1))
    assert np.abs(p0 - p1).max() < 1e-4 and p0.shape == (n_imp, C)
    assert np.abs(s0 - s1).max() < 1e-4 and s0.shape == (n_imp * C, 1)


def test_peer_table_gather_paths_match_local_gather(monkeypatch):
    """The rank-sharded-table gather (every 16-byte chunk read through its owner's mapping), its token-CSR form (every
    distinct id read once) and its chunked, GEMM-pipelined variant, exercised on ONE GPU by mapping both 'ranks' to the
    local table: loss and gradients must equal the plain local gather up to atomics order (same kernels downstream)."""
    import ctypes as C

    from ebrec.models.newsrec import _ebk

    V, E, nh, dh, att, B, H, C_, T = 3000, 32, 4, 8, 24, 32, 20, 5, 30     # R = 24 000 rows >= 4 * 4096
    rng = np.random.default_rng(8)
    P, his, pred, y = make_case(rng, V, E, nh, dh, att, B, H, C_, T)
    his[:, :, -6:] = 0                       # a padding-like id with thousands of positions (> 32: per-position path)
    pred[0, 0, :3], his[1, 2, 0] = V + 7, -1  # ids outside the table read a zero row
    lib = _ebk.lib()
    outs = []
    # "peers": one read per position; "peers_csr" (the default under data parallel): token CSR, one read per distinct id
    for mode in ("local", "peers", "peers_csr", "peers_chunked"):
        eng = make_engine(P, V, E, T, H, nh, dh, att, 0.2, 1e-3, 1, seed=2)
        tok, lab = eng.to_device_batch(his, pred, y)
        monkeypatch.setenv("EBK_DP_CSR_GATHER", "1" if mode == "peers_csr" else "0")
        if mode == "peers_chunked":
            monkeypatch.setenv("EBK_DP_CHUNKED_GATHER", "1")
        else:
            monkeypatch.delenv("EBK_DP_CHUNKED_GATHER", raising=False)
        if mode != "local":
            tbl = eng.params.offsets["news_Wqkv"]
            ptr = eng.params.theta.data_ptr()
            arr = (C.c_void_p * 2)(ptr, ptr)
            eng._force_peer_opts = (arr, 2, (tbl // 2) // 4 * 4)     # -> ebk_seqenc_opts.peer_tables of the forward
        loss, _ = eng.loss_and_grads_dev(tok, lab, B, C_, training=True, seeds=(11, 12))
        torch.cuda.synchronize()
        outs.append((float(loss), eng.params.grad.clone()))
    for l, g in outs[1:]:
        # loss and table gradient are summed with atomics (order noise); the gathered rows themselves are identical
        assert abs(l - outs[0][0]) < 1e-6 * abs(outs[0][0])
        assert float((g - outs[0][1]).abs().max()) < 1e-6
    # a forward that is not on the all-TMA path must refuse peer tables instead of reading a stale local shard
    eng = make_engine(P, V, E, T, H, nh, dh, att, 0.2, 1e-3, 0, seed=2)     # EBK_MATH_FP32
    tok, lab = eng.to_device_batch(his, pred, y)
    ptr = eng.params.theta.data_ptr()
    eng._force_peer_opts = ((C.c_void_p * 2)(ptr, ptr), 2, (eng.params.offsets["news_Wqkv"] // 2) // 4 * 4)
    with pytest.raises(_ebk.EbkError, match="peer tables need the all-TMA path"):
        eng.loss_and_grads_dev(tok, lab, B, C_, training=True, seeds=(11, 12))


def test_nrms_dummy_script_flow(capsys):
    """examples/quick_start/nrms_dummy.py statement for statement (default hparams_nrms, float64 1000x100 table,
    BATCH_SIZE 10, npratio 4): summary / fit / predict through the overlay package, nothing adapted."""
    from ebrec.models.newsrec.model_config import hparams_nrms
    from ebrec.models.newsrec.nrms import NRMSModel

    config = hparams_nrms
    BATCH_SIZE, HISTORY_SIZE, TITLE_SIZE, NPRATIO = 10, config.history_size, config.title_size, 4
    word_embeddings = np.random.rand(1000, 100)
    model = NRMSModel(hparams=config, word2vec_embedding=word_embeddings)
    model.model.summary()
    assert "860,800" in capsys.readouterr().out          # Keras' parameter count for this configuration
    his_input_title = np.random.randint(0, 1000, (BATCH_SIZE, HISTORY_SIZE, TITLE_SIZE))
    pred_input_title = np.random.randint(0, 1000, (BATCH_SIZE, NPRATIO + 1, TITLE_SIZE))
    label_data = np.zeros((BATCH_SIZE, NPRATIO + 1), dtype=int)
    for row in label_data:
        row[np.random.choice(NPRATIO + 1)] = 1
    input = (his_input_title, pred_input_title)
    hist = model.model.fit(input, label_data)
    out = model.model.predict(input)
    assert out.shape == (BATCH_SIZE, NPRATIO + 1) and np.allclose(out.sum(1), 1.0, atol=1e-5)
    assert np.isfinite(hist.history["loss"]).all()
