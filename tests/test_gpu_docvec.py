"""GPU parity of NRMSDocVec (dense + BatchNorm + dropout news encoder) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import docvec_oracle as DV, nrms_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def make(rng, Dd, units, nh, dh, att, B, H, C):
    P = DV.init_docvec_params(rng, Dd, units, nh, dh, att, dtype=np.float64)
    for i, u in enumerate(units):
        P[f"d{i}_b"] = rng.standard_normal(u) * 0.1
        P[f"d{i}_gamma"] = 1 + rng.standard_normal(u) * 0.1
        P[f"d{i}_beta"] = rng.standard_normal(u) * 0.1
        P[f"d{i}_mean"] = rng.standard_normal(u) * 0.1 + 0.3
        P[f"d{i}_var"] = 0.5 + rng.random(u)
    P["out_b"] = rng.standard_normal(nh * dh) * 0.1
    for k in ("user_WQ", "user_WK"):
        P[k] = P[k] * 6.0
    his = rng.standard_normal((B, H, Dd)).astype(np.float32)
    pred = rng.standard_normal((B, C, Dd)).astype(np.float32)
    y = np.zeros((B, C), np.float32)
    y[np.arange(B), rng.integers(0, C, B)] = 1
    return P, his, pred, y


def keras_order(P, n):
    w = []
    for i in range(n):
        w += [P[f"d{i}_W"], P[f"d{i}_b"], P[f"d{i}_gamma"], P[f"d{i}_beta"], P[f"d{i}_mean"], P[f"d{i}_var"]]
    return w + [P["out_W"], P["out_b"], P["user_WQ"], P["user_WK"], P["user_WV"], P["user_W"], P["user_b"], P["user_q"]]


def engine(P, Dd, units, H, nh, dh, att, dropout, math, l2=1e-4):
    from ebrec.models.newsrec._engine_docvec import DocVecEngine

    e = DocVecEngine(Ddoc=Dd, units=units, H=H, nh=nh, dh=dh, att=att, dropout=dropout, lr=1e-3, l2=l2, seed=3, math=math)
    e.set_weights(keras_order(P, len(units)))
    return e


CASES = [(768, [512, 512, 512], 16, 16, 200, 6, 20, 5), (64, [48, 32], 4, 8, 24, 5, 7, 3), (32, [], 2, 8, 12, 3, 4, 2)]


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_docvec_forward_scores(math, case):
    Dd, units, nh, dh, att, B, H, C = case
    P, his, pred, y = make(np.random.default_rng(sum(case[2:])), Dd, units, nh, dh, att, B, H, C)
    e = engine(P, Dd, units, H, nh, dh, att, 0.2, math)
    x, _ = e.to_device_batch(his, pred)
    probs = e.predict_dev(x, B, C).cpu().numpy()
    tol = 1e-4 if math == 0 else 1e-3
    assert rel(probs, DV.docvec_predict(his.astype(np.float64), pred.astype(np.float64), P, len(units), nh, dh)) < tol
    sig = e.predict_dev(x, B, C, head="sigmoid").cpu().numpy()
    assert rel(sig, DV.docvec_score(his.astype(np.float64), pred.astype(np.float64), P, len(units), nh, dh)) < tol


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("dropout", [0.0, 0.2])
@pytest.mark.parametrize("case", CASES[:2])
def test_docvec_loss_gradients_and_bn_stats(math, dropout, case):
    Dd, units, nh, dh, att, B, H, C = case
    P, his, pred, y = make(np.random.default_rng(sum(case[2:]) + 1), Dd, units, nh, dh, att, B, H, C)
    l2 = 1e-3
    e = engine(P, Dd, units, H, nh, dh, att, dropout, math, l2=l2)
    x, lab = e.to_device_batch(his, pred, y)
    sh, sc = 1000, 2000
    e.params.grad.zero_()
    loss, probs = e.loss_and_grads_dev(x, lab, B, C, training=True, seeds=(sh, sc))
    h64, p64 = his.astype(np.float64), pred.astype(np.float64)
    wl, wp, G, ns = DV.docvec_loss_and_grads(h64, p64, y, P, len(units), nh, dh, p_drop=dropout, seed_h=sh, seed_c=sc, l2=l2)
    ftol = 1e-4 if math == 0 else 2e-3
    assert abs(float(loss) - wl) < ftol * max(1.0, abs(wl)), (float(loss), wl)
    assert rel(probs.cpu().numpy(), wp) < 3 * ftol
    P32 = {k: v.astype(np.float32) for k, v in P.items()}
    _, _, G32, _ = DV.docvec_loss_and_grads(his, pred, y, P32, len(units), nh, dh, p_drop=dropout, seed_h=sh, seed_c=sc, l2=l2)
    btol = 2e-4 if math == 0 else 2e-2
    amp = 1.0 if math == 0 else 2.0 ** 13
    D = nh * dh
    got = {k: e.params.g(k).cpu().numpy() for k, _ in e.params.spec if not k.startswith("user_")}
    Wg = e.params.g("user_Wqkv").cpu().numpy()
    got.update(user_WQ=Wg[:, :D], user_WK=Wg[:, D:2 * D], user_WV=Wg[:, 2 * D:], user_W=e.params.g("user_attW").cpu().numpy(),
               user_b=e.params.g("user_attb").cpu().numpy(), user_q=e.params.g("user_attq").cpu().numpy().reshape(-1, 1))
    for k in DV.trainable_keys(len(units)):
        err = np.abs(got[k] - G[k]).max()
        allowed = btol * np.abs(G[k]).max() + 20 * amp * np.abs(G32[k].astype(np.float64) - G[k]).max()
        assert err <= allowed, (k, err, allowed)
    # BatchNorm moving statistics were updated twice (history call, then candidate call)
    for i in range(len(units)):
        assert rel(e.bn_mean[i].cpu().numpy(), ns[f"d{i}_mean"]) < 1e-3
        assert rel(e.bn_var[i].cpu().numpy(), ns[f"d{i}_var"]) < 1e-3


def test_docvec_facade_fit_predict_shapes():
    from ebrec.models.newsrec.model_config import hparams_nrms_docvec
    from ebrec.models.newsrec.nrms_docvec import NRMSDocVec

    class hp(hparams_nrms_docvec):
        history_size = 5
        title_size = 64
        newsencoder_units_per_layer = [32, 32]

    m = NRMSDocVec(hp, seed=1, newsencoder_units_per_layer=[32, 32])
    rng = np.random.default_rng(0)
    his, pred = rng.standard_normal((12, 5, 64)).astype(np.float32), rng.standard_normal((12, 4, 64)).astype(np.float32)
    y = np.zeros((12, 4), int)
    y[:, 0] = 1
    h = m.model.fit((his, pred), y, batch_size=4, epochs=2, verbose=0)
    assert len(h.history["loss"]) == 2 and h.history["loss"][1] < h.history["loss"][0] * 1.5
    assert m.model.predict((his, pred), batch_size=5).shape == (12, 4)
    assert m.scorer.predict((his, pred[:, :1]), batch_size=5).shape == (12, 1)
    assert len(m.model.get_weights()) == 6 * 2 + 8


def test_docvec_graph_replay_matches_eager_steps(monkeypatch):
    """The NRMSDocVec train step (about 90 launches) is replayed from a CUDA graph; the per-layer dropout seeds
    (seed1 / seed2 of the device-resident ebk_step_params + layer index) and Adam's alpha are read from device memory.
    5 steps -- eager warm-up, capture + replay, replays, with an lr change, and an evaluation on a LARGER batch in the
    middle (cached buffers are re-allocated: the captured graph must be dropped, not replayed on stale addresses) --
    must follow the eager engine (EBK_NO_GRAPH=1)."""
    Dd, units, nh, dh, att, B, H, C = 64, [48, 32], 4, 8, 24, 8, 7, 3
    rng = np.random.default_rng(11)
    P, _, _, y = make(rng, Dd, units, nh, dh, att, B, H, C)
    batches = [(rng.standard_normal((B, H, Dd)).astype(np.float32), rng.standard_normal((B, C, Dd)).astype(np.float32))
               for _ in range(6)]
    big = (rng.standard_normal((3 * B, H, Dd)).astype(np.float32), rng.standard_normal((3 * B, C, Dd)).astype(np.float32))
    res = {}
    for mode in ("graph", "eager"):
        if mode == "eager":
            monkeypatch.setenv("EBK_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("EBK_NO_GRAPH", raising=False)
        e = engine(P, Dd, units, H, nh, dh, att, 0.2, 1)
        e.loss_kind = 0
        losses = []
        for i, (his, pred) in enumerate(batches):
            if i == 2:
                e.lr = 5e-4
            if i == 4:
                xb, _ = e.to_device_batch(*big)
                e.predict_dev(xb, 3 * B, C)
            x, lab = e.to_device_batch(his, pred, y)
            loss, _ = e.train_step_dev(x, lab, B, C)
            losses.append(float(loss))
        res[mode] = (losses, e.get_weights(), getattr(e, "graph_steps", 0), e.step_count)
    (lg, wg, ng, tg), (le, we, ne, te) = res["graph"], res["eager"]
    assert ne == 0 and tg == te == 6
    assert ng == 3 + 1          # steps 1-3 replayed, the graph dropped by the big batch, step 4 eager again, step 5 replayed
    assert np.allclose(lg, le, rtol=0, atol=2e-5 * max(1.0, max(abs(v) for v in le))), (lg, le)
    for a, b in zip(wg, we):
        assert np.abs(a - b).max() < 2e-6 + 1e-5 * np.abs(b).max()
