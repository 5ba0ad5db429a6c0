"""GPU parity of NRMS with the optional Dense/BatchNorm/Dropout stack in the news encoder
(nrms.py:142-152, hparams.newsencoder_units_per_layer) against the oracle."""
import numpy as np
import pytest

from oracle import nrms_dense_oracle as ND, nrms_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def make(rng, V, E, units, nh, dh, att, B, H, C, T):
    P = ND.init_params(rng, V, E, units, nh, dh, att, dtype=np.float64)
    for i, u in enumerate(units):
        P[f"d{i}_b"] = rng.standard_normal(u) * 0.1
        P[f"d{i}_gamma"] = 1 + rng.standard_normal(u) * 0.1
        P[f"d{i}_beta"] = rng.standard_normal(u) * 0.1
        P[f"d{i}_mean"] = rng.standard_normal(u) * 0.1 + 0.2
        P[f"d{i}_var"] = 0.5 + rng.random(u)
    for k in ("news_b", "user_b"):
        P[k] = rng.standard_normal(att) * 0.1
    his = rng.integers(0, V, (B, H, T)).astype(np.int32)
    pred = rng.integers(0, V, (B, C, T)).astype(np.int32)
    y = np.zeros((B, C), np.float32)
    y[np.arange(B), rng.integers(0, C, B)] = 1
    return P, his, pred, y


def engine(P, V, E, T, H, units, nh, dh, att, dropout, math, l2=1e-4):
    from ebrec.models.newsrec._engine_nrms_dense import NRMSDenseEngine

    e = NRMSDenseEngine(V=V, E=E, T=T, H=H, nh=nh, dh=dh, att=att, units=units, l2=l2, dropout=dropout, lr=1e-3, seed=3,
                        math=math)
    e.set_weights([P[k] for k in ND.param_order(len(units))])
    return e


# V, E, units, nh, dh, att, B, H, C, T
CASES = [(300, 64, [48, 32], 4, 8, 24, 5, 7, 3, 12), (1000, 100, [512, 400], 20, 20, 200, 4, 20, 5, 30)]


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_forward_scores(math, case):
    V, E, units, nh, dh, att, B, H, C, T = case
    P, his, pred, y = make(np.random.default_rng(E + T), V, E, units, nh, dh, att, B, H, C, T)
    e = engine(P, V, E, T, H, units, nh, dh, att, 0.2, math)
    tok, _ = e.to_device_batch(his, pred)
    tol = 1e-4 if math == 0 else 1e-3
    assert rel(e.predict_dev(tok, B, C).cpu().numpy(), ND.predict(his, pred, P, len(units), nh, dh)) < tol
    assert rel(e.predict_dev(tok, B, C, head="sigmoid").cpu().numpy(), ND.score(his, pred, P, len(units), nh, dh)) < tol


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_loss_gradients_and_bn_stats(math, dropout):
    V, E, units, nh, dh, att, B, H, C, T = CASES[0]
    P, his, pred, y = make(np.random.default_rng(7), V, E, units, nh, dh, att, B, H, C, T)
    l2 = 1e-3
    e = engine(P, V, E, T, H, units, nh, dh, att, dropout, math, l2=l2)
    tok, lab = e.to_device_batch(his, pred, y)
    s1, sh, sc = 77, 1000, 2000
    e.params.grad.zero_()
    loss, probs = e.loss_and_grads_dev(tok, lab, B, C, training=True, seeds=(s1, sh, sc))
    ns = {}
    wl, wp, G = ND.loss_and_grads(his, pred, y, P, len(units), nh, dh, p_drop=dropout, seed1=s1, seed_h=sh, seed_c=sc, l2=l2,
                                  new_stats=ns)
    ftol = 1e-4 if math == 0 else 2e-3
    assert abs(float(loss) - wl) < ftol * max(1.0, abs(wl)), (float(loss), wl)
    assert rel(probs.cpu().numpy(), wp) < 3 * ftol
    P32 = {k: v.astype(np.float32) for k, v in P.items()}
    _, _, G32 = ND.loss_and_grads(his, pred, y, P32, len(units), nh, dh, p_drop=dropout, seed1=s1, seed_h=sh, seed_c=sc, l2=l2)
    btol = 2e-4 if math == 0 else 2e-2
    amp = 1.0 if math == 0 else 2.0 ** 13
    D = nh * dh
    got = {k: e.params.g(k).cpu().numpy() for k, _ in e.params.spec if k.startswith(("table", "d"))}
    for pre in ("news", "user"):
        Wg = e.params.g(f"{pre}_Wqkv").cpu().numpy()
        got.update({f"{pre}_WQ": Wg[:, :D], f"{pre}_WK": Wg[:, D:2 * D], f"{pre}_WV": Wg[:, 2 * D:],
                    f"{pre}_W": e.params.g(f"{pre}_attW").cpu().numpy(), f"{pre}_b": e.params.g(f"{pre}_attb").cpu().numpy(),
                    f"{pre}_q": e.params.g(f"{pre}_attq").cpu().numpy().reshape(-1, 1)})
    for k in ND.trainable_keys(len(units)):
        err = np.abs(got[k] - G[k]).max()
        allowed = btol * np.abs(G[k]).max() + 20 * amp * np.abs(G32[k].astype(np.float64) - G[k]).max()
        assert err <= allowed, (k, err, allowed)
    for i in range(len(units)):   # moving statistics updated by the history call, then by the candidate call
        assert rel(e.bn_mean[i].cpu().numpy(), ns[f"d{i}_mean"]) < 1e-3
        assert rel(e.bn_var[i].cpu().numpy(), ns[f"d{i}_var"]) < 1e-3


def test_facade_fit_predict_shapes():
    from ebrec.models.newsrec.model_config import hparams_nrms
    from ebrec.models.newsrec.nrms import NRMSModel

    class hp(hparams_nrms):
        history_size, title_size, head_num, head_dim, attention_hidden_dim = 5, 12, 4, 8, 24
        newsencoder_units_per_layer = [48, 32]
        newsencoder_l2_regularization = 1e-4

    rng = np.random.default_rng(0)
    m = NRMSModel(hp, word2vec_embedding=rng.random((200, 32)).astype(np.float32), seed=1)
    his, pred = rng.integers(0, 200, (12, 5, 12)), rng.integers(0, 200, (12, 4, 12))
    y = np.zeros((12, 4), int)
    y[:, 0] = 1
    h = m.model.fit((his, pred), y, batch_size=4, epochs=2, verbose=0)
    assert len(h.history["loss"]) == 2 and np.isfinite(h.history["loss"]).all()
    assert m.model.predict((his, pred), batch_size=5).shape == (12, 4)
    assert m.scorer.predict((his, pred[:, :1]), batch_size=5).shape == (12, 1)
    assert len(m.model.get_weights()) == 13 + 6 * 2
