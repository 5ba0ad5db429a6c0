"""GPU parity of NAML (conv text views + categorical views + AttLayer2 fusion) against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import naml_oracle as NA, nrms_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


# V, E, F, att, window, T, Tb, n_vert, d_vert, n_sub, d_sub, B, H, C
CASES = [
    (1000, 100, 400, 200, 3, 30, 40, 100, 10, 100, 10, 6, 20, 5),   # naml_dummy.py shapes (smaller batch)
    (60, 16, 24, 12, 3, 6, 9, 7, 3, 9, 5, 4, 5, 3),
    (50, 12, 16, 8, 4, 5, 7, 5, 2, 6, 3, 3, 4, 2),                  # even window: asymmetric 'same' padding
    (40, 8, 12, 8, 1, 4, 5, 4, 2, 4, 2, 2, 3, 2),                   # window 1 (pointwise conv)
]


def make(case, seed=0):
    V, E, F, att, w, T, Tb, nv, dv, ns, ds, B, H, C = case
    rng = np.random.default_rng(seed)
    P = NA.init_naml_params(rng, V, E, F, att, w, nv, dv, ns, ds, dtype=np.float64)
    for k in P:
        if k.endswith("_b") or k.endswith("convb") or k.endswith("denseb"):
            P[k] = rng.standard_normal(P[k].shape) * 0.1
    P["vert_emb"] *= 10
    P["subvert_emb"] *= 10
    arrays = [rng.integers(0, V, (B, H, T)), rng.integers(0, V, (B, H, Tb)), rng.integers(0, nv, (B, H, 1)),
              rng.integers(0, ns, (B, H, 1)), rng.integers(0, V, (B, C, T)), rng.integers(0, V, (B, C, Tb)),
              rng.integers(0, nv, (B, C, 1)), rng.integers(0, ns, (B, C, 1))]
    # out-of-range ids: zero embedding row, no gradient
    arrays[0][0, 0, 0] = V + 3
    arrays[5][0, 0, 1] = -1
    arrays[2][0, 1, 0] = nv + 1
    y = np.zeros((B, C), np.float32)
    y[np.arange(B), rng.integers(0, C, B)] = 1
    return P, arrays, y


def engine(case, P, dropout, math):
    from ebrec.models.newsrec._engine_naml import NAML_WEIGHT_ORDER, NAMLEngine

    V, E, F, att, w, T, Tb, nv, dv, ns, ds, B, H, C = case
    e = NAMLEngine(V=V, E=E, T=T, Tb=Tb, H=H, F=F, att=att, window=w, vert_num=nv, vert_dim=dv, subvert_num=ns,
                   subvert_dim=ds, dropout=dropout, lr=1e-3, seed=3, math=math)
    assert list(NAML_WEIGHT_ORDER) == list(NA.NAML_PARAM_ORDER)
    e.set_weights([P[k] for k in NA.NAML_PARAM_ORDER])
    return e


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_naml_forward_scores(math, case):
    P, arrays, y = make(case)
    B, H, C, T, Tb = case[11], case[12], case[13], case[5], case[6]
    e = engine(case, P, 0.2, math)
    x, _ = e.to_device_batch(arrays)
    art, *_ = NA.pack_inputs(*arrays)
    tol = 1e-4 if math == 0 else 1e-3  # north star: click scores within 1e-3 relative
    probs = e.predict_dev(x, B, C).cpu().numpy()
    assert rel(probs, NA.naml_predict(art, B, H, C, P, T, Tb)) < tol
    sig = e.predict_dev(x, B, C, head="sigmoid").cpu().numpy()
    assert rel(sig, NA.naml_score(art, B, H, C, P, T, Tb)) < tol


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("dropout", [0.0, 0.2])
@pytest.mark.parametrize("case", CASES)
def test_naml_loss_and_gradients(math, dropout, case):
    P, arrays, y = make(case, seed=1)
    B, H, C, T, Tb = case[11], case[12], case[13], case[5], case[6]
    e = engine(case, P, dropout, math)
    x, lab = e.to_device_batch(arrays, y)
    seeds = (101, 202, 303, 404)
    e.params.grad.zero_()
    loss, probs = e.loss_and_grads_dev(x, lab, B, C, training=True, seeds=seeds)
    art, *_ = NA.pack_inputs(*arrays)
    wl, wp, G = NA.naml_loss_and_grads(art, B, H, C, y, P, T, Tb, p_drop=dropout, seeds=seeds)
    ftol = 1e-4 if math == 0 else 2e-3
    assert abs(float(loss) - wl) < ftol * max(1.0, abs(wl)), (float(loss), wl)
    assert rel(probs.cpu().numpy(), wp) < 3 * ftol
    P32 = {k: v.astype(np.float32) for k, v in P.items()}
    _, _, G32 = NA.naml_loss_and_grads(art, B, H, C, y, P32, T, Tb, p_drop=dropout, seeds=seeds)
    btol = 2e-4 if math == 0 else 2e-2
    amp = 1.0 if math == 0 else 2.0 ** 13   # tf32 keeps 10 of fp32's 23 mantissa bits
    for k in NA.NAML_PARAM_ORDER:
        got = e.params.g(k).cpu().numpy().reshape(G[k].shape)
        err = np.abs(got - G[k]).max()
        allowed = btol * np.abs(G[k]).max() + 20 * amp * np.abs(G32[k].astype(np.float64) - G[k]).max()
        assert err <= allowed, (k, err, allowed)
    # bit-exact gather indices: rows of the table that no token touched have EXACTLY zero gradient
    gt = e.params.g("table").cpu().numpy()
    V = case[0]
    used = np.zeros(V, bool)
    toks = np.concatenate([art[:, :T + Tb].reshape(-1)])
    used[toks[(toks >= 0) & (toks < V)]] = True
    assert not gt[~used].any()


def test_naml_train_step_matches_oracle_adam():
    case = CASES[1]
    P, arrays, y = make(case, seed=2)
    B, H, C, T, Tb = case[11], case[12], case[13], case[5], case[6]
    e = engine(case, P, 0.0, 0)
    x, lab = e.to_device_batch(arrays, y)
    art, *_ = NA.pack_inputs(*arrays)
    P64 = {k: v.copy() for k, v in P.items()}
    m = {k: np.zeros_like(v) for k, v in P.items()}
    v_ = {k: np.zeros_like(v) for k, v in P.items()}
    for t in range(1, 4):
        e.train_step_dev(x, lab, B, C)
        _, _, G = NA.naml_loss_and_grads(art, B, H, C, y, P64, T, Tb)
        for k in P64:
            O.keras_adam_step(P64[k], G[k], m[k], v_[k], t, 1e-3)
    for k, got in zip(NA.NAML_PARAM_ORDER, e.get_weights()):
        # Adam's first steps move every touched weight by ~lr regardless of gradient size
        assert np.abs(got.reshape(P64[k].shape) - P64[k]).max() < 2e-4, k


def test_naml_facade_dummy_script_shapes():
    """examples/quick_start/naml_dummy.py with a smaller batch: fit + predict + scorer on the eight arrays."""
    from ebrec.models.newsrec import NAMLModel
    from ebrec.models.newsrec.model_config import hparams_naml

    rng = np.random.default_rng(0)
    emb = rng.random((1000, 100))
    m = NAMLModel(hparams=hparams_naml, word2vec_embedding=emb, seed=1)
    Bn, H, T, Tb, Cn = 24, hparams_naml.history_size, hparams_naml.title_size, hparams_naml.body_size, 5
    inp = (rng.integers(0, 1000, (Bn, H, T)), rng.integers(0, 1000, (Bn, H, Tb)), rng.integers(0, 100, (Bn, H, 1)),
           rng.integers(0, 100, (Bn, H, 1)), rng.integers(0, 1000, (Bn, Cn, T)), rng.integers(0, 1000, (Bn, Cn, Tb)),
           rng.integers(0, 100, (Bn, Cn, 1)), rng.integers(0, 100, (Bn, Cn, 1)))
    y = np.zeros((Bn, Cn), int)
    y[np.arange(Bn), rng.integers(0, Cn, Bn)] = 1
    m.model.summary(print_fn=lambda s: None)
    h = m.model.fit(inp, y, batch_size=8, epochs=2, verbose=0)
    assert len(h.history["loss"]) == 2 and np.isfinite(h.history["loss"]).all()
    p = m.model.predict(inp, batch_size=10)
    assert p.shape == (Bn, Cn) and np.allclose(p.sum(-1), 1, atol=1e-4)
    one = tuple(a[:, :1] if i >= 4 else a for i, a in enumerate(inp))
    s = m.scorer.predict(one, batch_size=10)
    assert s.shape == (Bn, 1) and ((s > 0) & (s < 1)).all()
    assert len(m.model.get_weights()) == 23
    nv = m.newsencoder.predict(np.concatenate([inp[4][:, 0], inp[5][:, 0], inp[6][:, 0], inp[7][:, 0]], axis=-1))
    assert nv.shape == (Bn, hparams_naml.filter_num)
    with pytest.raises(ValueError):
        class bad(hparams_naml):
            loss = "nope"
        NAMLModel(hparams=bad, word2vec_embedding=emb)


def test_attlayer_strided_output_and_errors():
    from ebrec.models.newsrec import _ebk

    lib = _ebk.lib()
    rng = np.random.default_rng(5)
    n, L, D, att = 7, 4, 16, 8
    X = rng.standard_normal((n, L, D)).astype(np.float32)
    W, b, q = (rng.standard_normal(s).astype(np.float32) for s in ((D, att), (att,), (att, 1)))
    want, _ = O.att_layer2_fwd(X.astype(np.float64), W.astype(np.float64), b.astype(np.float64), q.astype(np.float64))
    d = _ebk.AttLayerDesc(n, L, D, att, 0.0, 0)
    ws = torch.empty(lib.ebk_attlayer_workspace_bytes(C.byref(d)), dtype=torch.uint8, device="cuda")
    out = torch.full((n, 3, D), -7.0, device="cuda")
    t = [torch.from_numpy(a).cuda() for a in (X, W, b, q.reshape(-1))]
    _ebk.check(lib.ebk_attlayer_fwd(C.byref(d), _ebk.ptr(t[0]), _ebk.ptr(t[1]), _ebk.ptr(t[2]), _ebk.ptr(t[3]), 0, 0,
                                    _ebk.ptr(ws), ws.numel(), C.c_void_p(out.data_ptr() + 4 * D), 3 * D, _ebk.stream()))
    o = out.cpu().numpy()
    assert rel(o[:, 1], want) < 1e-5 and (o[:, 0] == -7).all() and (o[:, 2] == -7).all()
    assert lib.ebk_attlayer_fwd(C.byref(d), _ebk.ptr(t[0]), _ebk.ptr(t[1]), _ebk.ptr(t[2]), _ebk.ptr(t[3]), 0, 0,
                                _ebk.ptr(ws), 16, _ebk.ptr(out), 3 * D, _ebk.stream()) == -2  # EBK_ERR_WORKSPACE
    bad = _ebk.AttLayerDesc(n, 65, D, att, 0.0, 0)
    assert lib.ebk_attlayer_workspace_bytes(C.byref(bad)) == 0
