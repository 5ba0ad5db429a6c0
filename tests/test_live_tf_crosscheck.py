"""Live cross-check against the reference running on REAL TensorFlow (last row).

Skipped unless a genuine `tensorflow` (not oracle/tf_shim) is importable AND the reference source is reachable
(/root/reference/src in the build container, or baseline/_ref).  In the images used so far neither holds
(Python 3.12, no TF wheel; probed on the GPU box too: profiles/r02_tf_probe.txt), so the standing pin of the oracle
is tests/test_cpu_reference_golden.py (reference source over the torch-backed shim).  When TF is present this
test builds the reference NRMSModel / NRMSDocVec / NAMLModel, loads the seeded weights of tests/golden/ref_cases.py
with set_weights, and compares model.predict, scorer.predict and one dropout-free train_on_batch with the oracle.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(GOLD))


def _reference_src():
    for p in (Path("/root/reference/src"), ROOT / "baseline" / "_ref"):
        if (p / "ebrec" / "models" / "newsrec" / "nrms.py").exists():
            return p
    return None


def _load_reference():
    tf = pytest.importorskip("tensorflow")
    if getattr(tf, "IS_EBK_SHIM", False) or not hasattr(tf, "function"):
        pytest.skip("only the oracle's tf_shim is importable, not TensorFlow")
    src = _reference_src()
    if src is None:
        pytest.skip("reference source not available on this machine")
    # the product overlay shadows ebrec.models.newsrec: import the reference files under a private name
    import importlib.util

    mods = {}
    for name in ("layers", "model_config", "nrms", "nrms_docvec", "base_model", "naml"):
        path = src / "ebrec" / "models" / "newsrec" / f"{name}.py"
        text = path.read_text().replace("ebrec.models.newsrec.", "_ebk_ref_newsrec_")
        spec = importlib.util.spec_from_loader(f"_ebk_ref_newsrec_{name}", loader=None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        exec(compile(text, str(path), "exec"), mod.__dict__)
        mods[name] = mod
    return tf, mods


def test_nrms_reference_on_tensorflow_matches_oracle():
    tf, R = _load_reference()
    import ref_cases as RC
    from oracle import nrms_oracle as O

    os.environ.setdefault("TF_ENABLE_ONEDNN_OPTS", "0")
    for name in ("small", "c1"):
        (V, E, nh, dh, att, B, H, C, T), ws, his, pred, y = RC.nrms_case(name)
        hp = type("hp", (R["model_config"].hparams_nrms,), dict(
            history_size=H, title_size=T, head_num=nh, head_dim=dh, attention_hidden_dim=att, dropout=0.0,
            learning_rate=1e-3))
        m = R["nrms"].NRMSModel(hp, word2vec_embedding=ws[0].astype(np.float32), seed=1)
        m.model.set_weights([w.astype(np.float32) for w in ws])
        P = dict(zip(O.NRMS_PARAM_ORDER, ws))
        got = m.model.predict((his, pred), verbose=0)
        want = O.nrms_predict(his, pred, P, nh, dh)
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-4, name          # fp32 TF vs float64 oracle
        got = m.scorer.predict((his, pred[:, :1]), verbose=0)
        assert np.abs(got - O.nrms_score(his, pred[:, :1], P, nh, dh)).max() < 1e-4
        loss = m.model.train_on_batch((his, pred), y)
        wl, _, G = O.nrms_loss_and_grads(his, pred, y, P, nh, dh, training=False)
        assert abs(float(loss) - wl) < 1e-4 * max(1.0, abs(wl))
        Pm = {k: np.zeros_like(v) for k, v in P.items()}
        Pv = {k: np.zeros_like(v) for k, v in P.items()}
        for k in P:
            O.keras_adam_step(P[k], G[k], Pm[k], Pv[k], 1, 1e-3)
        for k, w in zip(O.NRMS_PARAM_ORDER, m.model.get_weights()):
            assert np.abs(w - P[k]).mean() < 0.05 * 1e-3, k       # one Adam step of travel ~lr: same direction and size


def test_docvec_and_naml_reference_on_tensorflow_match_oracle():
    tf, R = _load_reference()
    import ref_cases as RC
    from oracle import docvec_oracle as DV, naml_oracle as NA

    c, ws, his, pred, y = RC.docvec_case()
    hp = type("hp", (R["model_config"].hparams_nrms_docvec,), dict(
        title_size=c["Ddoc"], history_size=c["H"], head_num=c["nh"], head_dim=c["dh"], attention_hidden_dim=c["att"],
        dropout=0.0, newsencoder_units_per_layer=c["units"]))
    m = R["nrms_docvec"].NRMSDocVec(hp, seed=1)
    m.model.set_weights([w.astype(np.float32) for w in ws])
    keys = []
    for i in range(len(c["units"])):
        keys += [f"d{i}_W", f"d{i}_b", f"d{i}_gamma", f"d{i}_beta", f"d{i}_mean", f"d{i}_var"]
    keys += ["out_W", "out_b", "user_WQ", "user_WK", "user_WV", "user_W", "user_b", "user_q"]
    want = DV.docvec_predict(his, pred, dict(zip(keys, ws)), len(c["units"]), c["nh"], c["dh"])
    got = m.model.predict((his.astype(np.float32), pred.astype(np.float32)), verbose=0)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-4

    c, ws, x, y = RC.naml_case()
    hp = type("hp", (R["model_config"].hparams_naml,), dict(
        title_size=c["T"], body_size=c["Tb"], history_size=c["H"], vert_num=c["vert_num"], vert_emb_dim=c["vert_dim"],
        subvert_num=c["sub_num"], subvert_emb_dim=c["sub_dim"], attention_hidden_dim=c["att"], filter_num=c["F"],
        window_size=c["window"], dropout=0.0))
    m = R["naml"].NAMLModel(hp, word2vec_embedding=ws[0].astype(np.float32), seed=1)
    m.model.set_weights([w.astype(np.float32) for w in ws])
    art, B, H, C = NA.pack_inputs(*x)
    want = NA.naml_predict(art, B, H, C, dict(zip(NA.NAML_PARAM_ORDER, ws)), c["T"], c["Tb"])
    got = m.model.predict(x, verbose=0)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-4
