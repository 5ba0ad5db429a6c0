"""Generates tests/golden/ref_*.npz by running the REFERENCE's own model source.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_reference_fixtures.py          # EBK_FIXTURE_OUT=<dir>: write there instead of tests/golden

TensorFlow is not installable here, so the reference files
    /root/reference/src/ebrec/models/newsrec/{layers,nrms,nrms_docvec,naml,base_model,model_config}.py
are imported UNMODIFIED with oracle/tf_shim (a torch-float64 stand-in for the handful of tensorflow.keras
primitives they call) in front of them on sys.path.  Everything that defines the models -- AttLayer2.call,
SelfAttention.call (incl. adjoint_a=True), the wiring of _build_nrms/_build_naml/..., which Dropout sits where --
is the reference's code; losses and gradients are taken by torch autograd through that forward.  Inputs and
weights come from tests/golden/ref_cases.py (seeded), outputs are stored float64.

If a real TensorFlow is importable it is used instead of the shim for the inference outputs
(tests/test_live_tf_crosscheck.py does that comparison live).
"""
import sys
from pathlib import Path

import numpy as np

import os

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
OUT = Path(os.environ.get("EBK_FIXTURE_OUT", HERE))   # tests/test_cpu_reference_golden.py regenerates into a temp dir
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT / "oracle" / "tf_shim"))

import ref_cases as RC  # noqa: E402
import tensorflow as tf  # noqa: E402  (the shim)
from tensorflow import _impl as SH  # noqa: E402

assert getattr(tf, "IS_EBK_SHIM", False), "this generator expects oracle/tf_shim"

from ebrec.models.newsrec.model_config import hparams_naml, hparams_nrms, hparams_nrms_docvec  # noqa: E402
from ebrec.models.newsrec.naml import NAMLModel  # noqa: E402
from ebrec.models.newsrec.nrms import NRMSModel  # noqa: E402
from ebrec.models.newsrec.nrms_docvec import NRMSDocVec  # noqa: E402

from oracle.nrms_oracle import dropout_keep_mask  # noqa: E402  (the build's own counter-based mask function)


def hp_of(base, **kw):
    return type("hp", (base,), kw)


def mask_provider(rules):
    """rules: {(dim1, last_dim): (seed, total_rows, B*H)} -> keep mask of the rows of THIS call inside the
    build's element index space [all history rows ; all candidate rows]."""
    def provide(shape, layer):
        seed, n_his, n_all, p = rules[(shape[1], shape[-1])]
        per_row = int(np.prod(shape[1:]))
        keep = dropout_keep_mask(seed, n_all * per_row, p)
        off = 0 if shape[0] == n_his else n_his
        assert shape[0] in (n_his, n_all - n_his) and n_his != n_all - n_his
        return SH._t(keep[off * per_row: (off + shape[0]) * per_row].reshape(shape).astype(np.float64))
    return provide


def nrms_fixture(name, with_grads_full):
    (V, E, nh, dh, att, B, H, C, T), ws, his, pred, y = RC.nrms_case(name)
    hp = hp_of(hparams_nrms, history_size=H, title_size=T, head_num=nh, head_dim=dh, attention_hidden_dim=att,
               dropout=0.2, learning_rate=1e-3)
    m = NRMSModel(hp, word2vec_embedding=ws[0], seed=1)
    m.model.set_weights(ws)
    out = {"probs": m.model.predict((his, pred)), "scores": m.scorer.predict((his, pred[:, :1])),
           "news_vec": m.newsencoder.predict(pred.reshape(-1, T)), "user_vec": m.userencoder.predict(his)}
    z_news = out["news_vec"].reshape(B, C, -1)
    out["logits"] = np.einsum("bcd,bd->bc", z_news, out["user_vec"])
    # training-mode loss / gradients, dropout OFF (the graph without masks)
    SH._Phase.dropout_masks = lambda shape, layer: None   # Dropout off
    loss, _, grads = m.model.loss_and_grads((his, pred), y, training=True)
    out["loss_nodrop"] = loss
    keys = ["table", "news_WQ", "news_WK", "news_WV", "news_W", "news_b", "news_q",
            "user_WQ", "user_WK", "user_WV", "user_W", "user_b", "user_q"]
    for k, g in zip(keys, grads):
        if with_grads_full:
            out[f"g_nodrop_{k}"] = g
        out[f"gp_nodrop_{k}"] = RC.probe(k, g)
        out[f"gmax_nodrop_{k}"] = np.abs(g).max()
    # training-mode with the build's counter-based masks placed where the REFERENCE places its Dropout layers
    s1, s2 = RC.DROPOUT_SEEDS
    N = B * (H + C)
    SH._Phase.dropout_masks = mask_provider({(T, E): (s1, B * H, N, 0.2), (T, nh * dh): (s2, B * H, N, 0.2)})
    loss, _, grads = m.model.loss_and_grads((his, pred), y, training=True)
    out["loss_drop"] = loss
    for k, g in zip(keys, grads):
        if with_grads_full:
            out[f"g_drop_{k}"] = g
        out[f"gp_drop_{k}"] = RC.probe(k, g)
        out[f"gmax_drop_{k}"] = np.abs(g).max()
    if with_grads_full:
        # hparams.loss = "log_loss" -> binary_crossentropy on the softmax output (nrms.py:63-64)
        SH._Phase.dropout_masks = lambda shape, layer: None   # Dropout off
        hp2 = hp_of(hp, loss="log_loss")
        m2 = NRMSModel(hp2, word2vec_embedding=ws[0], seed=1)
        m2.model.set_weights(ws)
        assert m2.model.loss == "binary_crossentropy"
        loss, _, grads = m2.model.loss_and_grads((his, pred), y, training=True)
        out["loss_logloss"] = loss
        for k, g in zip(keys, grads):
            out[f"g_logloss_{k}"] = g
        # two optimizer steps (Keras-form Adam as restated in the shim) without dropout
        SH._Phase.dropout_masks = lambda shape, layer: None   # Dropout off
        m.model.optimizer = tf.keras.optimizers.Adam(learning_rate=1e-3)
        out["train_losses"] = np.array([m.model.train_on_batch((his, pred), y) for _ in range(2)])
        for k, w in zip(keys, m.model.get_weights()):
            out[f"w2_{k}"] = w
    SH._Phase.dropout_masks = None
    spread = float(np.ptp(out["logits"], axis=1).min())
    np.savez_compressed(OUT / f"ref_nrms_{name}.npz", **out)
    assert spread >= 1.0, f"{name}: logit spread {spread:.2f} < 1 would make the score gate vacuous"
    print(f"ref_nrms_{name}: loss {out['loss_nodrop']:.6f} / {out['loss_drop']:.6f}, min per-row logit spread {spread:.2f}")


def nrms_dense_fixture():
    c, ws, his, pred, y = RC.nrms_dense_case()
    hp = hp_of(hparams_nrms, history_size=c["H"], title_size=c["T"], head_num=c["nh"], head_dim=c["dh"],
               attention_hidden_dim=c["att"], dropout=0.0, learning_rate=1e-3,
               newsencoder_units_per_layer=c["units"], newsencoder_l2_regularization=1e-3)
    m = NRMSModel(hp, word2vec_embedding=ws[0], seed=1)
    m.model.set_weights(ws)
    out = {"probs": m.model.predict((his, pred)), "scores": m.scorer.predict((his, pred[:, :1]))}
    loss, _, grads = m.model.loss_and_grads((his, pred), y, training=True)     # BatchNorm in batch-statistics mode
    out["loss"] = loss
    for i, g in enumerate(grads):
        out[f"g_{i}"] = g
    for i, w in enumerate(m.model.get_weights()):                               # moving statistics after the step
        out[f"w_after_{i}"] = w
    np.savez_compressed(OUT / "ref_nrms_dense.npz", **out)
    print(f"ref_nrms_dense: loss {loss:.6f}, {len(grads)} arrays")


def docvec_fixture():
    c, ws, his, pred, y = RC.docvec_case()
    hp = hp_of(hparams_nrms_docvec, title_size=c["Ddoc"], history_size=c["H"], head_num=c["nh"], head_dim=c["dh"],
               attention_hidden_dim=c["att"], dropout=0.0, learning_rate=1e-3,
               newsencoder_units_per_layer=c["units"], newsencoder_l2_regularization=1e-3)
    m = NRMSDocVec(hp, seed=1)
    m.model.set_weights(ws)
    out = {"probs": m.model.predict((his, pred)), "scores": m.scorer.predict((his, pred[:, :1]))}
    loss, _, grads = m.model.loss_and_grads((his, pred), y, training=True)
    out["loss"] = loss
    for i, g in enumerate(grads):
        out[f"g_{i}"] = g
    for i, w in enumerate(m.model.get_weights()):
        out[f"w_after_{i}"] = w
    np.savez_compressed(OUT / "ref_docvec.npz", **out)
    print(f"ref_docvec: loss {loss:.6f}, {len(grads)} arrays, shapes {[g.shape for g in grads]}")


def naml_fixture():
    c, ws, x, y = RC.naml_case()
    hp = hp_of(hparams_naml, title_size=c["T"], body_size=c["Tb"], history_size=c["H"], vert_num=c["vert_num"],
               vert_emb_dim=c["vert_dim"], subvert_num=c["sub_num"], subvert_emb_dim=c["sub_dim"],
               attention_hidden_dim=c["att"], filter_num=c["F"], window_size=c["window"], dropout=0.2,
               learning_rate=1e-3)
    m = NAMLModel(hp, word2vec_embedding=ws[0], seed=1)
    print("NAML get_weights shapes:", [w.shape for w in m.model.get_weights()])
    m.model.set_weights(ws)
    out = {"probs": m.model.predict(x), "scores": m.scorer.predict(x[:4] + [a[:, :1] for a in x[4:]])}
    SH._Phase.dropout_masks = lambda shape, layer: None   # Dropout off
    loss, _, grads = m.model.loss_and_grads(x, y, training=True)
    out["loss_nodrop"] = loss
    for i, g in enumerate(grads):
        out[f"g_nodrop_{i}"] = g
    B, H, C, T, Tb, E, F = c["B"], c["H"], c["C"], c["T"], c["Tb"], c["E"], c["F"]
    N = B * (H + C)
    seeds = (101, 102, 103, 104)
    SH._Phase.dropout_masks = mask_provider({(T, E): (seeds[0], B * H, N, 0.2), (T, F): (seeds[1], B * H, N, 0.2),
                                             (Tb, E): (seeds[2], B * H, N, 0.2), (Tb, F): (seeds[3], B * H, N, 0.2)})
    loss, _, grads = m.model.loss_and_grads(x, y, training=True)
    out["loss_drop"] = loss
    out["drop_seeds"] = np.array(seeds)
    for i, g in enumerate(grads):
        out[f"g_drop_{i}"] = g
    SH._Phase.dropout_masks = None
    np.savez_compressed(OUT / "ref_naml.npz", **out)
    print(f"ref_naml: loss {out['loss_nodrop']:.6f} / {out['loss_drop']:.6f}")


if __name__ == "__main__":
    nrms_fixture("small", True)
    for n in ("c1", "e768", "h50"):
        nrms_fixture(n, False)
    nrms_dense_fixture()
    docvec_fixture()
    naml_fixture()
