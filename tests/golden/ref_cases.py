"""Seeded inputs / weights of the reference-generated golden cases.

Shared by tests/golden/make_reference_fixtures.py (which runs the REFERENCE's own model source over
oracle/tf_shim and stores its outputs in ref_*.npz) and by the tests that compare the oracle and the CUDA path
with those outputs.  Weights and inputs are regenerated from the seed (numpy Generator streams are stable), so the
committed fixtures only hold the reference's outputs.
"""
from __future__ import annotations

import numpy as np

# name: V, E, nh, dh, att, B, H, C, T
NRMS_CASES = {
    "small": (300, 48, 4, 8, 24, 6, 7, 5, 12),
    "c1": (1000, 100, 20, 20, 200, 4, 20, 5, 30),        # nrms_dummy.py shape (BASELINE config 1)
    "e768": (2000, 768, 20, 20, 200, 2, 20, 5, 30),      # xlm-roberta-base width (BASELINE config 3), small vocab
    "h50": (500, 64, 20, 20, 200, 2, 50, 5, 30),         # ebnerd_large history (BASELINE config 4)
}
# news-encoder WV scale: chosen so that every impression's logit spread (max - min over candidates) is >= 1
WV_SCALE = {"small": 1.0, "c1": 2.0, "e768": 2.5, "h50": 2.5}
DROPOUT_SEEDS = (0x1234567, 0x7654321)


def glorot(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[-1]))
    return rng.uniform(-lim, lim, size=shape)


def nrms_case(name):
    """-> (dims, weights in Keras get_weights order (float64), his, pred, y).  Scaled so that the per-impression
    logit spread is O(1) (a softmax over near-identical logits would make every score test vacuous)."""
    V, E, nh, dh, att, B, H, C, T = NRMS_CASES[name]
    rng = np.random.default_rng([7, V, E, H])
    D = nh * dh
    table = rng.standard_normal((V, E)) * (1.0 if E < 200 else 0.5)
    ws = [table]
    for din in (E, D):
        ws += [glorot(rng, (din, D)) * 3.0, glorot(rng, (din, D)) * 3.0,
               glorot(rng, (din, D)) * (WV_SCALE[name] if din == E else 1.0),
               glorot(rng, (D, att)), rng.standard_normal(att) * 0.05, glorot(rng, (att, 1)) * 2.0]
    his = rng.integers(0, V, (B, H, T))
    pred = rng.integers(0, V, (B, C, T))
    y = np.zeros((B, C))
    y[np.arange(B), rng.integers(0, C, B)] = 1.0
    return (V, E, nh, dh, att, B, H, C, T), ws, his, pred, y


# NRMS with the optional Dense/BatchNorm/Dropout stack (nrms.py:142-152)
DENSE_CASE = dict(V=200, E=32, nh=4, dh=8, att=24, B=5, H=6, C=4, T=10, units=[40, 32])


def nrms_dense_case():
    c = DENSE_CASE
    rng = np.random.default_rng(41)
    D = c["nh"] * c["dh"]
    ws = [rng.standard_normal((c["V"], c["E"]))]
    ws += [glorot(rng, (c["E"], D)) * 3.0, glorot(rng, (c["E"], D)) * 3.0, glorot(rng, (c["E"], D))]
    din = D
    for u in c["units"]:
        ws += [glorot(rng, (din, u)) * 2.0, rng.standard_normal(u) * 0.1, 1.0 + 0.1 * rng.standard_normal(u),
               0.1 * rng.standard_normal(u), 0.1 * rng.standard_normal(u), 1.0 + 0.2 * rng.random(u)]
        din = u
    ws += [glorot(rng, (din, c["att"])), rng.standard_normal(c["att"]) * 0.05, glorot(rng, (c["att"], 1)) * 2.0]
    ws += [glorot(rng, (D, D)) * 3.0, glorot(rng, (D, D)) * 3.0, glorot(rng, (D, D)) * 0.5,
           glorot(rng, (D, c["att"])), rng.standard_normal(c["att"]) * 0.05, glorot(rng, (c["att"], 1)) * 2.0]
    his = rng.integers(0, c["V"], (c["B"], c["H"], c["T"]))
    pred = rng.integers(0, c["V"], (c["B"], c["C"], c["T"]))
    y = np.zeros((c["B"], c["C"]))
    y[np.arange(c["B"]), rng.integers(0, c["C"], c["B"])] = 1.0
    return c, ws, his, pred, y


# NRMSDocVec (nrms_docvec.py): Ddoc -> units -> D
DOCVEC_CASE = dict(Ddoc=48, units=[64, 40], nh=4, dh=8, att=24, B=6, H=5, C=4)


def docvec_case():
    c = DOCVEC_CASE
    rng = np.random.default_rng(43)
    D = c["nh"] * c["dh"]
    ws, din = [], c["Ddoc"]
    for u in c["units"]:
        ws += [glorot(rng, (din, u)) * 2.0, rng.standard_normal(u) * 0.1, 1.0 + 0.1 * rng.standard_normal(u),
               0.1 * rng.standard_normal(u), 0.1 * rng.standard_normal(u), 1.0 + 0.2 * rng.random(u)]
        din = u
    ws += [glorot(rng, (din, D)) * 0.6, rng.standard_normal(D) * 0.1]
    ws += [glorot(rng, (D, D)) * 3.0, glorot(rng, (D, D)) * 3.0, glorot(rng, (D, D)),
           glorot(rng, (D, c["att"])), rng.standard_normal(c["att"]) * 0.05, glorot(rng, (c["att"], 1)) * 2.0]
    his = rng.standard_normal((c["B"], c["H"], c["Ddoc"]))
    pred = rng.standard_normal((c["B"], c["C"], c["Ddoc"]))
    y = np.zeros((c["B"], c["C"]))
    y[np.arange(c["B"]), rng.integers(0, c["C"], c["B"])] = 1.0
    return c, ws, his, pred, y


# NAML (naml.py)
NAML_CASE = dict(V=250, E=24, F=32, att=20, window=3, vert_num=12, vert_dim=6, sub_num=15, sub_dim=6,
                 B=4, H=6, C=5, T=10, Tb=14)


def naml_case():
    c = NAML_CASE
    rng = np.random.default_rng(47)
    E, F, att, w = c["E"], c["F"], c["att"], c["window"]
    ws = [rng.random((c["V"], E))]                                   # shared word table (base_model.py:44)
    for _view in ("title", "body"):                                   # Conv1D kernel, bias, AttLayer2 W, b, q
        ws += [glorot(rng, (w * E, F)).reshape(w, E, F) * 1.5, rng.standard_normal(F) * 0.1,
               glorot(rng, (F, att)), rng.standard_normal(att) * 0.05, glorot(rng, (att, 1)) * 2.0]
    for n, d in ((c["vert_num"], c["vert_dim"]), (c["sub_num"], c["sub_dim"])):   # Embedding, Dense kernel, bias
        ws += [rng.uniform(-0.5, 0.5, (n, d)), glorot(rng, (d, F)) * 2.0, rng.standard_normal(F) * 0.1]
    for _v in ("news", "user"):                                       # fusion AttLayer2, user AttLayer2
        ws += [glorot(rng, (F, att)), rng.standard_normal(att) * 0.05, glorot(rng, (att, 1)) * 2.0]
    B, H, C, T, Tb = c["B"], c["H"], c["C"], c["T"], c["Tb"]
    x = [rng.integers(0, c["V"], (B, H, T)), rng.integers(0, c["V"], (B, H, Tb)),
         rng.integers(0, c["vert_num"], (B, H, 1)), rng.integers(0, c["sub_num"], (B, H, 1)),
         rng.integers(0, c["V"], (B, C, T)), rng.integers(0, c["V"], (B, C, Tb)),
         rng.integers(0, c["vert_num"], (B, C, 1)), rng.integers(0, c["sub_num"], (B, C, 1))]
    y = np.zeros((B, C))
    y[np.arange(B), rng.integers(0, C, B)] = 1.0
    return c, ws, x, y


def probe(name: str, arr: np.ndarray) -> float:
    """Fixed random projection of an array (stores one float instead of a large gradient)."""
    rng = np.random.default_rng([len(name), *[ord(ch) for ch in name[:8]], *arr.shape])
    return float((arr * rng.standard_normal(arr.shape)).sum())
