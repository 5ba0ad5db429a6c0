"""NumPy restatement of NAML (reference src/ebrec/models/newsrec/naml.py:62-374).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED (no reference golden vectors).

News encoder (naml.py:91-141) per article, input [title(T) | body(Tb) | vert | subvert] int32:
  title/body (naml.py:143-203): shared Embedding -> Dropout -> Conv1D(F, window, 'same', relu) -> Dropout -> AttLayer2
  vert/subvert (naml.py:205-252): Embedding(n,10) -> Dense(F, relu)
  concat the four [F] views -> AttLayer2 (naml.py:133-138)
User encoder (naml.py:62-89): TimeDistributed(news) -> AttLayer2.   Score/loss as NRMS (naml.py:254-374).
Keras Conv1D 'same' with odd window w pads (w-1)/2 zeros on both sides; kernel shape [w, E, F].
"""
from __future__ import annotations

import numpy as np

from . import nrms_oracle as O

VIEWS = ("title", "body")


def init_naml_params(rng, V, E, F, att, window, vert_num, vert_dim, sub_num, sub_dim, dtype=np.float32, table=None) -> dict:
    P = {"table": (table if table is not None else rng.random((V, E))).astype(dtype)}  # base_model.py:44
    for v in VIEWS:
        P[f"{v}_convW"] = O.glorot_uniform(rng, (window * E, F), dtype).reshape(window, E, F)
        P[f"{v}_convb"] = np.zeros((F,), dtype)
        P[f"{v}_W"] = O.glorot_uniform(rng, (F, att), dtype)
        P[f"{v}_b"] = np.zeros((att,), dtype)
        P[f"{v}_q"] = O.glorot_uniform(rng, (att, 1), dtype)
    for v, n, d in (("vert", vert_num, vert_dim), ("subvert", sub_num, sub_dim)):
        P[f"{v}_emb"] = rng.uniform(-0.05, 0.05, (n, d)).astype(dtype)  # Keras Embedding default init
        P[f"{v}_denseW"] = O.glorot_uniform(rng, (d, F), dtype)
        P[f"{v}_denseb"] = np.zeros((F,), dtype)
    for v in ("news", "user"):
        P[f"{v}_W"] = O.glorot_uniform(rng, (F, att), dtype)
        P[f"{v}_b"] = np.zeros((att,), dtype)
        P[f"{v}_q"] = O.glorot_uniform(rng, (att, 1), dtype)
    return P


NAML_PARAM_ORDER = (["table"] + [f"{v}_{s}" for v in VIEWS for s in ("convW", "convb", "W", "b", "q")]
                    + [f"{v}_{s}" for v in ("vert", "subvert") for s in ("emb", "denseW", "denseb")]
                    + [f"{v}_{s}" for v in ("news", "user") for s in ("W", "b", "q")])


def conv1d_same_fwd(X, Wc, bc):
    """X [N, L, E], Wc [w, E, F] -> relu(conv) [N, L, F]."""
    w = Wc.shape[0]
    pad = (w - 1) // 2
    N, L, E = X.shape
    Xp = np.zeros((N, L + w - 1, E), X.dtype)
    Xp[:, pad:pad + L] = X
    z = sum(Xp[:, j:j + L] @ Wc[j] for j in range(w)) + bc
    return np.maximum(z, 0), Xp


def conv1d_same_bwd(dy, y, Xp, Wc):
    w = Wc.shape[0]
    pad = (w - 1) // 2
    N, L, F = dy.shape
    dz = dy * (y > 0)
    dWc = np.stack([np.einsum("nle,nlf->ef", Xp[:, j:j + L], dz) for j in range(w)])
    dbc = dz.sum(axis=(0, 1))
    dXp = np.zeros_like(Xp)
    for j in range(w):
        dXp[:, j:j + L] += dz @ Wc[j].T
    return dXp[:, pad:pad + L], dWc, dbc


def text_view_fwd(tok, P, view, *, training, p_drop, seed1, seed2):
    table = P["table"]
    V = table.shape[0]
    inb = (tok >= 0) & (tok < V)
    X = table[np.where(inb, tok, 0)] * inb[..., None].astype(table.dtype)
    keep1 = keep2 = None
    if training and p_drop > 0:
        X, keep1 = O.dropout_fwd(X, seed1, p_drop)  # naml.py:158 / 189
    y, Xp = conv1d_same_fwd(X, P[f"{view}_convW"], P[f"{view}_convb"])  # naml.py:159-166
    yd = y
    if training and p_drop > 0:
        yd, keep2 = O.dropout_fwd(y, seed2, p_drop)  # naml.py:167
    out, c_att = O.att_layer2_fwd(yd, P[f"{view}_W"], P[f"{view}_b"], P[f"{view}_q"])  # naml.py:168
    return out, (tok, inb, keep1, keep2, y, Xp, c_att, p_drop, view)


def text_view_bwd(dout, cache, P, grads):
    tok, inb, keep1, keep2, y, Xp, c_att, p_drop, view = cache
    dyd, dW, db, dq = O.att_layer2_bwd(dout, c_att)
    grads[f"{view}_W"] += dW
    grads[f"{view}_b"] += db
    grads[f"{view}_q"] += dq
    s = dyd.dtype.type(1.0 / (1.0 - p_drop)) if keep2 is not None else None
    dy = dyd * keep2 * s if keep2 is not None else dyd
    dX, dWc, dbc = conv1d_same_bwd(dy, y, Xp, P[f"{view}_convW"])
    grads[f"{view}_convW"] += dWc
    grads[f"{view}_convb"] += dbc
    if keep1 is not None:
        dX = dX * keep1 * s
    dX = dX * inb[..., None].astype(dX.dtype)
    np.add.at(grads["table"], np.where(inb, tok, 0).reshape(-1), dX.reshape(-1, dX.shape[-1]))


def cat_view_fwd(ids, P, view):
    emb = P[f"{view}_emb"]
    n = emb.shape[0]
    inb = (ids >= 0) & (ids < n)
    x = emb[np.where(inb, ids, 0)] * inb[:, None].astype(emb.dtype)
    y = np.maximum(x @ P[f"{view}_denseW"] + P[f"{view}_denseb"], 0)  # naml.py:217-223
    return y, (ids, inb, x, y, view)


def cat_view_bwd(dy, cache, P, grads):
    ids, inb, x, y, view = cache
    dz = dy * (y > 0)
    grads[f"{view}_denseW"] += x.T @ dz
    grads[f"{view}_denseb"] += dz.sum(0)
    dx = (dz @ P[f"{view}_denseW"].T) * inb[:, None].astype(dz.dtype)
    np.add.at(grads[f"{view}_emb"], np.where(inb, ids, 0), dx)


def news_encoder_fwd(art, P, T, Tb, *, training=False, p_drop=0.0, seeds=(0, 0, 0, 0)):
    """art [N, T+Tb+2] int -> [N, F]."""
    t_out, c_t = text_view_fwd(art[:, :T], P, "title", training=training, p_drop=p_drop, seed1=seeds[0], seed2=seeds[1])
    b_out, c_b = text_view_fwd(art[:, T:T + Tb], P, "body", training=training, p_drop=p_drop, seed1=seeds[2], seed2=seeds[3])
    v_out, c_v = cat_view_fwd(art[:, T + Tb], P, "vert")
    s_out, c_s = cat_view_fwd(art[:, T + Tb + 1], P, "subvert")
    cat = np.stack([t_out, b_out, v_out, s_out], axis=1)  # Concatenate(axis=-2), naml.py:133-135
    out, c_att = O.att_layer2_fwd(cat, P["news_W"], P["news_b"], P["news_q"])
    return out, (c_t, c_b, c_v, c_s, c_att)


def news_encoder_bwd(dout, cache, P, grads):
    c_t, c_b, c_v, c_s, c_att = cache
    dcat, dW, db, dq = O.att_layer2_bwd(dout, c_att)
    grads["news_W"] += dW
    grads["news_b"] += db
    grads["news_q"] += dq
    text_view_bwd(dcat[:, 0], c_t, P, grads)
    text_view_bwd(dcat[:, 1], c_b, P, grads)
    cat_view_bwd(dcat[:, 2], c_v, P, grads)
    cat_view_bwd(dcat[:, 3], c_s, P, grads)


def pack_inputs(his_title, his_body, his_vert, his_subvert, pred_title, pred_body, pred_vert, pred_subvert):
    """The 8 arrays of NAMLDataLoader -> article matrix [B*H + B*C, T+Tb+2] (history rows first)."""
    B, H, _ = his_title.shape
    C = pred_title.shape[1]
    h = np.concatenate([his_title, his_body, his_vert, his_subvert], axis=-1).reshape(B * H, -1)
    c = np.concatenate([pred_title, pred_body, pred_vert, pred_subvert], axis=-1).reshape(B * C, -1)
    return np.concatenate([h, c], axis=0), B, H, C


def naml_forward(art, B, H, C, P, T, Tb, *, training=False, p_drop=0.0, seeds=(0, 0, 0, 0)):
    n_all, c_news = news_encoder_fwd(art, P, T, Tb, training=training, p_drop=p_drop, seeds=seeds)
    F = n_all.shape[-1]
    Nh, Nc = n_all[:B * H].reshape(B, H, F), n_all[B * H:].reshape(B, C, F)
    u, c_user = O.att_layer2_fwd(Nh, P["user_W"], P["user_b"], P["user_q"])  # naml.py:79-84
    z = O.click_logits(Nc, u)
    return z, (c_news, c_user, Nc, u, F)


def naml_predict(art, B, H, C, P, T, Tb):
    return O.softmax(naml_forward(art, B, H, C, P, T, Tb)[0])


def naml_score(art, B, H, C, P, T, Tb):
    return O.sigmoid(naml_forward(art, B, H, C, P, T, Tb)[0])


def naml_loss_and_grads(art, B, H, C, y, P, T, Tb, *, p_drop=0.0, seeds=(0, 0, 0, 0), loss_scale=1.0):
    z, (c_news, c_user, Nc, u, F) = naml_forward(art, B, H, C, P, T, Tb, training=True, p_drop=p_drop, seeds=seeds)
    loss, prob, dz = O.softmax_ce(z, y)
    dz = dz * z.dtype.type(loss_scale)
    grads = {k: np.zeros_like(v) for k, v in P.items()}
    dNc = dz[..., None] * u[:, None, :]
    du = np.einsum("bc,bcd->bd", dz, Nc)
    dNh, dW, db, dq = O.att_layer2_bwd(du, c_user)
    grads["user_W"] += dW
    grads["user_b"] += db
    grads["user_q"] += dq
    dn_all = np.concatenate([dNh.reshape(B * H, F), dNc.reshape(B * C, F)], axis=0)
    news_encoder_bwd(dn_all, c_news, P, grads)
    return loss, prob, grads
