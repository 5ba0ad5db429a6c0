"""NumPy restatement of NRMS with the optional Dense/BatchNorm/Dropout stack in the news encoder
(reference src/ebrec/models/newsrec/nrms.py:116-159 with ``newsencoder_units_per_layer`` set, nrms.py:142-152).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED (no reference golden vectors).

News encoder per article: Embedding -> Dropout(p) -> SelfAttention -> for u in units:
[Dense(u, relu, kernel_regularizer=l2(newsencoder_l2_regularization)) -> BatchNormalization -> Dropout(p)]
-> AttLayer2.  NO Dropout directly after the SelfAttention in this branch (nrms.py:153-154 is the else).
The Dense/BN layers act on [articles, T, units] tensors: BatchNormalization(axis=-1) normalises over
articles x tokens, and because the news encoder is applied through two TimeDistributed calls
(history nrms.py:105-107, candidates nrms.py:196-198) the batch statistics are per call and the moving
averages are updated twice per step -- the same Keras semantics as NRMSDocVec (docvec_oracle.py).
The news vector has width units[-1], which must equal head_num*head_dim for the Dot of nrms.py:201.

Dropout element indices (the mask function is ours, nrms_oracle.dropout_keep_mask): embedded tokens over the
concatenated [history ; candidates] rows with ``seed1`` (as nrms_oracle); layer i of the history call uses
``seed_h + i`` over its [B*H*T, u_i] output, of the candidate call ``seed_c + i`` over [B*C*T, u_i].
"""
from __future__ import annotations

import numpy as np

from . import nrms_oracle as O
from .docvec_oracle import BN_EPS, BN_MOMENTUM


def init_params(rng, V, E, units, nh, dh, att, dtype=np.float32, table=None) -> dict:
    """Keras get_weights() order: table, WQ, WK, WV, [kernel, bias, gamma, beta, moving_mean, moving_var] per
    layer, AttLayer2 W, b, q, then the user encoder."""
    D = nh * dh
    assert units and units[-1] == D, "units[-1] must equal head_num*head_dim (Dot of nrms.py:201)"
    P = {"table": (table if table is not None else O.glorot_uniform(rng, (V, E))).astype(dtype)}
    for k in ("WQ", "WK", "WV"):
        P[f"news_{k}"] = O.glorot_uniform(rng, (E, D), dtype)
    din = D
    for i, u in enumerate(units):
        P[f"d{i}_W"] = O.glorot_uniform(rng, (din, u), dtype)
        P[f"d{i}_b"] = np.zeros((u,), dtype)
        P[f"d{i}_gamma"] = np.ones((u,), dtype)
        P[f"d{i}_beta"] = np.zeros((u,), dtype)
        P[f"d{i}_mean"] = np.zeros((u,), dtype)
        P[f"d{i}_var"] = np.ones((u,), dtype)
        din = u
    P["news_W"] = O.glorot_uniform(rng, (din, att), dtype)
    P["news_b"] = np.zeros((att,), dtype)
    P["news_q"] = O.glorot_uniform(rng, (att, 1), dtype)
    for k in ("WQ", "WK", "WV"):
        P[f"user_{k}"] = O.glorot_uniform(rng, (D, D), dtype)
    P["user_W"] = O.glorot_uniform(rng, (D, att), dtype)
    P["user_b"] = np.zeros((att,), dtype)
    P["user_q"] = O.glorot_uniform(rng, (att, 1), dtype)
    return P


def param_order(n_layers: int) -> list[str]:
    ks = ["table", "news_WQ", "news_WK", "news_WV"]
    for i in range(n_layers):
        ks += [f"d{i}_W", f"d{i}_b", f"d{i}_gamma", f"d{i}_beta", f"d{i}_mean", f"d{i}_var"]
    return ks + ["news_W", "news_b", "news_q", "user_WQ", "user_WK", "user_WV", "user_W", "user_b", "user_q"]


def trainable_keys(n_layers: int) -> list[str]:
    return [k for k in param_order(n_layers) if not k.endswith(("_mean", "_var"))]


def _stack_fwd(Y, P, n_layers, *, training, p_drop, seed, new_stats):
    """Y [n, T, D] -> [n, T, u_last]: Dense(relu) -> BN (statistics over n*T rows) -> Dropout, per layer."""
    f = Y.dtype.type
    n, T, _ = Y.shape
    x = Y.reshape(n * T, -1)
    caches = []
    for i in range(n_layers):
        a = np.maximum(x @ P[f"d{i}_W"] + P[f"d{i}_b"], 0)  # nrms.py:144-150
        if training:
            mean, var = a.mean(axis=0), a.var(axis=0)
            if new_stats is not None:
                mm = new_stats.get(f"d{i}_mean", P[f"d{i}_mean"])
                mv = new_stats.get(f"d{i}_var", P[f"d{i}_var"])
                new_stats[f"d{i}_mean"] = mm * f(BN_MOMENTUM) + mean * f(1 - BN_MOMENTUM)
                new_stats[f"d{i}_var"] = mv * f(BN_MOMENTUM) + var * f(1 - BN_MOMENTUM)
        else:
            mean, var = P[f"d{i}_mean"], P[f"d{i}_var"]
        invstd = 1.0 / np.sqrt(var + f(BN_EPS))
        xhat = (a - mean) * invstd
        y = xhat * P[f"d{i}_gamma"] + P[f"d{i}_beta"]  # nrms.py:151
        keep = None
        if training and p_drop > 0:
            y, keep = O.dropout_fwd(y, seed + i, p_drop)  # nrms.py:152
        caches.append((x, a, xhat, invstd, keep))
        x = y
    return x.reshape(n, T, -1), (caches, p_drop, n_layers, training)


def _stack_bwd(dZ, cache, P, grads):
    caches, p_drop, n_layers, training = cache
    n, T, _ = dZ.shape
    dy = dZ.reshape(n * T, -1)
    for i in reversed(range(n_layers)):
        x, a, xhat, invstd, keep = caches[i]
        N = a.shape[0]
        if keep is not None:
            dy = dy * keep * dy.dtype.type(1.0 / (1.0 - p_drop))
        grads[f"d{i}_gamma"] += (dy * xhat).sum(0)
        grads[f"d{i}_beta"] += dy.sum(0)
        dxhat = dy * P[f"d{i}_gamma"]
        if training:
            da = invstd / N * (N * dxhat - dxhat.sum(0) - xhat * (dxhat * xhat).sum(0))
        else:
            da = dxhat * invstd
        dz = da * (a > 0)
        grads[f"d{i}_W"] += x.T @ dz
        grads[f"d{i}_b"] += dz.sum(0)
        dy = dz @ P[f"d{i}_W"].T
    return dy.reshape(n, T, -1)


def forward(his, pred, P, n_layers, nh, dh, *, training=False, p_drop=0.0, seed1=0, seed_h=0, seed_c=0,
            new_stats=None):
    B, H, T = his.shape
    C = pred.shape[1]
    tok = np.concatenate([his.reshape(B * H, T), pred.reshape(B * C, T)], axis=0)
    table = P["table"]
    V = table.shape[0]
    inb = (tok >= 0) & (tok < V)
    E0 = table[np.where(inb, tok, 0)] * inb[..., None].astype(table.dtype)
    keep1 = None
    X = E0
    if training and p_drop > 0:
        X, keep1 = O.dropout_fwd(E0, seed1, p_drop)  # nrms.py:136
    Y0, c_sa = O.self_attention_fwd(X, P["news_WQ"], P["news_WK"], P["news_WV"], nh, dh)  # nrms.py:137-139
    BH = B * H
    Zh, c_h = _stack_fwd(Y0[:BH], P, n_layers, training=training, p_drop=p_drop, seed=seed_h, new_stats=new_stats)
    Zc, c_c = _stack_fwd(Y0[BH:], P, n_layers, training=training, p_drop=p_drop, seed=seed_c, new_stats=new_stats)
    Z = np.concatenate([Zh, Zc], axis=0)
    n_all, c_att = O.att_layer2_fwd(Z, P["news_W"], P["news_b"], P["news_q"])  # nrms.py:156
    D = n_all.shape[-1]
    Nh, Nc = n_all[:BH].reshape(B, H, D), n_all[BH:].reshape(B, C, D)
    u, c_user = O.user_encoder_fwd(Nh, P, nh, dh)
    z = O.click_logits(Nc, u)
    return z, (B, H, C, D, tok, inb, keep1, p_drop, c_sa, c_h, c_c, c_att, c_user, Nc, u)


def predict(his, pred, P, n_layers, nh, dh):
    return O.softmax(forward(his, pred, P, n_layers, nh, dh)[0])


def score(his, pred_one, P, n_layers, nh, dh):
    return O.sigmoid(forward(his, pred_one, P, n_layers, nh, dh)[0])


def loss_and_grads(his, pred, y, P, n_layers, nh, dh, *, training=True, p_drop=0.0, seed1=0, seed_h=0, seed_c=0,
                   l2=1e-4, loss_scale=1.0, new_stats=None):
    """loss = mean CE + l2 * sum_i ||d{i}_W||^2 (Keras adds each kernel_regularizer once per layer object)."""
    z, (B, H, C, D, tok, inb, keep1, p, c_sa, c_h, c_c, c_att, c_user, Nc, u) = forward(
        his, pred, P, n_layers, nh, dh, training=training, p_drop=p_drop, seed1=seed1, seed_h=seed_h, seed_c=seed_c,
        new_stats=new_stats)
    loss, prob, dz = O.softmax_ce(z, y)
    f = z.dtype.type
    dz = dz * f(loss_scale)
    grads = {k: np.zeros_like(v) for k, v in P.items()}
    dNc = dz[..., None] * u[:, None, :]
    du = np.einsum("bc,bcd->bd", dz, Nc)
    dNh = O.user_encoder_bwd(du, c_user, grads)
    dn_all = np.concatenate([dNh.reshape(B * H, D), dNc.reshape(B * C, D)], axis=0)
    dZ, dW, db, dq = O.att_layer2_bwd(dn_all, c_att)
    grads["news_W"] += dW
    grads["news_b"] += db
    grads["news_q"] += dq
    BH = B * H
    dY0 = np.concatenate([_stack_bwd(dZ[:BH], c_h, P, grads), _stack_bwd(dZ[BH:], c_c, P, grads)], axis=0)
    dX, dWQ, dWK, dWV = O.self_attention_bwd(dY0, c_sa)
    grads["news_WQ"] += dWQ
    grads["news_WK"] += dWK
    grads["news_WV"] += dWV
    if keep1 is not None:
        dX = dX * keep1 * dX.dtype.type(1.0 / (1.0 - p))
    dX = dX * inb[..., None].astype(dX.dtype)
    np.add.at(grads["table"], np.where(inb, tok, 0).reshape(-1), dX.reshape(-1, dX.shape[-1]))
    reg = 0.0
    for i in range(n_layers):
        reg += l2 * float((P[f"d{i}_W"].astype(np.float64) ** 2).sum())
        grads[f"d{i}_W"] += f(2 * l2 * loss_scale) * P[f"d{i}_W"]
    return loss + reg, prob, grads
