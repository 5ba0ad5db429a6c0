"""NumPy restatement of NRMSDocVec (reference src/ebrec/models/newsrec/nrms_docvec.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED (no reference golden vectors).

News encoder (nrms_docvec.py:99-137): x[768] -> [Dense(u, relu, L2 1e-4) -> BatchNormalization -> Dropout(p)]
for u in units_per_layer -> Dense(D, relu).  It is applied through TimeDistributed separately to the
history tensor and to the candidate tensor (nrms_docvec.py:88-90, 174-176), so in training mode the
BatchNorm batch statistics are computed per call (over B*H rows, then over B*C rows) and the moving
averages are updated twice per step.  User encoder / score / loss are those of NRMS
(nrms_docvec.py:75-97, 139-188) -> reused from nrms_oracle.

Keras semantics used: BatchNormalization(momentum=0.99, epsilon=1e-3), biased batch variance in both the
normalisation and the moving-average update (non-fused 2-D path); kernel_regularizer l2(l): loss += l*sum(W^2).
"""
from __future__ import annotations

import numpy as np

from . import nrms_oracle as O

BN_MOMENTUM = 0.99
BN_EPS = 1e-3


def init_docvec_params(rng, Ddoc, units, nh, dh, att, dtype=np.float32) -> dict:
    D = nh * dh
    P = {}
    din = Ddoc
    for i, u in enumerate(units):
        P[f"d{i}_W"] = O.glorot_uniform(rng, (din, u), dtype)
        P[f"d{i}_b"] = np.zeros((u,), dtype)
        P[f"d{i}_gamma"] = np.ones((u,), dtype)
        P[f"d{i}_beta"] = np.zeros((u,), dtype)
        P[f"d{i}_mean"] = np.zeros((u,), dtype)  # moving mean (non-trainable)
        P[f"d{i}_var"] = np.ones((u,), dtype)    # moving variance (non-trainable)
        din = u
    P["out_W"] = O.glorot_uniform(rng, (din, D), dtype)
    P["out_b"] = np.zeros((D,), dtype)
    for k in ("WQ", "WK", "WV"):
        P[f"user_{k}"] = O.glorot_uniform(rng, (D, D), dtype)
    P["user_W"] = O.glorot_uniform(rng, (D, att), dtype)
    P["user_b"] = np.zeros((att,), dtype)
    P["user_q"] = O.glorot_uniform(rng, (att, 1), dtype)
    return P


def trainable_keys(n_layers: int) -> list[str]:
    ks = []
    for i in range(n_layers):
        ks += [f"d{i}_W", f"d{i}_b", f"d{i}_gamma", f"d{i}_beta"]
    return ks + ["out_W", "out_b", "user_WQ", "user_WK", "user_WV", "user_W", "user_b", "user_q"]


def news_encoder_fwd(X, P, n_layers, *, training=False, p_drop=0.0, seed=0, new_stats=None):
    """X [N, Ddoc] -> [N, D].  In training mode BN uses the batch statistics of THIS call and, if
    ``new_stats`` is a dict, the updated moving averages are written to it (Keras updates them per call)."""
    f = X.dtype.type
    caches = []
    x = X
    for i in range(n_layers):
        W, b = P[f"d{i}_W"], P[f"d{i}_b"]
        a = np.maximum(x @ W + b, 0)  # Dense(relu), nrms_docvec.py:119-125
        if training:
            mean, var = a.mean(axis=0), a.var(axis=0)  # biased
            if new_stats is not None:
                mm = new_stats.get(f"d{i}_mean", P[f"d{i}_mean"])
                mv = new_stats.get(f"d{i}_var", P[f"d{i}_var"])
                new_stats[f"d{i}_mean"] = mm * f(BN_MOMENTUM) + mean * f(1 - BN_MOMENTUM)
                new_stats[f"d{i}_var"] = mv * f(BN_MOMENTUM) + var * f(1 - BN_MOMENTUM)
        else:
            mean, var = P[f"d{i}_mean"], P[f"d{i}_var"]
        invstd = 1.0 / np.sqrt(var + f(BN_EPS))
        xhat = (a - mean) * invstd
        y = xhat * P[f"d{i}_gamma"] + P[f"d{i}_beta"]  # BatchNormalization, nrms_docvec.py:126
        keep = None
        if training and p_drop > 0:
            y, keep = O.dropout_fwd(y, seed + i, p_drop)  # Dropout, nrms_docvec.py:127 (one seed per layer)
        caches.append((x, a, xhat, invstd, keep))
        x = y
    out = np.maximum(x @ P["out_W"] + P["out_b"], 0)  # nrms_docvec.py:130
    return out, (caches, x, out, p_drop, n_layers)


def news_encoder_bwd(dout, cache, P, grads, l2=0.0):
    caches, x_last, out, p_drop, n_layers = cache
    dz = dout * (out > 0)
    grads["out_W"] += x_last.T @ dz
    grads["out_b"] += dz.sum(0)
    dy = dz @ P["out_W"].T
    for i in reversed(range(n_layers)):
        x, a, xhat, invstd, keep = caches[i]
        N = a.shape[0]
        if keep is not None:
            dy = dy * keep * dy.dtype.type(1.0 / (1.0 - p_drop))
        grads[f"d{i}_gamma"] += (dy * xhat).sum(0)
        grads[f"d{i}_beta"] += dy.sum(0)
        dxhat = dy * P[f"d{i}_gamma"]
        da = invstd / N * (N * dxhat - dxhat.sum(0) - xhat * (dxhat * xhat).sum(0))  # training-mode BN backward
        dz = da * (a > 0)
        grads[f"d{i}_W"] += x.T @ dz
        grads[f"d{i}_b"] += dz.sum(0)
        dy = dz @ P[f"d{i}_W"].T
    return dy


def docvec_forward(his, pred, P, n_layers, nh, dh, *, training=False, p_drop=0.0, seed_h=0, seed_c=0, new_stats=None):
    """his [B,H,Ddoc], pred [B,C,Ddoc] float -> logits [B,C].  Two encoder calls (history, candidates)."""
    B, H, Dd = his.shape
    C = pred.shape[1]
    nh_out, c_h = news_encoder_fwd(his.reshape(B * H, Dd), P, n_layers, training=training, p_drop=p_drop, seed=seed_h,
                                   new_stats=new_stats)
    nc_out, c_c = news_encoder_fwd(pred.reshape(B * C, Dd), P, n_layers, training=training, p_drop=p_drop, seed=seed_c,
                                   new_stats=new_stats)
    D = nh * dh
    Nh, Nc = nh_out.reshape(B, H, D), nc_out.reshape(B, C, D)
    u, c_user = O.user_encoder_fwd(Nh, P, nh, dh)
    z = O.click_logits(Nc, u)
    return z, (B, H, C, D, c_h, c_c, c_user, Nc, u)


def docvec_predict(his, pred, P, n_layers, nh, dh):
    z, _ = docvec_forward(his, pred, P, n_layers, nh, dh)
    return O.softmax(z)


def docvec_score(his, pred_one, P, n_layers, nh, dh):
    z, _ = docvec_forward(his, pred_one, P, n_layers, nh, dh)
    return O.sigmoid(z)


def docvec_loss_and_grads(his, pred, y, P, n_layers, nh, dh, *, p_drop=0.0, seed_h=0, seed_c=0, l2=1e-4, loss_scale=1.0):
    """Training-mode loss (CE + l2 * sum ||W_dense||^2, nrms_docvec.py:122-124) and gradients; also returns the
    updated BatchNorm moving statistics."""
    new_stats = {}
    z, (B, H, C, D, c_h, c_c, c_user, Nc, u) = docvec_forward(his, pred, P, n_layers, nh, dh, training=True, p_drop=p_drop,
                                                               seed_h=seed_h, seed_c=seed_c, new_stats=new_stats)
    ce, prob, dz = O.softmax_ce(z, y)
    reg = sum(float((P[f"d{i}_W"].astype(np.float64) ** 2).sum()) for i in range(n_layers)) * l2
    loss = ce + reg
    dz = dz * z.dtype.type(loss_scale)
    grads = {k: np.zeros_like(P[k]) for k in trainable_keys(n_layers)}
    dNc = dz[..., None] * u[:, None, :]
    du = np.einsum("bc,bcd->bd", dz, Nc)
    dNh = O.user_encoder_bwd(du, c_user, grads)
    news_encoder_bwd(dNh.reshape(B * H, D), c_h, P, grads)
    news_encoder_bwd(dNc.reshape(B * C, D), c_c, P, grads)
    for i in range(n_layers):
        grads[f"d{i}_W"] += P[f"d{i}_W"] * P[f"d{i}_W"].dtype.type(2.0 * l2 * loss_scale)
    return loss, prob, grads, new_stats
