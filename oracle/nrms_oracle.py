"""NumPy restatement of the NRMS hot path (forward, loss, analytic backward, Keras Adam).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the
reference ships no golden vectors for the model math; every function cites the
reference lines it restates (paths relative to the reference repository root).

Layouts: token ids ``[N, T]`` int32, table ``[V, E]``, sequences ``[N, L, Din]``.
All functions are dtype-generic: pass float64 arrays for the checker, float32 to
mimic the reference's Keras default floatx.
"""
from __future__ import annotations

import numpy as np

K_EPSILON = 1e-7  # keras.backend.epsilon(), used by AttLayer2 (layers.py:75-77)

# ---------------------------------------------------------------------------
# Dropout mask.  TensorFlow's RNG stream cannot be reproduced, so the build
# defines its own counter-based mask (same function in csrc/ebk_common.cuh);
# semantics are Keras' inverted dropout: y = x * keep / (1 - p)  (nrms.py:136,153).
# ---------------------------------------------------------------------------
_U32 = np.uint64(0xFFFFFFFF)


def _mul32(a, c):
    return (a * np.uint64(c)) & _U32


def _group_bits(seed: int, g: np.ndarray):
    """Two 32-bit avalanche hashes of (seed, group) -> (x, y), each uint64 holding 32 bits."""
    s0 = np.uint64(seed & 0xFFFFFFFF)
    s1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    g = g.astype(np.uint64)
    x = (_mul32(g & _U32, 0x9E3779B1) + _mul32(g >> np.uint64(32), 0x85EBCA77) + s0) & _U32
    x ^= x >> np.uint64(16); x = _mul32(x, 0x7FEB352D)
    x ^= x >> np.uint64(15); x = _mul32(x, 0x846CA68B)
    x ^= x >> np.uint64(16)
    y = _mul32(x ^ s1, 0x9E3779B1)
    y ^= y >> np.uint64(15); y = _mul32(y, 0x2C1B3C6D)
    y ^= y >> np.uint64(12); y = _mul32(y, 0x297A2D39)
    y ^= y >> np.uint64(15)
    return x, y


def dropout_threshold(p: float) -> int:
    """16-bit keep threshold: element kept iff its 16 random bits >= threshold."""
    return int(np.floor(np.float32(p) * np.float32(65536.0) + np.float32(0.5)))


def dropout_keep_mask(seed: int, n_elems: int, p: float) -> np.ndarray:
    """Boolean keep mask over a flat element index space [0, n_elems).

    Elements are grouped by 4 (g = idx >> 2); group g draws two 32-bit words (x, y) from
    (seed, g) with lowbias32-style hashes; lane j = idx & 3 uses x[0:16], x[16:32], y[0:16],
    y[16:32].  Same function as csrc/ebk_common.cuh::dropout_group_bits.
    """
    idx = np.arange(n_elems, dtype=np.uint64)
    x, y = _group_bits(int(seed) & ((1 << 64) - 1), idx >> np.uint64(2))
    lane = idx & np.uint64(3)
    w = np.where(lane >= np.uint64(2), y, x)
    bits = np.where((lane & np.uint64(1)) == np.uint64(1), w >> np.uint64(16), w & np.uint64(0xFFFF))
    return bits >= np.uint64(dropout_threshold(p))


def dropout_fwd(x: np.ndarray, seed: int, p: float) -> tuple[np.ndarray, np.ndarray]:
    keep = dropout_keep_mask(seed, x.size, p).reshape(x.shape)
    scale = x.dtype.type(1.0 / (1.0 - p))
    return x * keep * scale, keep


# ---------------------------------------------------------------------------
# Initialisers
# ---------------------------------------------------------------------------
def glorot_uniform(rng: np.random.Generator, shape, dtype=np.float32) -> np.ndarray:
    """keras.initializers.glorot_uniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out))
    (layers.py:38,50,158,164,170; nrms.py:42)."""
    fan_in, fan_out = shape[0], shape[-1]
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(dtype)


def init_nrms_params(rng, V, E, nh, dh, att, dtype=np.float32, table=None) -> dict:
    """Parameter set of NRMSModel in Keras get_weights() order."""
    D = nh * dh
    p = {}
    p["table"] = (table if table is not None else glorot_uniform(rng, (V, E))).astype(dtype)
    for pre, din in (("news", E), ("user", D)):
        p[f"{pre}_WQ"] = glorot_uniform(rng, (din, D), dtype)
        p[f"{pre}_WK"] = glorot_uniform(rng, (din, D), dtype)
        p[f"{pre}_WV"] = glorot_uniform(rng, (din, D), dtype)
        p[f"{pre}_W"] = glorot_uniform(rng, (D, att), dtype)
        p[f"{pre}_b"] = np.zeros((att,), dtype)
        p[f"{pre}_q"] = glorot_uniform(rng, (att, 1), dtype)
    return p


NRMS_PARAM_ORDER = [
    "table",
    "news_WQ", "news_WK", "news_WV", "news_W", "news_b", "news_q",
    "user_WQ", "user_WK", "user_WV", "user_W", "user_b", "user_q",
]


# ---------------------------------------------------------------------------
# SelfAttention  (layers.py:200-254; weights layers.py:155-172, no bias, no mask)
# ---------------------------------------------------------------------------
def self_attention_fwd(X, WQ, WK, WV, nh, dh):
    N, L, _ = X.shape
    D = nh * dh

    def split(Y):  # [N,L,D] -> [N,nh,L,dh]   (layers.py:215-218)
        return Y.reshape(N, L, nh, dh).transpose(0, 2, 1, 3)

    Q = split(X @ WQ)  # layers.py:214
    Kh = split(X @ WK)  # layers.py:220
    Vh = split(X @ WV)  # layers.py:226
    S = np.einsum("nhqd,nhkd->nhqk", Q, Kh) / np.sqrt(X.dtype.type(dh))  # layers.py:231-233
    S = S - S.max(axis=-1, keepdims=True)
    A = np.exp(S)
    A = A / A.sum(axis=-1, keepdims=True)  # K.softmax over keys, layers.py:247
    # layers.py:249  tf.matmul(A, V, adjoint_a=True):  O[k,:] = sum_q A[q,k] V[q,:]
    O = np.einsum("nhqk,nhqd->nhkd", A, Vh)
    out = O.transpose(0, 2, 1, 3).reshape(N, L, D)  # layers.py:250-252
    cache = (X, WQ, WK, WV, Q, Kh, Vh, A, nh, dh)
    return out, cache


def self_attention_bwd(dout, cache):
    X, WQ, WK, WV, Q, Kh, Vh, A, nh, dh = cache
    N, L, _ = X.shape
    D = nh * dh
    dO = dout.reshape(N, L, nh, dh).transpose(0, 2, 1, 3)  # [N,nh,k,dh]
    dV = np.einsum("nhqk,nhkd->nhqd", A, dO)
    dA = np.einsum("nhqd,nhkd->nhqk", Vh, dO)
    dS = A * (dA - (dA * A).sum(axis=-1, keepdims=True))
    inv = X.dtype.type(1.0) / np.sqrt(X.dtype.type(dh))
    dQ = np.einsum("nhqk,nhkd->nhqd", dS, Kh) * inv
    dK = np.einsum("nhqk,nhqd->nhkd", dS, Q) * inv

    def merge(Y):
        return Y.transpose(0, 2, 1, 3).reshape(N * L, D)

    dQm, dKm, dVm = merge(dQ), merge(dK), merge(dV)
    Xf = X.reshape(N * L, -1)
    dWQ, dWK, dWV = Xf.T @ dQm, Xf.T @ dKm, Xf.T @ dVm
    dX = (dQm @ WQ.T + dKm @ WK.T + dVm @ WV.T).reshape(X.shape)
    return dX, dWQ, dWK, dWV


# ---------------------------------------------------------------------------
# AttLayer2  (layers.py:55-81; weights layers.py:35-52)
# ---------------------------------------------------------------------------
def att_layer2_fwd(X, W, b, q):
    h = np.tanh(X @ W + b)  # layers.py:65
    a = (h @ q)[..., 0]  # layers.py:66-68
    e = np.exp(a)  # layers.py:70-71 (mask is None on the NRMS path; NO max subtraction)
    w = e / (e.sum(axis=-1, keepdims=True) + X.dtype.type(K_EPSILON))  # layers.py:75-77
    y = (X * w[..., None]).sum(axis=1)  # layers.py:79-81
    return y, (X, W, q, h, w)


def att_layer2_bwd(dy, cache):
    X, W, q, h, w = cache
    N, L, D = X.shape
    dX = w[..., None] * dy[:, None, :]
    dw = np.einsum("nld,nd->nl", X, dy)
    da = w * (dw - (w * dw).sum(axis=-1, keepdims=True))
    dq = np.einsum("nla,nl->a", h, da)[:, None]
    dpre = (da[..., None] * q[:, 0][None, None, :]) * (1 - h * h)
    dW = X.reshape(N * L, D).T @ dpre.reshape(N * L, -1)
    db = dpre.sum(axis=(0, 1))
    dX = dX + dpre @ W.T
    return dX, dW, db, dq


# ---------------------------------------------------------------------------
# News encoder (nrms.py:116-159, units_per_layer=None branch) and user encoder
# (nrms.py:92-114)
# ---------------------------------------------------------------------------
def news_encoder_fwd(tok, P, nh, dh, *, training=False, p_drop=0.0, seed1=0, seed2=0, prefix="news"):
    table = P["table"]
    V = table.shape[0]
    tok = np.asarray(tok)
    inb = (tok >= 0) & (tok < V)
    # nrms.py:125-134 Embedding gather.  Out-of-range ids -> zero row, no gradient
    # (TF-GPU behaviour; SURVEY.md section 7 hard part 8).
    E0 = table[np.where(inb, tok, 0)] * inb[..., None].astype(table.dtype)
    keep1 = keep2 = None
    X = E0
    if training and p_drop > 0:
        X, keep1 = dropout_fwd(E0, seed1, p_drop)  # nrms.py:136
    Y0, c_sa = self_attention_fwd(X, P[f"{prefix}_WQ"], P[f"{prefix}_WK"], P[f"{prefix}_WV"], nh, dh)  # nrms.py:137-139
    Y = Y0
    if training and p_drop > 0:
        Y, keep2 = dropout_fwd(Y0, seed2, p_drop)  # nrms.py:153-154
    out, c_att = att_layer2_fwd(Y, P[f"{prefix}_W"], P[f"{prefix}_b"], P[f"{prefix}_q"])  # nrms.py:156
    return out, (tok, inb, keep1, keep2, c_sa, c_att, p_drop, prefix, V)


def news_encoder_bwd(dout, cache, grads):
    tok, inb, keep1, keep2, c_sa, c_att, p_drop, prefix, V = cache
    dY, dW, db, dq = att_layer2_bwd(dout, c_att)
    grads[f"{prefix}_W"] += dW
    grads[f"{prefix}_b"] += db
    grads[f"{prefix}_q"] += dq
    if keep2 is not None:
        dY = dY * keep2 * dY.dtype.type(1.0 / (1.0 - p_drop))
    dX, dWQ, dWK, dWV = self_attention_bwd(dY, c_sa)
    grads[f"{prefix}_WQ"] += dWQ
    grads[f"{prefix}_WK"] += dWK
    grads[f"{prefix}_WV"] += dWV
    if keep1 is not None:
        dX = dX * keep1 * dX.dtype.type(1.0 / (1.0 - p_drop))
    dX = dX * inb[..., None].astype(dX.dtype)
    E = dX.shape[-1]
    np.add.at(grads["table"], np.where(inb, tok, 0).reshape(-1), dX.reshape(-1, E))
    return grads


def user_encoder_fwd(Nh, P, nh, dh):
    """Nh: [B, H, D] encoded history (TimeDistributed(newsencoder), nrms.py:105-107)."""
    Y, c_sa = self_attention_fwd(Nh, P["user_WQ"], P["user_WK"], P["user_WV"], nh, dh)  # nrms.py:108-110
    u, c_att = att_layer2_fwd(Y, P["user_W"], P["user_b"], P["user_q"])  # nrms.py:111
    return u, (c_sa, c_att)


def user_encoder_bwd(du, cache, grads):
    c_sa, c_att = cache
    dY, dW, db, dq = att_layer2_bwd(du, c_att)
    grads["user_W"] += dW
    grads["user_b"] += db
    grads["user_q"] += dq
    dNh, dWQ, dWK, dWV = self_attention_bwd(dY, c_sa)
    grads["user_WQ"] += dWQ
    grads["user_WK"] += dWK
    grads["user_WV"] += dWV
    return dNh


# ---------------------------------------------------------------------------
# Score, loss (nrms.py:201-205; nrms.py:61-62 categorical_crossentropy)
# ---------------------------------------------------------------------------
def click_logits(news, user):
    """Dot(axes=-1): z[b,c] = news[b,c,:] . user[b,:]  (nrms.py:201)."""
    return np.einsum("bcd,bd->bc", news, user)


def softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def softmax_ce(z, y):
    """Keras categorical_crossentropy on an Activation('softmax') output = CE from the
    cached logits, mean over the batch.  Labels are
    used as given (Keras does not renormalise them on the from-logits path)."""
    m = z.max(axis=-1, keepdims=True)
    lse = m[..., 0] + np.log(np.exp(z - m).sum(axis=-1))
    y = y.astype(z.dtype)
    per = lse * y.sum(axis=-1) - (y * z).sum(axis=-1)
    loss = per.mean()
    p = softmax(z)
    dz = (p * y.sum(axis=-1, keepdims=True) - y) / z.dtype.type(z.shape[0])
    return loss, p, dz


def sigmoid_ce_on_softmax_logits(z, y):
    """hparams.loss == "log_loss" -> Keras "binary_crossentropy" (nrms.py:63-64) on the softmax Activation output of
    nrms.py:202.  keras.backend.binary_crossentropy looks for logits cached on its argument (`_keras_logits`, set by
    BOTH keras.activations.softmax and .sigmoid) and, finding the softmax's, evaluates
    tf.nn.sigmoid_cross_entropy_with_logits on them: max(z,0) - z*y + log(1 + exp(-|z|)), mean over the candidate
    axis, then mean over the batch.  The model's predictions stay the softmax probabilities."""
    y = y.astype(z.dtype)
    per = (np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))).mean(axis=-1)
    dz = (sigmoid(z) - y) / z.dtype.type(z.shape[-1] * z.shape[0])
    return per.mean(), softmax(z), dz


# ---------------------------------------------------------------------------
# Whole model
# ---------------------------------------------------------------------------
def nrms_forward(his, pred, P, nh, dh, *, training=False, p_drop=0.0, seed1=0, seed2=0):
    """NRMSModel.model forward (nrms.py:161-210).  News encoder is evaluated once over
    the concatenation [history articles ; candidate articles] (row order: all B*H
    history rows, then all B*C candidate rows) -- there is no cross-article coupling
    in this encoder, so this equals the two TimeDistributed calls (nrms.py:105,196);
    the concatenated order defines the dropout element indices."""
    B, H, T = his.shape
    C = pred.shape[1]
    tok = np.concatenate([his.reshape(B * H, T), pred.reshape(B * C, T)], axis=0)
    n_all, c_news = news_encoder_fwd(tok, P, nh, dh, training=training, p_drop=p_drop, seed1=seed1, seed2=seed2)
    D = n_all.shape[-1]
    Nh = n_all[: B * H].reshape(B, H, D)
    Nc = n_all[B * H:].reshape(B, C, D)
    u, c_user = user_encoder_fwd(Nh, P, nh, dh)
    z = click_logits(Nc, u)
    return z, (B, H, C, D, c_news, c_user, Nc, u)


def nrms_predict(his, pred, P, nh, dh):
    """model.predict: softmax over candidates (nrms.py:202), inference mode."""
    z, _ = nrms_forward(his, pred, P, nh, dh)
    return softmax(z)


def nrms_score(his, pred_one, P, nh, dh):
    """scorer.predict: sigmoid(news . user), [N,1] (nrms.py:204-205,208)."""
    z, _ = nrms_forward(his, pred_one, P, nh, dh)
    return sigmoid(z)


def nrms_loss_and_grads(his, pred, y, P, nh, dh, *, training=True, p_drop=0.0, seed1=0, seed2=0, loss_scale=1.0,
                        loss_kind="cross_entropy_loss"):
    z, (B, H, C, D, c_news, c_user, Nc, u) = nrms_forward(
        his, pred, P, nh, dh, training=training, p_drop=p_drop, seed1=seed1, seed2=seed2)
    loss, prob, dz = softmax_ce(z, y) if loss_kind == "cross_entropy_loss" else sigmoid_ce_on_softmax_logits(z, y)
    dz = dz * z.dtype.type(loss_scale)
    grads = {k: np.zeros_like(v) for k, v in P.items()}
    dNc = dz[..., None] * u[:, None, :]
    du = np.einsum("bc,bcd->bd", dz, Nc)
    dNh = user_encoder_bwd(du, c_user, grads)
    dn_all = np.concatenate([dNh.reshape(B * H, D), dNc.reshape(B * C, D)], axis=0)
    news_encoder_bwd(dn_all, c_news, grads)
    return loss, prob, grads


# ---------------------------------------------------------------------------
# tf.keras.optimizers.Adam (non-legacy, TF 2.12-2.15), nrms.py:76-77.
# Sparse (IndexedSlices) gradients are de-duplicated (summed) and then every
# row's m, v and theta are updated -> identical to the dense form below
# (SURVEY.md section 3.6 item 2).
# ---------------------------------------------------------------------------
def keras_adam_step(theta, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """In-place update; ``t`` is the 1-based step number; arithmetic in theta.dtype."""
    f = theta.dtype.type
    alpha = f(lr) * np.sqrt(f(1.0) - np.power(f(beta2), f(t))) / (f(1.0) - np.power(f(beta1), f(t)))
    m += (g - m) * f(1.0 - beta1)
    v += (g * g - v) * f(1.0 - beta2)
    theta -= (m * alpha) / (np.sqrt(v) + f(eps))
    return theta, m, v


def count_params(P) -> int:
    return int(sum(v.size for v in P.values()))
