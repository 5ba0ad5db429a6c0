"""torch-CPU float32 port of the reference NRMS graph, differentiated by autograd.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.
Two uses: (1) tests cross-check the numpy oracle's *analytic* backward against
autograd of the same forward; (2) bench.py's ``cpu_baseline`` / ``--impl reference``
leg times it on the host cores ("port": CPU restatement, not TensorFlow --
TensorFlow cannot be installed here).

It follows the same reference lines as nrms_oracle.py:
SelfAttention layers.py:200-254, AttLayer2 layers.py:55-81, wiring nrms.py:92-210,
loss nrms.py:61-62, Adam nrms.py:76-77 (Keras form, eps outside the bias correction).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .nrms_oracle import NRMS_PARAM_ORDER, K_EPSILON


def params_to_torch(P: dict, requires_grad=True) -> dict:
    return {k: torch.tensor(np.asarray(v, dtype=np.float32), requires_grad=requires_grad) for k, v in P.items()}


def self_attention(X, WQ, WK, WV, nh, dh):
    N, L, _ = X.shape
    Q = (X @ WQ).view(N, L, nh, dh).permute(0, 2, 1, 3)
    K = (X @ WK).view(N, L, nh, dh).permute(0, 2, 1, 3)
    V = (X @ WV).view(N, L, nh, dh).permute(0, 2, 1, 3)
    A = torch.softmax(Q @ K.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    O = A.transpose(-1, -2) @ V  # adjoint_a=True, layers.py:249
    return O.permute(0, 2, 1, 3).reshape(N, L, nh * dh)


def att_layer2(X, W, b, q):
    a = (torch.tanh(X @ W + b) @ q).squeeze(-1)
    e = torch.exp(a)
    w = e / (e.sum(dim=-1, keepdim=True) + K_EPSILON)
    return (X * w.unsqueeze(-1)).sum(dim=1)


def nrms_logits(his, pred, P, nh, dh, keep1=None, keep2=None, p_drop=0.0):
    B, H, T = his.shape
    C = pred.shape[1]
    tok = torch.cat([his.reshape(B * H, T), pred.reshape(B * C, T)], dim=0).long()
    X = P["table"][tok]
    if keep1 is not None:
        X = X * keep1 / (1.0 - p_drop)
    Y = self_attention(X, P["news_WQ"], P["news_WK"], P["news_WV"], nh, dh)
    if keep2 is not None:
        Y = Y * keep2 / (1.0 - p_drop)
    n_all = att_layer2(Y, P["news_W"], P["news_b"], P["news_q"])
    D = nh * dh
    Nh = n_all[: B * H].view(B, H, D)
    Nc = n_all[B * H:].view(B, C, D)
    Yu = self_attention(Nh, P["user_WQ"], P["user_WK"], P["user_WV"], nh, dh)
    u = att_layer2(Yu, P["user_W"], P["user_b"], P["user_q"])
    return torch.einsum("bcd,bd->bc", Nc, u)


def nrms_loss(his, pred, y, P, nh, dh, **kw):
    z = nrms_logits(his, pred, P, nh, dh, **kw)
    yf = y.to(z.dtype)
    lse = torch.logsumexp(z, dim=-1)
    return (lse * yf.sum(-1) - (yf * z).sum(-1)).mean(), z


class KerasAdam:
    """Dense Keras-form Adam over a dict of tensors (see nrms_oracle.keras_adam_step)."""

    def __init__(self, P, lr, beta1=0.9, beta2=0.999, eps=1e-7):
        self.P, self.lr, self.b1, self.b2, self.eps, self.t = P, lr, beta1, beta2, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in P.items()}

    @torch.no_grad()
    def step(self):
        self.t += 1
        alpha = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, th in self.P.items():
            g = th.grad
            if g is None:
                continue
            m, v = self.m[k], self.v[k]
            m.add_((g - m) * (1.0 - self.b1))
            v.add_((g * g - v) * (1.0 - self.b2))
            th.sub_((m * alpha) / (v.sqrt() + self.eps))
            th.grad = None


def train_step(his, pred, y, P, opt: KerasAdam, nh, dh, p_drop=0.0, rng: torch.Generator | None = None):
    """One reference-shaped train step (forward with dropout, backward, dense Adam)."""
    keep1 = keep2 = None
    if p_drop > 0:
        B, H, T = his.shape
        N = B * (H + pred.shape[1])
        E = P["table"].shape[1]
        keep1 = (torch.rand((N, T, E), generator=rng) >= p_drop).float()
        keep2 = (torch.rand((N, T, nh * dh), generator=rng) >= p_drop).float()
    loss, _ = nrms_loss(his, pred, y, P, nh, dh, keep1=keep1, keep2=keep2, p_drop=p_drop)
    loss.backward()
    opt.step()
    return float(loss)
