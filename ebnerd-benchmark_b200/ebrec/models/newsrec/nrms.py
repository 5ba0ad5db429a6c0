"""NRMSModel -- B200-native drop-in for the reference's src/ebrec/models/newsrec/nrms.py.

Same constructor, attributes (``model``, ``scorer``, ``newsencoder``, ``userencoder``,
``hparams``, ``seed``, ``word2vec_embedding``) and error behaviour (ValueError for unknown
loss / optimizer, nrms.py:56-80) as the reference class (nrms.py:12-210); the Keras graph
is replaced by an :class:`NRMSEngine` that keeps all parameters in HBM and runs the whole
per-impression forward/backward/Adam through the ebk CUDA library.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _ebk
from ._engine import NRMSEngine
from ._keraslike import KerasLikeModel


def glorot_uniform(seed, shape):
    """keras GlorotUniform(seed)(shape).  Keras returns the SAME tensor for the same (seed, shape)
    (WQ == WK == WV at init); with seed=None draws are independent."""
    rng = np.random.default_rng(None if seed is None else [int(seed), int(shape[0]), int(shape[-1])])
    limit = np.sqrt(6.0 / (shape[0] + shape[-1]))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


class _NRMSTrainModel(KerasLikeModel):
    """``NRMSModel.model`` ([his, pred] -> softmax over candidates) and, with head='sigmoid',
    ``NRMSModel.scorer`` ([his, pred_one] -> sigmoid)  (nrms.py:201-208)."""

    def _pack(self, inputs, y=None):
        his, pred = (np.asarray(a) for a in inputs)
        if his.ndim != pred.ndim or his.ndim not in (2, 3):
            raise ValueError(f"expected his [B,H,T] and pred [B,C,T] (or [B,H] / [B,C] article indices of a device-feed "
                             f"loader), got {his.shape} and {pred.shape}")
        B, C_ = pred.shape[0], pred.shape[1]
        tok, lab = self._engine.to_device_batch(his, pred, y)
        return tok, lab, B, C_

    def _train_batch(self, inputs, y):
        his, pred = (np.asarray(a) for a in inputs)
        if his.ndim == 3 and pred.ndim == 3 and hasattr(self._engine, "train_step_host"):
            loss, probs = self._engine.train_step_host(his, pred, y)   # graph-replayed step: H2D into its static inputs
            return loss, probs, pred.shape[0]
        tok, lab, B, C_ = self._pack(inputs, y)
        loss, probs = self._engine.train_step_dev(tok, lab, B, C_)
        return loss, probs, B

    def _eval_batch(self, inputs, y):
        tok, lab, B, C_ = self._pack(inputs, y)
        loss, probs = self._engine.eval_loss_dev(tok, lab, B, C_)
        return float(loss), probs.cpu().numpy(), B

    def _predict_batch(self, inputs):
        # distinct articles / histories are encoded once (the eval-mode loader repeats the history per candidate)
        his, pred = (np.asarray(a) for a in inputs)
        if his.ndim != pred.ndim or his.ndim not in (2, 3):
            raise ValueError(f"expected his [B,H,T] and pred [B,C,T] (or article indices), got {his.shape} and {pred.shape}")
        return self._engine.predict_host_dedup(his, pred, head=self._head).cpu().numpy()


class _EncoderView:
    """``model.newsencoder`` / ``model.userencoder``: predict-only views on the shared engine."""

    def __init__(self, engine, kind):
        self._engine, self._kind, self.name = engine, kind, f"{kind}_encoder"

    def predict(self, x, batch_size=None, verbose=0, **_):
        x = np.asarray(x)
        outs = []
        bs = int(batch_size or 32)
        for s in range(0, x.shape[0], bs):
            outs.append(self._engine.encode_host(self._kind, x[s: s + bs]))
        return np.concatenate(outs, axis=0) if outs else np.zeros((0, self._engine.D), np.float32)

    __call__ = predict


class NRMSModel:
    """NRMS model (Neural News Recommendation with Multi-Head Self-Attention, Wu et al. 2019).

    Args mirror nrms.py:23-30.  Extra keyword ``math`` selects the contraction arithmetic
    (default: tcgen05 TF32 tensor cores; ``_ebk.MATH_FP32`` for the CUDA-core fp32 path).
    """

    def __init__(self, hparams, word2vec_embedding: np.ndarray = None, word_emb_dim: int = 300,
                 vocab_size: int = 32000, seed: int = None, math: int = _ebk.MATH_TF32):
        self.hparams = hparams
        self.seed = seed
        self._math = math
        np.random.seed(seed)  # nrms.py:36-37
        if word2vec_embedding is None:
            self.word2vec_embedding = glorot_uniform(seed, (vocab_size, word_emb_dim))  # nrms.py:40-43
        else:
            self.word2vec_embedding = word2vec_embedding
        data_loss = self._get_loss(hparams.loss)
        self._get_opt(hparams.optimizer, hparams.learning_rate)
        self.model, self.scorer = self._build_graph()
        self.model.compile(loss=data_loss)

    def _get_loss(self, loss: str):
        if loss == "cross_entropy_loss":
            return "categorical_crossentropy"
        elif loss == "log_loss":
            return "binary_crossentropy"   # nrms.py:63-64 (ebk_score_loss, EBK_LOSS_BINARY_CE)
        raise ValueError(f"this loss not defined {loss}")

    def _get_opt(self, optimizer: str, lr: float):
        if optimizer != "adam":
            raise ValueError(f"this optimizer not defined {optimizer}")
        return optimizer

    def _build_graph(self):
        hp = self.hparams
        table = np.asarray(self.word2vec_embedding, dtype=np.float32)
        V, E = table.shape
        D, A = hp.head_num * hp.head_dim, hp.attention_hidden_dim
        s = self.seed
        units = list(getattr(hp, "newsencoder_units_per_layer", None) or [])
        if units:
            # optional Dense/BatchNorm/Dropout stack between SelfAttention and AttLayer2 (nrms.py:142-152)
            from ._engine_nrms_dense import NRMSDenseEngine

            self._engine = NRMSDenseEngine(V=V, E=E, T=hp.title_size, H=hp.history_size, nh=hp.head_num, dh=hp.head_dim,
                                           att=A, units=units, l2=getattr(hp, "newsencoder_l2_regularization", 1e-4),
                                           dropout=hp.dropout, lr=hp.learning_rate, seed=self.seed,
                                           math=getattr(self, "_math", _ebk.MATH_TF32))
            weights = [table, glorot_uniform(s, (E, D)), glorot_uniform(s, (E, D)), glorot_uniform(s, (E, D))]
            din = D
            for u in units:  # Keras: Dense kernel GlorotUniform (unseeded), bias 0; BN gamma 1, beta 0, mean 0, var 1
                weights += [glorot_uniform(None, (din, u)), np.zeros((u,), np.float32), np.ones((u,), np.float32),
                            np.zeros((u,), np.float32), np.zeros((u,), np.float32), np.ones((u,), np.float32)]
                din = u
            weights += [glorot_uniform(s, (din, A)), np.zeros((A,), np.float32), glorot_uniform(s, (A, 1))]
            weights += [glorot_uniform(s, (D, D)), glorot_uniform(s, (D, D)), glorot_uniform(s, (D, D)),
                        glorot_uniform(s, (D, A)), np.zeros((A,), np.float32), glorot_uniform(s, (A, 1))]
        else:
            self._engine = NRMSEngine(V=V, E=E, T=hp.title_size, H=hp.history_size, nh=hp.head_num, dh=hp.head_dim,
                                      att=A, dropout=hp.dropout, lr=hp.learning_rate, seed=self.seed,
                                      math=getattr(self, "_math", _ebk.MATH_TF32))
            weights = [table]
            for din in (E, D):  # news encoder, then user encoder (Keras get_weights order)
                weights += [glorot_uniform(s, (din, D)), glorot_uniform(s, (din, D)), glorot_uniform(s, (din, D)),
                            glorot_uniform(s, (D, A)), np.zeros((A,), np.float32), glorot_uniform(s, (A, 1))]
        self._engine.set_weights(weights)
        self._engine.loss_kind = (_ebk.LOSS_BINARY_CE if self._get_loss(hp.loss) == "binary_crossentropy"
                                  else _ebk.LOSS_CATEGORICAL_CE)
        model = _NRMSTrainModel(self, self._engine, "model", "softmax")
        scorer = _NRMSTrainModel(self, self._engine, "scorer", "sigmoid")
        self.newsencoder = _EncoderView(self._engine, "news")
        self.userencoder = _EncoderView(self._engine, "user")
        return model, scorer
