"""NRMSDocVec -- B200-native drop-in for the reference's src/ebrec/models/newsrec/nrms_docvec.py:8-188.

NRMS whose news encoder is an MLP over a precomputed document vector (no token path).  Same constructor,
attributes and error behaviour as the reference class; the shipped-broken callers are tolerated
(SURVEY.md section 7 item 9): an extra ``newsencoder_units_per_layer=`` keyword is accepted
(examples/quick_start/nrms_docvec_dummy.py:17) and ``hparams.newsencoder_units_per_layer = None`` means
"no hidden layers".
"""
from __future__ import annotations

import numpy as np

from . import _ebk
from ._engine_docvec import DocVecEngine
from ._keraslike import KerasLikeModel
from .nrms import _EncoderView, glorot_uniform


class _DocVecTrainModel(KerasLikeModel):
    def _pack(self, inputs, y=None):
        his, pred = (np.asarray(a) for a in inputs)
        if his.ndim != pred.ndim or his.ndim not in (2, 3):
            raise ValueError(f"expected his [B,H,Ddoc] and pred [B,C,Ddoc] (or [B,H] / [B,C] article indices of a "
                             f"device-feed loader), got {his.shape} and {pred.shape}")
        x, lab = self._engine.to_device_batch(his, pred, y)
        return x, lab, pred.shape[0], pred.shape[1]

    def _train_batch(self, inputs, y):
        x, lab, B, C_ = self._pack(inputs, y)
        loss, probs = self._engine.train_step_dev(x, lab, B, C_)
        return loss, probs, B

    def _eval_batch(self, inputs, y):
        x, lab, B, C_ = self._pack(inputs, y)
        loss, probs = self._engine.eval_loss_dev(x, lab, B, C_)
        return float(loss), probs.cpu().numpy(), B

    def _predict_batch(self, inputs):
        x, _, B, C_ = self._pack(inputs)
        return self._engine.predict_dev(x, B, C_, head=self._head).cpu().numpy()


class NRMSDocVec:
    """Modified NRMS (Wu et al. 2019): the news encoder embeds a document vector (nrms_docvec.py:8-19)."""

    def __init__(self, hparams, seed: int = None, math: int = _ebk.MATH_TF32, **kwargs):
        self.hparams = hparams
        self.seed = seed
        self._math = math
        np.random.seed(seed)
        units = kwargs.get("newsencoder_units_per_layer", getattr(hparams, "newsencoder_units_per_layer", None))
        self._units = list(units) if units else []
        data_loss = self._get_loss(hparams.loss)
        self._get_opt(hparams.optimizer, hparams.learning_rate)
        self.model, self.scorer = self._build_graph()
        self.model.compile(loss=data_loss)

    def _get_loss(self, loss: str):
        if loss == "cross_entropy_loss":
            return "categorical_crossentropy"
        elif loss == "log_loss":
            return "binary_crossentropy"   # nrms_docvec.py:46-47 (ebk_score_loss, EBK_LOSS_BINARY_CE)
        raise ValueError(f"this loss not defined {loss}")  # nrms_docvec.py:49

    def _get_opt(self, optimizer: str, lr: float):
        if optimizer != "adam":
            raise ValueError(f"this optimizer not defined {optimizer}")  # nrms_docvec.py:62
        return optimizer

    def _build_graph(self):
        hp = self.hparams
        Dd, D, A = hp.title_size, hp.head_num * hp.head_dim, hp.attention_hidden_dim
        self._engine = DocVecEngine(Ddoc=Dd, units=self._units, H=hp.history_size, nh=hp.head_num, dh=hp.head_dim, att=A,
                                    dropout=hp.dropout, lr=hp.learning_rate, l2=hp.newsencoder_l2_regularization,
                                    seed=self.seed, math=self._math)
        s, w, din = self.seed, [], Dd
        for u in self._units:  # Dense default init: GlorotUniform kernel, zero bias; BN: gamma 1, beta 0, mean 0, var 1
            w += [glorot_uniform(None if s is None else s + din + u, (din, u)), np.zeros(u, np.float32), np.ones(u, np.float32),
                  np.zeros(u, np.float32), np.zeros(u, np.float32), np.ones(u, np.float32)]
            din = u
        w += [glorot_uniform(None if s is None else s + 7, (din, D)), np.zeros(D, np.float32)]
        w += [glorot_uniform(s, (D, D)), glorot_uniform(s, (D, D)), glorot_uniform(s, (D, D)), glorot_uniform(s, (D, A)),
              np.zeros(A, np.float32), glorot_uniform(s, (A, 1))]
        self._engine.set_weights(w)
        self._engine.loss_kind = (_ebk.LOSS_BINARY_CE if self._get_loss(hp.loss) == "binary_crossentropy"
                                  else _ebk.LOSS_CATEGORICAL_CE)
        model = _DocVecTrainModel(self, self._engine, "model", "softmax")
        scorer = _DocVecTrainModel(self, self._engine, "scorer", "sigmoid")
        self.newsencoder = _EncoderView(self._engine, "news")
        self.userencoder = _EncoderView(self._engine, "user")
        return model, scorer
