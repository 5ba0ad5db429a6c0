"""NAMLModel -- B200-native drop-in for the reference's src/ebrec/models/newsrec/naml.py:13-374
(constructor path of base_model.py:19-86).

Same constructor arguments (``hparams, n_users=50000, word2vec_embedding=None, seed=None, **kwargs`` with the
BaseModel extras ``word_emb_dim`` / ``vocab_size``), attributes (``model``, ``scorer``, ``newsencoder``,
``userencoder``, ``hparams``, ``seed``, ``word2vec_embedding``, ``n_users``, ``loss``, ``train_optimizer``) and
error behaviour (ValueError for an unknown loss / optimizer, base_model.py:61-86).  ``model`` takes the eight
arrays of ``NAMLDataLoader`` (his title/body/vert/subvert, pred title/body/vert/subvert), ``scorer`` the same
with one candidate per row (naml.py:346-372).
"""
from __future__ import annotations

import numpy as np

from . import _ebk
from ._engine_naml import NAMLEngine
from ._keraslike import KerasLikeModel
from .nrms import _EncoderView, glorot_uniform

__all__ = ["NAMLModel"]


class _NAMLTrainModel(KerasLikeModel):
    @staticmethod
    def _n_samples(x):
        return int(np.asarray(x[0]).shape[0])

    def _pack(self, inputs, y=None):
        if len(inputs) != 8:
            raise ValueError(f"NAML expects 8 input arrays (naml.py:346-357), got {len(inputs)}")
        pred_title = np.asarray(inputs[4])
        B, C_ = pred_title.shape[0], pred_title.shape[1]
        x, lab = self._engine.to_device_batch(inputs, y)
        return x, lab, B, C_

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


class NAMLModel:
    """NAML (Neural News Recommendation with Attentive Multi-View Learning, Wu et al., IJCAI 2019)."""

    def __init__(self, hparams, n_users: int = 50000, word2vec_embedding=None, seed=None, word_emb_dim: int = 300,
                 vocab_size: int = 32000, math: int = _ebk.MATH_TF32, **kwargs):
        self.n_users = n_users
        self.seed = seed
        self._math = math
        np.random.seed(seed)  # base_model.py:35-37
        self.hparams = hparams
        if word2vec_embedding is None:
            self.word2vec_embedding = np.random.rand(vocab_size, word_emb_dim)  # base_model.py:43-44
        else:
            self.word2vec_embedding = word2vec_embedding
        self.loss = self._get_loss(hparams.loss)
        self.train_optimizer = self._get_opt(hparams.optimizer, hparams.learning_rate)
        self.model, self.scorer = self._build_graph()
        self.model.compile(loss=self.loss)

    def _get_loss(self, loss: str):
        if loss == "cross_entropy_loss":
            return "categorical_crossentropy"
        elif loss == "log_loss":
            return "binary_crossentropy"   # base_model.py:63-66 (ebk_score_loss, EBK_LOSS_BINARY_CE)
        raise ValueError(f"this loss not defined {loss}")  # base_model.py:72

    def _get_opt(self, optimizer: str, lr: float):
        if optimizer != "adam":
            raise ValueError(f"this optimizer not defined {optimizer}")  # base_model.py:84
        return optimizer

    @staticmethod
    def _is_relu(name) -> bool:
        if name in ("relu",):
            return True
        if name in (None, "linear"):
            return False
        raise NotImplementedError(f"activation {name!r} is not on the B200 path (relu / linear only)")

    def _build_graph(self):
        hp = self.hparams
        table = np.asarray(self.word2vec_embedding, dtype=np.float32)
        V, E = table.shape
        F, A, w = hp.filter_num, hp.attention_hidden_dim, hp.window_size
        self._engine = NAMLEngine(V=V, E=E, T=hp.title_size, Tb=hp.body_size, H=hp.history_size, F=F, att=A, window=w,
                                  vert_num=hp.vert_num, vert_dim=hp.vert_emb_dim, subvert_num=hp.subvert_num,
                                  subvert_dim=hp.subvert_emb_dim, dropout=hp.dropout, lr=hp.learning_rate,
                                  cnn_relu=self._is_relu(hp.cnn_activation), dense_relu=self._is_relu(hp.dense_activation),
                                  seed=self.seed, math=self._math)
        s = self.seed
        rng = np.random.default_rng(None if s is None else [int(s), 0xE3B])
        weights = [table]
        for _ in ("title", "body"):  # Conv1D: glorot over fan_in = w*E, fan_out = w*F (Keras receptive-field rule)
            limit = np.sqrt(6.0 / (w * E + w * F))
            cw = np.random.default_rng(None if s is None else [int(s), w * E, F]).uniform(-limit, limit, (w, E, F))
            weights += [cw.astype(np.float32), np.zeros(F, np.float32), glorot_uniform(s, (F, A)), np.zeros(A, np.float32),
                        glorot_uniform(s, (A, 1))]
        for n, d in ((hp.vert_num, hp.vert_emb_dim), (hp.subvert_num, hp.subvert_emb_dim)):
            weights += [rng.uniform(-0.05, 0.05, (n, d)).astype(np.float32),  # Keras Embedding default initializer
                        glorot_uniform(s, (d, F)), np.zeros(F, np.float32)]
        for _ in ("news", "user"):
            weights += [glorot_uniform(s, (F, A)), np.zeros(A, np.float32), glorot_uniform(s, (A, 1))]
        self._engine.set_weights(weights)
        self._engine.loss_kind = (_ebk.LOSS_BINARY_CE if self._get_loss(hp.loss) == "binary_crossentropy"
                                  else _ebk.LOSS_CATEGORICAL_CE)
        model = _NAMLTrainModel(self, self._engine, "model", "softmax")
        scorer = _NAMLTrainModel(self, self._engine, "scorer", "sigmoid")
        self.newsencoder = _EncoderView(self._engine, "news")
        self.userencoder = _EncoderView(self._engine, "user")
        return model, scorer
