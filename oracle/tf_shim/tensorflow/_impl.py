"""Minimal eager stand-in for the `tensorflow.keras` primitives the reference's NRMS-family
model files use.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Purpose: TensorFlow cannot be installed in the build container (Python 3.12, no wheel, no
network), so the reference's model math could not be executed at all.  With this package
on `sys.path` the reference's OWN source files

    /root/reference/src/ebrec/models/newsrec/{layers,nrms,nrms_docvec,naml,base_model}.py

import and run UNMODIFIED: their `AttLayer2.call`, `SelfAttention.call`, the graph wiring
of `_build_nrms` / `_build_naml` / ... are executed line by line; only the primitive ops
they call (K.dot, tf.matmul(adjoint_a=...), K.softmax, Embedding, Dense, Conv1D,
BatchNormalization, TimeDistributed, Dot, Activation, ...) are supplied here, each a few
lines over torch float64 with the documented TF/Keras semantics.  tests/golden/
make_reference_fixtures.py uses it to produce golden vectors (predictions, losses, and
gradients by torch autograd THROUGH the reference's forward code) that pin oracle/*.py
and the CUDA path.  What this does NOT pin: TensorFlow's own kernels and the Keras
training loop (loss reduction, Adam) -- those are restated here from their documentation.

Design: symbolic tensors (`KTensor`) carry a small concrete dummy value, so shape
inference is simply "run the layer on the dummy"; a functional `Model` re-evaluates the
recorded node graph on real inputs.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch

DT = torch.float64
_node_ids = itertools.count()
_DUMMY_BATCH = 2
_DUMMY_NONE = 3


class _Phase:
    training = False
    dropout_masks = None      # callable(shape, layer) -> keep-mask tensor or None (identity)
    bn_updates = True


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype.kind in "iub":
        return torch.from_numpy(a.astype(np.int64))
    return torch.from_numpy(a.astype(np.float64))


class KTensor:
    """Symbolic tensor: producing layer + its symbolic inputs + a concrete dummy value."""

    def __init__(self, dummy, layer=None, inputs=None, shape=None, call_kwargs=None):
        self.dummy, self.layer, self.inputs = dummy, layer, inputs
        self.id = next(_node_ids)
        self.call_kwargs = call_kwargs or {}
        self._shape = shape if shape is not None else (None,) + tuple(dummy.shape[1:])

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self.dummy.dtype


def _is_sym(x):
    if isinstance(x, KTensor):
        return True
    if isinstance(x, (list, tuple)):
        return any(_is_sym(a) for a in x)
    return False


def _map(f, x):
    if isinstance(x, (list, tuple)):
        return [_map(f, a) for a in x]
    return f(x)


def _flat(x):
    if isinstance(x, (list, tuple)):
        for a in x:
            yield from _flat(a)
    else:
        yield x


def Input(shape=None, dtype="float32", name=None, **_):
    if isinstance(shape, int):          # nrms_docvec.py:113 passes shape=(DOCUMENT_VECTOR_DIM) == an int
        shape = (shape,)
    dims = [_DUMMY_NONE if d is None else int(d) for d in shape]
    td = torch.int64 if "int" in str(dtype) else DT
    return KTensor(torch.zeros([_DUMMY_BATCH] + dims, dtype=td), shape=(None,) + tuple(shape))


# --------------------------------------------------------------------------------------------
# initializers / regularizers
# --------------------------------------------------------------------------------------------
class _Init:
    def __init__(self, seed=None, **_):
        self.seed = seed


class GlorotUniform(_Init):
    def __call__(self, shape, dtype=None):
        rng = np.random.default_rng(self.seed)
        fan_in = int(np.prod(shape[:-1])) if len(shape) > 1 else shape[0]
        fan_out = shape[-1]
        if len(shape) == 3:                      # Conv1D kernel [window, in, out]
            fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return torch.from_numpy(rng.uniform(-lim, lim, size=tuple(shape)))


class Zeros(_Init):
    def __call__(self, shape, dtype=None):
        return torch.zeros(tuple(shape), dtype=DT)


class Ones(_Init):
    def __call__(self, shape, dtype=None):
        return torch.ones(tuple(shape), dtype=DT)


class RandomUniform(_Init):
    def __call__(self, shape, dtype=None):
        rng = np.random.default_rng(self.seed)
        return torch.from_numpy(rng.uniform(-0.05, 0.05, size=tuple(shape)))


def _get_init(x, default):
    if x is None:
        return default()
    if isinstance(x, str):
        return {"zeros": Zeros, "ones": Ones, "glorot_uniform": GlorotUniform, "uniform": RandomUniform}[x]()
    return x


class L2:
    def __init__(self, l2=0.01):
        self.l2 = float(l2)

    def __call__(self, w):
        return self.l2 * (w * w).sum()


# --------------------------------------------------------------------------------------------
# Layer base, Model
# --------------------------------------------------------------------------------------------
class Layer:
    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        self.name = name or type(self).__name__.lower()
        self.trainable = trainable
        self.built = False
        self._weights: list[torch.Tensor] = []
        self._weight_names: list[str] = []
        self._reg: list = []

    # -- Keras API used by the reference's custom layers (layers.py:25-52, 144-173)
    def add_weight(self, name=None, shape=None, initializer=None, trainable=True, regularizer=None, **_):
        w = _get_init(initializer, GlorotUniform)(tuple(int(s) for s in shape)).to(DT).clone()
        w.requires_grad_(bool(trainable))
        w._keras_trainable = bool(trainable)
        self._weights.append(w)
        self._weight_names.append(name)
        if regularizer is not None:
            self._reg.append((regularizer, w))
        return w

    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        return inputs

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    def compute_mask(self, inputs, mask=None):
        return None

    @property
    def weights(self):
        return list(self._weights)

    def _sublayers(self):
        return []

    def _all_weights(self, seen):
        out = []
        for w in self._weights:
            if id(w) not in seen:
                seen.add(id(w))
                out.append(w)
        for l in self._sublayers():
            out += l._all_weights(seen)
        return out

    def _all_reg(self, seen):
        out = []
        if id(self) not in seen:
            seen.add(id(self))
            out += self._reg
            for l in self._sublayers():
                out += l._all_reg(seen)
        return out

    def _run(self, x, **kw):
        if not self.built:
            shp = _map(lambda a: (None,) + tuple(a.shape[1:]), x)
            if isinstance(x, (list, tuple)):
                shp = [tuple(s) for s in shp]
            self.build(shp)
            self.built = True
        return self.call(x, **kw)

    def __call__(self, inputs, **kwargs):
        if _is_sym(inputs):
            dummy_in = _map(lambda a: a.dummy, inputs)
            with torch.no_grad():
                dummy_out = self._run(dummy_in, **kwargs)
            return KTensor(dummy_out, layer=self, inputs=inputs, call_kwargs=kwargs)
        return self._run(_map(_t, inputs), **kwargs)


def _evaluate(out, feed: dict):
    """Value of symbolic tensor(s) `out` given {id(KTensor): value} for the graph inputs.  Nodes run in CREATION
    order (keras' functional executor walks nodes by depth and, inside a depth, in the order the layer calls were
    made): for NRMS that is user_encoder(history) before TimeDistributed(newsencoder)(candidates) -- it matters only
    for the order of the two BatchNormalization moving-average updates of a training step."""
    need, seen = [], set()

    def walk(t):
        if t.id in seen or t.id in feed:
            return
        seen.add(t.id)
        if t.layer is None:
            raise ValueError("graph input without a value")
        for i in _flat(t.inputs):
            walk(i)
        need.append(t)
    for o in _flat(out):
        walk(o)
    for t in sorted(need, key=lambda n: n.id):
        feed[t.id] = t.layer._run(_map(lambda a: feed[a.id], t.inputs), **t.call_kwargs)
    return _map(lambda a: feed[a.id], out)


class Model(Layer):
    """Functional model: keras.Model(inputs, outputs, name=...)."""

    def __init__(self, inputs=None, outputs=None, name=None, **kwargs):
        super().__init__(name=name or "model")
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = outputs
        self.built = True
        self.optimizer = None
        self.loss = None
        # layers in creation order of their (first) node, like keras' model.layers for these graphs
        nodes, seen = [], set()

        def walk(t):
            if not isinstance(t, KTensor) or t.id in seen:
                return
            seen.add(t.id)
            if t.layer is not None:
                for i in _flat(t.inputs):
                    walk(i)
                nodes.append(t)
        for o in _flat(outputs):
            walk(o)
        nodes.sort(key=lambda n: n.id)
        self.layers, lseen = [], set()
        for n in nodes:
            if id(n.layer) not in lseen:
                lseen.add(id(n.layer))
                self.layers.append(n.layer)

    def _sublayers(self):
        return self.layers

    @property
    def weights(self):
        return self._all_weights(set())

    @property
    def trainable_weights(self):
        return [w for w in self.weights if getattr(w, "_keras_trainable", True)]

    def get_weights(self):
        return [w.detach().numpy().copy() for w in self.weights]

    def set_weights(self, ws):
        cur = self.weights
        if len(ws) != len(cur):
            raise ValueError(f"expected {len(cur)} arrays, got {len(ws)}")
        with torch.no_grad():
            for w, a in zip(cur, ws):
                a = _t(np.asarray(a, dtype=np.float64))
                if tuple(a.shape) != tuple(w.shape):
                    raise ValueError(f"shape {tuple(a.shape)} != {tuple(w.shape)}")
                w.copy_(a)

    def count_params(self):
        return int(sum(w.numel() for w in self.weights))

    def call(self, x, **kw):
        xs = x if isinstance(x, (list, tuple)) else [x]
        if len(xs) != len(self.inputs):
            raise ValueError(f"model {self.name}: expected {len(self.inputs)} inputs, got {len(xs)}")
        feed = {}
        for s, v in zip(self.inputs, xs):
            v = _t(v)
            feed[s.id] = v.long() if s.dummy.dtype == torch.int64 else v.to(DT)
        return _evaluate(self.outputs, feed)

    # -- eager API
    def forward(self, x, training=False):
        old = _Phase.training
        _Phase.training = training
        try:
            return self.call(x)
        finally:
            _Phase.training = old

    def predict(self, x, batch_size=None, verbose=0, **_):
        with torch.no_grad():
            return self.forward(x, training=False).numpy()

    def compile(self, loss=None, optimizer=None, metrics=None, **_):
        self.loss, self.optimizer = loss, optimizer

    def regularization_loss(self):
        tot = torch.zeros((), dtype=DT)
        for reg, w in self._all_reg(set()):
            tot = tot + reg(w)
        return tot

    def loss_value(self, x, y, training=True):
        """Keras: mean over the batch of the per-sample loss of the compiled name + regularisers."""
        p = self.forward(x, training=training)
        y = _t(np.asarray(y, dtype=np.float64))
        eps = 1e-7
        z = getattr(p, "_keras_logits", None)
        if self.loss == "categorical_crossentropy":
            if z is not None:
                # keras backend: output of a softmax Activation -> softmax_cross_entropy_with_logits on its cached logits
                per = (torch.logsumexp(z, dim=-1, keepdim=True) * y - y * z).sum(dim=-1)
            else:  # on bare probabilities: renormalise, clip, -sum y log p
                q = p / p.sum(dim=-1, keepdim=True)
                per = -(y * torch.log(q.clamp(eps, 1.0 - eps))).sum(dim=-1)
        elif self.loss == "binary_crossentropy":
            if z is not None:
                # sigmoid_cross_entropy_with_logits: max(z,0) - z*y + log(1 + exp(-|z|)), mean over the last axis
                per = (z.clamp(min=0) - z * y + torch.log1p(torch.exp(-z.abs()))).mean(dim=-1)
            else:
                q = p.clamp(eps, 1.0 - eps)
                per = -(y * torch.log(q) + (1.0 - y) * torch.log(1.0 - q)).mean(dim=-1)
        else:
            raise ValueError(f"loss {self.loss!r}")
        return per.mean() + self.regularization_loss(), p

    def loss_and_grads(self, x, y, training=True):
        ws = self.trainable_weights
        for w in ws:
            w.grad = None
        loss, p = self.loss_value(x, y, training)
        loss.backward()
        grads = {id(w): (w.grad.clone() if w.grad is not None else torch.zeros_like(w)) for w in ws}
        return float(loss), p.detach().numpy(), [grads.get(id(w), torch.zeros_like(w)).numpy() for w in self.weights]

    def train_on_batch(self, x, y):
        loss, _, grads = self.loss_and_grads(x, y, training=True)
        self.optimizer.apply(self.weights, grads)
        return loss

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"  params: {self.count_params():,}')


class Adam:
    """tf.keras.optimizers.Adam (2.11+): eps outside the bias correction, dense (non-lazy) update."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **_):
        self.learning_rate, self.b1, self.b2, self.eps, self.t = float(learning_rate), beta_1, beta_2, epsilon, 0
        self.m, self.v = {}, {}

    lr = property(lambda self: self.learning_rate)

    def apply(self, weights, grads):
        self.t += 1
        alpha = self.learning_rate * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for w, g in zip(weights, grads):
                if not getattr(w, "_keras_trainable", True):
                    continue
                g = _t(g)
                m = self.m.setdefault(id(w), torch.zeros_like(w))
                v = self.v.setdefault(id(w), torch.zeros_like(w))
                m += (g - m) * (1.0 - self.b1)
                v += (g * g - v) * (1.0 - self.b2)
                w -= (m * alpha) / (v.sqrt() + self.eps)


# --------------------------------------------------------------------------------------------
# stock layers
# --------------------------------------------------------------------------------------------
def _activation(name):
    if name is None or name == "linear":
        return lambda x: x
    if callable(name):
        return name
    return {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid,
            "softmax": lambda x: torch.softmax(x, dim=-1)}[name]


class InputLayer(Layer):
    pass


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, weights=None, trainable=True, embeddings_initializer="uniform",
                 mask_zero=False, **kw):
        super().__init__(trainable=trainable, **kw)
        self.input_dim, self.output_dim, self._init_w, self._einit = int(input_dim), int(output_dim), weights, embeddings_initializer

    def build(self, input_shape):
        self.embeddings = self.add_weight("embeddings", (self.input_dim, self.output_dim),
                                          _get_init(self._einit, RandomUniform), trainable=self.trainable)
        if self._init_w is not None:
            with torch.no_grad():
                self.embeddings.copy_(_t(np.asarray(self._init_w[0], dtype=np.float64)))

    def call(self, x, **_):
        return self.embeddings[x.long()]


class Dropout(Layer):
    def __init__(self, rate, seed=None, **kw):
        super().__init__(**kw)
        self.rate = float(rate)

    def call(self, x, training=None, **_):
        if not _Phase.training or self.rate <= 0.0:
            return x
        if _Phase.dropout_masks is None:
            raise RuntimeError("training-mode Dropout needs an explicit mask provider (set _Phase.dropout_masks)")
        keep = _Phase.dropout_masks(tuple(x.shape), self)
        if keep is None:          # provider says: this Dropout is switched off (rate 0)
            return x
        return x * keep.to(x.dtype) / (1.0 - self.rate)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, **kw):
        super().__init__(**kw)
        self.units, self.act, self.use_bias = int(units), _activation(activation), use_bias
        self._ki, self._bi, self._kr = kernel_initializer, bias_initializer, kernel_regularizer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units), _get_init(self._ki, GlorotUniform),
                                      regularizer=self._kr)
        if self.use_bias:
            self.bias = self.add_weight("bias", (self.units,), _get_init(self._bi, Zeros))

    def call(self, x, **_):
        y = torch.matmul(x, self.kernel)
        if self.use_bias:
            y = y + self.bias
        return self.act(y)


class Conv1D(Layer):
    """padding='same', stride 1: out[t] = act(b + sum_w x[t + w - (K-1)//2] . kernel[w])  (TF pads the extra
    element of an even window on the right)."""

    def __init__(self, filters, kernel_size, activation=None, padding="valid", kernel_initializer=None,
                 bias_initializer=None, **kw):
        super().__init__(**kw)
        self.filters, self.k, self.act, self.padding = int(filters), int(kernel_size), _activation(activation), padding
        self._ki, self._bi = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (self.k, int(input_shape[-1]), self.filters), _get_init(self._ki, GlorotUniform))
        self.bias = self.add_weight("bias", (self.filters,), _get_init(self._bi, Zeros))

    def call(self, x, **_):
        N, L, E = x.shape
        if self.padding == "same":
            left = (self.k - 1) // 2
            xp = torch.cat([x.new_zeros(N, left, E), x, x.new_zeros(N, self.k - 1 - left, E)], dim=1)
            Lo = L
        else:
            xp, Lo = x, L - self.k + 1
        y = self.bias.expand(N, Lo, self.filters)
        for w in range(self.k):
            y = y + torch.matmul(xp[:, w: w + Lo, :], self.kernel[w])
        return self.act(y)


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, input_shape):
        n = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (n,), Ones())
        self.beta = self.add_weight("beta", (n,), Zeros())
        self.moving_mean = self.add_weight("moving_mean", (n,), Zeros(), trainable=False)
        self.moving_variance = self.add_weight("moving_variance", (n,), Ones(), trainable=False)

    def call(self, x, training=None, **_):
        if _Phase.training:
            red = tuple(range(x.dim() - 1))
            mean = x.mean(dim=red)
            var = ((x - mean) ** 2).mean(dim=red)          # biased, as Keras
            if _Phase.bn_updates:
                with torch.no_grad():
                    self.moving_mean.mul_(self.momentum).add_(mean.detach() * (1 - self.momentum))
                    self.moving_variance.mul_(self.momentum).add_(var.detach() * (1 - self.momentum))
        else:
            mean, var = self.moving_mean, self.moving_variance
        return (x - mean) / torch.sqrt(var + self.epsilon) * self.gamma + self.beta


class TimeDistributed(Layer):
    """Batch dim unknown -> keras reshapes [B, T, ...] to [B*T, ...], applies the layer, reshapes back."""

    def __init__(self, layer, **kw):
        super().__init__(**kw)
        self.layer = layer

    def _sublayers(self):
        return [self.layer]

    def call(self, x, **_):
        B, T = x.shape[0], x.shape[1]
        y = self.layer._run(x.reshape((B * T,) + tuple(x.shape[2:])))
        return y.reshape((B, T) + tuple(y.shape[1:]))


class Reshape(Layer):
    def __init__(self, target_shape, **kw):
        super().__init__(**kw)
        self.target_shape = tuple(target_shape)

    def call(self, x, **_):
        return x.reshape((x.shape[0],) + self.target_shape)


class Dot(Layer):
    """Dot(axes=-1)([a [B, C, D], b [B, D]]) -> [B, C] (batch dot over the last axes)."""

    def __init__(self, axes, **kw):
        super().__init__(**kw)
        assert axes == -1
        self.axes = axes

    def call(self, xs, **_):
        a, b = xs
        if a.dim() == 3 and b.dim() == 2:
            return torch.einsum("bcd,bd->bc", a, b)
        if a.dim() == 2 and b.dim() == 2:
            return (a * b).sum(dim=-1, keepdim=True)
        raise ValueError((a.shape, b.shape))


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.act = _activation(activation)
        self.name_of_act = activation if isinstance(activation, str) else None

    def call(self, x, **_):
        y = self.act(x)
        if self.name_of_act in ("softmax", "sigmoid"):
            y._keras_logits = x     # keras caches the logits of a softmax / sigmoid Activation for its losses
        return y


class Lambda(Layer):
    def __init__(self, function, **kw):
        super().__init__(**kw)
        self.fn = function

    def call(self, x, **_):
        return self.fn(x)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, xs, **_):
        return torch.cat(list(xs), dim=self.axis)


class _Unsupported(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is outside the NRMS/NAML/NRMSDocVec path (tf_shim)")


# --------------------------------------------------------------------------------------------
# keras.backend subset used by layers.py
# --------------------------------------------------------------------------------------------
class backend:
    @staticmethod
    def epsilon():
        return 1e-7

    @staticmethod
    def dot(x, y):
        return torch.matmul(x, y)

    tanh = staticmethod(torch.tanh)
    exp = staticmethod(torch.exp)
    sqrt = staticmethod(lambda x: torch.sqrt(_t(x)) if not isinstance(x, float) else math.sqrt(x))

    @staticmethod
    def squeeze(x, axis):
        return x.squeeze(axis)

    @staticmethod
    def expand_dims(x, axis=-1):
        return x.unsqueeze(axis)

    @staticmethod
    def cast(x, dtype="float32"):
        if isinstance(x, torch.Tensor):
            return x.to(DT) if "float" in str(dtype) else x.long()
        return torch.tensor(float(x), dtype=DT) if "float" in str(dtype) else torch.tensor(int(x))

    @staticmethod
    def sum(x, axis=None, keepdims=False):
        return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)

    @staticmethod
    def shape(x):
        return tuple(x.shape)

    @staticmethod
    def reshape(x, shape):
        return x.reshape(tuple(int(s) for s in shape))

    @staticmethod
    def permute_dimensions(x, pattern):
        return x.permute(*pattern)

    @staticmethod
    def softmax(x, axis=-1):
        return torch.softmax(x, dim=axis)

    @staticmethod
    def ones_like(x):
        return torch.ones_like(x)

    @staticmethod
    def one_hot(indices, num_classes):
        return torch.nn.functional.one_hot(indices.long(), int(num_classes)).to(DT)

    @staticmethod
    def cumsum(x, axis=0):
        return torch.cumsum(x, dim=axis)

    @staticmethod
    def concatenate(xs, axis=-1):
        return torch.cat(list(xs), dim=axis)


def matmul(a, b, adjoint_a=False, adjoint_b=False, transpose_a=False, transpose_b=False):
    """tf.matmul on real tensors: adjoint == transpose of the last two axes."""
    if adjoint_a or transpose_a:
        a = a.transpose(-1, -2)
    if adjoint_b or transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)
