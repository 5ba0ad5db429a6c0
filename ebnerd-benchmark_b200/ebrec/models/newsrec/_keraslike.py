"""Minimal Keras-shaped host facade over a device engine.

The reference drives its models through ``keras.Model`` (third-party): ``summary``,
``compile``, ``fit``, ``predict``, ``save_weights`` ... (call sites:
examples/quick_start/nrms_dummy.py:16,46-47; examples/reproducibility_scripts/ebnerd_nrms.py:244-260,302,338).
These classes keep that surface -- same method names, argument meaning and return types
(numpy arrays, ``History`` with ``.history``) -- while every FLOP runs in the ebk CUDA
library.  No TensorFlow import; callbacks are duck-typed (``set_model``, ``on_train_begin``,
``on_epoch_begin``, ``on_epoch_end(epoch, logs)``, ``on_train_end``).
"""
from __future__ import annotations

import queue
import threading
import time

import numpy as np
import torch


class History:
    def __init__(self):
        self.history: dict[str, list[float]] = {}
        self.epoch: list[int] = []
        self.model = None

    def _append(self, epoch: int, logs: dict):
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(v)


class _LR:
    """Mutable learning-rate handle; behaves like a float and like a tf.Variable (assign/numpy)."""

    def __init__(self, engine):
        self._e = engine

    def numpy(self):
        return np.float32(self._e.lr)

    def assign(self, v):
        self._e.lr = float(v)

    def __float__(self):
        return float(self._e.lr)

    def __repr__(self):
        return f"{self._e.lr}"


class AdamHandle:
    """What ``model.optimizer`` returns: Keras-form Adam state lives in the engine."""

    def __init__(self, engine):
        self._e = engine
        self._lr = _LR(engine)

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, v):
        self._e.lr = float(v)

    learning_rate = lr

    @property
    def iterations(self):
        return self._e.step_count

    def get_config(self):
        e = self._e
        return {"name": "Adam", "learning_rate": e.lr, "beta_1": e.beta1, "beta_2": e.beta2, "epsilon": e.eps}


def keras_auc(y_true: np.ndarray, y_pred: np.ndarray, num_thresholds: int = 200) -> float:
    """keras.metrics.AUC(num_thresholds=200, curve='ROC', summation_method='interpolation') over
    flattened predictions -- what ``metrics=['AUC']`` means in ebnerd_nrms.py:244-248."""
    yt = np.asarray(y_true).reshape(-1) > 0.5
    yp = np.asarray(y_pred, dtype=np.float64).reshape(-1)
    eps = 1e-7
    th = np.array([0.0 - eps] + [(i + 1) / (num_thresholds - 1) for i in range(num_thresholds - 2)] + [1.0 + eps])
    order = np.sort(yp[yt])
    tp = order.size - np.searchsorted(order, th, side="right")
    order_n = np.sort(yp[~yt])
    fp = order_n.size - np.searchsorted(order_n, th, side="right")
    fn, tn = yt.sum() - tp, (~yt).sum() - fp
    with np.errstate(divide="ignore", invalid="ignore"):
        tpr = np.where(tp + fn > 0, tp / (tp + fn), 0.0)
        fpr = np.where(fp + tn > 0, fp / (fp + tn), 0.0)
    return float(np.sum((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0))


class _Prefetcher:
    """Background-thread batch fetch (Keras' OrderedEnqueuer with workers=1 does the same)."""

    def __init__(self, fetch, order, depth=3):
        self.q: queue.Queue = queue.Queue(maxsize=depth)
        self.t = threading.Thread(target=self._run, args=(fetch, order), daemon=True)
        self.t.start()

    def _run(self, fetch, order):
        try:
            for i in order:
                self.q.put(fetch(i))
        except BaseException as e:  # surface loader errors in the consumer
            self.q.put(e)
        self.q.put(None)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item


def shard_for_rank(order: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Data parallel: the batches (or samples) of one epoch that THIS rank trains on.  Every rank draws the same
    shuffled `order` (same seed) and takes every world-th entry; the tail that cannot be split evenly is dropped
    so that all ranks run the same number of optimizer steps (each step ends in a collective).  The reference has
    no data parallelism; with world == 1 this is the identity."""
    if world <= 1:
        return order
    n = (len(order) // world) * world
    return order[:n][rank::world]


def _unpack_xy(data):
    """Keras accepts validation_data / x as (inputs, y): split such a pair, leave loaders and bare inputs alone."""
    if isinstance(data, (tuple, list)) and len(data) == 2 and isinstance(data[0], (tuple, list)):
        return tuple(data[0]), data[1]
    return data, None


def _is_sequence(x) -> bool:
    return hasattr(x, "__getitem__") and hasattr(x, "__len__") and not isinstance(x, (tuple, list, np.ndarray))


class KerasLikeModel:
    """Common fit/predict loop.  Subclasses provide ``_train_batch``, ``_eval_batch``,
    ``_predict_batch`` (host arrays in, numpy out) and weight accessors via ``self._engine``."""

    def __init__(self, owner, engine, name, head):
        self._owner, self._engine, self.name, self._head = owner, engine, name, head
        self.optimizer = AdamHandle(engine)
        self.loss = None
        self._metrics: list[str] = []
        self.stop_training = False
        self.history = None

    # ---- Keras surface ------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, metrics=None, **_):
        if loss is not None:
            self.loss = loss
            kinds = {"categorical_crossentropy": 0, "binary_crossentropy": 1}   # _ebk.LOSS_*
            if isinstance(loss, str):
                if loss not in kinds:
                    raise ValueError(f"loss {loss!r} is not on the B200 path (categorical_crossentropy / binary_crossentropy)")
                self._engine.loss_kind = kinds[loss]
        if metrics is not None:
            self._metrics = [str(m).lower() for m in metrics]
        if optimizer is not None and not isinstance(optimizer, AdamHandle):
            lr = getattr(optimizer, "learning_rate", getattr(optimizer, "lr", None))
            if lr is not None:
                self._engine.lr = float(lr.numpy() if hasattr(lr, "numpy") else lr)

    def count_params(self) -> int:
        return self._engine.count_params()

    def get_weights(self):
        return self._engine.get_weights()

    def set_weights(self, weights):
        self._engine.set_weights(weights)

    @property
    def variables(self):
        return self.get_weights()

    def save_weights(self, filepath, overwrite=True, **_):
        """Checkpoint = {weights in Keras get_weights() order, Adam m / v / iteration, lr} as a torch.save file
        (NOT Keras' h5 / TF-checkpoint format: there is no TensorFlow here; `get_weights()` arrays are in the
        reference's order, so `reference_model.set_weights(model.get_weights())` moves weights across).
        Data parallel: COLLECTIVE -- every rank calls it (ModelCheckpoint runs on all ranks); the rank-sharded
        optimizer state is gathered, rank 0 alone writes, and a barrier follows."""
        import os

        e = self._engine
        world = getattr(e, "world", 1)
        if hasattr(e, "sync_optimizer_state"):
            e.sync_optimizer_state()
        weights = e.get_weights()
        if world <= 1 or e.rank == 0:
            os.makedirs(os.path.dirname(str(filepath)) or ".", exist_ok=True)
            if not overwrite and os.path.exists(str(filepath)):
                raise FileExistsError(str(filepath))
            state = {"weights": [torch.from_numpy(np.ascontiguousarray(w)) for w in weights], "step_count": e.step_count,
                     "adam_m": e.params.m.cpu(), "adam_v": e.params.v.cpu(), "lr": e.lr, "format": "ebk-1"}
            tmp = f"{filepath}.tmp{os.getpid()}"
            torch.save(state, tmp)
            os.replace(tmp, str(filepath))
        if world > 1:
            torch.distributed.barrier()

    def load_weights(self, filepath, **_):
        """Restores weights, Adam moments / iteration count and the learning rate.  A checkpoint whose optimizer
        state does not fit this model (different vocabulary / widths) raises instead of silently dropping it."""
        e = self._engine
        state = torch.load(str(filepath), map_location="cpu", weights_only=True)
        e.set_weights([w.numpy() for w in state["weights"]])
        if "adam_m" in state:
            if state["adam_m"].numel() != e.params.m.numel():
                raise ValueError(f"{filepath}: optimizer state of {state['adam_m'].numel()} parameters does not fit this "
                                 f"model ({e.params.m.numel()})")
            e.params.m.copy_(state["adam_m"])
            e.params.v.copy_(state["adam_v"])
            e.step_count = int(state["step_count"])
        if "lr" in state:
            e.lr = float(state["lr"])
        if hasattr(e, "_table_stale"):
            e._table_stale = False   # every replica now holds the full table

    def summary(self, print_fn=print):
        e = self._engine
        print_fn(f'Model: "{self.name}"')
        print_fn("_" * 65)
        print_fn(f"{'Parameter (HBM, fp32)':<40}{'Shape':<18}{'Param #':>7}")
        print_fn("=" * 65)
        for n, s in e.params.spec:
            print_fn(f"{n:<40}{str(tuple(s)):<18}{int(np.prod(s)):>7}")
        print_fn("=" * 65)
        print_fn(f"Total params: {e.count_params():,}")
        print_fn(f"Trainable params: {e.trainable_params():,}")
        print_fn("_" * 65)

    # ---- batching -------------------------------------------------------------------
    @staticmethod
    def _n_samples(x):
        return int(np.asarray(x[0]).shape[0])

    def _batches(self, x, y, batch_size, shuffle, rng, shard=False):
        """Yield (inputs_tuple, y_or_None).  Arrays: Keras shuffles samples; Sequence: batch order.
        shard (training under data parallel): this rank's share of every epoch, see shard_for_rank."""
        if _is_sequence(x):
            if getattr(x, "device_feed", False):
                # device-resident feed: upload the loader's lookup matrix (token ids, doc vectors) once; batches
                # are article row indices
                if getattr(self._engine, "_article_matrix_src", None) is not x.lookup_article_matrix:
                    if hasattr(x, "lookup_article_matrix_body"):      # NAML: title + body token matrices
                        self._engine.set_article_matrices(x.lookup_article_matrix, x.lookup_article_matrix_body)
                    else:
                        self._engine.set_article_matrix(x.lookup_article_matrix)
                    self._engine._article_matrix_src = x.lookup_article_matrix
            order = np.arange(len(x))
            if shuffle:
                rng.shuffle(order)
            if shard:
                order = shard_for_rank(order, self._engine.rank, self._engine.world)

            def fetch(i):
                item = x[int(i)]
                return (item[0], item[1]) if isinstance(item, tuple) and len(item) == 2 else (item, None)

            yield from _Prefetcher(fetch, order)
            return
        xs = tuple(np.asarray(a) for a in x)
        n = self._n_samples(xs)
        bs = int(batch_size or 32)
        idx = np.arange(n)
        if shuffle:
            rng.shuffle(idx)
        world = self._engine.world if shard else 1
        if world > 1:
            # a global step = world consecutive batches of bs samples, one per rank; the uneven tail is dropped
            steps = n // (bs * world)
            for g in range(steps):
                lo = (g * world + self._engine.rank) * bs
                sel = idx[lo: lo + bs]
                yield tuple(a[sel] for a in xs), (None if y is None else np.asarray(y)[sel])
            return
        for s in range(0, n, bs):
            sel = idx[s: s + bs]
            if not shuffle:
                sel = slice(s, s + bs)
            yield tuple(a[sel] for a in xs), (None if y is None else np.asarray(y)[sel])

    # ---- fit / evaluate / predict -----------------------------------------------------
    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
            shuffle=True, initial_epoch=0, **_):
        cbs = list(callbacks or [])
        hist = History()
        hist.model = self
        self.history = hist
        self.stop_training = False
        for cb in cbs:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        for cb in cbs:
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin({})
        # the shuffle stream lives on the model: fit(epochs=1) called in a loop sees a new order every time
        rng = self.__dict__.setdefault("_shuffle_rng", np.random.default_rng(self._engine.seed + 7919))
        world = getattr(self._engine, "world", 1)
        if y is None:
            x, y = _unpack_xy(x)
        for epoch in range(initial_epoch, epochs):
            for cb in cbs:
                if hasattr(cb, "on_epoch_begin"):
                    cb.on_epoch_begin(epoch, {})
            t0 = time.time()
            tot_loss = torch.zeros(1, device=self._engine.device)
            n_seen = 0
            ys, ps = [], []
            nb = 0
            loss_host = self.__dict__.get("_loss_host")
            if loss_host is None:   # pinned ring the per-step losses are copied to (what a progress bar reads)
                loss_host = self._loss_host = torch.zeros(64, dtype=torch.float32).pin_memory()
            for inputs, yb in self._batches(x, y, batch_size, shuffle, rng, shard=True):
                loss_dev, probs_dev, bsz = self._train_batch(inputs, yb)
                tot_loss += loss_dev * bsz  # stays on device: no per-step host sync
                loss_host[nb % 64: nb % 64 + 1].copy_(loss_dev.reshape(-1)[:1], non_blocking=True)   # async D2H, 4 bytes
                n_seen += bsz
                nb += 1
                if "auc" in self._metrics:
                    ys.append(np.asarray(yb))
                    ps.append(probs_dev.clone())
            if world > 1:   # every rank logs the GLOBAL epoch mean (callbacks then decide identically on all ranks)
                agg = torch.cat([tot_loss.double(), torch.tensor([float(n_seen)], device=tot_loss.device, dtype=torch.float64)])
                torch.distributed.all_reduce(agg)
                logs = {"loss": float(agg[0]) / max(float(agg[1]), 1.0)}
            else:
                logs = {"loss": float(tot_loss) / max(n_seen, 1)}
            if "auc" in self._metrics and ys:
                yt, yp = np.concatenate(ys), torch.cat(ps).cpu().numpy()
                if world > 1:
                    parts = [None] * world
                    torch.distributed.all_gather_object(parts, (yt, yp))
                    yt, yp = np.concatenate([a for a, _ in parts]), np.concatenate([b for _, b in parts])
                logs["auc"] = keras_auc(yt, yp)
            if validation_data is not None:
                vlogs = self.evaluate(validation_data, verbose=0, return_dict=True)
                logs.update({f"val_{k}": v for k, v in vlogs.items()})
            logs["lr"] = float(self._engine.lr)
            if world > 1:
                # every rank evaluated the validation set itself; atomics-order noise in the last bits must not let
                # callbacks (EarlyStopping, ReduceLROnPlateau) decide differently per rank: rank 0's logs are THE logs
                box = [logs]
                torch.distributed.broadcast_object_list(box, src=0)
                logs = box[0]
            if verbose:
                dt = time.time() - t0
                msg = " - ".join(f"{k}: {v:.4f}" for k, v in logs.items() if k != "lr")
                print(f"Epoch {epoch + 1}/{epochs}\n{nb}/{nb} - {dt:.0f}s {1e3 * dt / max(nb, 1):.0f}ms/step - {msg}")
            hist._append(epoch, logs)
            for cb in cbs:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in cbs:
            if hasattr(cb, "on_train_end"):
                cb.on_train_end({})
        return hist

    def evaluate(self, x=None, y=None, batch_size=None, verbose=0, return_dict=False, **_):
        """Data parallel: every rank evaluates the WHOLE set (collective-free apart from the table sync), so all
        ranks see identical val_* logs."""
        if y is None:
            x, y = _unpack_xy(x)
        tot, n_seen, ys, ps = 0.0, 0, [], []
        rng = np.random.default_rng(0)
        for inputs, yb in self._batches(x, y, batch_size, False, rng):
            loss, probs, bsz = self._eval_batch(inputs, yb)
            tot += loss * bsz
            n_seen += bsz
            ys.append(np.asarray(yb))
            ps.append(probs)
        logs = {"loss": tot / max(n_seen, 1)}
        if "auc" in self._metrics and ys:
            logs["auc"] = keras_auc(np.concatenate(ys), np.concatenate(ps))
        if return_dict:
            return logs
        return logs["loss"] if len(logs) == 1 else list(logs.values())

    def predict(self, x, batch_size=None, verbose=0, **_):
        outs = []
        rng = np.random.default_rng(0)
        for inputs, _y in self._batches(x, None, batch_size, False, rng):
            outs.append(self._predict_batch(inputs))
        if not outs:
            return np.zeros((0, 1), np.float32)
        return np.concatenate(outs, axis=0)

    def train_on_batch(self, x, y):
        loss, _, _ = self._train_batch(tuple(np.asarray(a) for a in x), np.asarray(y))
        return float(loss)

    def test_on_batch(self, x, y):
        loss, _, _ = self._eval_batch(tuple(np.asarray(a) for a in x), np.asarray(y))
        return float(loss)

    def predict_on_batch(self, x):
        return self._predict_batch(tuple(np.asarray(a) for a in x))

    def __call__(self, x, training=False):
        return self.predict_on_batch(x)
