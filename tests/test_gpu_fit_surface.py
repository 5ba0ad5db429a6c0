"""The Keras training-loop surface the reproducibility scripts use
(/root/reference/examples/reproducibility_scripts/ebnerd_nrms.py:212-260, 287-348): fit with validation_data and
callbacks (EarlyStopping-restore, ModelCheckpoint save_weights, ReduceLROnPlateau through optimizer.lr), save_weights ->
load_weights round trips including the Adam state, device-resident feeds for NRMSDocVec and NAML, and the same under
data parallel (world 2, needs two GPUs: run by `gpurun --gpus 2`, skipped on one).

No TensorFlow here, so the callbacks below are small restatements of what the tf.keras callbacks DO to a model:
they only touch `model.get_weights/set_weights`, `model.stop_training`, `model.optimizer.lr` (read with `.numpy()`,
written with `.assign()` -- what keras.backend.get_value / set_value call) and `model.save_weights(path, overwrite=True)`.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


class EarlyStoppingLike:
    def __init__(self, monitor="val_auc", mode="max", patience=1, restore_best_weights=True):
        self.monitor, self.sign, self.patience, self.restore = monitor, (1 if mode == "max" else -1), patience, restore_best_weights
        self.best, self.wait, self.best_weights, self.stopped_epoch, self.best_epoch = None, 0, None, None, None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs):
        cur = logs[self.monitor] * self.sign
        if self.best is None or cur > self.best:
            self.best, self.wait, self.best_epoch = cur, 0, epoch
            if self.restore:
                self.best_weights = self.model.get_weights()
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
                if self.restore and self.best_weights is not None:
                    self.model.set_weights(self.best_weights)


class ReduceLROnPlateauLike:
    def __init__(self, monitor="val_loss", factor=0.2, patience=0, min_delta=1e9):
        self.monitor, self.factor, self.patience, self.min_delta, self.best, self.wait = monitor, factor, patience, min_delta, None, 0

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs):
        cur = logs[self.monitor]
        if self.best is None or cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
            return
        self.wait += 1
        if self.wait > self.patience:
            old = float(self.model.optimizer.lr.numpy())          # backend.get_value(optimizer.lr)
            self.model.optimizer.lr.assign(old * self.factor)     # backend.set_value(optimizer.lr, new_lr)
            self.wait = 0


class ModelCheckpointLike:
    def __init__(self, filepath):
        self.filepath, self.saved = filepath, []

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs):
        self.model.save_weights(self.filepath, overwrite=True)
        self.saved.append(epoch)


def synthetic_frames(rng, n_imp, n_art=150, H=6, C=5, T=10, V=300):
    """dict-of-columns behaviours + article_id -> token list; the clicked article shares a 'topic' with the history."""
    topics = rng.integers(0, 6, n_art)
    articles = {1000 + i: np.where(rng.random(T) < 0.6, topics[i] * 40 + rng.integers(0, 40, T), rng.integers(0, V, T)).tolist()
                for i in range(n_art)}
    ids = np.array(list(articles))
    by_topic = [ids[topics == t] for t in range(6)]
    beh = {"user_id": [], "hist": [], "article_ids_inview": [], "labels": []}
    for u in range(n_imp):
        t = rng.integers(0, 6)
        beh["user_id"].append(u)
        beh["hist"].append(rng.choice(by_topic[t], H).tolist())
        neg = rng.choice(ids[topics != t], C - 1, replace=False).tolist()
        pos = int(rng.choice(by_topic[t]))
        k = int(rng.integers(0, C))
        beh["article_ids_inview"].append(neg[:k] + [pos] + neg[k:])
        beh["labels"].append([1 if j == k else 0 for j in range(C)])
    return beh, articles


def make_model(seed=3, dropout=0.0, lr=2e-3):
    from ebrec.models.newsrec.model_config import hparams_nrms
    from ebrec.models.newsrec.nrms import NRMSModel

    class hp(hparams_nrms):
        history_size, title_size, head_num, head_dim, attention_hidden_dim = 6, 10, 4, 8, 24

    hp.dropout, hp.learning_rate = dropout, lr
    rng = np.random.default_rng(11)
    table = (rng.standard_normal((300, 32)) * 0.3).astype(np.float32)
    return NRMSModel(hp, word2vec_embedding=table, seed=seed)


def loaders(rng, n_train=192, n_val=96, cls=None, bs=32):
    from ebrec.models.newsrec.dataloader import NRMSDataLoader

    cls = cls or NRMSDataLoader
    beh, articles = synthetic_frames(rng, n_train + n_val)
    tr = {k: v[:n_train] for k, v in beh.items()}
    va = {k: v[n_train:] for k, v in beh.items()}
    kw = dict(article_dict=articles, history_column="hist", unknown_representation="zeros", batch_size=bs)
    return cls(behaviors=tr, **kw), cls(behaviors=va, **kw)


def test_fit_with_validation_and_callbacks(tmp_path):
    rng = np.random.default_rng(0)
    train, val = loaders(rng)
    m = make_model()
    m.model.compile(optimizer=m.model.optimizer, loss=m.model.loss, metrics=["AUC"])      # ebnerd_nrms.py:244-248
    ckpt = ModelCheckpointLike(str(tmp_path / "w" / "weights.ckpt"))
    rlr = ReduceLROnPlateauLike(monitor="val_loss", factor=0.2, patience=0, min_delta=1e9)     # never "improves": reduce every epoch after the first
    hist = m.model.fit(train, validation_data=val, epochs=3, callbacks=[ckpt, rlr], verbose=0)
    h = hist.history
    assert set(h) >= {"loss", "auc", "val_loss", "val_auc", "lr"} and all(len(v) == 3 for v in h.values())
    assert h["loss"][-1] < h["loss"][0] and h["val_auc"][-1] > 0.6                  # it learns the topic task
    assert ckpt.saved == [0, 1, 2] and os.path.exists(ckpt.filepath)
    assert np.allclose(h["lr"], [2e-3, 2e-3, 4e-4]) and abs(float(m.model.optimizer.lr.numpy()) - 8e-5) < 1e-9
    # validation_data as an ((his, pred), y) tuple of arrays gives the same val metrics as the loader
    (his, pred), y = val[0]
    for i in range(1, len(val)):
        (h2, p2), y2 = val[i]
        his, pred, y = np.concatenate([his, h2]), np.concatenate([pred, p2]), np.concatenate([y, y2])
    a = m.model.evaluate(val, return_dict=True)
    b = m.model.evaluate(((his, pred), y), batch_size=32, return_dict=True)
    assert abs(a["loss"] - b["loss"]) < 1e-6 and abs(a["auc"] - b["auc"]) < 1e-6


def test_early_stopping_restores_best_weights():
    """EarlyStopping(restore_best_weights=True): monitor the TRAINING loss in 'max' mode, so epoch 0 is the 'best'
    epoch by construction, epoch 1 counts as no improvement, patience 1 stops there and the epoch-0 weights return."""
    rng = np.random.default_rng(1)
    train, val = loaders(rng)
    m = make_model()
    m.model.compile(metrics=["AUC"])
    es = EarlyStoppingLike(monitor="loss", mode="max", patience=1, restore_best_weights=True)
    hist = m.model.fit(train, validation_data=val, epochs=6, callbacks=[es], verbose=0)
    h = hist.history
    assert h["loss"][1] < h["loss"][0]                          # premise of the construction
    assert es.stopped_epoch == 1 and len(h["loss"]) == 2 and m.model.stop_training
    again = m.model.evaluate(val, return_dict=True)
    assert abs(again["loss"] - h["val_loss"][0]) < 1e-6 and abs(again["auc"] - h["val_auc"][0]) < 1e-6
    for w, b in zip(m.model.get_weights(), es.best_weights):
        assert np.array_equal(w, b)
    assert m._engine.step_count == 2 * len(train)               # optimizer iterations are not rolled back (as in Keras)


@pytest.mark.parametrize("kind", ["nrms", "docvec"])
def test_save_load_round_trip_with_adam_state(tmp_path, kind):
    rng = np.random.default_rng(2)
    if kind == "nrms":
        train, val = loaders(rng)
        build = lambda: make_model(dropout=0.0)
        batch = train[0]
    else:
        from ebrec.models.newsrec.model_config import hparams_nrms_docvec
        from ebrec.models.newsrec.nrms_docvec import NRMSDocVec

        class hp(hparams_nrms_docvec):
            history_size, title_size, head_num, head_dim, attention_hidden_dim = 5, 48, 4, 8, 24
            newsencoder_units_per_layer, dropout, learning_rate = [32, 32], 0.0, 2e-3

        build = lambda: NRMSDocVec(hp, seed=4)
        his, pred = rng.standard_normal((64, 5, 48)).astype(np.float32), rng.standard_normal((64, 4, 48)).astype(np.float32)
        y = np.eye(4, dtype=np.float32)[rng.integers(0, 4, 64)]
        batch = ((his, pred), y)
        train = batch
    a = build()
    if kind == "nrms":
        a.model.fit(train, epochs=2, verbose=0)
    else:
        a.model.fit(train[0], train[1], batch_size=16, epochs=2, verbose=0)
    a.model.optimizer.lr.assign(7e-4)
    path = str(tmp_path / "ck" / "model.weights")
    a.model.save_weights(path)
    b = build()
    b.model.load_weights(path)
    ea, eb = a._engine, b._engine
    assert eb.step_count == ea.step_count > 0 and abs(eb.lr - 7e-4) < 1e-12
    assert torch.equal(ea.params.m, eb.params.m) and torch.equal(ea.params.v, eb.params.v)
    for wa, wb in zip(a.model.get_weights(), b.model.get_weights()):
        assert np.array_equal(wa, wb)
    # continuing from the checkpoint is the same trajectory as never having stopped
    (xa, ya) = batch
    la, lb = a.model.train_on_batch(xa, ya), b.model.train_on_batch(xa, ya)
    assert abs(la - lb) < 1e-6 * max(1.0, abs(la))
    for wa, wb in zip(a.model.get_weights(), b.model.get_weights()):
        assert np.abs(wa - wb).max() < 1e-6
    # an optimizer state that does not fit raises instead of being dropped
    if kind == "nrms":
        from ebrec.models.newsrec.model_config import hparams_nrms
        from ebrec.models.newsrec.nrms import NRMSModel

        class hp2(hparams_nrms):
            history_size, title_size, head_num, head_dim, attention_hidden_dim = 6, 10, 4, 8, 24

        other = NRMSModel(hp2, word2vec_embedding=np.zeros((500, 32), np.float32), seed=1)
        with pytest.raises((ValueError, RuntimeError)):
            other.model.load_weights(path)


def test_docvec_device_feed_matches_host_feed():
    """NRMSDocVecDataLoaderDevice: the float doc-vector matrix lives in HBM, batches carry row indices."""
    from ebrec.models.newsrec.dataloader import NRMSDataLoader, NRMSDocVecDataLoaderDevice
    from ebrec.models.newsrec.model_config import hparams_nrms_docvec
    from ebrec.models.newsrec.nrms_docvec import NRMSDocVec

    class hp(hparams_nrms_docvec):
        history_size, title_size, head_num, head_dim, attention_hidden_dim = 6, 48, 4, 8, 24
        newsencoder_units_per_layer, dropout, learning_rate = [32, 32], 0.0, 1e-3

    rng = np.random.default_rng(5)
    beh, articles = synthetic_frames(rng, 96)
    docs = {k: rng.standard_normal(48).astype(np.float32).tolist() for k in articles}
    kw = dict(behaviors=beh, article_dict=docs, history_column="hist", unknown_representation="zeros", batch_size=32)
    runs = []
    for cls in (NRMSDataLoader, NRMSDocVecDataLoaderDevice):
        m = NRMSDocVec(hp, seed=2)
        h = m.model.fit(cls(**kw), epochs=2, verbose=0, shuffle=False)
        runs.append((h.history["loss"], m.model.predict(cls(**kw)), m.scorer.predict(cls(**dict(kw, eval_mode=True)))))
    (l0, p0, s0), (l1, p1, s1) = runs
    assert np.allclose(l0, l1, rtol=0, atol=2e-5)
    assert np.abs(p0 - p1).max() < 1e-4 and p0.shape == (96, 5)
    assert np.abs(s0 - s1).max() < 1e-4 and s0.shape == (96 * 5, 1)


def test_naml_device_feed_matches_host_feed():
    from ebrec.models.newsrec.dataloader import NAMLDataLoader, NAMLDataLoaderDevice
    from ebrec.models.newsrec.model_config import hparams_naml
    from ebrec.models.newsrec.naml import NAMLModel

    class hp(hparams_naml):
        history_size, title_size, body_size, filter_num, attention_hidden_dim = 6, 10, 14, 32, 20
        vert_num, subvert_num, dropout, learning_rate = 9, 9, 0.0, 1e-3

    rng = np.random.default_rng(6)
    beh, articles = synthetic_frames(rng, 64)
    body = {k: rng.integers(1, 300, 14).tolist() for k in articles}
    cat = {k: (k % 7) + 1 for k in articles}
    table = (rng.standard_normal((300, 24)) * 0.3).astype(np.float32)
    kw = dict(behaviors=beh, article_dict=articles, body_mapping=body, category_mapping=cat, subcategory_mapping=cat,
              history_column="hist", unknown_representation="zeros", batch_size=32)
    runs = []
    for cls in (NAMLDataLoader, NAMLDataLoaderDevice):
        m = NAMLModel(hp, word2vec_embedding=table.copy(), seed=2)
        m._engine.eps = 1e-3   # keeps atomics-order noise of the scatter un-amplified (see test_gpu_nrms.py)
        h = m.model.fit(cls(**kw), epochs=2, verbose=0, shuffle=False)
        runs.append((h.history["loss"], m.model.predict(cls(**kw))))
    (l0, p0), (l1, p1) = runs
    assert np.allclose(l0, l1, rtol=0, atol=2e-5)
    assert np.abs(p0 - p1).max() < 1e-4 and p0.shape == (64, 5)


def test_dense_stack_newsencoder_predict():
    """newsencoder.predict of the Dense/BN variant (nrms.py:142-152) equals the news vectors the model scores with."""
    from ebrec.models.newsrec.model_config import hparams_nrms
    from ebrec.models.newsrec.nrms import NRMSModel

    class hp(hparams_nrms):
        history_size, title_size, head_num, head_dim, attention_hidden_dim = 4, 8, 4, 8, 16
        newsencoder_units_per_layer, dropout = [48, 32], 0.0

    rng = np.random.default_rng(7)
    m = NRMSModel(hp, word2vec_embedding=rng.standard_normal((100, 16)).astype(np.float32), seed=1)
    his, pred = rng.integers(0, 100, (3, 4, 8)), rng.integers(0, 100, (3, 2, 8))
    nv = m.newsencoder.predict(pred.reshape(-1, 8))
    uv = m.userencoder.predict(his)
    z = np.einsum("bcd,bd->bc", nv.reshape(3, 2, -1), uv)
    p = np.exp(z - z.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    assert np.abs(p - m.model.predict((his, pred))).max() < 1e-5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="data-parallel fit needs two GPUs (gpurun --gpus 2)")
def test_data_parallel_fit_shards_the_loader(tmp_path):
    """world 2: fit() gives every rank its share of each epoch's batches, logs identical global metrics on both
    ranks, checkpoints from rank 0 only, and lands on the single-GPU trajectory of the merged batches."""
    out = tmp_path / "dp"
    out.mkdir()
    env = dict(os.environ, EBK_DP_OUT=str(out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", str(ROOT / "tools" / "dp_fit_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DP_FIT_CHECK OK" in r.stdout
