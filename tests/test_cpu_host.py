"""CPU tests of the host side: reference helper known-answers, the dataloader contract (ported from the
reference's test/dataloader/test_newsrec.py), hparams surface, C-ABI symbol export, facade error behaviour,
and the data-parallel gradient identity over gloo (world_size 2)."""
import ctypes
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLD = Path(__file__).parent / "golden"


# ---- reference docstring known answers -------------------------------------------------------------------
def test_repeat_by_list_values_from_matrix_docstring_example():
    from ebrec.models.newsrec.dataloader import repeat_by_list_values_from_matrix

    out = repeat_by_list_values_from_matrix(np.array([[1, 0], [0, 0]]), np.array([[7, 8, 9], [10, 11, 12]]), np.array([1, 2]))
    want = np.array([[[10, 11, 12], [7, 8, 9]], [[7, 8, 9], [7, 8, 9]], [[7, 8, 9], [7, 8, 9]]])  # _python.py:376-386
    assert np.array_equal(out, want)


def test_create_lookup_objects_docstring_example():
    from ebrec.models.newsrec.dataloader import create_lookup_objects

    data = {10: np.array([0.1, 0.2, 0.3]), 20: np.array([0.4, 0.5, 0.6]), 30: np.array([0.7, 0.8, 0.9])}
    idx, mat = create_lookup_objects(data, "zeros")  # _python.py:440-465
    assert idx == {10: 1, 20: 2, 30: 3}
    np.testing.assert_allclose(mat, [[0, 0, 0], [0.1, 0.2, 0.3], [0.4, 0.5, 0.6], [0.7, 0.8, 0.9]])
    _, mat_mean = create_lookup_objects(data, "mean")
    np.testing.assert_allclose(mat_mean[0], [0.4, 0.5, 0.6])
    with pytest.raises(ValueError):
        create_lookup_objects(data, "median")


# ---- dataloader contract (test/dataloader/test_newsrec.py:66-105) ---------------------------------------------
@pytest.fixture(scope="module")
def sample():
    d = json.loads((GOLD / "ebnerd_sample.json").read_text())
    beh = d["behaviors"]
    mapping = {int(k): v for k, v in d["article_tokens"].items()}
    return beh, mapping


def test_nrms_dataloader_contract(sample):
    from ebrec.models.newsrec.dataloader import NRMSDataLoader, NRMSDataLoaderPretransform

    beh, mapping = sample
    BATCH = 100
    nmin = min(len(r) for r in beh["article_ids_inview"])
    keep = [i for i, r in enumerate(beh["article_ids_inview"]) if len(r) == nmin]
    train = {k: [v[i] for i in keep] for k, v in beh.items()}
    for cls in (NRMSDataLoader, NRMSDataLoaderPretransform):
        dl = cls(behaviors=train, article_dict=mapping, history_column="article_id_fixed",
                 unknown_representation="zeros", eval_mode=False, batch_size=BATCH)
        batch = next(iter(dl))
        assert len(dl) == int(np.ceil(len(keep) / BATCH))
        assert len(batch) == 2 and len(batch[0]) == 2
        his, pred = batch[0]
        assert isinstance(his.ravel()[0], np.integer) and isinstance(batch[1].ravel()[0], np.integer)
        n0 = min(BATCH, len(keep))
        assert his.shape == (n0, 3, 10) and pred.shape == (n0, nmin, 10) and batch[1].shape == (n0, nmin)
    test = NRMSDataLoader(behaviors=beh, article_dict=mapping, history_column="article_id_fixed",
                          unknown_representation="zeros", eval_mode=True, batch_size=BATCH)
    (his, pred), y = next(iter(test))
    n = sum(len(r) for r in beh["article_ids_inview"][:BATCH])
    assert len(y) == n and y.shape == (n, 1) and his.shape == (n, 3, 10) and pred.shape == (n, 1, 10)
    # history rows are repeated once per candidate of the impression
    n_first = len(beh["article_ids_inview"][0])
    assert np.array_equal(his[0], his[n_first - 1])
    # last batch may be short
    last = test[len(test) - 1]
    assert len(last[1]) == sum(len(r) for r in beh["article_ids_inview"][(len(test) - 1) * BATCH:])


def test_device_feed_loader_indices_reproduce_token_batches(sample):
    """NRMSDataLoaderDevice: lookup_article_matrix[batch indices] == NRMSDataLoader's batch,
    train and eval mode, so gather indices stay bit-exact when the gather moves to the device."""
    from ebrec.models.newsrec.dataloader import NRMSDataLoader, NRMSDataLoaderDevice

    beh, mapping = sample
    nmin = min(len(r) for r in beh["article_ids_inview"])
    keep = [i for i, r in enumerate(beh["article_ids_inview"]) if len(r) == nmin]
    train = {k: [v[i] for i in keep] for k, v in beh.items()}
    kw = dict(article_dict=mapping, history_column="article_id_fixed", unknown_representation="zeros", batch_size=64)
    for frame, eval_mode in ((train, False), (beh, True)):
        host = NRMSDataLoader(behaviors=frame, eval_mode=eval_mode, **kw)
        dev = NRMSDataLoaderDevice(behaviors=frame, eval_mode=eval_mode, **kw)
        assert dev.device_feed and len(dev) == len(host)
        for i in (0, len(host) - 1):
            (his, pred), y = host[i]
            (hi, pi), yi = dev[i]
            assert hi.dtype == np.int32 and hi.ndim == 2 and pi.ndim == 2
            assert np.array_equal(dev.lookup_article_matrix[hi], his)
            assert np.array_equal(dev.lookup_article_matrix[pi], pred)
            assert np.array_equal(y, yi)


def test_unknown_article_ids_map_to_row_zero(sample):
    from ebrec.models.newsrec.dataloader import NRMSDataLoader

    beh, mapping = sample
    two = {k: [v[0], v[1]] for k, v in beh.items()}
    two["article_id_fixed"] = [[-1, -2, -3], two["article_id_fixed"][1]]
    n = min(len(r) for r in two["article_ids_inview"])
    two["article_ids_inview"] = [r[:n] for r in two["article_ids_inview"]]
    two["labels"] = [r[:n] for r in two["labels"]]
    (his, _), _ = NRMSDataLoader(behaviors=two, article_dict=mapping, history_column="article_id_fixed",
                                 unknown_representation="zeros", batch_size=2)[0]
    assert not his[0].any() and his[1].any()


def test_naml_dataloader_contract(sample):
    from ebrec.models.newsrec.dataloader import NAMLDataLoader

    beh, mapping = sample
    nmin = min(len(r) for r in beh["article_ids_inview"])
    keep = [i for i, r in enumerate(beh["article_ids_inview"]) if len(r) == nmin][:40]
    train = {k: [v[i] for i in keep] for k, v in beh.items()}
    cat = {a: (a % 7) + 1 for a in mapping}
    kw = dict(behaviors=train, article_dict=mapping, body_mapping=mapping, category_mapping=cat, subcategory_mapping=cat,
              history_column="article_id_fixed", unknown_representation="zeros", batch_size=16)
    dl = NAMLDataLoader(**kw)
    x, y = dl[0]
    assert len(x) == 8  # test_newsrec.py:175-190
    assert x[0].shape == (16, 3, 10) and x[2].shape == (16, 3, 1) and x[4].shape == (16, nmin, 10) and x[7].shape == (16, nmin, 1)
    with pytest.raises(ValueError):
        NAMLDataLoader(**{**kw, "eval_mode": True})  # dataloader.py:289-290


# ---- hparams surface ----------------------------------------------------------------------------------------
def test_hparams_defaults_match_reference():
    from ebrec.models.newsrec.model_config import hparams_naml, hparams_nrms, hparams_nrms_docvec, hparams_to_dict

    d = hparams_to_dict(hparams_nrms)  # model_config.py:82-97
    assert d == {"title_size": 30, "history_size": 20, "head_num": 20, "head_dim": 20, "attention_hidden_dim": 200,
                 "optimizer": "adam", "loss": "cross_entropy_loss", "dropout": 0.2, "learning_rate": 1e-4,
                 "newsencoder_units_per_layer": None, "newsencoder_l2_regularization": 1e-4}
    assert hparams_nrms_docvec.title_size == 768 and hparams_nrms_docvec.newsencoder_units_per_layer == [512, 512, 512]
    assert hparams_naml.filter_num == 400 and hparams_naml.window_size == 3 and hparams_naml.body_size == 40


# ---- C-ABI library ------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "ebk.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ebk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib_path = ROOT / "ebnerd-benchmark_b200" / "csrc" / "libebk.so"
    if not lib_path.exists():
        subprocess.run(["make", "-C", str(lib_path.parent), "-j8"], check=True, capture_output=True)
    lib = ctypes.CDLL(str(lib_path))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from ebrec.models.newsrec import _ebk

    assert set(_ebk.SYMBOLS) <= declared
    assert lib.ebk_version() >= 100  # pure host call; no compute without a GPU


def test_c_abi_header_is_plain_c():
    """include/ebk.h is the drop-in boundary: it must compile as C99 (no C++ / torch types in the signatures)."""
    import shutil

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", str(ROOT / "include" / "ebk.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The ctypes.Structure mirrors in _ebk.py against the C structs of include/ebk.h: same size, same field offsets
    (a C program prints sizeof / offsetof; a field added on one side only shows up here, on CPU)."""
    import ctypes as C
    import shutil

    from ebrec.models.newsrec import _ebk

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    pairs = {"ebk_seqenc_desc": _ebk.SeqEncDesc, "ebk_seqenc_opts": _ebk.SeqEncOpts, "ebk_dense_desc": _ebk.DenseDesc,
             "ebk_attlayer_desc": _ebk.AttLayerDesc, "ebk_conv1d_desc": _ebk.Conv1dDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "ebk.h"}"', "int main(void) {"]
    for cname, mirror in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf(" {fname}=%zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines.append('  printf("ebk_step_params %zu seed1=%zu seed2=%zu alpha=%zu\\n", sizeof(ebk_step_params), '
                 'offsetof(ebk_step_params, seed1), offsetof(ebk_step_params, seed2), offsetof(ebk_step_params, alpha));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run([gcc, "-std=c99", "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    seen = {}
    for line in out:
        name, size, *fields = line.split()
        seen[name] = (int(size), {f.split("=")[0]: int(f.split("=")[1]) for f in fields})
    for cname, mirror in pairs.items():
        size, offs = seen[cname]
        assert C.sizeof(mirror) == size, (cname, C.sizeof(mirror), size)
        for fname, _ in mirror._fields_:
            assert getattr(mirror, fname).offset == offs[fname], (cname, fname)
    # the engine writes ebk_step_params as three 8-byte words: seed1 | seed2 | alpha (float) + padding
    assert seen["ebk_step_params"] == (24, {"seed1": 0, "seed2": 8, "alpha": 16})


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ebrec.models.newsrec import _ebk
    from ebrec.models.newsrec.model_config import hparams_nrms
    from ebrec.models.newsrec.nrms import NRMSModel

    with pytest.raises(_ebk.EbkError):
        NRMSModel(hparams_nrms, word2vec_embedding=np.random.rand(50, 8))


def test_unique_rows_helper_of_the_dedup_scorer():
    """Host half of the deduplicating predict path: distinct[inverse] reproduces the rows."""
    from ebrec.models.newsrec._engine import unique_rows

    rng = np.random.default_rng(0)
    pool = rng.integers(0, 250002, (50, 30)).astype(np.int32)
    rows = pool[rng.integers(0, 50, 700)]
    uniq, inv = unique_rows(rows)
    assert uniq.shape[0] == len({r.tobytes() for r in rows}) <= 50 and inv.shape == (700,)
    assert np.array_equal(uniq[inv], rows)
    u2, i2 = unique_rows(np.array([[3, 1], [3, 1], [0, 2]], dtype=np.int64))
    assert u2.shape == (2, 2) and np.array_equal(u2[i2], [[3, 1], [3, 1], [0, 2]])
    u3, i3 = unique_rows(np.zeros((0, 5), np.int32))
    assert u3.shape[0] == 0 and i3.shape == (0,)


def test_facade_rejects_unknown_loss_and_optimizer():
    from ebrec.models.newsrec.model_config import hparams_nrms
    from ebrec.models.newsrec.nrms import NRMSModel

    class bad_loss(hparams_nrms):
        loss = "hinge"

    class bad_opt(hparams_nrms):
        optimizer = "sgd"

    with pytest.raises(ValueError, match="this loss not defined"):  # nrms.py:66
        NRMSModel(bad_loss, word2vec_embedding=np.random.rand(50, 8))
    with pytest.raises(ValueError, match="this optimizer not defined"):  # nrms.py:79
        NRMSModel(bad_opt, word2vec_embedding=np.random.rand(50, 8))


def test_keras_auc_matches_trapezoid_definition():
    from ebrec.models.newsrec._keraslike import keras_auc

    rng = np.random.default_rng(0)
    y = rng.integers(0, 2, 5000)
    p = np.clip(0.35 * y + rng.random(5000) * 0.65, 0, 1)
    from sklearn.metrics import roc_auc_score

    assert abs(keras_auc(y, p) - roc_auc_score(y, p)) < 5e-3  # 200-threshold approximation
    assert abs(keras_auc(y, np.full(5000, 0.5)) - 0.5) < 1e-9


# ---- data parallel: sum over ranks of (1/world)-scaled shard gradients == global-batch mean gradient -------------
_DP_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["EBK_ROOT"])
from oracle import nrms_oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(0)
V, E, nh, dh, att, B, H, C, T = 40, 8, 2, 4, 6, 4, 3, 3, 5
P = O.init_nrms_params(rng, V, E, nh, dh, att, dtype=np.float64)
his = rng.integers(0, V, (B, H, T)); pred = rng.integers(0, V, (B, C, T))
y = np.zeros((B, C)); y[np.arange(B), rng.integers(0, C, B)] = 1
sl = slice(rank * B // world, (rank + 1) * B // world)          # contiguous shard of the global batch
_, _, G = O.nrms_loss_and_grads(his[sl], pred[sl], y[sl], P, nh, dh, training=False, loss_scale=1.0 / world)
flat = torch.from_numpy(np.concatenate([G[k].ravel() for k in O.NRMS_PARAM_ORDER]))
dist.all_reduce(flat)                                            # the one collective of a step
_, _, Gfull = O.nrms_loss_and_grads(his, pred, y, P, nh, dh, training=False)
want = np.concatenate([Gfull[k].ravel() for k in O.NRMS_PARAM_ORDER])
err = float(np.abs(flat.numpy() - want).max())
assert err < 1e-12, err
if rank == 0: print("DP_OK", err)
dist.destroy_process_group()
'''


def test_data_parallel_gradient_identity_gloo_world2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    env = dict(os.environ, EBK_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


_DP_ADAM_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["EBK_ROOT"])
from oracle import nrms_oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(1)
V, E, nh, dh, att, B, H, C, T = 32, 8, 2, 4, 6, 4, 3, 3, 5
P = O.init_nrms_params(rng, V, E, nh, dh, att, dtype=np.float64)
names = O.NRMS_PARAM_ORDER
sizes = [P[k].size for k in names]
n = sum(sizes); pad = (-n) % world; shard = (n + pad) // world          # flat buffer divisible into rank shards
flat = lambda d: np.concatenate([d[k].ravel() for k in names] + [np.zeros(pad)])
theta = flat(P); m = np.zeros_like(theta); v = np.zeros_like(theta)      # replicated state of the DP run
Pd = {k: P[k].copy() for k in names}; md = {k: np.zeros_like(P[k]) for k in names}; vd = {k: np.zeros_like(P[k]) for k in names}
lr = 1e-2
for t in range(1, 4):
    his = rng.integers(0, V, (B, H, T)); pred = rng.integers(0, V, (B, C, T))
    y = np.zeros((B, C)); y[np.arange(B), rng.integers(0, C, B)] = 1
    # ---- data parallel: shard of the batch, loss scaled by 1/world, reduce-scatter, Adam on the shard, all-gather
    Pcur = {}; o = 0
    for k, sz in zip(names, sizes):
        Pcur[k] = theta[o:o + sz].reshape(P[k].shape); o += sz
    sl = slice(rank * B // world, (rank + 1) * B // world)
    _, _, G = O.nrms_loss_and_grads(his[sl], pred[sl], y[sl], Pcur, nh, dh, training=False, loss_scale=1.0 / world)
    g = torch.from_numpy(flat(G)); dist.all_reduce(g)                    # gloo: reduce-scatter = all-reduce + own slice
    lo = rank * shard
    th_s, m_s, v_s = theta[lo:lo + shard].copy(), m[lo:lo + shard], v[lo:lo + shard]
    O.keras_adam_step(th_s, g.numpy()[lo:lo + shard], m_s, v_s, t, lr)
    parts = [torch.empty(shard, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(th_s))
    theta = torch.cat(parts).numpy()
    # ---- single process, whole batch, dense Adam
    _, _, Gd = O.nrms_loss_and_grads(his, pred, y, Pd, nh, dh, training=False)
    for k in names:
        O.keras_adam_step(Pd[k], Gd[k], md[k], vd[k], t, lr)
err = float(np.abs(theta[:n] - flat(Pd)[:n]).max())
assert err < 1e-12, err
if rank == 0: print("DP_ADAM_OK", err)
dist.destroy_process_group()
'''


def test_data_parallel_sharded_adam_equals_dense_gloo_world2(tmp_path):
    """The optimizer step of the data-parallel engine (reduce-scatter of the flat gradient, Keras Adam on this
    rank's 1/world slice of theta/m/v, all-gather of theta) against a single-process dense Adam on the whole batch:
    3 steps, world 2 over gloo, float64 oracle arithmetic."""
    script = tmp_path / "dp_adam_worker.py"
    script.write_text(_DP_ADAM_WORKER)
    env = dict(os.environ, EBK_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29519", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "DP_ADAM_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ---- data-parallel fit: which batches a rank trains on ---------------------------------------------------------
def test_shard_for_rank_gives_disjoint_equal_shares():
    from ebrec.models.newsrec._keraslike import _unpack_xy, shard_for_rank

    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 37):
        order = rng.permutation(n)
        assert np.array_equal(shard_for_rank(order, 0, 1), order)           # one GPU: identity
        for world in (2, 3, 8):
            shares = [shard_for_rank(order, r, world) for r in range(world)]
            assert len({len(s) for s in shares}) == 1 and len(shares[0]) == n // world   # same number of steps
            merged = np.concatenate(shares) if n >= world else np.array([], dtype=order.dtype)
            assert len(set(merged.tolist())) == len(merged)                                # disjoint
            assert set(merged.tolist()) == set(order[: (n // world) * world].tolist())     # only the uneven tail is dropped
            # global step g = the batches order[g*world : (g+1)*world], one per rank
            for g in range(n // world):
                assert [int(s[g]) for s in shares] == order[g * world:(g + 1) * world].tolist()
    # Keras-style (inputs, y) pairs are split; loaders and bare input tuples pass through
    his, pred, y = np.zeros((2, 3, 4)), np.zeros((2, 5, 4)), np.zeros((2, 5))
    x2, y2 = _unpack_xy(((his, pred), y))
    assert x2[0] is his and y2 is y
    assert _unpack_xy((his, pred)) == ((his, pred), None)
