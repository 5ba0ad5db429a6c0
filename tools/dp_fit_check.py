"""Data-parallel `fit` on real GPUs (run under torchrun, world 2+; tests/test_gpu_fit_surface.py drives it):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_fit_check.py

Every rank calls model.fit on the SAME loader; fit hands rank r the batches order[r::world] of every epoch
(_keraslike.shard_for_rank), logs global metrics, and ModelCheckpoint-style save_weights is collective with rank 0
writing.  Checked: (1) both ranks saw disjoint, complete shares; (2) logs identical on all ranks; (3) weights
identical on all ranks after the run; (4) rank 0 re-runs the same epochs on ONE GPU with merged batches (global batch
= world x loader batch, dropout 0) and lands on the same weights up to summation order; (5) the checkpoint written
during the DP run loads into a single-GPU model with the full (gathered) Adam state.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200"), str(ROOT / "tests")]
from ebrec.models.newsrec.dataloader import NRMSDataLoader  # noqa: E402
from ebrec.models.newsrec.model_config import hparams_nrms  # noqa: E402
from ebrec.models.newsrec.nrms import NRMSModel  # noqa: E402
from test_gpu_fit_surface import ModelCheckpointLike, synthetic_frames  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
out = Path(os.environ.get("EBK_DP_OUT", "/tmp"))


class hp(hparams_nrms):
    history_size, title_size, head_num, head_dim, attention_hidden_dim, dropout, learning_rate = 6, 10, 4, 8, 24, 0.0, 2e-3


rng = np.random.default_rng(0)
# 4 global steps per epoch + ONE more full batch that cannot be shared out and is dropped.  (All batches are full: the
# per-rank loss scale 1/(B*world) makes the summed gradient the global-batch mean only when every rank's B is the same;
# a ragged batch inside a global step would be weighted like a full one.)
beh, articles = synthetic_frames(rng, 32 * (4 * world + 1) + 40)
n_train = 32 * (4 * world + 1)
tr = {k: v[:n_train] for k, v in beh.items()}
va = {k: v[n_train:] for k, v in beh.items()}
kw = dict(article_dict=articles, history_column="hist", unknown_representation="zeros", batch_size=32)
table = (np.random.default_rng(11).standard_normal((300, 32)) * 0.3).astype(np.float32)

seen = []


class Spy(NRMSDataLoader):
    def __getitem__(self, idx):
        seen.append(int(idx))
        return super().__getitem__(idx)


m = NRMSModel(hp, word2vec_embedding=table.copy(), seed=3)
m._engine.eps = 1e-3        # keeps atomics-order noise un-amplified (see tests/test_gpu_nrms.py)
m.model.compile(metrics=["AUC"])
ck = ModelCheckpointLike(str(out / "dp.weights"))
train = Spy(behaviors=tr, **kw)
hist = m.model.fit(train, validation_data=NRMSDataLoader(behaviors=va, **kw), epochs=2, callbacks=[ck], verbose=0)
ok = True
# (1) disjoint + complete shares: 2 epochs x (len // world) batches per rank
all_seen = [None] * world
dist.all_gather_object(all_seen, seen)
per_epoch = (len(train) // world)
checks = {}
checks["shares_len"] = all(len(s) == 2 * per_epoch for s in all_seen)
ok &= checks["shares_len"]
for e in range(2):
    got = sorted(i for s in all_seen for i in s[e * per_epoch:(e + 1) * per_epoch])
    checks[f"disjoint_epoch{e}"] = len(set(got)) == len(got) == per_epoch * world
    ok &= checks[f"disjoint_epoch{e}"]
# (2) identical logs
logs = [None] * world
dist.all_gather_object(logs, hist.history)
checks["logs_identical"] = all(l == logs[0] for l in logs)
ok &= checks["logs_identical"]
# (3) identical weights
w = m.model.get_weights()
flat = torch.from_numpy(np.concatenate([a.ravel() for a in w])).cuda()
others = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(others, flat)
checks["weights_identical"] = all(bool(torch.equal(others[0], o)) for o in others)
ok &= checks["weights_identical"]
msg = ""
if rank == 0:
    # (4) single-GPU run over the merged batches
    solo = NRMSModel(hp, word2vec_embedding=table.copy(), seed=3)
    e = solo._engine
    e.world, e.rank, e.eps, e.sparse_table_grad = 1, 0, 1e-3, False
    plain = NRMSDataLoader(behaviors=tr, **kw)
    order_rng = np.random.default_rng(e.seed + 7919)            # fit's shuffle stream
    for _ in range(2):
        order = np.arange(len(plain))
        order_rng.shuffle(order)
        order = order[:(len(order) // world) * world]
        for g in range(0, len(order), world):
            parts = [plain[int(i)] for i in order[g:g + world]]
            his = np.concatenate([p[0][0] for p in parts])
            pred = np.concatenate([p[0][1] for p in parts])
            y = np.concatenate([p[1] for p in parts])
            solo.model.train_on_batch((his, pred), y)
    dev = max(float(np.abs(a - b).mean()) for a, b in zip(w, solo.model.get_weights()))
    checks["solo_match"] = dev < 2e-6
    ok &= checks["solo_match"]
    # (5) the DP checkpoint restores a complete single-GPU model
    back = NRMSModel(hp, word2vec_embedding=table.copy(), seed=3)
    back._engine.world, back._engine.rank = 1, 0
    back.model.load_weights(ck.filepath)
    checks["ckpt_weights"] = all(np.array_equal(a, b) for a, b in zip(w, back.model.get_weights()))
    checks["ckpt_adam"] = back._engine.step_count == 2 * per_epoch and float(back._engine.params.m.abs().sum()) > 0
    nz = float((back._engine.params.v[: 300 * 32] != 0).float().mean())     # gathered from BOTH shards
    checks["ckpt_adam_coverage"] = nz > 0.5
    ok &= checks["ckpt_weights"] and checks["ckpt_adam"] and checks["ckpt_adam_coverage"]
    msg = f"checks {checks} mean|dp - solo| {dev:.2e}, adam-v coverage {nz:.2f}, loss {hist.history['loss']}, val_auc {hist.history['val_auc']}"
print(f"rank {rank}: ok={ok} {checks if rank else msg}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DP_FIT_CHECK {'OK' if int(flag) else 'FAIL'} world={world} {msg}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
