"""Device engine of NRMSDocVec (reference src/ebrec/models/newsrec/nrms_docvec.py:8-188).

News encoder = MLP over a precomputed document vector: [Dense(u, relu, l2) + BatchNorm + Dropout] x len(units)
followed by Dense(D, relu) (nrms_docvec.py:109-130), applied separately to the history rows and to the
candidate rows of a batch, exactly as the reference's two TimeDistributed calls do (so BatchNorm batch
statistics are per call and the moving averages are updated twice per step).  User encoder, click score,
loss and optimizer are those of NRMS and reuse the same C-ABI calls (`ebk_seqenc_*`, `ebk_score_*`,
`ebk_adam_keras_step`); the MLP layers go through `ebk_dense_fwd/bwd`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _ebk
from ._engine import FlatParams, NRMSEngine, _mix

BN_MOMENTUM, BN_EPS = 0.99, 1e-3  # Keras BatchNormalization defaults


class DocVecEngine(NRMSEngine):
    def __init__(self, *, Ddoc, units, H, nh, dh, att, dropout, lr, l2, seed=None, math=_ebk.MATH_TF32, device=None,
                 beta1=0.9, beta2=0.999, eps=1e-7):
        _ebk.require_device()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.Ddoc, self.units, self.H = int(Ddoc), [int(u) for u in (units or [])], int(H)
        self.nh, self.dh, self.att = int(nh), int(dh), int(att)
        self.D = self.nh * self.dh
        self.dropout, self.l2 = float(dropout), float(l2)
        self.lr, self.beta1, self.beta2, self.eps = float(lr), beta1, beta2, eps
        self.math = int(math)
        self.math_infer = _ebk.MATH_TF32X3 if self.math == _ebk.MATH_TF32 else self.math
        self.seed = 0 if seed is None else int(seed)
        self.step_count = 0
        D, A = self.D, self.att
        spec, din = [], self.Ddoc
        for i, u in enumerate(self.units):
            spec += [(f"d{i}_W", (din, u)), (f"d{i}_b", (u,)), (f"d{i}_gamma", (u,)), (f"d{i}_beta", (u,))]
            din = u
        spec += [("out_W", (din, D)), ("out_b", (D,)),
                 ("user_Wqkv", (D, 3 * D)), ("user_attW", (D, A)), ("user_attb", (A,)), ("user_attq", (A,))]
        self.params = FlatParams(spec, self.device)
        # BatchNorm moving statistics: state, not parameters (Adam never touches them)
        self.bn_mean = [torch.zeros(u, device=self.device) for u in self.units]
        self.bn_var = [torch.ones(u, device=self.device) for u in self.units]
        self._ws, self._bufs = {}, {}
        self.world, self.rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world, self.rank = torch.distributed.get_world_size(), torch.distributed.get_rank()

    # ------------------------------------------------------------------ weights (Keras get_weights order)
    def set_weights(self, weights):
        n = len(self.units)
        if len(weights) != 6 * n + 2 + 6:
            raise ValueError(f"NRMSDocVec expects {6 * n + 8} weight arrays, got {len(weights)}")
        w = [torch.as_tensor(np.asarray(a, dtype=np.float32)) for a in weights]
        P = self.params
        with torch.no_grad():
            for i in range(n):
                W, b, g, be, mm, mv = w[6 * i: 6 * i + 6]
                P.p(f"d{i}_W").copy_(W); P.p(f"d{i}_b").copy_(b); P.p(f"d{i}_gamma").copy_(g); P.p(f"d{i}_beta").copy_(be)
                self.bn_mean[i].copy_(mm); self.bn_var[i].copy_(mv)
            o = 6 * n
            P.p("out_W").copy_(w[o]); P.p("out_b").copy_(w[o + 1])
            P.p("user_Wqkv").copy_(torch.cat([w[o + 2], w[o + 3], w[o + 4]], dim=1))
            P.p("user_attW").copy_(w[o + 5]); P.p("user_attb").copy_(w[o + 6].reshape(-1)); P.p("user_attq").copy_(w[o + 7].reshape(-1))

    def get_weights(self):
        P, D, out = self.params, self.D, []
        for i in range(len(self.units)):
            out += [P.p(f"d{i}_W").cpu().numpy(), P.p(f"d{i}_b").cpu().numpy(), P.p(f"d{i}_gamma").cpu().numpy(),
                    P.p(f"d{i}_beta").cpu().numpy(), self.bn_mean[i].cpu().numpy(), self.bn_var[i].cpu().numpy()]
        Wqkv = P.p("user_Wqkv").cpu().numpy()
        out += [P.p("out_W").cpu().numpy(), P.p("out_b").cpu().numpy(), Wqkv[:, :D].copy(), Wqkv[:, D:2 * D].copy(),
                Wqkv[:, 2 * D:].copy(), P.p("user_attW").cpu().numpy(), P.p("user_attb").cpu().numpy(),
                P.p("user_attq").cpu().numpy().reshape(-1, 1)]
        return out

    def count_params(self):
        return super().count_params() + 2 * sum(self.units)

    def trainable_params(self):
        return NRMSEngine.count_params(self)

    # ------------------------------------------------------------------ MLP news encoder
    def _layers(self):
        din = self.Ddoc
        for i, u in enumerate(self.units):
            yield i, f"d{i}", din, u, True
            din = u
        yield len(self.units), "out", din, self.D, False

    def _dense_desc(self, n_rows, K, U, bn, training):
        math = self.math if training else self.math_infer
        return _ebk.DenseDesc(n_rows, K, U, 1, 1 if bn else 0, BN_MOMENTUM, BN_EPS, self.dropout if bn else 0.0,
                              self.l2 if bn else 0.0, math)

    def _dense_ws(self, key, desc):
        need = _ebk.lib().ebk_dense_workspace_bytes(C.byref(desc))
        cur = self._ws.get(key)
        if cur is None or cur.numel() < need:
            if cur is not None:
                self._drop_graphs()
            cur = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = cur
        return cur

    def _mlp_fwd(self, call, x, out, training, seed, step_dev=None):
        """x [n, Ddoc] -> out [n, D]; keeps per-layer inputs / descs / workspaces for backward.  Layer i drops with seed
        `seed + i`; with step_dev (CUDA-graph replay) the base seed is seed1 (call 0) / seed2 (call 1) of the
        device-resident ebk_step_params."""
        lib, P = _ebk.lib(), self.params
        n = x.shape[0]
        ctx, cur = [], x
        for i, name, K, U, bn in self._layers():
            desc = self._dense_desc(n, K, U, bn, training)
            ws = self._dense_ws((call, i), desc)
            y = out if not bn else self._buf(f"y{call}_{i}", (n, U))
            head = (C.byref(desc), _ebk.ptr(cur), _ebk.ptr(P.p(f"{name}_W")), _ebk.ptr(P.p(f"{name}_b")),
                    _ebk.ptr(P.p(f"{name}_gamma")) if bn else None, _ebk.ptr(P.p(f"{name}_beta")) if bn else None,
                    _ebk.ptr(self.bn_mean[i]) if bn else None, _ebk.ptr(self.bn_var[i]) if bn else None, int(training))
            tail = (_ebk.ptr(ws), ws.numel(), _ebk.ptr(y), _ebk.stream())
            if step_dev is not None:
                _ebk.check(lib.ebk_dense_fwd_p(*head, C.c_void_p(step_dev.data_ptr()), call & 1, i, *tail))
            else:
                _ebk.check(lib.ebk_dense_fwd(*head, (seed + i) & ((1 << 64) - 1), *tail))
            ctx.append((desc, ws, cur, y, name, bn, i))
            cur = y
        return ctx

    def _mlp_bwd(self, ctx, d_out, training, seed, l2_scale, step_dev=None, call=0):
        lib, P = _ebk.lib(), self.params
        dy = d_out
        for desc, ws, x_in, y, name, bn, i in reversed(ctx):
            first = i == 0
            dx = None if first else self._buf(f"dx_{i % 2}", (desc.N, desc.K))  # ping-pong: dy of layer i-1
            head = (C.byref(desc), _ebk.ptr(x_in), _ebk.ptr(P.p(f"{name}_W")), _ebk.ptr(P.p(f"{name}_gamma")) if bn else None,
                    _ebk.ptr(y), int(training))
            tail = (_ebk.ptr(ws), ws.numel(), _ebk.ptr(dy), l2_scale, _ebk.ptr(P.g(f"{name}_W")), _ebk.ptr(P.g(f"{name}_b")),
                    _ebk.ptr(P.g(f"{name}_gamma")) if bn else None, _ebk.ptr(P.g(f"{name}_beta")) if bn else None,
                    _ebk.ptr(dx) if dx is not None else None, _ebk.stream())
            if step_dev is not None:
                _ebk.check(lib.ebk_dense_bwd_p(*head, C.c_void_p(step_dev.data_ptr()), call & 1, i, *tail))
            else:
                _ebk.check(lib.ebk_dense_bwd(*head, (seed + i) & ((1 << 64) - 1), *tail))
            dy = dx
        return None

    # ------------------------------------------------------------------ forward / backward
    def _encode_vec(self, x_all, B, training, seeds, step_dev=None):
        """x_all [B*H + B*C, Ddoc] float -> n_all [N, D], u [B, D]."""
        lib, P = _ebk.lib(), self.params
        N, BH = x_all.shape[0], B * self.H
        n_all = self._buf("n_all", (N, self.D))
        ctx_h = self._mlp_fwd(0, x_all[:BH], n_all[:BH], training, seeds[0], step_dev)
        ctx_c = self._mlp_fwd(1, x_all[BH:], n_all[BH:], training, seeds[1], step_dev)
        du = self._desc("user", B, training)
        wu = self._workspace("user", du)
        u = self._buf("u", (B, self.D))
        _ebk.check(lib.ebk_seqenc_fwd(C.byref(du), None, _ebk.ptr(n_all), _ebk.ptr(P.p("user_Wqkv")),
                                      _ebk.ptr(P.p("user_attW")), _ebk.ptr(P.p("user_attb")), _ebk.ptr(P.p("user_attq")),
                                      0, 0, 0, _ebk.ptr(wu), wu.numel(), _ebk.ptr(u), _ebk.stream()))
        return n_all, u, (ctx_h, ctx_c, du, wu)

    def forward_logits_parts(self, x_all, B, C_, training=False, seeds=(0, 0), step_dev=None):
        n_all, u, ctx = self._encode_vec(x_all, B, training, seeds, step_dev)
        return n_all, n_all[B * self.H:].view(B, C_, self.D), u, ctx

    def step_seeds(self):
        base = _mix(self.seed, self.step_count * self.world + self.rank)
        return _mix(base, 1) & ((1 << 62) - 1), _mix(base, 2) & ((1 << 62) - 1)

    def loss_and_grads_dev(self, x_all, labels, B, C_, training=True, seeds=None, step_dev=None):
        lib, P = _ebk.lib(), self.params
        seeds = self.step_seeds() if seeds is None else seeds
        n_all, news_c, u, (ctx_h, ctx_c, du, wu) = self.forward_logits_parts(x_all, B, C_, training, seeds, step_dev)
        N, BH = n_all.shape[0], B * self.H
        probs = self._buf("probs", (B, C_))
        loss = self._buf("loss", (1,))
        loss.zero_()
        dn_all = self._buf("dn_all", (N, self.D))
        d_user = self._buf("d_user", (B, self.D))
        _ebk.check(lib.ebk_score_loss(self.loss_kind, B, C_, self.D, _ebk.ptr(news_c), _ebk.ptr(u), _ebk.ptr(labels),
                                      1.0 / (B * self.world), 1.0 / B, _ebk.ptr(probs), _ebk.ptr(loss),
                                      _ebk.ptr(dn_all[BH:]), _ebk.ptr(d_user), _ebk.stream()))
        for i in range(len(self.units)):  # + l2 * sum ||W||^2  (nrms_docvec.py:122-124)
            W = P.p(f"d{i}_W")
            _ebk.check(lib.ebk_sumsq_accum(_ebk.ptr(W), W.numel(), self.l2, _ebk.ptr(loss), _ebk.stream()))
        _ebk.check(lib.ebk_seqenc_bwd(C.byref(du), None, _ebk.ptr(n_all), _ebk.ptr(P.p("user_Wqkv")),
                                      _ebk.ptr(P.p("user_attW")), _ebk.ptr(P.p("user_attb")), _ebk.ptr(P.p("user_attq")),
                                      0, 0, 0, _ebk.ptr(wu), wu.numel(), _ebk.ptr(d_user), _ebk.ptr(P.g("user_Wqkv")),
                                      _ebk.ptr(P.g("user_attW")), _ebk.ptr(P.g("user_attb")), _ebk.ptr(P.g("user_attq")),
                                      None, _ebk.ptr(dn_all), _ebk.stream()))
        # the l2 gradient 2*l2*W is added once per step (by the history call), scaled like the loss
        self._mlp_bwd(ctx_h, dn_all[:BH], training, seeds[0], 1.0 / self.world, step_dev, 0)
        self._mlp_bwd(ctx_c, dn_all[BH:], training, seeds[1], 0.0, step_dev, 1)
        return loss, probs

    # ------------------------------------------------------------------ CUDA-graph replay of the training step
    def _graph_ok(self) -> bool:
        """One GPU, unless EBK_NO_GRAPH=1 or the library profiler is recording (every launch of this graph takes its
        per-step scalars -- two dropout seeds, the Adam alpha -- from a device-resident ebk_step_params)."""
        if os.environ.get("EBK_NO_GRAPH", "0") == "1" or type(self) is not DocVecEngine or self.world != 1:
            return False
        return not _ebk.lib().ebk_prof_is_enabled()

    def train_step_dev(self, x_all, labels, B, C_):
        """One optimizer iteration on device-resident rows x_all [B*H + B*C, Ddoc].  The step is about 90 small
        launches (5 120-12 800 rows through four Dense layers, twice): replayed from a CUDA graph per batch shape."""
        if not self._graph_ok():
            loss, probs = self.loss_and_grads_dev(x_all, labels, B, C_, training=True)
            self.apply_adam()
            return loss, probs
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (int(B), int(C_), tuple(x_all.shape), self.eps, self.dropout, self.beta1, self.beta2, self.l2, int(self.loss_kind))
        st = graphs.get(key)
        if st is None:
            st = {"step": torch.zeros(3, dtype=torch.int64, device=self.device), "x": torch.empty_like(x_all),
                  "lab": torch.empty_like(labels), "graph": None, "warm": 0}
            graphs[key] = st
        if st["graph"] is None and st["warm"] < 1:     # eager warm-up with the same shapes: allocates every workspace
            st["warm"] += 1
            loss, probs = self.loss_and_grads_dev(x_all, labels, B, C_, training=True)
            self.apply_adam()
            return loss, probs
        st["x"].copy_(x_all, non_blocking=True)
        st["lab"].copy_(labels, non_blocking=True)
        self._write_step_params(st)
        if st["graph"] is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = _ebk.lib().ebk_launch_count()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                loss, probs = self.loss_and_grads_dev(st["x"], st["lab"], B, C_, training=True, seeds=(0, 0),
                                                      step_dev=st["step"])
                P, sp = self.params, C.c_void_p(st["step"].data_ptr())
                _ebk.check(_ebk.lib().ebk_adam_keras_step_p(_ebk.ptr(P.theta), _ebk.ptr(P.grad), _ebk.ptr(P.m), _ebk.ptr(P.v),
                                                            P.n, 0.0, sp, self.beta1, self.beta2, self.eps, 1, _ebk.stream()))
            st["graph"], st["loss"], st["probs"] = g, loss, probs
            st["launches"] = int(_ebk.lib().ebk_launch_count() - l0)
        st["graph"].replay()
        self.step_count += 1
        self.graph_steps = getattr(self, "graph_steps", 0) + 1
        self.graph_launches = getattr(self, "graph_launches", 0) + st["launches"]
        return st["loss"], st["probs"]

    # ------------------------------------------------------------------ device-resident doc-vector matrix
    def set_article_matrix(self, matrix: np.ndarray) -> None:
        """Upload the [n_articles + 1, Ddoc] document-vector matrix of a dataloader once (row 0 = unknown article,
        create_lookup_objects, _python.py:412-484); batches then carry article ROW INDICES.  The reference gathers
        the float vectors on the host for every batch (dataloader.py:146-180) and ships [B, H+C, Ddoc] floats
        (39 MB per step at BASELINE config 2); here a step moves B*(H+C) int32."""
        m = np.ascontiguousarray(np.asarray(matrix), dtype=np.float32)
        if m.ndim != 2 or m.shape[1] != self.Ddoc:
            raise ValueError(f"doc-vector matrix must be [n_articles, title_size={self.Ddoc}], got {m.shape}")
        self.article_matrix = torch.from_numpy(m).to(self.device)

    def vectors_from_indices(self, his_idx, pred_idx) -> torch.Tensor:
        """[B,H] + [B,C] article row indices -> [B*H + B*C, Ddoc] float rows on the device (history first)."""
        if getattr(self, "article_matrix", None) is None:
            raise ValueError("index batches need set_article_matrix(lookup_article_matrix) first")
        idx = np.concatenate([np.asarray(his_idx).reshape(-1), np.asarray(pred_idx).reshape(-1)]).astype(np.int32)
        n = self.article_matrix.shape[0]
        if idx.size and (idx.min() < 0 or idx.max() >= n):
            raise IndexError(f"article row index outside [0, {n})")
        return self.article_matrix.index_select(0, self._h2d("idx", idx))

    # ------------------------------------------------------------------ host convenience
    def to_device_batch(self, his, pred, y=None):
        if np.asarray(his).ndim == 2:   # article row indices of a device-feed loader
            xd = self.vectors_from_indices(his, pred)
            lab = None
            if y is not None:
                lab = self._h2d("lab", np.ascontiguousarray(y, dtype=np.float32))
            return xd, lab
        his, pred = np.asarray(his, dtype=np.float32), np.asarray(pred, dtype=np.float32)
        B, H, Dd = his.shape
        C_ = pred.shape[1]
        x = np.empty((B * H + B * C_, Dd), dtype=np.float32)
        x[: B * H] = his.reshape(B * H, Dd)
        x[B * H:] = pred.reshape(B * C_, Dd)
        xd = torch.from_numpy(x).to(self.device, non_blocking=True)
        lab = None
        if y is not None:
            lab = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(self.device, non_blocking=True)
        return xd, lab

    def encode_host(self, kind, x):
        x = np.asarray(x, dtype=np.float32)
        if kind == "news":
            xd = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, self.Ddoc))).to(self.device)
            out = torch.empty((xd.shape[0], self.D), device=self.device)
            self._mlp_fwd(2, xd, out, False, 0)
            return out.cpu().numpy()
        B = x.shape[0]
        xd = torch.from_numpy(np.ascontiguousarray(x.reshape(B * self.H, self.Ddoc))).to(self.device)
        _, u, _ = self._encode_vec(xd, B, False, (0, 0))
        return u.clone().cpu().numpy()
