"""Device engine of NRMS with the optional Dense/BatchNorm/Dropout stack in the news encoder
(reference src/ebrec/models/newsrec/nrms.py:142-152, ``hparams.newsencoder_units_per_layer``).

News encoder = Embedding -> Dropout -> SelfAttention (`ebk_seqenc_*` with ``att = 0``: stop after the attention,
all N = B*(H+C) articles in one call) -> [Dense(u, relu, l2) + BatchNorm + Dropout] per layer (`ebk_dense_*`,
applied separately to the history rows and to the candidate rows, exactly as the reference's two
TimeDistributed calls do, so BatchNorm statistics are per call and the moving averages are updated twice per
step) -> AttLayer2 (`ebk_attlayer_*`).  User encoder, click score, loss and optimizer are those of NRMS.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ebk
from ._engine import FlatParams, NRMSEngine, _mix

BN_MOMENTUM, BN_EPS = 0.99, 1e-3  # Keras BatchNormalization defaults


class NRMSDenseEngine(NRMSEngine):
    def __init__(self, *, V, E, T, H, nh, dh, att, units, l2, dropout, lr, seed=None, math=_ebk.MATH_TF32, device=None,
                 beta1=0.9, beta2=0.999, eps=1e-7):
        _ebk.require_device()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.V, self.E, self.T, self.H = int(V), int(E), int(T), int(H)
        self.nh, self.dh, self.att = int(nh), int(dh), int(att)
        self.D = self.nh * self.dh
        self.units = [int(u) for u in units]
        if not self.units or self.units[-1] != self.D:
            raise ValueError(f"newsencoder_units_per_layer[-1] must equal head_num*head_dim = {self.D} "
                             f"(the Dot of nrms.py:201 pairs news and user vectors), got {units}")
        self.dropout, self.l2 = float(dropout), float(l2)
        self.lr, self.beta1, self.beta2, self.eps = float(lr), beta1, beta2, eps
        self.math = int(math)
        self.math_infer = _ebk.MATH_TF32X3 if self.math == _ebk.MATH_TF32 else self.math
        self.seed = 0 if seed is None else int(seed)
        self.step_count = 0
        D, A = self.D, self.att
        spec = [("table", (self.V, self.E)), ("news_Wqkv", (self.E, 3 * D))]
        din = D
        for i, u in enumerate(self.units):
            spec += [(f"d{i}_W", (din, u)), (f"d{i}_b", (u,)), (f"d{i}_gamma", (u,)), (f"d{i}_beta", (u,))]
            din = u
        spec += [("news_attW", (din, A)), ("news_attb", (A,)), ("news_attq", (A,)),
                 ("user_Wqkv", (D, 3 * D)), ("user_attW", (D, A)), ("user_attb", (A,)), ("user_attq", (A,))]
        self.params = FlatParams(spec, self.device)
        self.bn_mean = [torch.zeros(u, device=self.device) for u in self.units]
        self.bn_var = [torch.ones(u, device=self.device) for u in self.units]
        self._ws, self._bufs = {}, {}
        self.world, self.rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world, self.rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        self.sparse_table_grad = False
        self.launches_per_step = 0

    # ------------------------------------------------------------------ weights (Keras get_weights order)
    def set_weights(self, weights):
        n = len(self.units)
        if len(weights) != 4 + 6 * n + 3 + 6:
            raise ValueError(f"NRMS (dense stack) expects {13 + 6 * n} weight arrays, got {len(weights)}")
        w = [torch.as_tensor(np.asarray(a, dtype=np.float32)) for a in weights]
        P = self.params
        with torch.no_grad():
            P.p("table").copy_(w[0])
            P.p("news_Wqkv").copy_(torch.cat([w[1], w[2], w[3]], dim=1))
            for i in range(n):
                W, b, g, be, mm, mv = w[4 + 6 * i: 10 + 6 * i]
                P.p(f"d{i}_W").copy_(W); P.p(f"d{i}_b").copy_(b); P.p(f"d{i}_gamma").copy_(g); P.p(f"d{i}_beta").copy_(be)
                self.bn_mean[i].copy_(mm); self.bn_var[i].copy_(mv)
            o = 4 + 6 * n
            P.p("news_attW").copy_(w[o]); P.p("news_attb").copy_(w[o + 1].reshape(-1)); P.p("news_attq").copy_(w[o + 2].reshape(-1))
            P.p("user_Wqkv").copy_(torch.cat([w[o + 3], w[o + 4], w[o + 5]], dim=1))
            P.p("user_attW").copy_(w[o + 6]); P.p("user_attb").copy_(w[o + 7].reshape(-1)); P.p("user_attq").copy_(w[o + 8].reshape(-1))

    def get_weights(self):
        P, D = self.params, self.D
        Wn = P.p("news_Wqkv").cpu().numpy()
        out = [P.p("table").cpu().numpy(), Wn[:, :D].copy(), Wn[:, D:2 * D].copy(), Wn[:, 2 * D:].copy()]
        for i in range(len(self.units)):
            out += [P.p(f"d{i}_W").cpu().numpy(), P.p(f"d{i}_b").cpu().numpy(), P.p(f"d{i}_gamma").cpu().numpy(),
                    P.p(f"d{i}_beta").cpu().numpy(), self.bn_mean[i].cpu().numpy(), self.bn_var[i].cpu().numpy()]
        out += [P.p("news_attW").cpu().numpy(), P.p("news_attb").cpu().numpy(), P.p("news_attq").cpu().numpy().reshape(-1, 1)]
        Wu = P.p("user_Wqkv").cpu().numpy()
        out += [Wu[:, :D].copy(), Wu[:, D:2 * D].copy(), Wu[:, 2 * D:].copy(), P.p("user_attW").cpu().numpy(),
                P.p("user_attb").cpu().numpy(), P.p("user_attq").cpu().numpy().reshape(-1, 1)]
        return out

    def count_params(self):
        return super().count_params() + 2 * sum(self.units)

    def trainable_params(self):
        return NRMSEngine.count_params(self)

    # ------------------------------------------------------------------ descriptors / scratch
    def _desc(self, kind, n_seq, training=False):
        math = self.math if training else self.math_infer
        if kind == "news":  # att = 0: Embedding -> Dropout -> SelfAttention only
            return _ebk.SeqEncDesc(n_seq, self.T, self.E, self.nh, self.dh, 0, self.V, self.dropout, math)
        return _ebk.SeqEncDesc(n_seq, self.H, self.D, self.nh, self.dh, self.att, 0, 0.0, math)

    def _dense_desc(self, n_rows, K, U, training):
        math = self.math if training else self.math_infer
        return _ebk.DenseDesc(n_rows, K, U, 1, 1, BN_MOMENTUM, BN_EPS, self.dropout, self.l2, math)

    def _sized(self, key, need):
        cur = self._ws.get(key)
        if cur is None or cur.numel() < need:
            cur = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = cur
        return cur

    def _stack_fwd(self, call, x, training, seed):
        """x [rows, D] -> [rows, u_last]; keeps per-layer inputs / descs / workspaces for backward."""
        lib, P = _ebk.lib(), self.params
        n, ctx, cur, din = x.shape[0], [], x, self.D
        for i, u in enumerate(self.units):
            desc = self._dense_desc(n, din, u, training)
            ws = self._sized(("dense", call, i), lib.ebk_dense_workspace_bytes(C.byref(desc)))
            y = self._buf(f"z{call}_{i}", (n, u))
            _ebk.check(lib.ebk_dense_fwd(C.byref(desc), _ebk.ptr(cur), _ebk.ptr(P.p(f"d{i}_W")), _ebk.ptr(P.p(f"d{i}_b")),
                                         _ebk.ptr(P.p(f"d{i}_gamma")), _ebk.ptr(P.p(f"d{i}_beta")), _ebk.ptr(self.bn_mean[i]),
                                         _ebk.ptr(self.bn_var[i]), int(training), (seed + i) & ((1 << 64) - 1), _ebk.ptr(ws),
                                         ws.numel(), _ebk.ptr(y), _ebk.stream()))
            ctx.append((desc, ws, cur, y, i))
            cur, din = y, u
        return cur, ctx

    def _stack_bwd(self, ctx, d_out, dx0, training, seed, l2_scale):
        lib, P = _ebk.lib(), self.params
        dy = d_out
        for desc, ws, x_in, y, i in reversed(ctx):
            dx = dx0 if i == 0 else self._buf(f"dz_{i % 2}", (desc.N, desc.K))
            _ebk.check(lib.ebk_dense_bwd(C.byref(desc), _ebk.ptr(x_in), _ebk.ptr(P.p(f"d{i}_W")), _ebk.ptr(P.p(f"d{i}_gamma")),
                                         _ebk.ptr(y), int(training), (seed + i) & ((1 << 64) - 1), _ebk.ptr(ws), ws.numel(),
                                         _ebk.ptr(dy), l2_scale, _ebk.ptr(P.g(f"d{i}_W")), _ebk.ptr(P.g(f"d{i}_b")),
                                         _ebk.ptr(P.g(f"d{i}_gamma")), _ebk.ptr(P.g(f"d{i}_beta")), _ebk.ptr(dx), _ebk.stream()))
            dy = dx

    # ------------------------------------------------------------------ forward / backward
    def _encode(self, tok_all, B, Hh, training, seeds=(0, 0, 0)):
        lib, P = _ebk.lib(), self.params
        N = tok_all.shape[0]
        if Hh != self.H:
            raise ValueError(f"history length {Hh} != hparams.history_size {self.H}")
        R, RH = N * self.T, B * self.H * self.T
        dn = self._desc("news", N, training)
        wn = self._workspace("news", dn)
        y0 = self._buf("y0", (R, self.D))
        _ebk.check(lib.ebk_seqenc_fwd(C.byref(dn), _ebk.ptr(tok_all), _ebk.ptr(P.p("table")), _ebk.ptr(P.p("news_Wqkv")),
                                      None, None, None, int(training), seeds[0], 0, _ebk.ptr(wn), wn.numel(), _ebk.ptr(y0),
                                      _ebk.stream()))
        zh, ctx_h = self._stack_fwd(0, y0[:RH], training, seeds[1])
        zc, ctx_c = self._stack_fwd(1, y0[RH:], training, seeds[2])
        U = self.units[-1]
        z = self._buf("z_all", (R, U))
        z[:RH].copy_(zh)
        z[RH:].copy_(zc)
        math = self.math if training else self.math_infer
        da = _ebk.AttLayerDesc(N, self.T, U, self.att, 0.0, math)
        wa = self._sized(("att", 0), lib.ebk_attlayer_workspace_bytes(C.byref(da)))
        n_all = self._buf("n_all", (N, self.D))
        _ebk.check(lib.ebk_attlayer_fwd(C.byref(da), _ebk.ptr(z), _ebk.ptr(P.p("news_attW")), _ebk.ptr(P.p("news_attb")),
                                        _ebk.ptr(P.p("news_attq")), int(training), 0, _ebk.ptr(wa), wa.numel(), _ebk.ptr(n_all),
                                        self.D, _ebk.stream()))
        du = self._desc("user", B, training)
        wu = self._workspace("user", du)
        u = self._buf("u", (B, self.D))
        _ebk.check(lib.ebk_seqenc_fwd(C.byref(du), None, _ebk.ptr(n_all), _ebk.ptr(P.p("user_Wqkv")), _ebk.ptr(P.p("user_attW")),
                                      _ebk.ptr(P.p("user_attb")), _ebk.ptr(P.p("user_attq")), 0, 0, 0, _ebk.ptr(wu), wu.numel(),
                                      _ebk.ptr(u), _ebk.stream()))
        return n_all, u, (dn, wn, ctx_h, ctx_c, z, da, wa, du, wu, y0)

    def forward_logits_parts(self, tok_all, B, C_, training=False, seeds=(0, 0, 0)):
        n_all, u, ctx = self._encode(tok_all, B, self.H, training, seeds)
        return n_all, n_all[B * self.H:].view(B, C_, self.D), u, ctx

    def step_seeds(self):
        base = _mix(self.seed, self.step_count * self.world + self.rank)
        return _mix(base, 1), _mix(base, 2) & ((1 << 62) - 1), _mix(base, 3) & ((1 << 62) - 1)

    def loss_and_grads_dev(self, tok_all, labels, B, C_, training=True, seeds=None, sparse_table=False):
        lib, P = _ebk.lib(), self.params
        seeds = self.step_seeds() if seeds is None else seeds
        n_all, news_c, u, (dn, wn, ctx_h, ctx_c, z, da, wa, du, wu, y0) = self.forward_logits_parts(tok_all, B, C_, training, seeds)
        N, BH = n_all.shape[0], B * self.H
        R, RH, U = N * self.T, BH * self.T, self.units[-1]
        probs = self._buf("probs", (B, C_))
        loss = self._buf("loss", (1,))
        loss.zero_()
        dn_all = self._buf("dn_all", (N, self.D))
        d_user = self._buf("d_user", (B, self.D))
        _ebk.check(lib.ebk_score_loss(self.loss_kind, B, C_, self.D, _ebk.ptr(news_c), _ebk.ptr(u), _ebk.ptr(labels),
                                      1.0 / (B * self.world), 1.0 / B, _ebk.ptr(probs), _ebk.ptr(loss),
                                      _ebk.ptr(dn_all[BH:]), _ebk.ptr(d_user), _ebk.stream()))
        for i in range(len(self.units)):  # + l2 * sum ||W||^2  (nrms.py:147-149)
            W = P.p(f"d{i}_W")
            _ebk.check(lib.ebk_sumsq_accum(_ebk.ptr(W), W.numel(), self.l2, _ebk.ptr(loss), _ebk.stream()))
        _ebk.check(lib.ebk_seqenc_bwd(C.byref(du), None, _ebk.ptr(n_all), _ebk.ptr(P.p("user_Wqkv")), _ebk.ptr(P.p("user_attW")),
                                      _ebk.ptr(P.p("user_attb")), _ebk.ptr(P.p("user_attq")), 0, 0, 0, _ebk.ptr(wu), wu.numel(),
                                      _ebk.ptr(d_user), _ebk.ptr(P.g("user_Wqkv")), _ebk.ptr(P.g("user_attW")),
                                      _ebk.ptr(P.g("user_attb")), _ebk.ptr(P.g("user_attq")), None, _ebk.ptr(dn_all), _ebk.stream()))
        dz = self._buf("dz_all", (R, U))
        _ebk.check(lib.ebk_attlayer_bwd(C.byref(da), _ebk.ptr(z), _ebk.ptr(P.p("news_attW")), _ebk.ptr(P.p("news_attq")),
                                        int(training), 0, _ebk.ptr(wa), wa.numel(), _ebk.ptr(dn_all), self.D,
                                        _ebk.ptr(P.g("news_attW")), _ebk.ptr(P.g("news_attb")), _ebk.ptr(P.g("news_attq")),
                                        _ebk.ptr(dz), _ebk.stream()))
        dy0 = self._buf("dy0", (R, self.D))
        # the l2 gradient 2*l2*W is added once per step (by the history call), scaled like the loss
        self._stack_bwd(ctx_h, dz[:RH], dy0[:RH], training, seeds[1], 1.0 / self.world)
        self._stack_bwd(ctx_c, dz[RH:], dy0[RH:], training, seeds[2], 0.0)
        _ebk.check(lib.ebk_seqenc_bwd(C.byref(dn), _ebk.ptr(tok_all), _ebk.ptr(P.p("table")), _ebk.ptr(P.p("news_Wqkv")),
                                      None, None, None, int(training), seeds[0], 0, _ebk.ptr(wn), wn.numel(), _ebk.ptr(dy0),
                                      _ebk.ptr(P.g("news_Wqkv")), None, None, None, _ebk.ptr(P.g("table")), None, _ebk.stream()))
        return loss, probs

    def predict_host_dedup(self, his, pred, head="softmax"):
        tok, _ = self.to_device_batch(his, pred)
        return self.predict_dev(tok, np.asarray(pred).shape[0], np.asarray(pred).shape[1], head=head)

    def encode_host(self, kind, x):
        x = np.asarray(x)
        if kind == "news":
            # newsencoder.predict: Embedding -> SelfAttention -> Dense/BN stack (moving statistics) -> AttLayer2
            lib, P = _ebk.lib(), self.params
            tok = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, self.T), dtype=np.int32)).to(self.device)
            N = tok.shape[0]
            dn = self._desc("news", N, False)
            wn = self._workspace("news", dn)
            y0 = self._buf("y0", (N * self.T, self.D))
            _ebk.check(lib.ebk_seqenc_fwd(C.byref(dn), _ebk.ptr(tok), _ebk.ptr(P.p("table")), _ebk.ptr(P.p("news_Wqkv")),
                                          None, None, None, 0, 0, 0, _ebk.ptr(wn), wn.numel(), _ebk.ptr(y0), _ebk.stream()))
            z, _ = self._stack_fwd(2, y0, False, 0)
            U = self.units[-1]
            da = _ebk.AttLayerDesc(N, self.T, U, self.att, 0.0, self.math_infer)
            wa = self._sized(("att", 2), lib.ebk_attlayer_workspace_bytes(C.byref(da)))
            out = torch.empty((N, self.D), device=self.device)
            _ebk.check(lib.ebk_attlayer_fwd(C.byref(da), _ebk.ptr(z.contiguous()), _ebk.ptr(P.p("news_attW")),
                                            _ebk.ptr(P.p("news_attb")), _ebk.ptr(P.p("news_attq")), 0, 0, _ebk.ptr(wa),
                                            wa.numel(), _ebk.ptr(out), self.D, _ebk.stream()))
            return out.cpu().numpy()
        B = x.shape[0]
        tok = torch.from_numpy(np.ascontiguousarray(x.reshape(B * self.H, self.T), dtype=np.int32)).to(self.device)
        _, u, _ = self._encode(tok, B, self.H, False)
        return u.clone().cpu().numpy()
