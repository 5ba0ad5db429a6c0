"""Device engine of NAML (reference src/ebrec/models/newsrec/naml.py:13-374, base_model.py:19-86).

News encoder per article (naml.py:91-141): title and body views = shared Embedding -> Dropout ->
Conv1D(filter_num, window, same, relu) -> Dropout -> AttLayer2 (naml.py:143-203); vert and subvert views =
Embedding(n, 10) -> Dense(filter_num, relu) (naml.py:205-252); the four [F] views are stacked and pooled by
another AttLayer2 (naml.py:133-138).  User encoder = AttLayer2 over the history's news vectors
(naml.py:62-89).  Click score, loss and Keras Adam are those of NRMS.

History and candidate articles are encoded in ONE pass over N = B*(H+C) rows.  All parameters live in one
flat HBM buffer (see _engine.FlatParams); C-ABI calls: ebk_conv1d_*, ebk_attlayer_*, ebk_catview_*,
ebk_score_*, ebk_adam_keras_step.

Weight order of get_weights/set_weights (`NAML_WEIGHT_ORDER`): Keras' own order for this nested functional
graph cannot be reproduced without TensorFlow, so the order is the construction order of naml.py:318-330:
table, title (conv kernel [w,E,F], conv bias, att W, b, q), body (same), vert (emb, dense W, b),
subvert (same), news-fusion AttLayer2 (W, b, q), user AttLayer2 (W, b, q).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ebk
from ._engine import FlatParams, NRMSEngine, _mix

TEXT_VIEWS = ("title", "body")
CAT_VIEWS = ("vert", "subvert")
NAML_WEIGHT_ORDER = (["table"] + [f"{v}_{s}" for v in TEXT_VIEWS for s in ("convW", "convb", "W", "b", "q")]
                     + [f"{v}_{s}" for v in CAT_VIEWS for s in ("emb", "denseW", "denseb")]
                     + [f"{v}_{s}" for v in ("news", "user") for s in ("W", "b", "q")])


class NAMLEngine(NRMSEngine):
    def __init__(self, *, V, E, T, Tb, H, F, att, window, vert_num, vert_dim, subvert_num, subvert_dim, dropout, lr,
                 cnn_relu=True, dense_relu=True, seed=None, math=_ebk.MATH_TF32, device=None, beta1=0.9, beta2=0.999,
                 eps=1e-7):
        _ebk.require_device()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.V, self.E, self.T, self.Tb, self.H = int(V), int(E), int(T), int(Tb), int(H)
        self.F = self.D = int(F)  # D: width of news/user vectors (what the score kernels and encoder views use)
        self.att, self.window = int(att), int(window)
        self.cat = {"vert": (int(vert_num), int(vert_dim)), "subvert": (int(subvert_num), int(subvert_dim))}
        self.cnn_relu, self.dense_relu = int(bool(cnn_relu)), int(bool(dense_relu))
        self.dropout = float(dropout)
        self.lr, self.beta1, self.beta2, self.eps = float(lr), beta1, beta2, eps
        self.math = int(math)
        self.math_infer = _ebk.MATH_TF32X3 if self.math == _ebk.MATH_TF32 else self.math
        self.seed = 0 if seed is None else int(seed)
        self.step_count = 0
        F_, A, w = self.F, self.att, self.window
        spec = [("table", (self.V, self.E))]
        for v in TEXT_VIEWS:
            spec += [(f"{v}_convW", (w * self.E, F_)), (f"{v}_convb", (F_,)), (f"{v}_W", (F_, A)), (f"{v}_b", (A,)),
                     (f"{v}_q", (A,))]
        for v in CAT_VIEWS:
            n, d = self.cat[v]
            spec += [(f"{v}_emb", (n, d)), (f"{v}_denseW", (d, F_)), (f"{v}_denseb", (F_,))]
        for v in ("news", "user"):
            spec += [(f"{v}_W", (F_, A)), (f"{v}_b", (A,)), (f"{v}_q", (A,))]
        self.params = FlatParams(spec, self.device)
        self._ws, self._bufs = {}, {}
        self.world, self.rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world, self.rank = torch.distributed.get_world_size(), torch.distributed.get_rank()

    # ------------------------------------------------------------------ weights
    def set_weights(self, weights):
        if len(weights) != len(NAML_WEIGHT_ORDER):
            raise ValueError(f"NAML expects {len(NAML_WEIGHT_ORDER)} weight arrays, got {len(weights)}")
        P = self.params
        with torch.no_grad():
            for name, a in zip(NAML_WEIGHT_ORDER, weights):
                t = torch.as_tensor(np.asarray(a, dtype=np.float32))
                P.p(name).copy_(t.reshape(P.p(name).shape))

    def get_weights(self):
        P, out = self.params, []
        for name in NAML_WEIGHT_ORDER:
            a = P.p(name).cpu().numpy()
            if name.endswith("_convW"):
                a = a.reshape(self.window, self.E, self.F)  # Keras Conv1D kernel shape
            elif name.endswith("_q"):
                a = a.reshape(-1, 1)
            out.append(a)
        return out

    # ------------------------------------------------------------------ scratch
    def _ws_for(self, key, need):
        cur = self._ws.get(key)
        if cur is None or cur.numel() < need:
            cur = torch.empty(max(int(need), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = cur
        return cur

    def _att(self, key, n_seq, L, dropout, training):
        lib = _ebk.lib()
        d = _ebk.AttLayerDesc(n_seq, L, self.F, self.att, dropout, self.math if training else self.math_infer)
        need = lib.ebk_attlayer_workspace_bytes(C.byref(d))
        if need == 0 and n_seq > 0:
            raise _ebk.EbkError(f"bad attlayer descriptor: {lib.ebk_last_error().decode()}")
        return d, self._ws_for(("att", key), need)

    def _conv(self, key, n_seq, L, training):
        lib = _ebk.lib()
        d = _ebk.Conv1dDesc(n_seq, L, self.E, self.F, self.window, self.V, self.dropout, self.cnn_relu,
                            self.math if training else self.math_infer)
        need = lib.ebk_conv1d_workspace_bytes(C.byref(d))
        if need == 0 and n_seq > 0:
            raise _ebk.EbkError(f"bad conv1d descriptor: {lib.ebk_last_error().decode()}")
        return d, self._ws_for(("conv", key), need)

    # ------------------------------------------------------------------ forward
    def _encode_news(self, x, training, seeds):
        """x = (title [N,T], body [N,Tb], vert [N], subvert [N]) int32 -> n_all [N, F] (+ backward context)."""
        lib, P, st = _ebk.lib(), self.params, _ebk.stream()
        N, F_ = x[0].shape[0], self.F
        cat = self._buf("cat", (N, 4, F_))
        ctx = {"N": N, "x": x, "text": [], "cats": []}
        for vi, (v, L) in enumerate(zip(TEXT_VIEWS, (self.T, self.Tb))):
            tok = x[vi]
            dc, wc = self._conv(v, N, L, training)
            y = self._buf(f"y_{v}", (N * L, F_))
            _ebk.check(lib.ebk_conv1d_fwd(C.byref(dc), _ebk.ptr(tok), _ebk.ptr(P.p("table")), _ebk.ptr(P.p(f"{v}_convW")),
                                          _ebk.ptr(P.p(f"{v}_convb")), int(training), seeds[2 * vi], _ebk.ptr(wc), wc.numel(),
                                          _ebk.ptr(y), st))
            da, wa = self._att(v, N, L, self.dropout, training)
            _ebk.check(lib.ebk_attlayer_fwd(C.byref(da), _ebk.ptr(y), _ebk.ptr(P.p(f"{v}_W")), _ebk.ptr(P.p(f"{v}_b")),
                                            _ebk.ptr(P.p(f"{v}_q")), int(training), seeds[2 * vi + 1], _ebk.ptr(wa), wa.numel(),
                                            C.c_void_p(cat.data_ptr() + 4 * vi * F_), 4 * F_, st))
            ctx["text"].append((v, vi, tok, dc, wc, y, da, wa))
        for ci, v in enumerate(CAT_VIEWS):
            n_cat, dim = self.cat[v]
            wk = self._ws_for(("cat", v), lib.ebk_catview_workspace_bytes(n_cat, F_))
            _ebk.check(lib.ebk_catview_fwd(N, n_cat, dim, F_, self.dense_relu, _ebk.ptr(x[2 + ci]), _ebk.ptr(P.p(f"{v}_emb")),
                                           _ebk.ptr(P.p(f"{v}_denseW")), _ebk.ptr(P.p(f"{v}_denseb")), _ebk.ptr(wk), wk.numel(),
                                           C.c_void_p(cat.data_ptr() + 4 * (2 + ci) * F_), 4 * F_, st))
            ctx["cats"].append((v, ci, wk))
        dn, wn = self._att("news", N, 4, 0.0, training)
        n_all = self._buf("n_all", (N, F_))
        _ebk.check(lib.ebk_attlayer_fwd(C.byref(dn), _ebk.ptr(cat), _ebk.ptr(P.p("news_W")), _ebk.ptr(P.p("news_b")),
                                        _ebk.ptr(P.p("news_q")), 0, 0, _ebk.ptr(wn), wn.numel(), _ebk.ptr(n_all), F_, st))
        ctx.update(cat=cat, dn=dn, wn=wn)
        return n_all, ctx

    def forward_logits_parts(self, x, B, C_, training=False, seeds=(0, 0, 0, 0)):
        lib, P = _ebk.lib(), self.params
        N = x[0].shape[0]
        Hh = (N - B * C_) // B
        if Hh != self.H:
            raise ValueError(f"history length {Hh} != hparams.history_size {self.H}")
        n_all, ctx = self._encode_news(x, training, seeds)
        du, wu = self._att("user", B, self.H, 0.0, training)
        u = self._buf("u", (B, self.F))
        _ebk.check(lib.ebk_attlayer_fwd(C.byref(du), _ebk.ptr(n_all), _ebk.ptr(P.p("user_W")), _ebk.ptr(P.p("user_b")),
                                        _ebk.ptr(P.p("user_q")), 0, 0, _ebk.ptr(wu), wu.numel(), _ebk.ptr(u), self.F,
                                        _ebk.stream()))
        ctx.update(du=du, wu=wu)
        return n_all, n_all[B * self.H:].view(B, C_, self.F), u, ctx

    def step_seeds(self):
        base = _mix(self.seed, self.step_count * self.world + self.rank)
        return tuple(_mix(base, i + 1) for i in range(4))

    # ------------------------------------------------------------------ backward
    def loss_and_grads_dev(self, x, labels, B, C_, training=True, seeds=None):
        lib, P, st = _ebk.lib(), self.params, _ebk.stream()
        seeds = self.step_seeds() if seeds is None else seeds
        n_all, news_c, u, ctx = self.forward_logits_parts(x, B, C_, training, seeds)
        N, F_, BH = ctx["N"], self.F, B * self.H
        probs = self._buf("probs", (B, C_))
        loss = self._buf("loss", (1,))
        loss.zero_()
        dn_all = self._buf("dn_all", (N, F_))
        d_user = self._buf("d_user", (B, F_))
        _ebk.check(lib.ebk_score_loss(self.loss_kind, B, C_, F_, _ebk.ptr(news_c), _ebk.ptr(u), _ebk.ptr(labels),
                                      1.0 / (B * self.world), 1.0 / B, _ebk.ptr(probs), _ebk.ptr(loss),
                                      _ebk.ptr(dn_all[BH:]), _ebk.ptr(d_user), st))
        # user AttLayer2: its input gradient IS the gradient of the history rows of n_all
        _ebk.check(lib.ebk_attlayer_bwd(C.byref(ctx["du"]), _ebk.ptr(n_all), _ebk.ptr(P.p("user_W")), _ebk.ptr(P.p("user_q")), 0, 0,
                                        _ebk.ptr(ctx["wu"]), ctx["wu"].numel(), _ebk.ptr(d_user), F_, _ebk.ptr(P.g("user_W")),
                                        _ebk.ptr(P.g("user_b")), _ebk.ptr(P.g("user_q")), _ebk.ptr(dn_all), st))
        dcat = self._buf("dcat", (N, 4, F_))
        _ebk.check(lib.ebk_attlayer_bwd(C.byref(ctx["dn"]), _ebk.ptr(ctx["cat"]), _ebk.ptr(P.p("news_W")), _ebk.ptr(P.p("news_q")),
                                        0, 0, _ebk.ptr(ctx["wn"]), ctx["wn"].numel(), _ebk.ptr(dn_all), F_,
                                        _ebk.ptr(P.g("news_W")), _ebk.ptr(P.g("news_b")), _ebk.ptr(P.g("news_q")), _ebk.ptr(dcat), st))
        for v, vi, tok, dc, wc, y, da, wa in ctx["text"]:
            L = dc.L
            dy = self._buf("dy_text", (N * L, F_))
            _ebk.check(lib.ebk_attlayer_bwd(C.byref(da), _ebk.ptr(y), _ebk.ptr(P.p(f"{v}_W")), _ebk.ptr(P.p(f"{v}_q")),
                                            int(training), seeds[2 * vi + 1], _ebk.ptr(wa), wa.numel(),
                                            C.c_void_p(dcat.data_ptr() + 4 * vi * F_), 4 * F_, _ebk.ptr(P.g(f"{v}_W")),
                                            _ebk.ptr(P.g(f"{v}_b")), _ebk.ptr(P.g(f"{v}_q")), _ebk.ptr(dy), st))
            _ebk.check(lib.ebk_conv1d_bwd(C.byref(dc), _ebk.ptr(tok), _ebk.ptr(P.p(f"{v}_convW")), _ebk.ptr(y), int(training),
                                          seeds[2 * vi], seeds[2 * vi + 1], _ebk.ptr(wc), wc.numel(), _ebk.ptr(dy),
                                          _ebk.ptr(P.g(f"{v}_convW")), _ebk.ptr(P.g(f"{v}_convb")), _ebk.ptr(P.g("table")), st))
        for v, ci, wk in ctx["cats"]:
            n_cat, dim = self.cat[v]
            _ebk.check(lib.ebk_catview_bwd(N, n_cat, dim, F_, self.dense_relu, _ebk.ptr(ctx["x"][2 + ci]), _ebk.ptr(P.p(f"{v}_emb")),
                                           _ebk.ptr(P.p(f"{v}_denseW")), _ebk.ptr(wk), wk.numel(),
                                           C.c_void_p(dcat.data_ptr() + 4 * (2 + ci) * F_), 4 * F_, _ebk.ptr(P.g(f"{v}_emb")),
                                           _ebk.ptr(P.g(f"{v}_denseW")), _ebk.ptr(P.g(f"{v}_denseb")), st))
        return loss, probs

    # ------------------------------------------------------------------ host convenience
    @staticmethod
    def pack_inputs(arrays):
        """The 8 arrays of NAMLDataLoader (his title/body/vert/subvert, pred title/body/vert/subvert) ->
        (title [N,T], body [N,Tb], vert [N], subvert [N]) int32 with the B*H history rows first."""
        a = [np.asarray(v) for v in arrays]
        out = []
        for k in range(4):
            h, c = a[k], a[4 + k]
            h2 = h.reshape(h.shape[0] * h.shape[1], -1)
            c2 = c.reshape(c.shape[0] * c.shape[1], -1)
            m = np.concatenate([h2, c2], axis=0).astype(np.int32, copy=False)
            out.append(np.ascontiguousarray(m if k < 2 else m.reshape(-1)))
        return out

    def set_article_matrices(self, title_matrix: np.ndarray, body_matrix: np.ndarray) -> None:
        """Device-resident batch feed: the [n+1, title_size] and [n+1, body_size] token matrices of a
        NAMLDataLoaderDevice are uploaded once; batches then carry article row indices for the two text views."""
        t = np.ascontiguousarray(np.asarray(title_matrix), dtype=np.int32)
        b = np.ascontiguousarray(np.asarray(body_matrix), dtype=np.int32)
        if t.ndim != 2 or t.shape[1] != self.T or b.ndim != 2 or b.shape[1] != self.Tb:
            raise ValueError(f"token matrices must be [n, {self.T}] and [n, {self.Tb}], got {t.shape} and {b.shape}")
        self.title_matrix, self.body_matrix = torch.from_numpy(t).to(self.device), torch.from_numpy(b).to(self.device)

    def to_device_batch(self, arrays, y=None):
        if np.asarray(arrays[0]).ndim == 2:
            # index feed: (his_title_idx [B,H], his_body_idx [B,H], his_vert [B,H,1], his_subvert [B,H,1], pred_...)
            if getattr(self, "title_matrix", None) is None:
                raise ValueError("index batches need set_article_matrices(title_matrix, body_matrix) first")
            a = [np.asarray(v) for v in arrays]
            rows = []
            for k, mat in ((0, self.title_matrix), (1, self.body_matrix)):
                idx = np.concatenate([a[k].reshape(-1), a[4 + k].reshape(-1)]).astype(np.int64)
                if idx.size and (idx.min() < 0 or idx.max() >= mat.shape[0]):
                    raise IndexError(f"article row index outside [0, {mat.shape[0]})")
                rows.append(mat.index_select(0, torch.from_numpy(idx).to(self.device, non_blocking=True)))
            cats = [torch.from_numpy(np.concatenate([a[k].reshape(-1), a[4 + k].reshape(-1)]).astype(np.int32)).to(
                self.device, non_blocking=True) for k in (2, 3)]
            x = (rows[0], rows[1], cats[0], cats[1])
            lab = None
            if y is not None:
                lab = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(self.device, non_blocking=True)
            return x, lab
        x = tuple(torch.from_numpy(m).to(self.device, non_blocking=True) for m in self.pack_inputs(arrays))
        lab = None
        if y is not None:
            lab = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(self.device, non_blocking=True)
        return x, lab

    def encode_host(self, kind, x):
        """newsencoder.predict ([N, T+Tb+2] ids -> [N,F]) / userencoder.predict ([B,H,T+Tb+2] -> [B,F])."""
        x = np.asarray(x)
        W = self.T + self.Tb + 2
        rows = np.ascontiguousarray(x.reshape(-1, W), dtype=np.int32)
        parts = (rows[:, :self.T], rows[:, self.T:self.T + self.Tb], rows[:, self.T + self.Tb], rows[:, self.T + self.Tb + 1])
        xd = tuple(torch.from_numpy(np.ascontiguousarray(p)).to(self.device) for p in parts)
        n_all, _ = self._encode_news(xd, False, (0, 0, 0, 0))
        if kind == "news":
            return n_all.clone().cpu().numpy()
        lib, P = _ebk.lib(), self.params
        B = x.shape[0]
        du, wu = self._att("user", B, self.H, 0.0, False)
        u = torch.empty((B, self.F), device=self.device)
        _ebk.check(lib.ebk_attlayer_fwd(C.byref(du), _ebk.ptr(n_all), _ebk.ptr(P.p("user_W")), _ebk.ptr(P.p("user_b")),
                                        _ebk.ptr(P.p("user_q")), 0, 0, _ebk.ptr(wu), wu.numel(), _ebk.ptr(u), self.F, _ebk.stream()))
        return u.cpu().numpy()
