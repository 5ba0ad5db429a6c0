"""ctypes binding of libebk.so (the C-ABI declared in include/ebk.h).

There is deliberately NO fallback: if the CUDA library is missing or the device is not
a CC 10.x (sm_100a) part, every compute entry point raises.  Tensors cross the boundary
as raw device pointers (``tensor.data_ptr()``) plus explicit shapes; the stream is
torch's current stream.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG_ROOT = Path(__file__).resolve().parents[3]  # .../ebnerd-benchmark_b200
LIB_PATH = Path(os.environ.get("EBK_LIB", _PKG_ROOT / "csrc" / "libebk.so"))

MATH_FP32 = 0
MATH_TF32 = 1
MATH_TF32X3 = 2
LOSS_CATEGORICAL_CE = 0   # hparams.loss "cross_entropy_loss"
LOSS_BINARY_CE = 1        # hparams.loss "log_loss"

SYMBOLS = [
    "ebk_last_error", "ebk_version", "ebk_device_ok",
    "ebk_seqenc_workspace_bytes", "ebk_seqenc_fwd", "ebk_seqenc_bwd", "ebk_seqenc_fwd_opts", "ebk_seqenc_bwd_opts",
    "ebk_join_deferred", "ebk_seqenc_uses_tma", "ebk_ipc_export", "ebk_ipc_open", "ebk_memcpy_async",
    "ebk_score_softmax_ce", "ebk_score_loss", "ebk_score_sigmoid", "ebk_adam_keras_step", "ebk_adam_keras_step_p",
    "ebk_embed_adam_step_p", "ebk_dp_token_flags", "ebk_adam_pull_step",
    "ebk_embed_adam_workspace_bytes", "ebk_embed_adam_step", "ebk_token_csr_bytes",
    "ebk_dense_workspace_bytes", "ebk_dense_fwd", "ebk_dense_bwd", "ebk_dense_fwd_p", "ebk_dense_bwd_p", "ebk_sumsq_accum",
    "ebk_attlayer_workspace_bytes", "ebk_attlayer_fwd", "ebk_attlayer_bwd",
    "ebk_conv1d_workspace_bytes", "ebk_conv1d_fwd", "ebk_conv1d_bwd",
    "ebk_catview_workspace_bytes", "ebk_catview_fwd", "ebk_catview_bwd",
    "ebk_launch_count", "ebk_prof_enable", "ebk_prof_is_enabled", "ebk_prof_num_tags", "ebk_prof_tag_name", "ebk_prof_collect",
    "ebk_gemm", "ebk_gemm_tma", "ebk_attention_core_fwd", "ebk_attention_core_bwd", "ebk_dropout_mask",
]


class EbkError(RuntimeError):
    pass


class SeqEncDesc(C.Structure):
    """Mirror of ebk_seqenc_desc (include/ebk.h)."""
    _fields_ = [
        ("n_seq", C.c_int32), ("L", C.c_int32), ("Din", C.c_int32), ("nh", C.c_int32),
        ("dh", C.c_int32), ("att", C.c_int32), ("V", C.c_int32), ("dropout", C.c_float),
        ("math", C.c_int32),
    ]


class SeqEncOpts(C.Structure):
    """Mirror of ebk_seqenc_opts (include/ebk.h): per-call options, nothing sticky."""
    _fields_ = [("defer_wgrad", C.c_int32), ("table_grad_event", C.c_void_p), ("peer_tables", C.POINTER(C.c_void_p)),
                ("peer_world", C.c_int32), ("peer_shard_floats", C.c_size_t), ("step_dev", C.c_void_p),
                ("token_csr_ws", C.c_void_p), ("token_csr_ws_bytes", C.c_size_t)]


class DenseDesc(C.Structure):
    """Mirror of ebk_dense_desc (include/ebk.h)."""
    _fields_ = [
        ("N", C.c_int32), ("K", C.c_int32), ("U", C.c_int32), ("relu", C.c_int32), ("bn", C.c_int32),
        ("bn_momentum", C.c_float), ("bn_eps", C.c_float), ("dropout", C.c_float), ("l2", C.c_float),
        ("math", C.c_int32),
    ]


class AttLayerDesc(C.Structure):
    """Mirror of ebk_attlayer_desc (include/ebk.h)."""
    _fields_ = [("n_seq", C.c_int32), ("L", C.c_int32), ("D", C.c_int32), ("att", C.c_int32),
                ("dropout", C.c_float), ("math", C.c_int32)]


class Conv1dDesc(C.Structure):
    """Mirror of ebk_conv1d_desc (include/ebk.h)."""
    _fields_ = [("n_seq", C.c_int32), ("L", C.c_int32), ("E", C.c_int32), ("F", C.c_int32), ("window", C.c_int32),
                ("V", C.c_int32), ("dropout", C.c_float), ("relu", C.c_int32), ("math", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    """Load libebk.so once; raise loudly if it is absent (no CPU/PyTorch fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EbkError(
            f"CUDA extension {LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C {LIB_PATH.parent}`); there is no CPU fallback.")
    l = C.CDLL(str(LIB_PATH))
    vp, i32, u64, f32, f64, sz = C.c_void_p, C.c_int32, C.c_uint64, C.c_float, C.c_double, C.c_size_t
    dp = C.POINTER(SeqEncDesc)
    l.ebk_last_error.restype = C.c_char_p
    l.ebk_last_error.argtypes = []
    l.ebk_version.restype = C.c_int
    l.ebk_device_ok.restype = C.c_int
    l.ebk_seqenc_workspace_bytes.restype = sz
    l.ebk_seqenc_workspace_bytes.argtypes = [dp]
    l.ebk_seqenc_fwd.argtypes = [dp, vp, vp, vp, vp, vp, vp, C.c_int, u64, u64, vp, sz, vp, vp]
    l.ebk_seqenc_bwd.argtypes = [dp, vp, vp, vp, vp, vp, vp, C.c_int, u64, u64, vp, sz, vp,
                                 vp, vp, vp, vp, vp, vp, vp]
    op = C.POINTER(SeqEncOpts)
    l.ebk_seqenc_fwd_opts.argtypes = [dp, op, vp, vp, vp, vp, vp, vp, C.c_int, u64, u64, vp, sz, vp, vp]
    l.ebk_seqenc_bwd_opts.argtypes = [dp, op, vp, vp, vp, vp, vp, vp, C.c_int, u64, u64, vp, sz, vp,
                                      vp, vp, vp, vp, vp, vp, vp]
    l.ebk_join_deferred.argtypes = [vp]
    l.ebk_seqenc_uses_tma.argtypes = [dp]
    l.ebk_ipc_export.argtypes = [vp, vp, C.POINTER(sz)]
    l.ebk_ipc_open.argtypes = [vp, sz, C.POINTER(vp)]
    l.ebk_memcpy_async.argtypes = [vp, vp, sz, vp]
    ddp = C.POINTER(DenseDesc)
    l.ebk_dense_workspace_bytes.restype = sz
    l.ebk_dense_workspace_bytes.argtypes = [ddp]
    l.ebk_dense_fwd.argtypes = [ddp, vp, vp, vp, vp, vp, vp, vp, C.c_int, u64, vp, sz, vp, vp]
    l.ebk_dense_bwd.argtypes = [ddp, vp, vp, vp, vp, C.c_int, u64, vp, sz, vp, f32, vp, vp, vp, vp, vp, vp]
    l.ebk_dense_fwd_p.argtypes = [ddp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, u64, vp, sz, vp, vp]
    l.ebk_dense_bwd_p.argtypes = [ddp, vp, vp, vp, vp, C.c_int, vp, C.c_int, u64, vp, sz, vp, f32, vp, vp, vp, vp, vp, vp]
    l.ebk_sumsq_accum.argtypes = [vp, sz, f32, vp, vp]
    adp, cdp = C.POINTER(AttLayerDesc), C.POINTER(Conv1dDesc)
    l.ebk_attlayer_workspace_bytes.restype = sz
    l.ebk_attlayer_workspace_bytes.argtypes = [adp]
    l.ebk_attlayer_fwd.argtypes = [adp, vp, vp, vp, vp, C.c_int, u64, vp, sz, vp, i32, vp]
    l.ebk_attlayer_bwd.argtypes = [adp, vp, vp, vp, C.c_int, u64, vp, sz, vp, i32, vp, vp, vp, vp, vp]
    l.ebk_conv1d_workspace_bytes.restype = sz
    l.ebk_conv1d_workspace_bytes.argtypes = [cdp]
    l.ebk_conv1d_fwd.argtypes = [cdp, vp, vp, vp, vp, C.c_int, u64, vp, sz, vp, vp]
    l.ebk_conv1d_bwd.argtypes = [cdp, vp, vp, vp, C.c_int, u64, u64, vp, sz, vp, vp, vp, vp, vp]
    l.ebk_catview_workspace_bytes.restype = sz
    l.ebk_catview_workspace_bytes.argtypes = [i32, i32]
    l.ebk_catview_fwd.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp, i32, vp]
    l.ebk_catview_bwd.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp, i32, vp, vp, vp, vp]
    l.ebk_score_softmax_ce.argtypes = [i32, i32, i32, vp, vp, vp, f32, vp, vp, vp, vp, vp]
    l.ebk_score_loss.argtypes = [i32, i32, i32, i32, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp]
    l.ebk_score_sigmoid.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    l.ebk_adam_keras_step.argtypes = [vp, vp, vp, vp, sz, f32, f64, f64, f32, C.c_int, vp]
    l.ebk_adam_keras_step_p.argtypes = [vp, vp, vp, vp, sz, f32, vp, f64, f64, f32, C.c_int, vp]
    l.ebk_embed_adam_step_p.argtypes = [i32, i32, i32, vp, vp, f32, u64, vp, vp, vp, vp, f32, vp, f64, f64, f32, vp, sz, vp]
    l.ebk_dp_token_flags.argtypes = [i32, i32, vp, vp, sz, vp]
    l.ebk_adam_pull_step.argtypes = [vp, vp, vp, C.POINTER(vp), vp, sz, i32, i32, i32, sz, sz, f32, vp, f64, f64, f32, vp]
    l.ebk_embed_adam_workspace_bytes.restype = sz
    l.ebk_embed_adam_workspace_bytes.argtypes = [i32, i32]
    l.ebk_token_csr_bytes.restype = sz
    l.ebk_token_csr_bytes.argtypes = [i32, i32]
    l.ebk_embed_adam_step.argtypes = [i32, i32, i32, vp, vp, f32, u64, vp, vp, vp, vp, f32, f64, f64, f32, vp, sz, vp]
    l.ebk_gemm.argtypes = [i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, f32, vp]
    l.ebk_gemm_tma.argtypes = [i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, f32, f32, vp]
    l.ebk_attention_core_fwd.argtypes = [i32, i32, i32, i32, vp, vp, vp]
    l.ebk_attention_core_bwd.argtypes = [i32, i32, i32, i32, vp, vp, f32, u64, vp, vp]
    l.ebk_dropout_mask.argtypes = [u64, f32, sz, vp, vp]
    for name in SYMBOLS:
        fn = getattr(l, name)
        if name != "ebk_last_error" and not name.endswith("_workspace_bytes"):
            fn.restype = C.c_int
    l.ebk_launch_count.restype = C.c_longlong
    l.ebk_prof_tag_name.restype = C.c_char_p
    l.ebk_prof_tag_name.argtypes = [C.c_int]
    l.ebk_prof_enable.argtypes = [C.c_int]
    l.ebk_prof_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    _lib = l
    return l


def require_device() -> None:
    if not torch.cuda.is_available():
        raise EbkError("no CUDA device: the ebk hot path is sm_100a-only and has no CPU fallback")
    if not lib().ebk_device_ok():
        raise EbkError("current CUDA device is not compute capability 10.x (B200 / sm_100a)")


def check(status: int) -> None:
    if status != 0:
        raise EbkError(f"ebk status {status}: {lib().ebk_last_error().decode()}")


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "ebk needs contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def prof_collect() -> dict[str, tuple[float, int]]:
    """{kernel-group name: (total ms, count)} recorded since ebk_prof_enable(1)."""
    l = lib()
    n = l.ebk_prof_num_tags()
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(l.ebk_prof_collect(ms, cnt))
    return {l.ebk_prof_tag_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
