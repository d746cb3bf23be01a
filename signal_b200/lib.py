"""ctypes binding of libsignal_b200.so (C ABI: include/signal_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (plain nvcc,
sm_100a).  There is no CPU or PyTorch fallback: if the library is missing or a
call is rejected, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsignal_b200.so")

SIG_F32, SIG_BF16, SIG_F16 = 0, 1, 2
CTX_SIM, CTX_ALIGN, CTX_SELECT, CTX_DAS = 0, 1, 2, 3
FLAG_FORCE_SIMT = 1
SIG_FLAG_SHARE_SMS = 4   # sig_align_fwd/bwd: SIM runs concurrently on another stream, leave it SMs (include/signal_b200.h)
SIG_FLAG_PATCH_MEAN = 8  # sig_align_fwd: the GAM mean pool was deposited in ctx by the caller (include/signal_b200.h)
SIG_FLAG_EAGER_BWD = 2   # sig_align_fwd also runs the loss-weight-independent part of the backward (include/signal_b200.h)

_VP3 = C.c_void_p * 3
_I64x3 = C.c_int64 * 3


class SigTokens(C.Structure):
    _fields_ = [("patch", _VP3), ("cls", _VP3), ("patch_stride_b", _I64x3), ("patch_stride_l", _I64x3),
                ("cls_stride_b", _I64x3), ("dtype", C.c_int32), ("B", C.c_int32), ("L", C.c_int32), ("d", C.c_int32)]


class SigTokenGrads(C.Structure):
    _fields_ = [("dpatch", _VP3), ("dcls", _VP3), ("patch_stride_b", _I64x3), ("patch_stride_l", _I64x3),
                ("cls_stride_b", _I64x3), ("accumulate", C.c_int32), ("zero_cls", C.c_int32),
                ("wait_event", C.c_void_p), ("done_event", C.c_void_p),
                ("fuse_skip_dx", C.c_int32), ("reserved_", C.c_int32), ("fuse_pds", C.c_void_p), ("fuse_dxqt", C.c_void_p)]


SIM_PARAM_FIELDS = ["sel_wq", "sel_bq", "sel_wk", "sel_bk", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b",
                    "ffn0_w", "ffn0_b", "ffn2_w", "ffn2_b", "ln1_w", "ln1_b", "ln2_w", "ln2_b"]
SIM_GRAD_FIELDS = SIM_PARAM_FIELDS[4:]


class SigSelFold(C.Structure):
    _fields_ = [("m_hl", C.c_void_p), ("v", C.c_void_p), ("u", C.c_void_p), ("s0", C.c_void_p)]


class SigSimParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in SIM_PARAM_FIELDS] + [("sel_fold", C.POINTER(SigSelFold)), ("pool_out", C.c_void_p),
                                                                ("pool_event", C.c_void_p)]


class SigSimParamGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in SIM_GRAD_FIELDS] + [("early_event", C.c_void_p), ("late_event", C.c_void_p)]


ALIGN_MOD_FIELDS = ["proj_q_w", "proj_q_b", "off0_w", "off0_b", "off2_w", "off2_b", "off4_w"]


class SigAlignParams(C.Structure):
    _fields_ = [("contra_temp", C.c_void_p)] + [(n, _VP3) for n in ALIGN_MOD_FIELDS] + [("patch_mean_event", C.c_void_p)]


class SigAlignParamGrads(C.Structure):
    _fields_ = [("contra_temp", C.c_void_p)] + [(n, _VP3) for n in ALIGN_MOD_FIELDS] + [("done_event", C.c_void_p)]


class SigXchgPeers(C.Structure):
    _fields_ = [("buf", C.c_void_p * 8), ("flags", C.c_void_p * 8), ("multicast", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32)]


_lib = None

# every symbol include/signal_b200.h declares (tests check that the .so exports them all)
EXPORTS = [
    "sig_version", "sig_error_string", "sig_ctx_bytes",
    "sig_sim_fwd", "sig_sim_bwd", "sig_sim_fold_selection", "sig_sim_select_fwd", "sig_sim_select_from_scores", "sig_mask_mul_bwd",
    "sig_sim_attn_fwd", "sig_sim_attn_bwd", "sig_align_fwd", "sig_align_bwd", "sig_das_fwd", "sig_das_bwd",
    "sig_volume3_ws_bytes", "sig_volume3_fwd", "sig_volume3_bwd",
    "sig_debug_launch_count", "sig_profile_enable", "sig_profile_collect", "sig_debug_gemm_bf16", "sig_debug_tc_stamps", "sig_profile_timeline", "sig_profile_scope_begin", "sig_profile_scope_end", "sig_sim_dx_operands",
    "sig_convert_half", "sig_xchg_flag_bytes", "sig_xchg_allreduce_f32", "sig_infer_features", "sig_euclidean_distmat", "sig_rank_eval",
    "sig_loss_ws_bytes", "sig_xent_ls_fwd", "sig_xent_ls_bwd", "sig_triplet_fwd", "sig_triplet_bwd", "sig_bnneck_ws_bytes", "sig_bnneck_cls_fwd", "sig_bnneck_cls_bwd",
    "sig_tokens_ws_bytes", "sig_tokens_fwd", "sig_tokens_bwd", "sig_align_patch_mean_slot",
    "sig_volume_n_ws_bytes", "sig_volume_n_fwd", "sig_volume_n_bwd",
]


def load():
    """Load (once) and return the ctypes handle.  Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"signal_b200: CUDA extension not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, u, sz, i64 = C.c_void_p, C.c_int, C.c_uint, C.c_size_t, C.c_int64
    P = C.POINTER
    lib.sig_version.restype = i
    lib.sig_error_string.restype = C.c_char_p
    lib.sig_error_string.argtypes = [i]
    lib.sig_ctx_bytes.restype = sz
    lib.sig_ctx_bytes.argtypes = [i, i, i, i, i, u]
    lib.sig_sim_fwd.argtypes = [P(SigTokens), P(SigSimParams), i, i, i, vp, vp, vp, sz, u, i, vp]
    lib.sig_sim_bwd.argtypes = [P(SigTokens), P(SigSimParams), vp, P(SigTokenGrads), P(SigSimParamGrads), vp, sz, u, i, vp]
    lib.sig_sim_fold_selection.argtypes = [P(SigSimParams), i, vp, vp, vp, vp, vp, sz, i, vp]
    lib.sig_sim_select_fwd.argtypes = [P(SigTokens), P(SigSimParams), i, i, i, i, vp, vp, vp, sz, i, vp]
    lib.sig_sim_select_from_scores.argtypes = [vp, vp, vp, i, i, i, i, i, i, vp, i, vp]
    lib.sig_mask_mul_bwd.argtypes = [vp, vp, i, i, i, i, P(SigTokenGrads), i, vp]
    lib.sig_sim_attn_fwd.argtypes = [P(SigTokens), P(SigSimParams), vp, vp, vp, sz, u, i, vp]
    lib.sig_sim_attn_bwd.argtypes = [P(SigTokens), P(SigSimParams), vp, vp, P(SigTokenGrads), P(SigSimParamGrads), vp, sz, u, i, vp]
    lib.sig_align_fwd.argtypes = [P(SigTokens), P(SigAlignParams), i, i, i, vp, vp, sz, u, i, vp]
    lib.sig_align_bwd.argtypes = [P(SigTokens), P(SigAlignParams), i, i, i, vp, P(SigTokenGrads), P(SigAlignParamGrads), vp, sz, u, i, vp]
    lib.sig_das_fwd.argtypes = [vp, i64, i64, i, i, i, i, i, P(SigAlignParams), i, vp, vp, sz, u, i, vp]
    lib.sig_das_bwd.argtypes = [vp, i64, i64, i, i, i, i, i, P(SigAlignParams), i, vp, vp, P(SigAlignParamGrads), vp, sz, u, i, vp]
    lib.sig_volume3_ws_bytes.restype = sz
    lib.sig_volume3_ws_bytes.argtypes = [i, i]
    lib.sig_volume3_fwd.argtypes = [vp, vp, vp, i, i, i, vp, vp, sz, i, vp]
    lib.sig_volume3_bwd.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, sz, i, vp]
    lib.sig_debug_gemm_bf16.argtypes = [vp, i, P(i64), vp, i, P(i64), vp, i64, i, vp, i, i, i, C.c_float, i, i, i, i64, i64, vp, vp, i, i, vp]
    lib.sig_sim_dx_operands.argtypes = [vp, i, i, i, i, u, P(vp), P(vp)]
    lib.sig_align_patch_mean_slot.argtypes = [vp, i, i, i, i, u, P(vp)]
    lib.sig_infer_features.argtypes = [_VP3, _I64x3, vp, i64, i, i, i, i, vp, i, vp]
    lib.sig_euclidean_distmat.argtypes = [vp, vp, i, i, i, vp, vp, sz, i, vp]
    lib.sig_rank_eval.argtypes = [vp, i64, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp, i, vp]
    lib.sig_xchg_flag_bytes.restype = sz
    lib.sig_xchg_flag_bytes.argtypes = []
    lib.sig_xchg_allreduce_f32.argtypes = [P(SigXchgPeers), sz, sz, C.c_float, i, i, vp]
    lib.sig_convert_half.argtypes = [vp, i64, i64, i, vp, i64, i64, i, i, i, i, i, vp]
    lib.sig_profile_timeline.argtypes = [C.c_char_p, sz]
    f = C.c_float
    lib.sig_loss_ws_bytes.restype = sz
    lib.sig_loss_ws_bytes.argtypes = [i]
    lib.sig_xent_ls_fwd.argtypes = [vp, i, i64, vp, i, i, f, vp, vp, vp, sz, i, vp]
    lib.sig_xent_ls_bwd.argtypes = [vp, i, i64, vp, i, i, f, vp, vp, vp, i64, i, vp]
    lib.sig_triplet_fwd.argtypes = [vp, i, i64, vp, i, i, f, i, f, vp, vp, vp, vp, vp, vp, sz, i, vp]
    lib.sig_triplet_bwd.argtypes = [vp, i, i64, i, i, f, i, f, vp, vp, vp, vp, vp, vp, vp, vp, i64, i, vp]
    lib.sig_bnneck_ws_bytes.restype = sz
    lib.sig_bnneck_ws_bytes.argtypes = [i, i, i]
    lib.sig_bnneck_cls_fwd.argtypes = [vp, i, i64, i, i, i, vp, vp, vp, vp, f, f, i, vp, vp, i64, vp, i64, vp, vp, vp, vp, sz, i, vp]
    lib.sig_bnneck_cls_bwd.argtypes = [vp, i, i64, i, i, i, vp, vp, vp, vp, i, vp, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, sz, i, vp]
    lib.sig_volume_n_ws_bytes.restype = sz
    lib.sig_volume_n_ws_bytes.argtypes = [i, i, i]
    lib.sig_volume_n_fwd.argtypes = [i, P(vp), i, i, i, vp, vp, sz, i, vp]
    lib.sig_volume_n_bwd.argtypes = [i, P(vp), i, i, i, vp, P(vp), vp, sz, i, vp]
    lib.sig_tokens_ws_bytes.restype = sz
    lib.sig_tokens_ws_bytes.argtypes = [i, i, i, i, i, i]
    lib.sig_tokens_fwd.argtypes = [vp, i, i64, i64, i, i, i, i, vp, vp, f, vp, vp, i, vp, vp, sz, vp, sz, i, vp]
    lib.sig_tokens_bwd.argtypes = [vp, i, i64, i64, i, i, i, i, vp, vp, vp, i, i64, i64, vp, sz, vp, i64, i64, vp, vp, vp, vp, sz, i, vp]
    lib.sig_profile_scope_begin.restype = vp
    lib.sig_profile_scope_begin.argtypes = [C.c_char_p, vp]
    lib.sig_profile_scope_end.restype = None
    lib.sig_profile_scope_end.argtypes = [vp]
    lib.sig_debug_tc_stamps.argtypes = [P(C.c_longlong)]
    lib.sig_debug_launch_count.restype = C.c_ulonglong
    lib.sig_profile_enable.argtypes = [i]
    lib.sig_profile_collect.argtypes = [C.c_char_p, sz, P(C.c_float), P(C.c_int), i]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("sig_error_string", "sig_ctx_bytes", "sig_volume3_ws_bytes", "sig_debug_launch_count", "sig_xchg_flag_bytes",
                        "sig_loss_ws_bytes", "sig_bnneck_ws_bytes", "sig_tokens_ws_bytes", "sig_volume_n_ws_bytes", "sig_profile_scope_begin", "sig_profile_scope_end"):
            fn.restype = i
    if lib.sig_version() != 1:
        raise RuntimeError("signal_b200: ABI version mismatch between lib.py and libsignal_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sig_error_string(rc).decode()
        raise RuntimeError(f"signal_b200.{what} failed (code {rc}): {msg}")


def dtype_enum(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return SIG_F32
    if t.dtype == torch.bfloat16:
        return SIG_BF16
    raise RuntimeError(f"signal_b200: unsupported kernel dtype {t.dtype} (fp32 or bf16; the nn.Module shims bridge fp16 "
                       "token maps through functional.HalfBridge)")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"signal_b200: {what} must be a CUDA tensor (there is no CPU path)")


def tokens_struct(patches: Sequence[torch.Tensor], cls: Optional[Sequence[torch.Tensor]]) -> SigTokens:
    """Describe three strided [B,L,d] patch views (+ optional [B,d] CLS views) without copying."""
    t = SigTokens()
    B, L, d = patches[0].shape
    for m in range(3):
        p = patches[m]
        _require_cuda(p, "patch tokens")
        if p.shape != (B, L, d) or p.dtype != patches[0].dtype or p.stride(2) != 1:
            raise RuntimeError("signal_b200: patch maps must share shape/dtype and have unit channel stride")
        t.patch[m] = p.data_ptr()
        t.patch_stride_b[m] = p.stride(0)
        t.patch_stride_l[m] = p.stride(1)
        if cls is not None:
            g = cls[m]
            _require_cuda(g, "CLS tokens")
            if g.shape != (B, d) or g.dtype != p.dtype or g.stride(1) != 1:
                raise RuntimeError("signal_b200: CLS tokens must be [B,d] views with unit channel stride")
            t.cls[m] = g.data_ptr()
            t.cls_stride_b[m] = g.stride(0)
    t.dtype = dtype_enum(patches[0])
    t.B, t.L, t.d = B, L, d
    return t


def token_grads_struct(dpatch: Sequence[torch.Tensor], dcls: Optional[Sequence[torch.Tensor]],
                       accumulate: bool = False, zero_cls: bool = False, wait_event: Optional[int] = None,
                       done_event: Optional[int] = None, fuse_skip_dx: bool = False, fuse_ops=None) -> SigTokenGrads:
    g = SigTokenGrads()
    for m in range(3):
        g.dpatch[m] = dpatch[m].data_ptr()
        g.patch_stride_b[m] = dpatch[m].stride(0)
        g.patch_stride_l[m] = dpatch[m].stride(1)
        if dcls is not None:
            g.dcls[m] = dcls[m].data_ptr()
            g.cls_stride_b[m] = dcls[m].stride(0)
    g.accumulate = int(accumulate)
    g.zero_cls = int(zero_cls)
    g.wait_event = wait_event
    g.done_event = done_event
    g.fuse_skip_dx = int(fuse_skip_dx)
    if fuse_ops is not None:
        g.fuse_pds, g.fuse_dxqt = fuse_ops
    return g


def sim_dx_operands(ctx_buf: torch.Tensor, B: int, L: int, d: int, dtype: int, flags: int):
    """(pds, dxqt) device pointers inside a SIM ctx buffer, or None when this configuration does not run on
    the bf16 tensor-core path (see sig_token_grads.fuse_* in include/signal_b200.h)."""
    a, b = C.c_void_p(), C.c_void_p()
    rc = load().sig_sim_dx_operands(ctx_buf.data_ptr(), B, L, d, dtype, flags, C.byref(a), C.byref(b))
    return (a.value, b.value) if rc == 0 else None


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise RuntimeError("signal_b200: parameters must be contiguous fp32 CUDA tensors (fp32 masters)")
    return t


def sim_params_struct(params: Sequence[torch.Tensor], fold: Optional[Sequence[torch.Tensor]] = None) -> SigSimParams:
    """params in SIM_PARAM_FIELDS order (16 tensors); fold = optional (m_hl, v, u, s0) selection cache."""
    s = SigSimParams()
    for name, t in zip(SIM_PARAM_FIELDS, params):
        setattr(s, name, _f32c(t).data_ptr())
    if fold is not None:
        f = SigSelFold(*(t.data_ptr() for t in fold))
        s._fold_keepalive = f
        s.sel_fold = C.pointer(f)
    return s


def fold_selection(sel_params: Sequence[torch.Tensor]):
    """(m_hl bf16 [d,2d], v, u, s0) for the four frozen token_selection tensors (W_q.w, W_q.b, W_k.w, W_k.b)."""
    lib = load()
    wq = sel_params[0]
    d, dev = wq.shape[0], wq.device
    prm = sim_params_struct([p.detach() for p in sel_params] + [sel_params[0].detach()] * 12)
    m_hl = torch.empty(d, 2 * d, dtype=torch.bfloat16, device=dev)
    v = torch.empty(d, dtype=torch.float32, device=dev)
    u = torch.empty(d, dtype=torch.float32, device=dev)
    s0 = torch.empty(1, dtype=torch.float32, device=dev)
    ws = torch.empty(d * d, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.sig_sim_fold_selection(C.byref(prm), d, m_hl.data_ptr(), v.data_ptr(), u.data_ptr(), s0.data_ptr(),
                                         ws.data_ptr(), ws.numel() * 4, dev.index, stream_ptr(dev)), "sig_sim_fold_selection")
    return m_hl, v, u, s0


def sim_grads_struct(grads: Sequence[torch.Tensor]) -> SigSimParamGrads:
    s = SigSimParamGrads()
    for name, t in zip(SIM_GRAD_FIELDS, grads):
        setattr(s, name, t.data_ptr())
    return s


def align_params_struct(contra_temp: torch.Tensor, mods: Sequence[Sequence[torch.Tensor]], cls=SigAlignParams):
    """mods[m] = 7 tensors in ALIGN_MOD_FIELDS order for modality m (r, n, t); entries may be None."""
    s = cls()
    s.contra_temp = contra_temp.data_ptr() if contra_temp is not None else None
    for m, mod in enumerate(mods):
        if mod is None:
            continue
        for name, t in zip(ALIGN_MOD_FIELDS, mod):
            getattr(s, name)[m] = t.data_ptr()
    return s


def ctx_bytes(kind: int, B: int, L: int, d: int, dtype: int = SIG_F32, flags: int = 0) -> int:
    n = load().sig_ctx_bytes(kind, B, L, d, dtype, flags)
    if n == 0:
        raise RuntimeError(f"signal_b200: unsupported shape B={B} L={L} d={d} (L <= 128, d % 64 == 0)")
    return n


def launch_count() -> int:
    """Kernel launches enqueued by the library since it was loaded."""
    return int(load().sig_debug_launch_count())


def profile_enable(on: bool):
    load().sig_profile_enable(int(on))


def profile_collect():
    """-> {phase: (summed ms, scopes)} for everything recorded since the last collect."""
    lib = load()
    names = C.create_string_buffer(4096)
    ms = (C.c_float * 64)()
    cnt = (C.c_int * 64)()
    n = lib.sig_profile_collect(names, 4096, ms, cnt, 64)
    keys = names.value.decode().split("\n")[:n]
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(keys)}


def debug_gemm_bf16(A, a_mode, B, b_mode, M, N, K, out_bf16=False, bias=None, alpha=1.0, act=0, ksplit=1, bn=128,
                    out=None, rowvec=None, rowvec_scale=None, accumulate=False):
    """Unit-test seam: run the tcgen05 GEMM core on bf16 CUDA tensors, returns C [M,N].
    out: optional [B,128,N] (possibly strided) destination for the token-row epilogue."""
    lib = load()

    def geom(t, mode):
        if mode in (0, 2):
            return (C.c_int64 * 5)(t.stride(0), 0, 0, t.shape[0], t.shape[1])
        return (C.c_int64 * 5)(0, t.stride(0), t.stride(1), t.shape[0], t.shape[2])

    dev = A.device
    csb = csl = 0
    if out is None:
        out = (torch.zeros if ksplit > 1 else torch.empty)(M, N, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
    else:
        csb, csl = out.stride(0), out.stride(1)
        out_bf16 = out.dtype == torch.bfloat16
    with torch.cuda.device(dev):
        check(lib.sig_debug_gemm_bf16(A.data_ptr(), a_mode, geom(A, a_mode), B.data_ptr(), b_mode, geom(B, b_mode),
                                      out.data_ptr(), N, int(out_bf16), None if bias is None else bias.data_ptr(), M, N, K,
                                      alpha, act, ksplit, bn, csb, csl, None if rowvec is None else rowvec.data_ptr(),
                                      None if rowvec_scale is None else rowvec_scale.data_ptr(), int(accumulate),
                                      dev.index, stream_ptr(dev)), "sig_debug_gemm_bf16")
    return out
