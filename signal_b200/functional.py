"""torch.autograd.Function wrappers over the C ABI (signal_b200.lib).

Host-side plumbing only: allocate outputs / saved-state buffers with torch, pass raw
pointers + strides to libsignal_b200.so on the current CUDA stream, wire gradients.
Every Function has a *packed* mode whose token inputs are the three [B,1+L,d] token maps
(CLS = row 0, patches = rows 1.., modeling/meta_arch.py:108-109): the backward kernels
then write one [B,1+L,d] gradient per modality directly (no per-view zero-fill + add).
"""
from __future__ import annotations

import os

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import lib as L_

_SIM_GRAD_SHAPES = lambda d: [(3 * d, d), (3 * d,), (d, d), (d,), (2 * d, d), (2 * d,), (d, 2 * d), (d,), (d,), (d,), (d,), (d,)]


def _carve(flat: torch.Tensor, shapes) -> List[torch.Tensor]:
    out, off = [], 0
    for s in shapes:
        n = 1
        for x in s:
            n *= x
        out.append(flat[off:off + n].view(s))
        off += (n + 3) // 4 * 4   # keep every view 16-byte aligned
    return out


def _arena(shapes, device, flat: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    total = sum(((int(torch.Size(s).numel()) + 3) // 4 * 4) for s in shapes)
    if flat is None:
        flat = torch.empty(total, dtype=torch.float32, device=device)
    else:        # caller-owned arena (e.g. symmetric memory of parallel.GradExchange)
        if flat.numel() < total or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise RuntimeError(f"signal_b200: the gradient arena must be a contiguous fp32 tensor of >= {total} elements")
        flat = flat[:total]
    return flat, _carve(flat, shapes)


def head_grad_numel(d: int) -> int:
    """fp32 elements of the flat gradient arena of SIM + AlignM at width d (FusionHead.grad_numel)."""
    return sum(((int(torch.Size(s).numel()) + 3) // 4 * 4) for s in _SIM_GRAD_SHAPES(d) + _align_grad_shapes(d))


def _head_arena(d: int, device, flat: Optional[torch.Tensor] = None):
    """ONE flat arena for the gradients of both modules, ``[SIM late part | SIM early part | AlignM]``, so that a
    data-parallel exchange needs two collectives: ``flat[cut:]`` (28.5 MB at d = 768) once SIM's early gradients and
    AlignM's gradients are final -- both happen in the first third of the backward when AlignM's weight-independent
    chain already ran in the forward call (SIG_FLAG_EAGER_BWD) -- and ``flat[:cut]`` (W_q/W_k/in_proj_bias, 4.7 MB)
    when SIM's token-side backward is through.  Returns (flat, SIM grads, AlignM grads, cut)."""
    sshapes, ashapes = _SIM_GRAD_SHAPES(d), _align_grad_shapes(d)
    order = [1, 0] + list(range(2, len(sshapes)))
    flat, views = _arena([sshapes[i] for i in order] + ashapes, device, flat)
    sv, pg_a = views[:len(sshapes)], views[len(sshapes):]
    pg_s = [None] * len(sshapes)
    for i, v in zip(order, sv):
        pg_s[i] = v
    cut = (3 * d + 3) // 4 * 4 + 2 * d * d
    return flat, pg_s, pg_a, cut, sum(((int(torch.Size(s).numel()) + 3) // 4 * 4) for s in sshapes)


def _sim_arena(d: int, device):
    """SIM's parameter-gradient arena laid out so that the gradients which are final LAST (in_proj_bias and
    in_proj_weight rows [0, 2d) = W_q, W_k: they need the token-side backward) are one contiguous prefix
    ``flat[:split]``; everything in ``flat[split:]`` is final when sig_sim_param_grads.early_event fires.
    Returns (flat, grads in SIM_GRAD_FIELDS order, split)."""
    shapes = _SIM_GRAD_SHAPES(d)
    order = [1, 0] + list(range(2, len(shapes)))          # in_proj_b first, then in_proj_w, then the rest
    flat, views = _arena([shapes[i] for i in order], device)
    pg = [None] * len(shapes)
    for i, v in zip(order, views):
        pg[i] = v
    split = (3 * d + 3) // 4 * 4 + 2 * d * d
    return flat, pg, split


def _split_tokens(packed: bool, toks: Sequence[torch.Tensor]):
    if packed:
        for t in toks:
            if t.dim() != 3 or not t.is_contiguous():
                raise RuntimeError("signal_b200: packed token maps must be contiguous [B,1+L,d]")
        return [t[:, 1:] for t in toks], [t[:, 0] for t in toks]
    return list(toks[:3]), list(toks[3:6])


def _alloc_token_grads(packed: bool, toks, need_cls: bool):
    """Returns (returned grads, dpatch views, dcls views or None)."""
    if packed:
        full = [torch.empty_like(t) for t in toks]
        return full, [t[:, 1:] for t in full], [t[:, 0] for t in full]
    B, L, d = toks[0].shape
    dp = [torch.empty(B, L, d, dtype=toks[0].dtype, device=toks[0].device) for _ in range(3)]
    dc = [torch.empty(B, d, dtype=toks[0].dtype, device=toks[0].device) for _ in range(3)] if need_cls else None
    return dp + (dc if dc else []), dp, dc


# ------------------------------------------------------------------------------------------
# fp16 boundary
# ------------------------------------------------------------------------------------------
def _convert_half(src: torch.Tensor, dst_dtype: torch.dtype) -> torch.Tensor:
    """fp16 <-> bf16 copy of a [.., d] tensor with unit channel stride (2-D [B,d] or 3-D [B,L,d] views are read through
    their strides, no .contiguous()) -> contiguous tensor of dst_dtype.  One streaming kernel (sig_convert_half)."""
    lib = L_.load()
    code = {torch.float16: L_.SIG_F16, torch.bfloat16: L_.SIG_BF16}
    x = src if src.dim() == 3 else src.unsqueeze(1)
    if x.dim() != 3 or x.stride(2) != 1 or (x.stride(0) | x.stride(1)) % 8 or x.data_ptr() % 16:
        x = x.reshape(-1, 1, src.shape[-1]).contiguous() if src.dim() != 3 else x.contiguous()
    nb, nl, d = x.shape
    dst = torch.empty(nb, nl, d, dtype=dst_dtype, device=x.device)
    with torch.cuda.device(x.device):
        L_.check(lib.sig_convert_half(x.data_ptr(), x.stride(0), x.stride(1), code[x.dtype], dst.data_ptr(), dst.stride(0),
                                      dst.stride(1), code[dst_dtype], nb, nl, d, x.device.index, L_.stream_ptr(x.device)),
                 "sig_convert_half")
    return dst.view(src.shape)


class HalfBridge(torch.autograd.Function):
    """y = x converted between fp16 and bf16; the backward converts the gradient the other way.

    The reference trains under fp16 autocast (engine/processor.py:165), so SIM / AlignM receive fp16 token maps; the
    kernels compute on bf16 operands with fp32 accumulation.  The module shims therefore convert each token map once
    on the way in (fp16 -> bf16, exact exponent, mantissa rounded to 8 bits) and each result / token gradient once on
    the way out (bf16 -> fp16)."""

    @staticmethod
    def forward(ctx, x, dst_dtype):
        ctx.src_dtype = x.dtype
        return _convert_half(x, dst_dtype)

    @staticmethod
    def backward(ctx, g):
        return _convert_half(g, ctx.src_dtype), None


# ------------------------------------------------------------------------------------------
# SIM
# ------------------------------------------------------------------------------------------
class SimFunction(torch.autograd.Function):
    """Select_Interactive_Module.forward (useA.py:454-476) -> (out [B,3d], masks fp32 [3,B,L])."""

    @staticmethod
    def forward(ctx, packed: bool, k1: int, k2: int, max_keep: int, flags: int, *tensors):
        # tensors: 3 (packed) or 6 token tensors, 16 parameters, then optionally the 4 tensors of the
        # folded-selection cache (lib.fold_selection)
        ntok = 3 if packed else 6
        toks, params, fold = tensors[:ntok], tensors[ntok:ntok + 16], tensors[ntok + 16:]
        lib = L_.load()
        patches, cls = _split_tokens(packed, toks)
        B, L, d = patches[0].shape
        dev = patches[0].device
        tok = L_.tokens_struct(patches, cls)
        prm = L_.sim_params_struct(params, fold if len(fold) == 4 else None)
        ctx.nfold = len(fold)
        out = torch.empty(B, 3 * d, dtype=patches[0].dtype, device=dev)
        masks = torch.empty(3, B, L, dtype=torch.float32, device=dev)
        nbytes = L_.ctx_bytes(L_.CTX_SIM, B, L, d, L_.dtype_enum(patches[0]), flags)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_sim_fwd(C.byref(tok), C.byref(prm), k1, k2, max_keep, out.data_ptr(), masks.data_ptr(),
                                     buf.data_ptr(), nbytes, flags, dev.index, L_.stream_ptr(dev)), "sig_sim_fwd")
        ctx.save_for_backward(*toks, *params, buf)
        ctx.packed, ctx.ntok, ctx.flags = packed, ntok, flags
        ctx.mark_non_differentiable(masks)
        return out, masks

    @staticmethod
    def backward(ctx, dout, _dmasks):
        saved = ctx.saved_tensors
        toks, params, buf = saved[:ctx.ntok], saved[ctx.ntok:-1], saved[-1]
        lib = L_.load()
        patches, cls = _split_tokens(ctx.packed, toks)
        B, L, d = patches[0].shape
        dev = patches[0].device
        dout = dout.to(patches[0].dtype).contiguous()
        ret, dpatch, dcls = _alloc_token_grads(ctx.packed, toks, True)
        _, pg = _arena(_SIM_GRAD_SHAPES(d), dev)
        tok = L_.tokens_struct(patches, cls)
        prm = L_.sim_params_struct(params)
        tg = L_.token_grads_struct(dpatch, dcls)
        gs = L_.sim_grads_struct(pg)
        with torch.cuda.device(dev):
            L_.check(lib.sig_sim_bwd(C.byref(tok), C.byref(prm), dout.data_ptr(), C.byref(tg), C.byref(gs), buf.data_ptr(),
                                     buf.numel(), ctx.flags, dev.index, L_.stream_ptr(dev)), "sig_sim_bwd")
        return (None,) * 5 + tuple(ret) + (None,) * 4 + tuple(pg) + (None,) * ctx.nfold


class AttnFunction(torch.autograd.Function):
    """ModalInteractive.forward (useA.py:364-411) on arbitrary K/V token maps (no masks)."""

    @staticmethod
    def forward(ctx, flags: int, *tensors):
        toks, params = tensors[:6], tensors[6:]
        lib = L_.load()
        patches, cls = list(toks[:3]), list(toks[3:6])
        B, L, d = patches[0].shape
        dev = patches[0].device
        tok = L_.tokens_struct(patches, cls)
        prm = L_.sim_params_struct(params)
        out = torch.empty(B, 3 * d, dtype=patches[0].dtype, device=dev)
        nbytes = L_.ctx_bytes(L_.CTX_SIM, B, L, d, L_.dtype_enum(patches[0]), flags)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_sim_attn_fwd(C.byref(tok), C.byref(prm), None, out.data_ptr(), buf.data_ptr(), nbytes, flags,
                                          dev.index, L_.stream_ptr(dev)), "sig_sim_attn_fwd")
        ctx.save_for_backward(*toks, *params, buf)
        ctx.flags = flags
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        toks, params, buf = saved[:6], saved[6:-1], saved[-1]
        lib = L_.load()
        patches, cls = list(toks[:3]), list(toks[3:6])
        B, L, d = patches[0].shape
        dev = patches[0].device
        dout = dout.to(patches[0].dtype).contiguous()
        ret, dpatch, dcls = _alloc_token_grads(False, toks, True)
        _, pg = _arena(_SIM_GRAD_SHAPES(d), dev)
        tok = L_.tokens_struct(patches, cls)
        prm = L_.sim_params_struct(params)
        tg = L_.token_grads_struct(dpatch, dcls)
        gs = L_.sim_grads_struct(pg)
        with torch.cuda.device(dev):
            L_.check(lib.sig_sim_attn_bwd(C.byref(tok), C.byref(prm), None, dout.data_ptr(), C.byref(tg), C.byref(gs),
                                          buf.data_ptr(), buf.numel(), ctx.flags, dev.index, L_.stream_ptr(dev)),
                     "sig_sim_attn_bwd")
        return (None,) + tuple(ret) + (None,) * 4 + tuple(pg)


class SelectFunction(torch.autograd.Function):
    """TokenSelection.forward (useA.py:223-325) -> (selected x3 [B,L,d], masks fp32 [3,B,L])."""

    @staticmethod
    def forward(ctx, k1: int, k2: int, max_keep: int, *tensors):
        toks, params = tensors[:6], tensors[6:]
        lib = L_.load()
        patches, cls = list(toks[:3]), list(toks[3:6])
        B, L, d = patches[0].shape
        dev = patches[0].device
        tok = L_.tokens_struct(patches, cls)
        prm = L_.sim_params_struct(list(params) + [params[0]] * 12)   # only the four selection tensors are read
        masks = torch.empty(3, B, L, dtype=torch.float32, device=dev)
        selected = torch.empty(3, B, L, d, dtype=patches[0].dtype, device=dev)
        nbytes = L_.ctx_bytes(L_.CTX_SELECT, B, L, d, L_.dtype_enum(patches[0]), 0)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_sim_select_fwd(C.byref(tok), C.byref(prm), 3, k1, k2, max_keep, masks.data_ptr(),
                                            selected.data_ptr(), buf.data_ptr(), nbytes, dev.index, L_.stream_ptr(dev)),
                     "sig_sim_select_fwd")
        ctx.save_for_backward(masks)
        ctx.dtype = patches[0].dtype
        ctx.mark_non_differentiable(masks)
        return selected[0], selected[1], selected[2], masks

    @staticmethod
    def backward(ctx, d0, d1, d2, _dm):
        (masks,) = ctx.saved_tensors
        lib = L_.load()
        _, B, L = masks.shape
        dev = masks.device
        d = next(g.shape[-1] for g in (d0, d1, d2) if g is not None)
        dsel = torch.stack([torch.zeros(B, L, d, dtype=ctx.dtype, device=dev) if g is None else g.to(ctx.dtype)
                            for g in (d0, d1, d2)]).contiguous()
        dpatch = [torch.empty(B, L, d, dtype=ctx.dtype, device=dev) for _ in range(3)]
        tg = L_.token_grads_struct(dpatch, None)
        with torch.cuda.device(dev):
            L_.check(lib.sig_mask_mul_bwd(dsel.data_ptr(), masks.data_ptr(), L_.SIG_BF16 if ctx.dtype == torch.bfloat16 else L_.SIG_F32,
                                          B, L, d, C.byref(tg), dev.index, L_.stream_ptr(dev)), "sig_mask_mul_bwd")
        return (None,) * 3 + tuple(dpatch) + (None,) * 7


def select_masks(which: int, patches, cls, sel_params, k1: int, k2: int, max_keep: int = -1) -> torch.Tensor:
    """intra (1) / inter (2) / union (3) selection masks, fp32 [3,B,L]; not differentiable."""
    lib = L_.load()
    patches = [p.detach() for p in patches]
    cls = [c.detach() for c in cls]
    B, L, d = patches[0].shape
    dev = patches[0].device
    tok = L_.tokens_struct(patches, cls)
    sel = [p.detach() for p in sel_params]
    prm = L_.sim_params_struct(sel + [sel[0]] * 12)
    masks = torch.empty(3, B, L, dtype=torch.float32, device=dev)
    nbytes = L_.ctx_bytes(L_.CTX_SELECT, B, L, d, L_.dtype_enum(patches[0]), 0)
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L_.check(lib.sig_sim_select_fwd(C.byref(tok), C.byref(prm), which, k1, k2, max_keep, masks.data_ptr(), None,
                                        buf.data_ptr(), nbytes, dev.index, L_.stream_ptr(dev)), "sig_sim_select_fwd")
    return masks


def select_from_scores(intra: Optional[torch.Tensor], inter: Optional[torch.Tensor], raw: Optional[torch.Tensor],
                       which: int, k1: int, k2: int, max_keep: int = -1) -> torch.Tensor:
    """Test seam: rank-select on caller-supplied fp32 scores ([3,B,L], [3,B,2L], [3,B,L])."""
    lib = L_.load()
    ref = intra if intra is not None else inter
    B = ref.shape[1]
    L = intra.shape[2] if intra is not None else inter.shape[2] // 2
    dev = ref.device
    keep = [t.contiguous().float() if t is not None else None for t in (intra, inter, raw)]
    masks = torch.empty(3, B, L, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L_.check(lib.sig_sim_select_from_scores(*(None if t is None else t.data_ptr() for t in keep), B, L, which, k1, k2,
                                                max_keep, masks.data_ptr(), dev.index, L_.stream_ptr(dev)),
                 "sig_sim_select_from_scores")
    return masks


# ------------------------------------------------------------------------------------------
# AlignmentM
# ------------------------------------------------------------------------------------------
def _align_grad_shapes(d):
    one = [(d, d, 1, 1), (d,), (d, d, 1, 1), (d,), (d, 1, 4, 4), (d,), (1, d, 1, 1)]
    return [()] + one * 3


def _deposit_patch_mean(buf, hint, B, L, d, dt, flags, stream=None):
    """SIG_FLAG_PATCH_MEAN: copy the caller's patch means (three [B,d] fp32 tensors or one [3,B,d], e.g.
    TokenProducer.last_patch_mean of the three modalities) into the slot of AlignM's ctx buffer.  Returns the flag bit, or 0
    when this configuration does not run on the tensor-core path (the pool pass then runs as usual)."""
    if hint is None:
        return 0
    slot = C.c_void_p()
    if L_.load().sig_align_patch_mean_slot(buf.data_ptr(), B, L, d, dt, flags, C.byref(slot)) != 0:
        return 0
    off = slot.value - buf.data_ptr()
    dst = buf[off: off + 3 * B * d * 4].view(torch.float32).view(3, B, d)
    src = hint if torch.is_tensor(hint) else torch.stack([h.detach() for h in hint])
    if tuple(src.shape) != (3, B, d) or src.dtype != torch.float32:
        raise RuntimeError("signal_b200: patch_mean hint must be fp32 [3,B,d] (or three [B,d] tensors)")
    if stream is None:
        dst.copy_(src.detach())
    else:
        with torch.cuda.stream(stream):
            dst.copy_(src.detach())
    return L_.SIG_FLAG_PATCH_MEAN


class AlignFunction(torch.autograd.Function):
    """AlignmentM.forward (useB.py:169-190) -> (gam, lam) 0-dim fp32 (lam = 0 when do_lam is False).

    tensors: 3 token maps (packed) or 3 patch maps, then contra_temp, then 7 tensors per
    modality in lib.ALIGN_MOD_FIELDS order (r, n, t), then optionally the fp32 [3,B,d] patch means (forward hint,
    not differentiated: the token gradient of the pool still flows through the token maps).
    """

    @staticmethod
    def forward(ctx, packed: bool, h: int, w: int, do_lam: bool, flags: int, *tensors):
        toks, params, hint = tensors[:3], tensors[3:25], (tensors[25] if len(tensors) > 25 else None)
        ctx.has_hint = hint is not None
        lib = L_.load()
        patches = [t[:, 1:] for t in toks] if packed else list(toks)
        B, L, d = patches[0].shape
        dev = patches[0].device
        tok = L_.tokens_struct(patches, None)
        mods = [params[1 + 7 * m: 8 + 7 * m] for m in range(3)]
        for mod in mods:
            for t in mod:
                L_._f32c(t)
        prm = L_.align_params_struct(L_._f32c(params[0]), mods)
        losses = torch.zeros(2, dtype=torch.float32, device=dev)
        nbytes = L_.ctx_bytes(L_.CTX_ALIGN, B, L, d, L_.dtype_enum(patches[0]), flags)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            fwd_flags = flags | _deposit_patch_mean(buf, hint, B, L, d, L_.dtype_enum(patches[0]), flags)
            L_.check(lib.sig_align_fwd(C.byref(tok), C.byref(prm), h, w, int(do_lam), losses.data_ptr(), buf.data_ptr(), nbytes,
                                       fwd_flags, dev.index, L_.stream_ptr(dev)), "sig_align_fwd")
        ctx.save_for_backward(*toks, *params, buf)
        ctx.cfg = (packed, h, w, do_lam, flags)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, dgam, dlam):
        packed, h, w, do_lam, flags = ctx.cfg
        saved = ctx.saved_tensors
        toks, params, buf = saved[:3], saved[3:-1], saved[-1]
        lib = L_.load()
        patches = [t[:, 1:] for t in toks] if packed else list(toks)
        B, L, d = patches[0].shape
        dev = patches[0].device
        z = torch.zeros((), dtype=torch.float32, device=dev)
        dl = torch.stack([z if dgam is None else dgam.float(), z if dlam is None else dlam.float()]).contiguous()
        if packed:
            ret = [torch.empty_like(t) for t in toks]
            dpatch, dcls = [t[:, 1:] for t in ret], [t[:, 0] for t in ret]
        else:
            ret = [torch.empty(B, L, d, dtype=toks[0].dtype, device=dev) for _ in range(3)]
            dpatch, dcls = ret, None
        flat, pg = _arena(_align_grad_shapes(d), dev)
        if not do_lam:
            flat.zero_()
        tok = L_.tokens_struct(patches, None)
        mods = [params[1 + 7 * m: 8 + 7 * m] for m in range(3)]
        prm = L_.align_params_struct(params[0], mods)
        gmods = [pg[1 + 7 * m: 8 + 7 * m] for m in range(3)]
        gs = L_.align_params_struct(pg[0], gmods, cls=L_.SigAlignParamGrads)
        tg = L_.token_grads_struct(dpatch, dcls, accumulate=False, zero_cls=packed)
        with torch.cuda.device(dev):
            L_.check(lib.sig_align_bwd(C.byref(tok), C.byref(prm), h, w, int(do_lam), dl.data_ptr(), C.byref(tg), C.byref(gs),
                                       buf.data_ptr(), buf.numel(), flags, dev.index, L_.stream_ptr(dev)), "sig_align_bwd")
        if not do_lam:      # stage == "CLS": the DAS parameters took no part (useB.py:181-183) -> no gradient, like the reference
            pg = [pg[0]] + [None] * (len(pg) - 1)
        return (None,) * 5 + tuple(ret) + tuple(pg) + ((None,) if ctx.has_hint else ())


class DasFunction(torch.autograd.Function):
    """DA_sample.forward (DAS.py:107-165) on a channels-last [B,h*w,d] view -> sampled fp32 [B,Hk*Wk,d].

    params: the 7 tensors of one modality in lib.ALIGN_MOD_FIELDS order.
    """

    @staticmethod
    def forward(ctx, h: int, w: int, flags: int, x, *params):
        lib = L_.load()
        B, L, d = x.shape
        dev = x.device
        if x.stride(2) != 1:
            raise RuntimeError("signal_b200: DA_sample needs a channels-last input view")
        for t in params:
            L_._f32c(t)
        prm = L_.align_params_struct(None, [params])
        P = (h // 4) * (w // 4)
        sampled = torch.empty(B, P, d, dtype=torch.float32, device=dev)
        nbytes = L_.ctx_bytes(L_.CTX_DAS, B, L, d, L_.dtype_enum(x), flags)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_das_fwd(x.data_ptr(), x.stride(0), x.stride(1), L_.dtype_enum(x), B, h, w, d, C.byref(prm), 0,
                                     sampled.data_ptr(), buf.data_ptr(), nbytes, flags, dev.index, L_.stream_ptr(dev)),
                     "sig_das_fwd")
        ctx.save_for_backward(x, *params, buf)
        ctx.cfg = (h, w, flags)
        return sampled

    @staticmethod
    def backward(ctx, dsampled):
        h, w, flags = ctx.cfg
        saved = ctx.saved_tensors
        x, params, buf = saved[0], saved[1:-1], saved[-1]
        lib = L_.load()
        B, L, d = x.shape
        dev = x.device
        dsampled = dsampled.float().contiguous()
        dx = torch.empty(B, L, d, dtype=x.dtype, device=dev)
        shapes = [(d, d, 1, 1), (d,), (d, d, 1, 1), (d,), (d, 1, 4, 4), (d,), (1, d, 1, 1)]
        _, pg = _arena(shapes, dev)
        prm = L_.align_params_struct(None, [params])
        gs = L_.align_params_struct(None, [pg], cls=L_.SigAlignParamGrads)
        with torch.cuda.device(dev):
            L_.check(lib.sig_das_bwd(x.data_ptr(), x.stride(0), x.stride(1), L_.dtype_enum(x), B, h, w, d, C.byref(prm), 0,
                                     dsampled.data_ptr(), dx.data_ptr(), C.byref(gs), buf.data_ptr(), buf.numel(), flags,
                                     dev.index, L_.stream_ptr(dev)), "sig_das_bwd")
        return (None,) * 3 + (dx,) + tuple(pg)


class VolumeFunction(torch.autograd.Function):
    """utils/volume.py:14 volume_computation3 -> [B1,B2] fp32."""

    @staticmethod
    def forward(ctx, l, v, a):
        lib = L_.load()
        lf, vf, af = (t.float().contiguous() for t in (l, v, a))
        B1, d = lf.shape
        B2 = vf.shape[0]
        dev = lf.device
        vol = torch.empty(B1, B2, dtype=torch.float32, device=dev)
        nbytes = lib.sig_volume3_ws_bytes(B1, B2)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_volume3_fwd(lf.data_ptr(), vf.data_ptr(), af.data_ptr(), B1, B2, d, vol.data_ptr(), ws.data_ptr(),
                                         nbytes, dev.index, L_.stream_ptr(dev)), "sig_volume3_fwd")
        ctx.save_for_backward(lf, vf, af)
        ctx.dtypes = (l.dtype, v.dtype, a.dtype)
        return vol

    @staticmethod
    def backward(ctx, dvol):
        lf, vf, af = ctx.saved_tensors
        lib = L_.load()
        B1, d = lf.shape
        B2 = vf.shape[0]
        dev = lf.device
        dvol = dvol.float().contiguous()
        dl, dv, da = torch.empty_like(lf), torch.empty_like(vf), torch.empty_like(af)
        nbytes = lib.sig_volume3_ws_bytes(B1, B2)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_volume3_bwd(lf.data_ptr(), vf.data_ptr(), af.data_ptr(), B1, B2, d, dvol.data_ptr(), dl.data_ptr(),
                                         dv.data_ptr(), da.data_ptr(), ws.data_ptr(), nbytes, dev.index, L_.stream_ptr(dev)),
                     "sig_volume3_bwd")
        return dl.to(ctx.dtypes[0]), dv.to(ctx.dtypes[1]), da.to(ctx.dtypes[2])


class VolumeNFunction(torch.autograd.Function):
    """utils/volume.py:65 volume_computation4 / :119 volume_computation5 -> [B1,B2] fp32 (language first, then the n-1
    modalities that share the second batch)."""

    @staticmethod
    def forward(ctx, *feats):
        lib = L_.load()
        n = len(feats)
        fs = [t.float().contiguous() for t in feats]
        for t in fs:
            L_._require_cuda(t, "features")
        B1, d = fs[0].shape
        B2 = fs[1].shape[0]
        if any(t.shape != (B2, d) for t in fs[1:]):
            raise RuntimeError("signal_b200: volume_computation: video / audio / ... must share one [B2, d] shape")
        dev = fs[0].device
        vol = torch.empty(B1, B2, dtype=torch.float32, device=dev)
        nbytes = lib.sig_volume_n_ws_bytes(n, B1, B2)
        if nbytes == 0:
            raise RuntimeError(f"signal_b200: volume_computation with {n} modalities is not supported (3, 4 or 5)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in fs])
        with torch.cuda.device(dev):
            L_.check(lib.sig_volume_n_fwd(n, ptrs, B1, B2, d, vol.data_ptr(), ws.data_ptr(), nbytes, dev.index, L_.stream_ptr(dev)),
                     "sig_volume_n_fwd")
        ctx.save_for_backward(*fs)
        ctx.dtypes = tuple(t.dtype for t in feats)
        return vol

    @staticmethod
    def backward(ctx, dvol):
        fs = ctx.saved_tensors
        lib = L_.load()
        n = len(fs)
        B1, d = fs[0].shape
        B2 = fs[1].shape[0]
        dev = fs[0].device
        dvol = dvol.float().contiguous()
        grads = [torch.empty_like(t) for t in fs]
        nbytes = lib.sig_volume_n_ws_bytes(n, B1, B2)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in fs])
        gptrs = (C.c_void_p * n)(*[t.data_ptr() for t in grads])
        with torch.cuda.device(dev):
            L_.check(lib.sig_volume_n_bwd(n, ptrs, B1, B2, d, dvol.data_ptr(), gptrs, ws.data_ptr(), nbytes, dev.index,
                                          L_.stream_ptr(dev)), "sig_volume_n_bwd")
        return tuple(g.to(dt) for g, dt in zip(grads, ctx.dtypes))


# ------------------------------------------------------------------------------------------
# Whole head: SIM and AlignM of one step as ONE autograd node
# ------------------------------------------------------------------------------------------
_POOL_EVENTS = {}


def _pool_event(dev):
    """one cudaEvent per device for the SIM -> AlignM hand-over of the mean pool (recorded once so that its handle exists)"""
    e = _POOL_EVENTS.get(dev)
    if e is None:
        e = torch.cuda.Event()
        e.record(torch.cuda.current_stream(dev))
        _POOL_EVENTS[dev] = e
    return e


class HeadFunction(torch.autograd.Function):
    """Select_Interactive_Module.forward + AlignmentM.forward on the same three [B,1+L,d] token maps.

    The two modules are independent until their gradients meet in the token maps, so AlignM runs on a
    side stream concurrently with SIM (forward and backward), and the backward writes ONE gradient
    map per modality: SIM's token-gradient kernel overwrites it, AlignM's dX GEMM accumulates on top
    (ordered by a CUDA event).  tensors: 3 token maps, 16 SIM parameters, 22 AlignM parameters and,
    optionally, the 4 tensors of the folded-selection cache.
    """

    @staticmethod
    def forward(ctx, h: int, w: int, do_lam: bool, k1: int, k2: int, max_keep: int, flags: int, side, event, *tensors):
        toks, sp, ap, fold = tensors[:3], tensors[3:19], tensors[19:41], tensors[41:]
        lib = L_.load()
        patches, cls = _split_tokens(True, toks)
        B, L, d = patches[0].shape
        dev = patches[0].device
        dt = L_.dtype_enum(patches[0])
        tok = L_.tokens_struct(patches, cls)
        tok_a = L_.tokens_struct(patches, None)
        sprm = L_.sim_params_struct(sp, fold if len(fold) == 4 else None)
        mods = [ap[1 + 7 * m: 8 + 7 * m] for m in range(3)]
        for t in ap:
            L_._f32c(t)
        aprm = L_.align_params_struct(ap[0], mods)
        out = torch.empty(B, 3 * d, dtype=patches[0].dtype, device=dev)
        masks = torch.empty(3, B, L, dtype=torch.float32, device=dev)
        losses = torch.zeros(2, dtype=torch.float32, device=dev)
        nb_s = L_.ctx_bytes(L_.CTX_SIM, B, L, d, dt, flags)
        nb_a = L_.ctx_bytes(L_.CTX_ALIGN, B, L, d, dt, flags)
        buf_s = torch.empty(nb_s, dtype=torch.uint8, device=dev)
        buf_a = torch.empty(nb_a, dtype=torch.uint8, device=dev)
        main = torch.cuda.current_stream(dev)
        hi = event[2] if len(event) > 2 and event[2] is not None else main    # SIM's (high-priority) stream
        # Training step: AlignM's forward call also runs the loss-weight-independent part of its backward (unit weight,
        # SIG_FLAG_EAGER_BWD) -- on the side stream, under SIM's forward chain of small kernels, where the GPU is
        # otherwise mostly idle; the backward call then starts at the weight-gradient GEMM.
        flags_a = flags | L_.SIG_FLAG_SHARE_SMS          # SIM's chain runs next to AlignM's kernels
        # N2 forward: SIM's selection-score pass has every token value in a register anyway, so it also delivers GAM's mean
        # pool straight into AlignM's ctx slot (sig_sim_params.pool_out -> SIG_FLAG_PATCH_MEAN); AlignM's own pass over the
        # tokens goes away and its GAM chain waits for the event SIM records.  (bf16 tensor-core path only.)
        # Measured on one B200, d = 768 (profiles/r2_fused_pool_experiment.json): 1.788 vs 1.805 ms at B = 512, where the step is
        # throughput-bound and a 75 MB pass less counts; 0.600 vs 0.592 ms at B = 128, where the score kernel sits on SIM's
        # latency-bound critical path and the pool ran in its shadow.  Hence: on from 256 samples up (SIG_FUSE_POOL=1 / 0 forces).
        pool_ev = None
        fuse_env = os.environ.get("SIG_FUSE_POOL", "auto")
        if fuse_env == "1" or (fuse_env == "auto" and B >= 256):
            slot = C.c_void_p()
            if lib.sig_align_patch_mean_slot(buf_a.data_ptr(), B, L, d, dt, flags, C.byref(slot)) == 0 and dt == L_.SIG_BF16 \
                    and not (flags & L_.FLAG_FORCE_SIMT):
                pool_ev = _pool_event(dev)
                sprm.pool_out = slot.value
                sprm.pool_event = pool_ev.cuda_event
                aprm.patch_mean_event = pool_ev.cuda_event
                flags_a = flags_a | L_.SIG_FLAG_PATCH_MEAN
        # Data parallel too (round 2): with AlignM's gradients final in the first third of the backward the in-backward
        # exchange starts earlier and hides under more compute -- measured with the NVLink exchange kernel 0.637 vs 0.676 ms
        # at N = 2 and 0.634 vs 0.675 ms at N = 4 (with ncclAllReduce 0.664 vs 0.676 at N = 2); round 1 had measured the
        # opposite with NCCL's SM-hungry kernels.  SIG_EAGER_BWD=0 switches it off.
        eager_ok = os.environ.get("SIG_EAGER_BWD", "auto") != "0"
        if do_lam and any(ctx.needs_input_grad[9:]) and eager_ok:
            flags_a = flags_a | L_.SIG_FLAG_EAGER_BWD
        with torch.cuda.device(dev):
            side.wait_stream(main)
            if hi is not main:
                hi.wait_stream(main)
            L_.check(lib.sig_sim_fwd(C.byref(tok), C.byref(sprm), k1, k2, max_keep, out.data_ptr(), masks.data_ptr(),
                                     buf_s.data_ptr(), nb_s, flags, dev.index, hi.cuda_stream), "sig_sim_fwd")
            L_.check(lib.sig_align_fwd(C.byref(tok_a), C.byref(aprm), h, w, int(do_lam), losses.data_ptr(), buf_a.data_ptr(), nb_a,
                                       flags_a, dev.index, side.cuda_stream), "sig_align_fwd")
            main.wait_stream(side)
            if hi is not main:
                main.wait_stream(hi)
        ctx.save_for_backward(*toks, *sp, *ap, buf_s, buf_a)
        ctx.cfg = (h, w, do_lam, flags, side, event, len(fold), flags_a)   # event: (torch.cuda.Event, grad_sync or None, ...)
        ctx.mark_non_differentiable(masks)
        return out, masks, losses[0], losses[1]

    @staticmethod
    def backward(ctx, dout, _dmasks, dgam, dlam):
        h, w, do_lam, flags, side, event, nfold, flags_a = ctx.cfg
        saved = ctx.saved_tensors
        toks, sp, ap, buf_s, buf_a = saved[:3], saved[3:19], saved[19:41], saved[41], saved[42]
        lib = L_.load()
        patches, cls = _split_tokens(True, toks)
        B, L, d = patches[0].shape
        dev = patches[0].device
        z = torch.zeros((), dtype=torch.float32, device=dev)
        dl = torch.stack([z if dgam is None else dgam.float(), z if dlam is None else dlam.float()]).contiguous()
        dout = (torch.zeros(B, 3 * d, dtype=patches[0].dtype, device=dev) if dout is None
                else dout.to(patches[0].dtype).contiguous())
        dtoks = [torch.empty_like(t) for t in toks]
        dpatch, dcls = [t[:, 1:] for t in dtoks], [t[:, 0] for t in dtoks]
        sync = event[3] if len(event) > 3 else None     # (comm stream, early event, AlignM event, pieces) of FusionHead
        if event[1] is None:
            sync = None
        if sync is not None and sync[3] in (2, 3):
            # ONE arena [SIM late | SIM early | AlignM]; exchanged as 2 pieces ([early + AlignM], [late]: few calls, NCCL) or
            # 3 pieces (early, AlignM, late: each as soon as it is final -- the NVLink kernel's calls are cheap)
            flat_h, pg_s, pg_a, cut_h, cut2_h = _head_arena(d, dev, event[4] if len(event) > 4 else None)
            if not do_lam:
                flat_h[cut2_h:].zero_()
        else:
            flat_s, pg_s, split_s = _sim_arena(d, dev)
            flat_a, pg_a = _arena(_align_grad_shapes(d), dev)
            if not do_lam:
                flat_a.zero_()
        tok = L_.tokens_struct(patches, cls)
        tok_a = L_.tokens_struct(patches, None)
        sprm = L_.sim_params_struct(sp)
        mods = [ap[1 + 7 * m: 8 + 7 * m] for m in range(3)]
        aprm = L_.align_params_struct(ap[0], mods)
        gs_s = L_.sim_grads_struct(pg_s)
        gs_a = L_.align_params_struct(pg_a[0], [pg_a[1 + 7 * m: 8 + 7 * m] for m in range(3)], cls=L_.SigAlignParamGrads)
        hi = event[2] if len(event) > 2 else None
        event, grad_sync = event[0], event[1]
        late_ev = sync[4] if sync is not None and len(sync) > 4 and os.environ.get("SIG_LATE_EVENT", "1") != "0" else None
        if sync is not None:
            gs_s.early_event = sync[1].cuda_event
            gs_a.done_event = sync[2].cuda_event
            if late_ev is not None:
                gs_s.late_event = late_ev.cuda_event
        evh = event.cuda_event
        # SIM's token-gradient kernel overwrites the shared map (it has no long GEMM in front of it and finishes
        # first); AlignM's dX GEMM, which can run its weight-gradient GEMM while it waits, adds on top
        fuse = None
        if do_lam and L == 128 and d <= 768 and os.environ.get("SIG_FUSE_DX", "1") != "0":   # (AlignM's tensor-core path limits)
            fuse = L_.sim_dx_operands(buf_s, B, L, d, L_.dtype_enum(patches[0]), flags)
        if fuse is not None:
            # bf16 tensor-core path: SIM hands the two operands of its token gradient to AlignM, whose dX GEMM
            # appends them as one more k-block and writes the shared map once (no SIM write, no read-back)
            tg_s = L_.token_grads_struct(dpatch, dcls, accumulate=False, done_event=evh, fuse_skip_dx=True)
            tg_a = L_.token_grads_struct(dpatch, dcls, accumulate=False, zero_cls=False, wait_event=evh, fuse_ops=fuse)
        else:
            tg_s = L_.token_grads_struct(dpatch, dcls, accumulate=False, done_event=evh)
            tg_a = L_.token_grads_struct(dpatch, dcls, accumulate=True, zero_cls=False, wait_event=evh)
        main = torch.cuda.current_stream(dev)
        if hi is None:
            hi = main
        with torch.cuda.device(dev):
            side.wait_stream(main)
            if hi is not main:
                hi.wait_stream(main)
            # (SIM first: its call records the event AlignM's call waits on)
            L_.check(lib.sig_sim_bwd(C.byref(tok), C.byref(sprm), dout.data_ptr(), C.byref(tg_s), C.byref(gs_s), buf_s.data_ptr(),
                                     buf_s.numel(), flags, dev.index, hi.cuda_stream), "sig_sim_bwd")
            # Data parallel: the exchange runs on a communication stream in pieces, each started by the event(s) the
            # library records when that piece is final -- SIM's early part (FFN, out_proj, norms, W_v: ready before
            # the token-side backward) together with AlignM's arena (ready before its dX GEMM; in the first third of
            # the backward with the eager chain) as ONE collective, then SIM's W_q/W_k part (a collective costs
            # ~25 us + 2.2 us/MB at N = 2); SIG_SYNC_CHUNKS=3: three collectives
            if sync is not None:
                comm = sync[0]
                if sync[3] == 3:
                    with torch.cuda.stream(comm):
                        comm.wait_event(sync[1])          # (recorded on SIM's stream, which is ordered behind `main`)
                        grad_sync(flat_h[cut_h:cut2_h])
                elif sync[3] != 2:
                    with torch.cuda.stream(comm):
                        comm.wait_event(sync[1])
                        grad_sync(flat_s[split_s:])
            elif grad_sync is not None:
                with torch.cuda.stream(hi):
                    grad_sync(flat_s)
            L_.check(lib.sig_align_bwd(C.byref(tok_a), C.byref(aprm), h, w, int(do_lam), dl.data_ptr(), C.byref(tg_a), C.byref(gs_a),
                                       buf_a.data_ptr(), buf_a.numel(), flags_a, dev.index, side.cuda_stream), "sig_align_bwd")
            if sync is not None:
                with torch.cuda.stream(comm):
                    comm.wait_event(sync[2])
                    if sync[3] == 2:      # SIM's early part and AlignM's arena are adjacent: one collective, then the late part
                        comm.wait_event(sync[1])
                        grad_sync(flat_h[cut_h:])
                        # the late piece (W_q / W_k / in_proj_bias): as soon as its last weight-gradient GEMM is through, not
                        # when SIM's stream has also finished the CLS-gradient GEMM and the token-gradient writes
                        comm.wait_event(late_ev) if late_ev is not None else comm.wait_stream(hi)
                        grad_sync(flat_h[:cut_h])
                    elif sync[3] == 3:
                        grad_sync(flat_h[cut2_h:])
                        comm.wait_event(late_ev) if late_ev is not None else comm.wait_stream(hi)
                        grad_sync(flat_h[:cut_h])
                    else:
                        grad_sync(flat_a)
                        comm.wait_stream(hi)
                        grad_sync(flat_s[:split_s])
                main.wait_stream(comm)
            elif grad_sync is not None:
                with torch.cuda.stream(side):
                    grad_sync(flat_a)
            main.wait_stream(side)
            if hi is not main:
                main.wait_stream(hi)
        if not do_lam:      # stage == "CLS": no gradient for the DAS parameters (useB.py:181-183)
            pg_a = [pg_a[0]] + [None] * (len(pg_a) - 1)
        return (None,) * 9 + tuple(dtoks) + (None,) * 4 + tuple(pg_s) + tuple(pg_a) + (None,) * nfold
