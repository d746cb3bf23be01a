"""Drop-in nn.Modules for the reference's fusion head.

Same class names, constructor signatures, sub-module / parameter names (so reference
checkpoints load with ``load_state_dict``) and attribute surface as

* modeling/AddModule/useA.py : TokenSelection, ModalInteractive, LayerNorm, Select_Interactive_Module
* modeling/AddModule/useB.py : AlignmentM
* modeling/AddModule/DAS.py  : DA_sample
* utils/volume.py            : volume_computation3

The nn.Linear / nn.MultiheadAttention / nn.Conv2d sub-modules are *parameter containers*
(constructed in the reference's order, so the default initialisation under a given seed is
identical); their forward is never called -- all arithmetic runs in libsignal_b200.so.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import functional as F_

__all__ = ["TokenSelection", "ModalInteractive", "LayerNorm", "Select_Interactive_Module", "AlignmentM", "DA_sample",
           "volume_computation3", "FusionHead"]


def _packed_base(patch: torch.Tensor, glob: torch.Tensor):
    """If (patch, glob) are x[:,1:] and x[:,0] of one contiguous [B,1+L,d] map x, return x."""
    base = patch._base
    if base is None or base is not glob._base or base.dim() != 3 or not base.is_contiguous():
        return None
    B, L, d = patch.shape
    if base.shape != (B, L + 1, d) or glob.shape != (B, d):
        return None
    if patch.storage_offset() != base.storage_offset() + d or glob.storage_offset() != base.storage_offset():
        return None
    if patch.stride() != (base.stride(0), d, 1) or glob.stride() != (base.stride(0), 1):
        return None
    if not (patch.requires_grad == glob.requires_grad == base.requires_grad):
        return None
    return base


def _packed_patch_base(patch: torch.Tensor):
    """If patch is x[:,1:] of a contiguous [B,1+L,d] map x, return x."""
    base = patch._base
    if base is None or base.dim() != 3 or not base.is_contiguous():
        return None
    B, L, d = patch.shape
    if base.shape != (B, L + 1, d) or patch.storage_offset() != base.storage_offset() + d:
        return None
    if patch.stride() != (base.stride(0), d, 1) or patch.requires_grad != base.requires_grad:
        return None
    return base


def _half_in(patches, globs=None):
    """fp16 token maps (the reference's amp.autocast, engine/processor.py:165) -> bf16 for the kernels; one conversion
    pass per [B,1+L,d] map when (patch, glob) are the two views of one map, else one per view.  Returns
    (patches, globs, was_half); the caller converts its floating-point outputs back with _half_out."""
    if patches[0].dtype != torch.float16:
        return list(patches), (list(globs) if globs is not None else None), False
    bf = torch.bfloat16
    new_p, new_g = [], []
    for m, p in enumerate(patches):
        g = globs[m] if globs is not None else None
        base = _packed_base(p, g) if g is not None else _packed_patch_base(p)
        if base is not None:
            nb = F_.HalfBridge.apply(base, bf)
            new_p.append(nb[:, 1:])
            new_g.append(nb[:, 0])
        else:
            new_p.append(F_.HalfBridge.apply(p, bf))
            new_g.append(F_.HalfBridge.apply(g, bf) if g is not None else None)
    return new_p, (new_g if globs is not None else None), True


def _half_out(t, was_half):
    return F_.HalfBridge.apply(t, torch.float16) if was_half else t


def _same_strides(patches):
    """The bf16 tensor-core path of AlignM reads the three modalities through one tensor-map geometry (the C entry
    rejects mixed strides with SIG_ERR_SHAPE): hand it contiguous copies in that (unusual) case."""
    if patches[0].dtype == torch.bfloat16 and len({(p.stride(0), p.stride(1)) for p in patches}) > 1:
        return [p.contiguous() for p in patches]
    return list(patches)


class LayerNorm(nn.LayerNorm):
    """fp32 LayerNorm that casts back to the input dtype (useA.py:414-423).  Parameter container
    for norm1/norm2; kept callable for code that uses the class on its own."""

    def forward(self, x: torch.Tensor):
        orig_type = x.dtype
        ret = super().forward(x.type(torch.float32))
        return ret.type(orig_type)


class TokenSelection(nn.Module):
    """useA.py:16-325."""

    def __init__(self, dim, k=112, keep_ratio=None):
        super().__init__()
        self.dim = dim
        self.k1 = k
        self.k2 = 2 * k
        self.keep_ratio = keep_ratio
        self.W_q = nn.Linear(dim, dim)
        self.W_k = nn.Linear(dim, dim)
        self.W_v = nn.Linear(dim, dim)   # unused by the reference forward as well (useA.py:48)
        self.register_load_state_dict_post_hook(lambda module, _incompatible: module.invalidate_fold())

    def invalidate_fold(self):
        """Drop the cached fold M = W_k^T W_q of the frozen selection parameters (used by the bf16 path).  The cache
        is keyed on (data_ptr, _version, device, dtype) of the four tensors and is dropped by load_state_dict and by
        .to()/.cuda()/.float(); in-place edits through ``p.data`` (EMA updates, manual re-initialisation) do NOT bump
        ``_version`` -- call this after them."""
        self.__dict__.pop("_fold_cache", None)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_fold()
        return super()._apply(fn, *args, **kwargs)

    def _sel_params(self):
        return [self.W_q.weight, self.W_q.bias, self.W_k.weight, self.W_k.bias]

    def _selection_fold(self):
        """Cached fold of the four frozen selection tensors (they never receive gradients; the cache is
        rebuilt when any of them is modified in place, replaced, or moved)."""
        sel = self._sel_params()
        key = tuple((p.data_ptr(), p._version, p.device, p.dtype) for p in sel)
        cache = self.__dict__.get("_fold_cache")
        if cache is None or cache[0] != key:
            with torch.no_grad():
                from . import lib as L_
                cache = (key, L_.fold_selection([p.detach() for p in sel]))
            self.__dict__["_fold_cache"] = cache
        return cache[1]

    def _max_keep(self, L):
        return -1 if self.keep_ratio is None else int(L * self.keep_ratio)

    def _masks(self, which, patches, cls):
        dt = patches[0].dtype
        with torch.no_grad():
            patches, cls, _ = _half_in(patches, cls)
        m = F_.select_masks(which, patches, cls, self._sel_params(), self.k1, self.k2, -1)
        return tuple(m[i].to(dt).unsqueeze(-1) for i in range(3))

    def intra_modal_token_selection(self, rgb_patches, nir_patches, tir_patches, rgb_global, nir_global, tir_global):
        return self._masks(1, [rgb_patches, nir_patches, tir_patches], [rgb_global, nir_global, tir_global])

    def inter_modal_token_selection(self, rgb_patches, nir_patches, tir_patches, rgb_global, nir_global, tir_global):
        return self._masks(2, [rgb_patches, nir_patches, tir_patches], [rgb_global, nir_global, tir_global])

    def forward(self, rgb_patches, nir_patches, tir_patches, rgb_global, nir_global, tir_global):
        L = rgb_patches.size(1)
        (rgb_patches, nir_patches, tir_patches), (rgb_global, nir_global, tir_global), half = _half_in(
            [rgb_patches, nir_patches, tir_patches], [rgb_global, nir_global, tir_global])
        r, n, t, masks = F_.SelectFunction.apply(self.k1, self.k2, self._max_keep(L), rgb_patches, nir_patches, tir_patches,
                                                 rgb_global, nir_global, tir_global, *[p.detach() for p in self._sel_params()])
        self.last_masks = {"RGB": masks[0].unsqueeze(-1), "NI": masks[1].unsqueeze(-1), "TI": masks[2].unsqueeze(-1)}
        return _half_out(r, half), _half_out(n, half), _half_out(t, half)


class ModalInteractive(nn.Module):
    """useA.py:328-411."""

    def __init__(self, dim, num_heads=8):
        super().__init__()
        if num_heads != 8:
            raise NotImplementedError("signal_b200: ModalInteractive kernels are built for 8 heads (useA.py:449)")
        self.dim = dim
        self.num_heads = num_heads
        self.cross_attn = nn.MultiheadAttention(dim, num_heads, batch_first=True)
        self.ffn = nn.Sequential(nn.Linear(dim, 2 * dim), nn.GELU(), nn.Linear(2 * dim, dim))
        self.norm1 = LayerNorm(dim)
        self.norm2 = LayerNorm(dim)

    def _attn_params(self):
        a = self.cross_attn
        return [a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                self.ffn[0].weight, self.ffn[0].bias, self.ffn[2].weight, self.ffn[2].bias,
                self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias]

    def forward(self, rgb_selected, nir_selected, tir_selected, rgb_global, nir_global, tir_global):
        p = self._attn_params()
        dummy = [p[0].detach()] * 4   # selection parameters are not read by the attention entry point
        (rgb_selected, nir_selected, tir_selected), (rgb_global, nir_global, tir_global), half = _half_in(
            [rgb_selected, nir_selected, tir_selected], [rgb_global, nir_global, tir_global])
        return _half_out(F_.AttnFunction.apply(0, rgb_selected, nir_selected, tir_selected, rgb_global, nir_global, tir_global,
                                               *dummy, *p), half)


class Select_Interactive_Module(nn.Module):
    """useA.py:426-476.  forward -> [B, 3*dim]."""

    def __init__(self, dim, k=112, keep_ratio=None):
        super().__init__()
        num_heads = 8
        self.token_selection = TokenSelection(dim, k, keep_ratio)
        self.modal_interactive = ModalInteractive(dim, num_heads)
        self.flags = 0          # lib.FLAG_FORCE_SIMT to cross-check the tcgen05 path
        self.fuse_views = True  # route x[:,1:], x[:,0] views through their common [B,1+L,d] map

    def forward(self, rgb_patches, nir_patches, tir_patches, rgb_global, nir_global, tir_global):
        ts, mi = self.token_selection, self.modal_interactive
        if ts._forward_hooks or ts._forward_pre_hooks or mi._forward_hooks or mi._forward_pre_hooks:
            # something observes the sub-modules (e.g. zablation/CAM.py:164): run them one by one
            sel = ts(rgb_patches, nir_patches, tir_patches, rgb_global, nir_global, tir_global)
            return mi(*sel, rgb_global, nir_global, tir_global)
        L = rgb_patches.size(1)
        (rgb_patches, nir_patches, tir_patches), (rgb_global, nir_global, tir_global), half = _half_in(
            [rgb_patches, nir_patches, tir_patches], [rgb_global, nir_global, tir_global])
        params = [p.detach() for p in ts._sel_params()] + mi._attn_params()
        if rgb_patches.dtype == torch.bfloat16 and not (self.flags & 1):
            params = params + list(ts._selection_fold())
        bases = None
        if self.fuse_views:
            bases = [_packed_base(p, g) for p, g in ((rgb_patches, rgb_global), (nir_patches, nir_global), (tir_patches, tir_global))]
            if any(b is None for b in bases):
                bases = None
        if bases is not None:
            out, masks = F_.SimFunction.apply(True, ts.k1, ts.k2, ts._max_keep(L), self.flags, *bases, *params)
        else:
            out, masks = F_.SimFunction.apply(False, ts.k1, ts.k2, ts._max_keep(L), self.flags, rgb_patches, nir_patches,
                                              tir_patches, rgb_global, nir_global, tir_global, *params)
        ts.last_masks = {"RGB": masks[0].unsqueeze(-1), "NI": masks[1].unsqueeze(-1), "TI": masks[2].unsqueeze(-1)}
        return _half_out(out, half)


class DA_sample(nn.Module):
    """DAS.py:17-165.  forward(x [B,C,H,W]) -> [B,C,H/4,W/4]."""

    def __init__(self, n_heads, n_head_channels, n_groups, stride, offset_range_factor, ksize):
        super().__init__()
        if n_groups != 1 or stride != 4 or ksize != 4 or offset_range_factor != 2:
            raise NotImplementedError("signal_b200: DA_sample kernels cover the configuration AlignmentM uses "
                                      "(n_groups=1, stride=4, ksize=4, offset_range_factor=2; useB.py:63-68)")
        self.n_head_channels = n_head_channels
        self.nc = n_head_channels * n_heads
        self.n_groups = n_groups
        self.n_group_channels = self.nc // self.n_groups
        self.kk = ksize
        self.stride = stride
        self.offset_range_factor = offset_range_factor
        c = self.n_group_channels
        self.conv_offset = nn.Sequential(
            nn.Conv2d(c, c, 1, 1, 0), nn.GELU(),
            nn.Conv2d(c, c, self.kk, stride, 0, groups=c), nn.GELU(),
            nn.Conv2d(c, 1, 1, 1, 0, bias=False))
        self.proj_q = nn.Conv2d(self.nc, self.nc, kernel_size=1, stride=1, padding=0)
        self.flags = 0

    def _params(self):
        co = self.conv_offset
        return [self.proj_q.weight, self.proj_q.bias, co[0].weight, co[0].bias, co[2].weight, co[2].bias, co[4].weight]

    def forward(self, x):
        B, Cc, H, W = x.shape
        xt = x.permute(0, 2, 3, 1)                     # [B,H,W,C]; free for the channels-last view AlignmentM passes
        if xt.stride(3) != 1 or xt.stride(1) != W * xt.stride(2):
            xt = xt.contiguous()
        xt = xt.reshape(B, H * W, Cc)
        if xt.dtype == torch.float16:
            xt = F_.HalfBridge.apply(xt, torch.bfloat16)
        s = F_.DasFunction.apply(H, W, self.flags, xt, *self._params())    # [B,P,C] fp32
        return s.to(x.dtype).reshape(B, H // 4, W // 4, Cc).permute(0, 3, 1, 2)


class AlignmentM(nn.Module):
    """useB.py:28-190.  forward(R, N, T, stage) -> gam | (gam, lam)."""

    def __init__(self, feat_dim, H, W):
        super().__init__()
        self.feat_dim = feat_dim
        self.contra_temp = nn.Parameter(torch.tensor(0.07))
        self.h, self.w = H, W
        self.mse = nn.MSELoss()
        # the reference hard-codes 512 head channels (useB.py:64) and therefore only runs at
        # feat_dim == 512; identical there, and feat_dim == 768 works here
        n_heads, n_head_channels, n_groups, stride, offset_range_factor, ksize = 1, feat_dim, 1, 4, 2, 4
        self.DAS_r = DA_sample(n_heads, n_head_channels, n_groups, stride, offset_range_factor, ksize)
        self.DAS_n = DA_sample(n_heads, n_head_channels, n_groups, stride, offset_range_factor, ksize)
        self.DAS_t = DA_sample(n_heads, n_head_channels, n_groups, stride, offset_range_factor, ksize)
        self.flags = 0
        self.fuse_views = True
        # Optional forward hint, consumed by the NEXT call: the fp32 mean over the patch rows of the three modalities
        # ([3,B,d] or three [B,d] tensors), e.g. TokenProducer.last_patch_mean -- the GAM pool pass over the tokens
        # (useB.py:84-86) is then skipped on the bf16 path.  It must be the mean of exactly the maps passed in.
        self.patch_mean_hint = None
        # Signal.forward computes AlignM in eval mode too and throws the result away (make_model.py:277-281).  Opt-in:
        # with skip_in_eval = True an eval-mode call launches nothing and returns fp32 zeros of the reference's shapes.
        self.skip_in_eval = False

    def _params(self):
        return [self.contra_temp] + self.DAS_r._params() + self.DAS_n._params() + self.DAS_t._params()

    def _run(self, RGB_patch, NI_patch, TI_patch, do_lam):
        if RGB_patch.dtype != torch.bfloat16:
            # the hint stands for the pool of the maps the kernels read: fp16 maps are converted to bf16 first and fp32
            # maps pool in fp64 on the exact path, so only bf16 callers may skip the pass
            self.patch_mean_hint = None
        (RGB_patch, NI_patch, TI_patch), _, _ = _half_in([RGB_patch, NI_patch, TI_patch])     # (the two losses are fp32)
        bases = None
        if self.fuse_views:
            bases = [_packed_patch_base(p) for p in (RGB_patch, NI_patch, TI_patch)]
            if any(b is None for b in bases):
                bases = None
        hint, self.patch_mean_hint = self.patch_mean_hint, None
        extra = ()
        if hint is not None:
            extra = (hint.detach() if torch.is_tensor(hint) else torch.stack([t.detach() for t in hint]),)
        if bases is not None:
            return F_.AlignFunction.apply(True, self.h, self.w, do_lam, self.flags, *bases, *self._params(), *extra)
        return F_.AlignFunction.apply(False, self.h, self.w, do_lam, self.flags, *_same_strides([RGB_patch, NI_patch, TI_patch]),
                                      *self._params(), *extra)

    def Cls_Align(self, RGB_patch, NI_patch, TI_patch):
        return self._run(RGB_patch, NI_patch, TI_patch, False)[0]

    def patch_Align(self, RGB_patch, NI_patch, TI_patch):
        return self._run(RGB_patch, NI_patch, TI_patch, True)[1]

    def forward(self, RGB_patch, NI_patch, TI_patch, stage):
        if self.skip_in_eval and not self.training:
            if not RGB_patch.is_cuda:
                raise RuntimeError("signal_b200: patch tokens must be CUDA tensors (there is no CPU path)")
            z = torch.zeros((), dtype=torch.float32, device=RGB_patch.device)
            return z if stage == "CLS" else (z, z.clone())
        if stage == "CLS":
            return self.Cls_Align(RGB_patch, NI_patch, TI_patch)
        if "Cls_Align" in self.__dict__ or "patch_Align" in self.__dict__:
            # a method was monkey-patched on the instance (zablation/offestvisual.py:209-214): honour it
            return self.Cls_Align(RGB_patch, NI_patch, TI_patch), self.patch_Align(RGB_patch, NI_patch, TI_patch)
        return self._run(RGB_patch, NI_patch, TI_patch, True)


def volume_computation3(language, video, audio):
    """utils/volume.py:14-62 -> [B1,B2] fp32."""
    return F_.VolumeFunction.apply(language, video, audio)


def volume_computation4(language, video, audio, subtitles):
    """utils/volume.py:65-116 -> [B1,B2] fp32 (never called by the reference; completes the utils/volume.py surface)."""
    return F_.VolumeNFunction.apply(language, video, audio, subtitles)


def volume_computation5(language, video, audio, subtitles, depth):
    """utils/volume.py:119-182 -> [B1,B2] fp32."""
    return F_.VolumeNFunction.apply(language, video, audio, subtitles, depth)


class FusionHead:
    """SIM + AlignM of one training step as a single call.

    ``head = FusionHead(model.SIM, model.AlignM)`` holds references to the two drop-in modules (it owns
    no parameters, so checkpoints are unchanged) and replaces the two consecutive calls of
    make_model.py:191,205::

        vars_total = self.SIM(rgb_patch, ni_patch, ti_patch, RGB_global, NI_global, TI_global)
        loss_area, patch_loss = self.AlignM(rgb_patch, ni_patch, ti_patch, stage=...)

    by ``vars_total, loss_area, patch_loss = head(rgb_patch, ..., TI_global, stage=...)``.  AlignM then runs
    on a side stream next to SIM, and the backward writes one token-gradient map per modality.
    Falls back to the two module calls when the six inputs are not views of three [B,1+L,d] maps.
    """

    def __init__(self, sim: "Select_Interactive_Module", align: "AlignmentM", grad_sync=None):
        self.sim, self.align = sim, align
        self._side = {}
        # Data parallel (one process per GPU): ``grad_sync(flat)`` is called inside the backward with contiguous
        # pieces of the flat fp32 parameter-gradient arena (two per step: SIM's early part, then AlignM's gradients
        # + SIM's W_q/W_k part), on a communication stream that waits for the event the library records when
        # that piece is final -- e.g. ``lambda a: dist.all_reduce(a, op=dist.ReduceOp.AVG)``.  The gradients
        # that reach ``.grad`` are then already averaged and the exchange overlaps the rest of the backward.
        self.grad_sync = grad_sync
        # optional caller-owned flat fp32 arena the backward carves the parameter gradients from (grad_numel() elements;
        # parallel.GradExchange.arena: symmetric memory, so the exchange kernel can reach the peers' gradients)
        self.grad_arena = None

    def grad_numel(self) -> int:
        return F_.head_grad_numel(self.sim.token_selection.dim)

    def make_graphed(self, rgb_tokens, ni_tokens, ti_tokens, stage="together_CLS_Patch"):
        """CUDA-graph replay speed through the ordinary autograd API.

        The eager step is host-bound (ctypes calls, allocations and ~110 kernel launches cost ~1.5 ms for 0.6 ms of GPU
        work at B = 128).  ``graphed = head.make_graphed(rgb, ni, ti)`` captures the forward and the backward of the
        head as two CUDA graphs (``torch.cuda.make_graphed_callables``: static input / output / gradient buffers,
        three warm-up iterations) and returns a callable

            vars_total, loss_area, patch_loss = graphed(rgb_tokens, ni_tokens, ti_tokens)      # [B,1+L,d] maps

        that takes part in autograd like the module calls it replaces: the losses can be combined with anything else,
        ``backward()`` replays the backward graph and delivers gradients to the three token maps and to the parameters
        of SIM / AlignM.  The sample maps fix shape, dtype and device; pass tensors that require grad if the token
        gradients are needed.  ``sim.token_selection.last_masks`` refers to static tensors that every replay refreshes.
        Parameters may be updated in place between calls (the graphs read them through their addresses); replacing a
        parameter tensor, changing ``keep_ratio`` or registering hooks needs a new ``make_graphed``."""
        head = self

        class _Head(nn.Module):
            def __init__(self):
                super().__init__()
                self.SIM, self.AlignM = head.sim, head.align

            def forward(self, rgb, ni, ti):
                out, gam, lam = head(rgb[:, 1:], ni[:, 1:], ti[:, 1:], rgb[:, 0], ni[:, 0], ti[:, 0], stage=stage)
                return (out, gam) if lam is None else (out, gam, lam)

        if self.grad_sync is not None:
            raise RuntimeError("signal_b200: make_graphed captures a rank-local step; capture the data-parallel step (grad_sync "
                               "set) yourself with torch.cuda.graph as bench.py does")
        sample = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in (rgb_tokens, ni_tokens, ti_tokens))
        graphed = torch.cuda.make_graphed_callables(_Head(), sample, allow_unused_input=True)
        ts = self.sim.token_selection
        static_masks = dict(ts.last_masks)      # written by the captured forward: the same memory on every replay

        def call(rgb, ni, ti):
            res = graphed(rgb, ni, ti)
            ts.last_masks = static_masks
            return res
        return call

    def _stream(self, dev):
        st = self._side.get(dev)
        if st is None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))       # materialise the underlying cudaEvent_t
            # SIG_PRIO=1 (experiment, off by default -- measured slower): SIM on a high-priority stream
            hi = torch.cuda.Stream(dev, priority=-1) if os.environ.get("SIG_PRIO", "0") == "1" else None
            # data parallel: a communication stream and the two events the library records when a piece of the
            # parameter gradients is final (sig_sim_param_grads.early_event, sig_align_param_grads.done_event)
            sync = [torch.cuda.Stream(dev, priority=-1), torch.cuda.Event(), torch.cuda.Event()]
            late = torch.cuda.Event()          # sig_sim_param_grads.late_event: W_q / W_k / in_proj_bias gradients enqueued
            for e in sync[1:] + [late]:
                e.record(torch.cuda.current_stream(dev))
            # SIG_SYNC_CHUNKS (diagnostic): 2 = default; 3 = AlignM's arena and SIM's late part as two collectives;
            # 0 = one exchange per module after its backward
            # 4 = separate arenas per module, three collectives (round-1 layout)
            pieces = int(os.environ.get("SIG_SYNC_CHUNKS", "-1"))
            sync = None if pieces == 0 else sync + [pieces, late]
            st = (torch.cuda.Stream(dev), ev, hi, sync)
            self._side[dev] = st
        return st

    def __call__(self, rgb_patch, ni_patch, ti_patch, rgb_global, ni_global, ti_global, stage="together_CLS_Patch"):
        sim, al = self.sim, self.align
        ts, mi = sim.token_selection, sim.modal_interactive
        (rgb_patch, ni_patch, ti_patch), (rgb_global, ni_global, ti_global), half = _half_in(
            [rgb_patch, ni_patch, ti_patch], [rgb_global, ni_global, ti_global])
        bases = [_packed_base(p, g) for p, g in ((rgb_patch, rgb_global), (ni_patch, ni_global), (ti_patch, ti_global))]
        hooked = ts._forward_hooks or ts._forward_pre_hooks or mi._forward_hooks or mi._forward_pre_hooks
        patched = "Cls_Align" in al.__dict__ or "patch_Align" in al.__dict__
        if any(b is None for b in bases) or hooked or patched or not rgb_patch.is_cuda:
            out = _half_out(sim(rgb_patch, ni_patch, ti_patch, rgb_global, ni_global, ti_global), half)
            res = al(rgb_patch, ni_patch, ti_patch, stage=stage)
            return (out, res, None) if stage == "CLS" else (out, res[0], res[1])
        L = rgb_patch.size(1)
        side, ev, hi, sync = self._stream(rgb_patch.device)
        if sync is not None and sync[3] < 0:
            # default: two pieces, [SIM early + AlignM] and [SIM late] (SIG_SYNC_CHUNKS=3: SIM early on its own as soon as
            # it is final -- measured slower at N = 2 and N = 8: every call costs ~25 us of cross-GPU barriers, and the early
            # piece then runs next to SIM's HBM-bound token passes instead of next to the GEMMs)
            sync = sync[:3] + [2] + sync[4:]
        params = [p.detach() for p in ts._sel_params()] + mi._attn_params() + al._params()
        flags = sim.flags | al.flags
        if rgb_patch.dtype == torch.bfloat16 and not (flags & 1):
            params = params + list(ts._selection_fold())
        out, masks, gam, lam = F_.HeadFunction.apply(al.h, al.w, stage != "CLS", ts.k1, ts.k2, ts._max_keep(L), flags, side,
                                                     (ev, self.grad_sync, hi, sync, self.grad_arena), *bases, *params)
        ts.last_masks = {"RGB": masks[0].unsqueeze(-1), "NI": masks[1].unsqueeze(-1), "TI": masks[2].unsqueeze(-1)}
        out = _half_out(out, half)
        return (out, gam, None) if stage == "CLS" else (out, gam, lam)
