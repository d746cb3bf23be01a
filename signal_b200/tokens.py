"""Token producer in front of the fusion head.

The reference's CLIP vision tower ends with ``x = self.ln_post(x); xproj = x @ self.proj`` (modeling/clip/model.py:485-487)
and ``build_transformer.forward`` splits the result into ``x_cash = x[:, 1:]`` / ``global_feat = x[:, 0]``
(modeling/meta_arch.py:108-110): that is where the three ``[B, 129, d]`` token maps of the head come from.
``TokenProducer`` runs that tail as CUDA kernels behind the C ABI (``sig_tokens_fwd/bwd``, csrc/tokens.cu): LayerNorm with
fp32 statistics -> bf16 operand -> tcgen05 GEMM, writing the ``[B, 1+L, d]`` map the head's TMA descriptors read, with the
GAM mean pool of the patch rows (useB.py:84-86) as a by-product.  Like ``BNNeckClassifier`` it owns no parameters: it
wraps the tower's ``ln_post`` module and ``proj`` parameter, so checkpoints and optimizer state are untouched.

    tp = TokenProducer(model.base.ln_post, model.base.proj)
    x_cash, global_feat = tp(x)            # x: [B, 1+L, W] (any batch / row strides, e.g. blocks_out.permute(1, 0, 2))
    tp.last_patch_mean                      # fp32 [B, d]: mean over the patch rows (detached by-product)

There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as L_

_DT = {torch.float32: L_.SIG_F32, torch.bfloat16: L_.SIG_BF16, torch.float16: L_.SIG_F16}


def _u8(n, dev):
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=dev)


class TokenProducerFunction(torch.autograd.Function):
    """(x [B,L1,W] strided, ln_w, ln_b, proj [W,D], eps, want_mean, tok_dtype) -> (tokens [B,L1,D] of tok_dtype,
    patch_mean [B,D] fp32 or empty).  tok_dtype = x.dtype, or bf16 / fp16 for fp32 x (autocast)."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, proj, eps: float, want_mean: bool, tok_dtype):
        lib = L_.load()
        if not x.is_cuda:
            raise RuntimeError("signal_b200: TokenProducer input must be a CUDA tensor (there is no CPU path)")
        if x.dim() != 3 or x.stride(2) != 1 or x.dtype not in _DT:
            raise RuntimeError("signal_b200: TokenProducer expects a [B, 1+L, W] fp32 / bf16 / fp16 map with unit channel stride")
        B, L1, W = x.shape
        D = proj.shape[1]
        dev, dt = x.device, _DT[x.dtype]
        if tok_dtype is None:
            tok_dtype = x.dtype
        if tok_dtype not in _DT or (tok_dtype != x.dtype and x.dtype != torch.float32):
            raise RuntimeError(f"signal_b200: TokenProducer cannot produce {tok_dtype} tokens from {x.dtype} input")
        tdt = _DT[tok_dtype]
        ln_w, ln_b, proj = L_._f32c(ln_w.detach()), L_._f32c(ln_b.detach()), L_._f32c(proj.detach())
        nb = [lib.sig_tokens_ws_bytes(k, B, L1, W, D, tdt) for k in range(3)]
        if nb[0] == 0:
            raise RuntimeError(f"signal_b200: unsupported token producer shape B={B} L1={L1} W={W} D={D} (W, D % 8 == 0, <= 1024)")
        tokens = torch.empty(B, L1, D, dtype=tok_dtype, device=dev)
        mean = torch.empty(B, D, dtype=torch.float32, device=dev) if want_mean else None
        saved = _u8(nb[0], dev)
        scratch = _u8(nb[1], dev) if nb[1] else None
        with torch.cuda.device(dev):
            L_.check(lib.sig_tokens_fwd(x.data_ptr(), dt, x.stride(0), x.stride(1), B, L1, W, D, ln_w.data_ptr(), ln_b.data_ptr(),
                                        float(eps), proj.data_ptr(), tokens.data_ptr(), tdt, None if mean is None else mean.data_ptr(),
                                        saved.data_ptr(), saved.numel(), None if scratch is None else scratch.data_ptr(),
                                        0 if scratch is None else scratch.numel(), dev.index, L_.stream_ptr(dev)), "sig_tokens_fwd")
        ctx.save_for_backward(x, ln_w, proj, saved)
        ctx.geom = (B, L1, W, D, dt, tdt, tok_dtype, nb[2])
        if mean is None:
            mean = torch.empty(0, dtype=torch.float32, device=dev)
        ctx.mark_non_differentiable(mean)
        return tokens, mean

    @staticmethod
    def backward(ctx, dtokens, _dmean):
        lib = L_.load()
        x, ln_w, proj, saved = ctx.saved_tensors
        B, L1, W, D, dt, tdt, tok_dtype, nscratch = ctx.geom
        dev = x.device
        if dtokens.dtype != tok_dtype:
            dtokens = dtokens.to(tok_dtype)
        if dtokens.stride(2) != 1 or (tdt == L_.SIG_F32 and dtokens.stride(0) != L1 * dtokens.stride(1)):
            dtokens = dtokens.contiguous()
        dx = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=dev) if _dense(x) else torch.empty_like(x, memory_format=torch.contiguous_format)
        d_ln_w = torch.empty(W, dtype=torch.float32, device=dev)
        d_ln_b = torch.empty(W, dtype=torch.float32, device=dev)
        d_proj = torch.empty(W, D, dtype=torch.float32, device=dev)
        scratch = _u8(nscratch, dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_tokens_bwd(x.data_ptr(), dt, x.stride(0), x.stride(1), B, L1, W, D, ln_w.data_ptr(), proj.data_ptr(),
                                        dtokens.data_ptr(), tdt, dtokens.stride(0), dtokens.stride(1), saved.data_ptr(), saved.numel(),
                                        dx.data_ptr(), dx.stride(0), dx.stride(1), d_ln_w.data_ptr(), d_ln_b.data_ptr(),
                                        d_proj.data_ptr(), scratch.data_ptr(), scratch.numel(), dev.index, L_.stream_ptr(dev)),
                     "sig_tokens_bwd")
        return dx, d_ln_w, d_ln_b, d_proj, None, None, None


def _dense(x: torch.Tensor) -> bool:
    """True when x's strides are a permutation of a dense layout (so a gradient with the same strides has no holes)."""
    n = 1
    for size, stride in sorted(zip(x.shape, x.stride()), key=lambda p: p[1]):
        if size == 1:
            continue
        if stride != n:
            return False
        n *= size
    return True


class TokenProducer(nn.Module):
    """``ln_post`` + ``@ proj`` + CLS / patch split of the reference's vision tower (clip/model.py:485-487,
    meta_arch.py:108-110) on the B200 kernels.  ``forward(x) -> (x_cash [B,L,d], global_feat [B,d])``, both views of one
    ``[B, 1+L, d]`` map -- exactly the strided views the head consumes without a copy.  ``tokens(x)`` returns the whole map."""

    def __init__(self, ln_post: nn.LayerNorm, proj: torch.Tensor, patch_mean: bool = True):
        super().__init__()
        if tuple(ln_post.normalized_shape) != (proj.shape[0],) or not ln_post.elementwise_affine:
            raise ValueError("TokenProducer: ln_post must be an affine LayerNorm over proj.shape[0] channels")
        # plain attributes, not registered sub-modules / parameters: the tower keeps owning them (state_dict unchanged)
        object.__setattr__(self, "_ln", ln_post)
        object.__setattr__(self, "_proj", proj)
        self.want_patch_mean = patch_mean
        self.last_patch_mean = None

    def tokens(self, x: torch.Tensor) -> torch.Tensor:
        ln = self._ln
        # under torch.autocast the reference's tail is layer_norm in fp32 followed by a half matmul: fp32 input (a residual
        # stream that autocast keeps in fp32) then yields tokens of the autocast dtype, exactly like `ln_post(x) @ proj`
        tok_dtype = None
        if x.is_cuda and x.dtype == torch.float32 and torch.is_autocast_enabled("cuda"):
            tok_dtype = torch.get_autocast_dtype("cuda")
        tok, mean = TokenProducerFunction.apply(x, ln.weight, ln.bias, self._proj, ln.eps, self.want_patch_mean, tok_dtype)
        self.last_patch_mean = mean if self.want_patch_mean else None
        return tok

    def forward(self, x: torch.Tensor):
        tok = self.tokens(x)
        return tok[:, 1:], tok[:, 0]
