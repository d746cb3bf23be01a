"""CPU oracle for the token producer in front of the fusion head.

TEST INFRASTRUCTURE ONLY (same rules as oracle/signal_oracle.py: only tests/, __graft_entry__.smoke() and the
cpu_baseline legs of bench.py may import this; never the product path).

Independent restatement, from the math, with explicit forward AND backward formulas (no autograd, no nn.LayerNorm) of
    x = self.ln_post(x); xproj = x @ self.proj          modeling/clip/model.py:485-487
    LayerNorm: statistics in fp32, result cast back     modeling/clip/model.py:154-160
    global_feat = x[:, 0]; x_cash = x[:, 1:]            modeling/meta_arch.py:108-110
Parity pin: tests/golden/make_tokens_golden.py runs the LIVE reference `VisionTransformer` (imported from /root/reference in
the build container), captures the input of its `ln_post`, its output and the gradients autograd returns for a seeded
cotangent, and stores them in tests/golden/tokens_*.npz; tests/test_tokens_oracle_golden.py replays this file against them.
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor


def layernorm_stats(x: Tensor, eps: float):
    """mean and 1/sqrt(biased variance + eps) over the last axis (nn.LayerNorm semantics)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return mu, 1.0 / torch.sqrt(var + eps)


def tokens_fwd(x: Tensor, ln_w: Tensor, ln_b: Tensor, proj: Tensor, eps: float = 1e-5, operand_dtype=None):
    """x [B, 1+L, W] -> (tokens [B, 1+L, D], patch_mean [B, D], xn).  `operand_dtype` (e.g. torch.bfloat16) rounds the GEMM
    operands and the result the way autocast does (clip/model.py:159-160 returns the input dtype; :487 is a half matmul)."""
    mu, rstd = layernorm_stats(x, eps)
    xn = (x - mu) * rstd * ln_w + ln_b
    if operand_dtype is not None:
        xn = xn.to(operand_dtype).to(x.dtype)
        proj = proj.to(operand_dtype).to(x.dtype)
    tok = torch.einsum("blw,wd->bld", xn, proj)
    if operand_dtype is not None:
        tok = tok.to(operand_dtype).to(x.dtype)
    return tok, tok[:, 1:].mean(dim=1), xn


def split(tokens: Tensor):
    """meta_arch.py:108-110 -> (x_cash, global_feat)"""
    return tokens[:, 1:], tokens[:, 0]


def tokens_bwd(x: Tensor, ln_w: Tensor, proj: Tensor, dtok: Tensor, eps: float = 1e-5, xn: Tensor = None, operand_dtype=None):
    """-> (dx, d_ln_w, d_ln_b, d_proj) for cotangent dtok of the tokens.
    d xn = dtok proj^T; d proj = xn^T dtok; with xhat = (x - mu) rstd and g = d xn * gamma:
    dx = rstd (g - mean(g) - xhat mean(g xhat)); d gamma = sum d xn * xhat; d beta = sum d xn."""
    W = x.shape[-1]
    mu, rstd = layernorm_stats(x, eps)
    xhat = (x - mu) * rstd
    if xn is None:
        raise ValueError("pass the xn returned by tokens_fwd (the saved GEMM operand)")
    pj = proj
    if operand_dtype is not None:
        pj = proj.to(operand_dtype).to(x.dtype)
        dtok = dtok.to(operand_dtype).to(x.dtype)
    dxn = torch.einsum("bld,wd->blw", dtok, pj)
    if operand_dtype is not None:
        dxn = dxn.to(operand_dtype).to(x.dtype)
    d_proj = torch.einsum("blw,bld->wd", xn, dtok)
    g = dxn * ln_w
    c1 = g.sum(dim=-1, keepdim=True) / W
    c2 = (g * xhat).sum(dim=-1, keepdim=True) / W
    dx = rstd * (g - c1 - xhat * c2)
    d_ln_w = (dxn * xhat).sum(dim=(0, 1))
    d_ln_b = dxn.sum(dim=(0, 1))
    return dx, d_ln_w, d_ln_b, d_proj
