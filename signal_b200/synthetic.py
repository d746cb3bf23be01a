"""Seeded synthetic tokens and parameters for tests, smoke() and bench.py.

Everything is generated on the CPU with ``torch.Generator`` so that the build
container (where the golden vectors are produced from the live reference) and
the GPU box (where ``/root/reference`` does not exist) see identical values.

Parameter dicts use the reference's ``state_dict`` key names
(``Select_Interactive_Module`` -> modeling/AddModule/useA.py:33-48,340-361,442-452;
``AlignmentM`` -> modeling/AddModule/useB.py:44-74, DAS.py:30-72).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Optional, Tuple

import torch

Shapes = Dict[str, Tuple[int, ...]]


def sim_param_shapes(d: int) -> Shapes:
    """state_dict layout of Select_Interactive_Module(dim=d)."""
    return {
        "token_selection.W_q.weight": (d, d), "token_selection.W_q.bias": (d,),
        "token_selection.W_k.weight": (d, d), "token_selection.W_k.bias": (d,),
        "token_selection.W_v.weight": (d, d), "token_selection.W_v.bias": (d,),
        "modal_interactive.cross_attn.in_proj_weight": (3 * d, d),
        "modal_interactive.cross_attn.in_proj_bias": (3 * d,),
        "modal_interactive.cross_attn.out_proj.weight": (d, d),
        "modal_interactive.cross_attn.out_proj.bias": (d,),
        "modal_interactive.ffn.0.weight": (2 * d, d), "modal_interactive.ffn.0.bias": (2 * d,),
        "modal_interactive.ffn.2.weight": (d, 2 * d), "modal_interactive.ffn.2.bias": (d,),
        "modal_interactive.norm1.weight": (d,), "modal_interactive.norm1.bias": (d,),
        "modal_interactive.norm2.weight": (d,), "modal_interactive.norm2.bias": (d,),
    }


def align_param_shapes(d: int) -> Shapes:
    """state_dict layout of AlignmentM(feat_dim=d, H, W) (with d-channel DAS)."""
    s: Shapes = {"contra_temp": ()}
    for m in ("DAS_r", "DAS_n", "DAS_t"):
        s[f"{m}.conv_offset.0.weight"] = (d, d, 1, 1)
        s[f"{m}.conv_offset.0.bias"] = (d,)
        s[f"{m}.conv_offset.2.weight"] = (d, 1, 4, 4)
        s[f"{m}.conv_offset.2.bias"] = (d,)
        s[f"{m}.conv_offset.4.weight"] = (1, d, 1, 1)
        s[f"{m}.proj_q.weight"] = (d, d, 1, 1)
        s[f"{m}.proj_q.bias"] = (d,)
    return s


def _key_seed(seed: int, key: str) -> int:
    return (seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1)


def make_params(shapes: Shapes, seed: int, offset_gain: float = 1.0,
                dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic parameters, one independent stream per key.

    weights ~ N(0, (0.6/sqrt(fan_in))^2), biases ~ N(0, 0.1^2), LayerNorm weight
    1 + 0.1 N(0,1), contra_temp = 0.07 (useB.py:56).  ``offset_gain`` scales
    ``conv_offset.4.weight`` so tests can push the DAS offsets into the
    tanh / clamp saturated regime.
    """
    out = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed(_key_seed(seed, key))
        if key == "contra_temp":
            t = torch.tensor(0.07, dtype=torch.float64)
        elif key.endswith("norm1.weight") or key.endswith("norm2.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif key.endswith("bias"):
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g, dtype=torch.float64) * (0.6 / math.sqrt(fan_in))
            if key.endswith("conv_offset.4.weight"):
                t = t * offset_gain
        out[key] = t.to(dtype)
    return out


def smooth_patches(toks, h: int, w: int, passes: int = 2):
    """Low-pass the patch rows of each token map over the h x w grid ([1,2,1]/4 separable filter,
    edge-replicated, `passes` times) and rescale to unit variance.  White-noise features make the
    deformable sampling of LAM chaotic w.r.t. its offsets; smooth maps give a well-conditioned
    problem for reduced-precision parity checks."""
    out = []
    for t in toks:
        B, L1, d = t.shape
        x = t[:, 1:].reshape(B, h, w, d).float()
        for _ in range(passes):
            xp = torch.cat([x[:, :1], x, x[:, -1:]], dim=1)
            x = 0.25 * xp[:, :-2] + 0.5 * xp[:, 1:-1] + 0.25 * xp[:, 2:]
            xp = torch.cat([x[:, :, :1], x, x[:, :, -1:]], dim=2)
            x = 0.25 * xp[:, :, :-2] + 0.5 * xp[:, :, 1:-1] + 0.25 * xp[:, :, 2:]
        x = x / x.std()
        t2 = t.clone().float()
        t2[:, 1:] = x.reshape(B, h * w, d)
        out.append(t2.to(t.dtype))
    return out


def make_tokens(B: int, d: int, seed: int = 1234, L: int = 128, structured: bool = False,
                dtype: torch.dtype = torch.float32):
    """Three [B, 1+L, d] token maps in (RGB, NI, TI) order.

    iid: randn.  structured: a shared per-sample direction ``u`` plus noise, which
    makes the modalities correlated (GAM conditioning) and the selections overlap.
    Values are generated in fp32 and rounded once to ``dtype``.
    """
    g = torch.Generator().manual_seed(seed)
    toks = []
    if not structured:
        for _ in range(3):
            toks.append(torch.randn(B, 1 + L, d, generator=g))
    else:
        u = torch.randn(B, d, generator=g)
        a = torch.rand(B, L, generator=g)
        for _ in range(3):
            t = torch.empty(B, 1 + L, d)
            t[:, 0] = u + 0.25 * torch.randn(B, d, generator=g)
            t[:, 1:] = a[..., None] * u[:, None] + 0.5 * torch.randn(B, L, d, generator=g)
            toks.append(t)
    return [t.to(dtype) for t in toks]


def make_cotangent(B: int, d: int, seed: int = 4321) -> torch.Tensor:
    """Random cotangent for the [B, 3d] SIM output (an all-ones cotangent has an
    identically zero gradient through the final LayerNorm)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3 * d, generator=g)


def probe_vector(key: str, numel: int, seed: int = 7) -> torch.Tensor:
    """[4, numel] fixed random probes used to fingerprint a gradient tensor."""
    g = torch.Generator().manual_seed(_key_seed(seed, "probe:" + key))
    return torch.randn(4, numel, generator=g, dtype=torch.float64)


def token_projection(d: int, seed: int = 11, cols: int = 4) -> torch.Tensor:
    """[d, cols] projection used to fingerprint per-token gradients."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(d, cols, generator=g, dtype=torch.float64) / math.sqrt(d)
