"""CPU oracle for the Signal fusion head (SIM + GAM + LAM).

TEST INFRASTRUCTURE ONLY.  Nothing in ``signal_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or as the
reported CPU baseline -- never as the product path.

This is an independent restatement, from the math, of the reference's fusion
head.  Every function cites the reference file:line it follows (paths relative
to the upstream repository root).  It is written with plain torch CPU tensor
ops (matmul / exp / erf / sort / gather) instead of the high level ``nn``
modules the reference uses, so that it pins *semantics* (explicit multi-head
attention, explicit bilinear sampling, closed-form 3x3 Gram determinant,
rank-based top-k) rather than re-calling the same library entry points.

Parity pin: the reference ships no golden vectors for this path (its
``tests/`` exercise unrelated packages), so the oracle is pinned against the
reference *itself*, run in the build container:
``tests/golden/make_golden.py`` imports the live reference modules, feeds them
seeded inputs/parameters and stores outputs + gradient projections in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays the oracle on
the same seeds and compares.  Tie handling is the one place the oracle is
*stricter* than the reference: ``torch.topk`` leaves the order of equal
scores unspecified, the oracle (and the CUDA path) select the lowest index.

Parameters are passed as a flat ``dict`` that uses the reference's
``state_dict`` key names (``token_selection.W_q.weight`` ...,
``DAS_r.proj_q.weight`` ...), so a reference checkpoint can be fed directly.
All functions are differentiable by torch autograd (used as the gradient
oracle) and dtype generic (fp32 for parity, fp64 for gradient checks).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

NUM_HEADS = 8          # useA.py:449 (num_heads = 8)
LN_EPS = 1e-5          # nn.LayerNorm default, useA.py:414
LABEL_SMOOTHING = 0.1  # useB.py:121-122
DAS_STRIDE = 4         # useB.py:66
DAS_KSIZE = 4          # useB.py:68
DAS_RANGE_FACTOR = 2   # useB.py:67


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def gelu_exact(x: Tensor) -> Tensor:
    """nn.GELU() (erf form) -- useA.py:356, DAS.py:59,63."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """x W^T + b with the leading dims folded into one 2-D matmul (what nn.Linear / a 1x1 conv compute)."""
    y = x.reshape(-1, x.shape[-1]) @ weight.t()
    if bias is not None:
        y = y + bias
    return y.reshape(*x.shape[:-1], weight.shape[0])


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """fp32 LayerNorm over the last dim, biased variance -- useA.py:414-423."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * weight + bias


def topk_mask_lowest_index(scores: Tensor, k: int) -> Tensor:
    """Boolean mask of the k largest entries per row, ties -> lowest index.

    rank_i = #{j : s_j > s_i or (s_j == s_i and j < i)}; keep rank_i < k.
    Restates ``torch.topk(scores, k)`` + index scatter (useA.py:79-93,155-218)
    with a *defined* tie order (the reference's is unspecified).
    """
    n = scores.shape[-1]
    k = min(k, n)
    order = torch.sort(scores, dim=-1, descending=True, stable=True).indices
    mask = torch.zeros_like(scores, dtype=torch.bool)
    mask.scatter_(-1, order[..., :k], True)
    return mask


# --------------------------------------------------------------------------
# SIM -- token selection (useA.py:16-325)
# --------------------------------------------------------------------------
def intra_modal_scores(patches: Tensor, cls: Tensor) -> Tensor:
    """softmax_L(cls . patch^T / sqrt(d)) -- useA.py:72-74."""
    d = patches.shape[-1]
    logits = torch.einsum("bd,bld->bl", cls, patches) / math.sqrt(d)
    return torch.softmax(logits, dim=1)


def intra_modal_masks(patches: Sequence[Tensor], cls: Sequence[Tensor], k1: int):
    """useA.py:50-96 -> three boolean masks [B, L]."""
    return [topk_mask_lowest_index(intra_modal_scores(p, c), k1)
            for p, c in zip(patches, cls)]


def inter_modal_scores(params: Params, patches: Sequence[Tensor],
                       cls: Sequence[Tensor], prefix: str = "token_selection.") -> Tensor:
    """S = softmax_j((W_q Q)(W_k K)^T / sqrt(d)) over all 3L keys -- useA.py:116-129.

    Returns [B, 3, 3L].
    """
    d = patches[0].shape[-1]
    queries = torch.stack(list(cls), dim=1)                       # useA.py:116
    keys = torch.cat(list(patches), dim=1)                        # useA.py:120
    q = linear(queries, params[prefix + "W_q.weight"], params[prefix + "W_q.bias"])   # :123
    k = linear(keys, params[prefix + "W_k.weight"], params[prefix + "W_k.bias"])      # :124
    s = torch.einsum("bqd,bjd->bqj", q, k) / math.sqrt(d)         # :128
    return torch.softmax(s, dim=2)                                # :129


def inter_modal_masks_from_scores(scores: Tensor, L: int, k2: int):
    """useA.py:136-218: per query keep the other two modalities' 2L scores,
    top-k2, and scatter back onto the *selected* modalities' masks."""
    s = scores
    d_rgb = torch.cat([s[:, 0, L:2 * L], s[:, 0, 2 * L:]], dim=1)   # RGB cls -> NIR, TIR
    d_nir = torch.cat([s[:, 1, :L], s[:, 1, 2 * L:]], dim=1)        # NIR cls -> RGB, TIR
    d_tir = torch.cat([s[:, 2, :L], s[:, 2, L:2 * L]], dim=1)       # TIR cls -> RGB, NIR
    sel_rgb = topk_mask_lowest_index(d_rgb, k2)
    sel_nir = topk_mask_lowest_index(d_nir, k2)
    sel_tir = topk_mask_lowest_index(d_tir, k2)
    rgb_mask = sel_nir[:, :L] | sel_tir[:, :L]      # useA.py:199-202, 216-219
    nir_mask = sel_rgb[:, :L] | sel_tir[:, L:]      # useA.py:182-185, 218-221
    tir_mask = sel_rgb[:, L:] | sel_nir[:, L:]      # useA.py:184-185, 201-202
    return [rgb_mask, nir_mask, tir_mask]


def keep_ratio_adjust(mask: Tensor, raw_score: Tensor, max_keep: int) -> Tensor:
    """useA.py:254-314: force exactly ``max_keep`` kept tokens per row.

    Too many -> keep the top-``max_keep`` *of the selected* by the raw dot
    ``cls . patch`` (no 1/sqrt(d), no softmax, :259-261); too few -> add the
    best *unselected*.  Equivalent: order by (selected desc, raw desc, index
    asc) and keep the first ``max_keep``.
    """
    B, L = mask.shape
    out = torch.zeros_like(mask)
    for b in range(B):                       # small, test-only loop
        sel = mask[b]
        idx = torch.arange(L)
        key = sorted(idx.tolist(), key=lambda i: (0 if sel[i] else 1, -float(raw_score[b, i]), i))
        out[b, torch.tensor(key[:max_keep], dtype=torch.long)] = True
    return out


def token_selection_masks(params: Params, patches: Sequence[Tensor], cls: Sequence[Tensor],
                          k: int, keep_ratio: Optional[float] = None,
                          prefix: str = "token_selection."):
    """TokenSelection.forward up to the union -- useA.py:223-314.  Boolean [B, L] x3."""
    L = patches[0].shape[1]
    with torch.no_grad():
        inter = inter_modal_masks_from_scores(
            inter_modal_scores(params, patches, cls, prefix), L, 2 * k)
        intra = intra_modal_masks(patches, cls, k)
        masks = [a | b for a, b in zip(inter, intra)]                 # :249-251
        if keep_ratio is not None:
            max_keep = int(L * keep_ratio)                            # :256
            masks = [keep_ratio_adjust(m, torch.einsum("bd,bld->bl", c, p), max_keep)
                     for m, p, c in zip(masks, patches, cls)]
    return masks


def token_selection(params, patches, cls, k, keep_ratio=None, prefix="token_selection."):
    """TokenSelection.forward -- useA.py:223-325.  Returns (selected x3, masks x3 [B,L,1])."""
    masks = token_selection_masks(params, patches, cls, k, keep_ratio, prefix)
    fmasks = [m.to(patches[0].dtype).unsqueeze(-1) for m in masks]
    selected = [p * m for p, m in zip(patches, fmasks)]               # :318-320
    return selected, fmasks


# --------------------------------------------------------------------------
# SIM -- modal interaction (useA.py:328-411)
# --------------------------------------------------------------------------
def multi_head_cross_attention(params: Params, queries: Tensor, kv: Tensor,
                               prefix: str = "modal_interactive.cross_attn.") -> Tensor:
    """nn.MultiheadAttention(dim, 8, batch_first=True)(Q, KV, KV)[0] -- useA.py:351,388.

    Packed in-proj rows 0:d = q, d:2d = k, 2d:3d = v; q scaled by
    sqrt(1/head_dim) after projection (torch/nn/functional.py
    multi_head_attention_forward); no dropout, no masks.
    """
    B, nq, d = queries.shape
    hd = d // NUM_HEADS
    w = params[prefix + "in_proj_weight"]
    b = params[prefix + "in_proj_bias"]
    q = linear(queries, w[:d], b[:d])
    k = linear(kv, w[d:2 * d], b[d:2 * d])
    v = linear(kv, w[2 * d:], b[2 * d:])
    q = q.reshape(B, nq, NUM_HEADS, hd).transpose(1, 2) * math.sqrt(1.0 / hd)
    k = k.reshape(B, -1, NUM_HEADS, hd).transpose(1, 2)
    v = v.reshape(B, -1, NUM_HEADS, hd).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, nq, d)
    return linear(o, params[prefix + "out_proj.weight"], params[prefix + "out_proj.bias"])


def modal_interactive(params: Params, selected: Sequence[Tensor], cls: Sequence[Tensor],
                      prefix: str = "modal_interactive.") -> Tensor:
    """ModalInteractive.forward -- useA.py:364-411 -> [B, 3d]."""
    queries = torch.stack(list(cls), dim=1)                            # :379
    kv = torch.cat(list(selected), dim=1)                              # :383
    attn = multi_head_cross_attention(params, queries, kv, prefix + "cross_attn.")  # :388
    y1 = layer_norm(queries + attn, params[prefix + "norm1.weight"], params[prefix + "norm1.bias"])  # :393
    h = gelu_exact(linear(y1, params[prefix + "ffn.0.weight"], params[prefix + "ffn.0.bias"]))
    f = linear(h, params[prefix + "ffn.2.weight"], params[prefix + "ffn.2.bias"])      # :397
    y2 = layer_norm(y1 + f, params[prefix + "norm2.weight"], params[prefix + "norm2.bias"])  # :401
    return torch.cat([y2[:, 0], y2[:, 1], y2[:, 2]], dim=1)            # :408


def sim_forward(params: Params, patches, cls, k: int, keep_ratio=None):
    """Select_Interactive_Module.forward -- useA.py:454-476.  Returns (out [B,3d], masks)."""
    selected, masks = token_selection(params, patches, cls, k, keep_ratio)
    return modal_interactive(params, selected, cls), masks


# --------------------------------------------------------------------------
# GAM -- Gram volume + contrastive CE (utils/volume.py:14-62, useB.py:76-126)
# --------------------------------------------------------------------------
def _volume3_f64(language: Tensor, video: Tensor, audio: Tensor) -> Tensor:
    """sqrt|det G| evaluated in fp64 whatever the input precision (inputs are up-cast exactly).

    The checker deliberately does NOT reproduce the reference's fp32 determinant (``torch.det(G.float())``, volume.py:57):
    the determinant cancels when the modalities align, and d(gam)/d(contra_temp) cancels to ~1e-3 of its terms, so an
    fp32 evaluation carries ~3e-4 of rounding noise on that gradient -- more than the 1e-4 parity tolerance it is used
    to check.  The algorithm is the reference's; only the arithmetic is wider."""
    l, v, a = language.double(), video.double(), audio.double()
    ll = (l * l).sum(-1)[:, None]
    vv = (v * v).sum(-1)[None, :]
    aa = (a * a).sum(-1)[None, :]
    va = (v * a).sum(-1)[None, :]
    lv = l @ v.T
    la = l @ a.T
    det = ll * (vv * aa - va * va) - lv * (lv * aa - va * la) + la * (lv * va - vv * la)
    return torch.sqrt(torch.abs(det))


def volume3(language: Tensor, video: Tensor, audio: Tensor) -> Tensor:
    """sqrt|det G| for G = Gram(l_i, v_j, a_j), closed form -- volume.py:35-60.

    det = ll(vv*aa - va^2) - lv(lv*aa - va*la) + la(lv*va - vv*la)
    (cofactor expansion of the symmetric 3x3 the reference stacks at :47-52).
    Output [B1, B2], fp32 like ``G.float()`` at :57 (fp64 for fp64 inputs); evaluated in fp64 (_volume3_f64).
    """
    V = _volume3_f64(language, video, audio)
    return V if language.dtype == torch.float64 else V.float()


def volume_n(language: Tensor, *others: Tensor) -> Tensor:
    """volume_computation4 / 5 -- volume.py:65-116, :119-182: sqrt|det G| of the n x n Gram matrix of (language_i, o_1_j, ...,
    o_{n-1}_j), n = 1 + len(others).  Restated without torch.det: Laplace expansion along the first row, recursively, in fp64
    (like _volume3_f64 the checker keeps the arithmetic wider than the reference's ``G.float()``).  Differentiable."""
    feats = [language.double()] + [o.double() for o in others]
    n = len(feats)
    B1, B2 = feats[0].shape[0], feats[1].shape[0]
    G = [[None] * n for _ in range(n)]
    G[0][0] = (feats[0] * feats[0]).sum(-1)[:, None].expand(B1, B2)
    for k in range(1, n):
        G[0][k] = G[k][0] = feats[0] @ feats[k].T
        for l in range(k, n):
            G[k][l] = G[l][k] = (feats[k] * feats[l]).sum(-1)[None, :].expand(B1, B2)

    def det(rows, cols):
        if len(rows) == 1:
            return G[rows[0]][cols[0]]
        total = 0.0
        for idx, c in enumerate(cols):
            minor = det(rows[1:], cols[:idx] + cols[idx + 1:])
            total = total + (-1.0) ** idx * G[rows[0]][c] * minor
        return total

    V = torch.sqrt(torch.abs(det(list(range(n)), list(range(n)))))
    return V if language.dtype == torch.float64 else V.float()


def _ce_label_smoothing(logits: Tensor) -> Tensor:
    """F.cross_entropy(logits, arange(B), label_smoothing=0.1), mean reduction."""
    B = logits.shape[0]
    logp = torch.log_softmax(logits, dim=1)
    nll = -logp.diagonal().mean()
    smooth = -logp.mean(dim=1).mean()
    return (1.0 - LABEL_SMOOTHING) * nll + LABEL_SMOOTHING * smooth


def gam_loss(patches: Sequence[Tensor], contra_temp: Tensor) -> Tensor:
    """AlignmentM.Cls_Align -- useB.py:76-126."""
    feats = []
    for p in patches:
        m = p.mean(dim=1)                                              # :92-94
        feats.append(m / m.norm(dim=-1, keepdim=True).clamp_min(1e-12))  # :98-100
    V = _volume3_f64(*feats)                                           # :106 (fp64 arithmetic, see _volume3_f64)
    vol = V / contra_temp.double()                                     # :107
    loss = 0.5 * (_ce_label_smoothing(-vol) + _ce_label_smoothing(-vol.T))  # :120-124
    return loss.to(patches[0].dtype if patches[0].dtype == torch.float64 else torch.float32)


# --------------------------------------------------------------------------
# LAM -- deformable sampling + pairwise MSE (DAS.py:17-165, useB.py:128-167)
# --------------------------------------------------------------------------
def das_offsets(params: Params, x_tok: Tensor, h: int, w: int, prefix: str) -> Tensor:
    """Offset net of DA_sample on channels-last tokens ``x_tok`` [B, h*w, d].

    proj_q 1x1 (DAS.py:129) -> conv_offset: 1x1, GELU, depthwise 4x4 s4 p0,
    GELU, 1x1 -> 1 channel, no bias (DAS.py:55-66,136).  Returns the raw
    (pre-tanh) offset logit ``o`` [B, Hk, Wk].
    """
    B, L, d = x_tok.shape
    wq = params[prefix + "proj_q.weight"].reshape(d, d)
    q = linear(x_tok, wq, params[prefix + "proj_q.bias"])
    w0 = params[prefix + "conv_offset.0.weight"].reshape(d, d)
    g = gelu_exact(linear(q, w0, params[prefix + "conv_offset.0.bias"]))   # [B, L, d]
    Hk, Wk = h // DAS_STRIDE, w // DAS_STRIDE
    g = g.reshape(B, Hk, DAS_KSIZE, Wk, DAS_KSIZE, d)
    wdw = params[prefix + "conv_offset.2.weight"].reshape(d, DAS_KSIZE, DAS_KSIZE)
    u = torch.einsum("bpiqjc,cij->bpqc", g, wdw) + params[prefix + "conv_offset.2.bias"]
    u = gelu_exact(u)
    return torch.einsum("bpqc,c->bpq", u, params[prefix + "conv_offset.4.weight"].reshape(d))


def das_positions(o: Tensor) -> Tuple[Tensor, Tensor]:
    """Normalised sample coordinates (y, x) in [-1, 1] -- DAS.py:140-153, 91-105.

    One scalar ``o`` per sample point drives *both* axes (1->2 channel
    broadcast at :145-146): off_y = 2 tanh(o)/(Hk-1), off_x = 2 tanh(o)/(Wk-1).
    ref = ((i + 0.5)/(n-1))*2 - 1, which exceeds +1 for the last cell; clamp.
    """
    B, Hk, Wk = o.shape
    t = torch.tanh(o)
    ref_y = (torch.arange(Hk, dtype=o.dtype, device=o.device) + 0.5) / (Hk - 1.0) * 2.0 - 1.0
    ref_x = (torch.arange(Wk, dtype=o.dtype, device=o.device) + 0.5) / (Wk - 1.0) * 2.0 - 1.0
    py = (t * (DAS_RANGE_FACTOR / (Hk - 1.0)) + ref_y[None, :, None]).clamp(-1.0, 1.0)
    px = (t * (DAS_RANGE_FACTOR / (Wk - 1.0)) + ref_x[None, None, :]).clamp(-1.0, 1.0)
    return py, px


def bilinear_sample_tokens(x_tok: Tensor, py: Tensor, px: Tensor, h: int, w: int) -> Tensor:
    """F.grid_sample(bilinear, zeros padding, align_corners=True) -- DAS.py:158-163,
    coordinate un-normalisation ((c+1)/2)*(size-1) as in ATen GridSampler.h.

    x_tok [B, h*w, d] channels-last; py/px [B, Hk, Wk] -> [B, Hk*Wk, d].
    """
    B, L, d = x_tok.shape
    iy = (py + 1.0) * 0.5 * (h - 1)
    ix = (px + 1.0) * 0.5 * (w - 1)
    y0 = torch.floor(iy.detach())
    x0 = torch.floor(ix.detach())
    wy = iy - y0
    wx = ix - x0
    out = 0
    for dy, dx, wgt in ((0, 0, (1 - wy) * (1 - wx)), (0, 1, (1 - wy) * wx),
                        (1, 0, wy * (1 - wx)), (1, 1, wy * wx)):
        yy = (y0 + dy).long()
        xx = (x0 + dx).long()
        inb = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        lin = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).reshape(B, -1)
        v = torch.gather(x_tok, 1, lin[..., None].expand(-1, -1, d))
        out = out + v * (wgt * inb.to(x_tok.dtype)).reshape(B, -1, 1)
    return out


def da_sample(params: Params, x_tok: Tensor, h: int, w: int, prefix: str) -> Tensor:
    """DA_sample.forward on channels-last tokens -- DAS.py:107-165 -> [B, Hk*Wk, d]."""
    o = das_offsets(params, x_tok, h, w, prefix)
    py, px = das_positions(o)
    return bilinear_sample_tokens(x_tok, py, px, h, w)


def lam_loss(params: Params, patches: Sequence[Tensor], h: int, w: int) -> Tensor:
    """AlignmentM.patch_Align -- useB.py:128-167."""
    s_r = da_sample(params, patches[0], h, w, "DAS_r.")
    s_n = da_sample(params, patches[1], h, w, "DAS_n.")
    s_t = da_sample(params, patches[2], h, w, "DAS_t.")
    mse = lambda a, b: ((a - b) ** 2).mean()
    return (mse(s_n, s_r) + mse(s_t, s_r) + mse(s_t, s_n)) / 3.0      # :161-165


# --------------------------------------------------------------------------
# whole head (make_model.py:181-207 call sites)
# --------------------------------------------------------------------------
def head_forward(sim_params: Params, align_params: Params, tokens: Sequence[Tensor],
                 k: int, h: int, w: int, keep_ratio=None, stage: str = "together_CLS_Patch"):
    """tokens: three [B, 1+h*w, d] maps; cls = row 0, patches = rows 1.. (meta_arch.py:108-109).

    Returns (sim_out [B,3d], gam_loss, lam_loss or None, masks).
    """
    patches = [t[:, 1:] for t in tokens]
    cls = [t[:, 0] for t in tokens]
    out, masks = sim_forward(sim_params, patches, cls, k, keep_ratio)
    gam = gam_loss(patches, align_params["contra_temp"])
    lam = None if stage == "CLS" else lam_loss(align_params, patches, h, w)  # useB.py:181-190
    return out, gam, lam, masks
