"""Full Signal training step around the B200 fusion head (BASELINE.json configs[3]: MSVR310).

What the reference runs per iteration (modeling/make_model.py:148-255, engine/processor.py:165-261,
configs/MSVR310/Signal.yml): three images per sample -> ONE shared CLIP ViT-B/16 vision tower (clip/model.py:419-487,
SIE camera embedding, meta_arch.py:95-110) -> per modality [B,129,512] tokens -> SIM (vars_total [B,1536]) + AlignM (GAM, LAM)
-> BNNeck + classifier on each CLS and on vars_total (DIRECT: 0) -> sum over the four heads of
0.25 * label-smoothed ID loss + 1.0 * soft-margin triplet loss (layers/make_loss.py:109-161) + 0.2 * GAM + 0.01 * LAM -> Adam.

Scope: "Backbone stays stock PyTorch; only the head is ours".  ``ClipViT`` below is the
standard CLIP vision transformer written with stock torch modules (random init: there is no checkpoint and no network);
it is NOT part of the product path and exists so that the head can be timed inside a complete training step.  bench.py
times the step twice: with the B200 head (FusionHead + the loss kernels, the default here) and with a stock-PyTorch head
that it plugs in through ``torch_head`` (the reference's algorithm in plain torch ops; bench.py's comparison arm).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import losses as LS
from . import modules as M


class _QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class _Block(nn.Module):
    """pre-LN residual attention block of CLIP (clip/model.py:168-180, the 'nothing' pattern the shipped configs use)"""

    def __init__(self, width, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(width, heads, batch_first=True)
        self.ln_1 = nn.LayerNorm(width)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(width, 4 * width)), ("gelu", _QuickGELU()),
                                              ("c_proj", nn.Linear(4 * width, width))]))
        self.ln_2 = nn.LayerNorm(width)

    def forward(self, x):
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class ClipViT(nn.Module):
    """CLIP ViT vision tower: conv patch embedding, class token, positional embedding, ln_pre, `layers` blocks, ln_post,
    projection to `out_dim` (clip/model.py:419-487).  forward(x [B,3,H,W], cv_emb [B,1,width] or None) -> [B, 1+h*w, out_dim]."""

    def __init__(self, h, w, patch=16, width=768, layers=12, heads=12, out_dim=512):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, patch, patch, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(h * w + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.blocks = nn.Sequential(*[_Block(width, heads) for _ in range(layers)])
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, out_dim))
        self.token_producer = None   # signal_b200.tokens.TokenProducer: the tail on the B200 kernels (N3), set by the caller

    def forward(self, x, cv_emb=None):
        x = self.conv1(x).flatten(2).transpose(1, 2)
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        if cv_emb is not None:
            cls = cls + cv_emb.to(x.dtype)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.blocks(self.ln_pre(x))
        if self.token_producer is not None:
            return self.token_producer.tokens(x)
        return self.ln_post(x) @ self.proj.to(x.dtype)


class SignalTrainStep(nn.Module):
    """Backbone + fusion head + BNNeck heads + losses of one training iteration (DIRECT: 0, USE_A, USE_B,
    stage 'together_CLS_Patch').  Default: signal_b200 modules and loss kernels.  ``torch_head`` (optional): a callable
    ``(step, patches, cls, target) -> (vars_total, gam, lam, head_loss_fn)`` that replaces the head and the per-head loss
    with stock torch ops -- bench.py's comparison arm."""

    def __init__(self, grid=(8, 16), topk=64, num_classes=155, camera_num=8, feat_dim=512, w_id=0.25, w_tri=1.0, w_gam=0.2, w_lam=0.01,
                 torch_head=None, layers=12, width=768, heads=12, b200_tokens=True):
        super().__init__()
        self.h, self.w = grid
        self.torch_head = torch_head
        self.cfg = dict(topk=topk, w_id=w_id, w_tri=w_tri, w_gam=w_gam, w_lam=w_lam, num_classes=num_classes)
        self.backbone = ClipViT(self.h, self.w, 16, width, layers, heads, feat_dim)
        self.cv_embed = nn.Parameter(torch.zeros(camera_num, 1, width))
        nn.init.trunc_normal_(self.cv_embed, std=0.02)
        self.SIM = M.Select_Interactive_Module(feat_dim, k=topk)
        self.AlignM = M.AlignmentM(feat_dim, self.h, self.w)
        mk_bn = lambda n: nn.BatchNorm1d(n)
        mk_cls = lambda n: nn.Linear(n, num_classes, bias=False)
        self.bottleneck_r, self.bottleneck_n, self.bottleneck_t, self.bottleneck_var = mk_bn(feat_dim), mk_bn(feat_dim), mk_bn(feat_dim), mk_bn(3 * feat_dim)
        self.classifier_r, self.classifier_n, self.classifier_t, self.classifier_var = mk_cls(feat_dim), mk_cls(feat_dim), mk_cls(feat_dim), mk_cls(3 * feat_dim)
        for bn in (self.bottleneck_r, self.bottleneck_n, self.bottleneck_t, self.bottleneck_var):
            bn.bias.requires_grad_(False)                       # make_model.py:98-100,119
        for c in (self.classifier_r, self.classifier_n, self.classifier_t, self.classifier_var):
            nn.init.normal_(c.weight, std=0.001)                # weights_init_classifier
        self.fusion = M.FusionHead(self.SIM, self.AlignM)
        if torch_head is None and b200_tokens:
            from .tokens import TokenProducer
            # object.__setattr__: the producer wraps ln_post / proj of the backbone and must not register them twice
            object.__setattr__(self.backbone, "token_producer", TokenProducer(self.backbone.ln_post, self.backbone.proj, patch_mean=False))
        self.xent = LS.CrossEntropyLabelSmooth(num_classes)
        self.triplet = LS.TripletLoss()

    def tokens(self, imgs, cam_label):
        """three image batches -> three contiguous [B,129,feat_dim] token maps (shared backbone, one pass over 3B images)"""
        B = imgs[0].shape[0]
        cv = self.cv_embed[cam_label].repeat(3, 1, 1)
        tok = self.backbone(torch.cat(imgs, dim=0), cv)
        return [tok[m * B:(m + 1) * B] for m in range(3)]       # contiguous slices of one [3B,129,d] tensor

    def forward(self, imgs, target, cam_label):
        c = self.cfg
        toks = self.tokens(imgs, cam_label)
        patches, cls = [t[:, 1:] for t in toks], [t[:, 0] for t in toks]
        if self.torch_head is None:
            vars_total, gam, lam = self.fusion(*patches, *cls, stage="together_CLS_Patch")
            heads = [(LS.BNNeckClassifier(bn, cl)(f)[1], f) for bn, cl, f in
                     ((self.bottleneck_r, self.classifier_r, cls[0]), (self.bottleneck_n, self.classifier_n, cls[1]),
                      (self.bottleneck_t, self.classifier_t, cls[2]), (self.bottleneck_var, self.classifier_var, vars_total))]
            loss = sum(c["w_id"] * self.xent(s, target) + c["w_tri"] * self.triplet(f, target)[0] for s, f in heads)
        else:
            vars_total, gam, lam, head_loss = self.torch_head(self, patches, cls)
            heads = [(cl(bn(f.float())).float(), f) for bn, cl, f in
                     ((self.bottleneck_r, self.classifier_r, cls[0]), (self.bottleneck_n, self.classifier_n, cls[1]),
                      (self.bottleneck_t, self.classifier_t, cls[2]), (self.bottleneck_var, self.classifier_var, vars_total))]
            loss = sum(head_loss(s, f, target, c) for s, f in heads)
        return loss + c["w_gam"] * gam + c["w_lam"] * lam
