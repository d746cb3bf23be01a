"""CPU oracle for the ID / metric losses behind the fusion head.

TEST INFRASTRUCTURE ONLY (same rules as oracle/signal_oracle.py: only tests/, __graft_entry__.smoke() and the
cpu_baseline legs of bench.py may import this; never the product path).

Independent restatement, from the math, in plain torch CPU ops; differentiable by autograd (gradient oracle), dtype
generic.  Parity pin: tests/golden/make_loss_golden.py imports the LIVE reference classes
(layers/softmax_loss.py, layers/triplet_loss.py) in the build container and stores their outputs and gradients for
seeded inputs in tests/golden/losses_*.npz; tests/test_loss_oracle_golden.py replays this file against them.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

Tensor = torch.Tensor


def xent_label_smooth(logits: Tensor, targets: Tensor, num_classes: int, epsilon: float = 0.1) -> Tensor:
    """layers/softmax_loss.py:23-34.  mean over the batch of  sum_k -t_k log p_k,
    t = (1 - eps) onehot + eps / C, written as  lse - (1 - eps) z_y - (eps / C) sum_k z_k."""
    z = logits
    lse = torch.logsumexp(z, dim=1)
    zy = z[torch.arange(z.shape[0]), targets.long()]
    row = lse - (1.0 - epsilon) * zy - (epsilon / num_classes) * z.sum(dim=1)
    return row.sum() / z.shape[0]


def pairwise_euclid(x: Tensor) -> Tensor:
    """layers/triplet_loss.py:16-31 with y = x:  sqrt(clamp(|xi|^2 + |xj|^2 - 2 xi.xj, 1e-12))."""
    sq = (x * x).sum(dim=1)
    q = sq[:, None] + sq[None, :] - 2.0 * (x @ x.t())
    return q.clamp(min=1e-12).sqrt()


def hard_mining(dist: Tensor, labels: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """layers/triplet_loss.py:51-104: hardest positive (max over same label, the anchor included) and hardest negative
    (min over other labels) of every anchor; ties -> lowest index.  Returns (dist_ap, dist_an, p_idx, n_idx)."""
    same = labels[:, None] == labels[None, :]
    neg_inf = torch.full_like(dist, float("-inf"))
    pos_inf = torch.full_like(dist, float("inf"))
    dp = torch.where(same, dist, neg_inf)
    dn = torch.where(same, pos_inf, dist)
    ap = dp.max(dim=1).values
    an = dn.min(dim=1).values
    n = dist.shape[0]
    ar = torch.arange(n)
    big = torch.full((n, n), n, dtype=torch.long)
    p_idx = torch.where(dp == ap[:, None], ar[None, :].expand(n, n), big).min(dim=1).values
    n_idx = torch.where(dn == an[:, None], ar[None, :].expand(n, n), big).min(dim=1).values
    return dist[ar, p_idx], dist[ar, n_idx], p_idx, n_idx


def triplet_loss(feat: Tensor, labels: Tensor, margin: Optional[float] = None, hard_factor: float = 0.0):
    """layers/triplet_loss.py:121-135 (normalize_feature=False).  Returns (loss, dist_ap, dist_an)."""
    dist = pairwise_euclid(feat)
    ap, an, _, _ = hard_mining(dist, labels)
    ap = ap * (1.0 + hard_factor)
    an = an * (1.0 - hard_factor)
    s = an - ap
    if margin is None:
        loss = torch.nn.functional.softplus(-s).mean()        # SoftMarginLoss(s, y=1) = log(1 + exp(-s))
    else:
        loss = (margin - s).clamp(min=0).mean()               # MarginRankingLoss(margin)(an, ap, y=1)
    return loss, ap, an


def bnneck_classifier(x: Tensor, gamma: Tensor, beta: Tensor, W: Tensor, eps: float = 1e-5,
                      running: Optional[Tuple[Tensor, Tensor]] = None, momentum: float = 0.1, training: bool = True):
    """modeling/make_model.py:128-131,194-195: BatchNorm1d (batch statistics when training, biased variance for the
    normalisation, unbiased for the running update) followed by a bias-free Linear.
    Returns (feat_bn, logits, new_running_mean, new_running_var)."""
    B = x.shape[0]
    if training:
        mean = x.sum(dim=0) / B
        var = ((x - mean) ** 2).sum(dim=0) / B
        new_rm = new_rv = None
        if running is not None:
            new_rm = (1 - momentum) * running[0] + momentum * mean.detach()
            new_rv = (1 - momentum) * running[1] + momentum * var.detach() * B / (B - 1)
    else:
        mean, var = running
        new_rm, new_rv = running
    y = (x - mean) / torch.sqrt(var + eps) * gamma + beta
    return y, y @ W.t(), new_rm, new_rv
