"""Data-parallel plumbing for the fusion head (one process per GPU, torch.distributed).

The head shards by batch only: SIM and LAM are per-sample, the B x B GAM grid is
computed on the local shard exactly like the reference under DDP (useB.py:76-126 has no all-gather).
The only exchange step is the gradient all-reduce.  Each backward of signal_b200.functional returns
its parameter gradients as views of ONE flat fp32 arena, which autograd adopts as ``.grad``; so the
exchange is one all-reduce per arena (two per step: SIM and AlignM) instead of one per parameter.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def grad_arenas(params: Iterable[torch.nn.Parameter]) -> List[torch.Tensor]:
    """Distinct storages behind the parameters' gradients (the flat arenas), in first-seen order."""
    seen, out = set(), []
    for p in params:
        g = p.grad
        if g is None:
            continue
        base = g._base if g._base is not None else g
        if id(base) not in seen:
            seen.add(id(base))
            out.append(base)
    return out


def allreduce_param_grads(params: Iterable[torch.nn.Parameter], world_size: int, group=None) -> int:
    """Average gradients over ranks (what DDP does, engine/processor.py:100-105).  Returns the number
    of collectives issued."""
    arenas = grad_arenas(params)
    for a in arenas:
        dist.all_reduce(a, group=group)
        a.div_(world_size)
    return len(arenas)


def shard_batch(n: int, rank: int, world_size: int) -> slice:
    """Contiguous batch shard of rank `rank` (identity-aware samplers hand each rank whole P x K
    groups, data/datasets/sampler_ddp.py:165-175; n must divide evenly)."""
    if n % world_size:
        raise ValueError(f"batch {n} does not divide over {world_size} ranks")
    per = n // world_size
    return slice(rank * per, (rank + 1) * per)
