"""Data-parallel plumbing for the fusion head (one process per GPU, torch.distributed).

The head shards by batch only: SIM and LAM are per-sample, the B x B GAM grid is
computed on the local shard exactly like the reference under DDP (useB.py:76-126 has no all-gather).
The only exchange step is the gradient all-reduce.  Each backward of signal_b200.functional returns
its parameter gradients as views of ONE flat fp32 arena, which autograd adopts as ``.grad``; so the
exchange is one all-reduce per arena (SIM and AlignM called separately: two per step) instead of one per
parameter.  ``FusionHead`` goes further: one arena for both modules, exchanged INSIDE the backward in two
pieces as soon as each is final (``FusionHead.grad_sync``, functional.HeadFunction.backward).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def grad_arenas(params: Iterable[torch.nn.Parameter]) -> List[torch.Tensor]:
    """The flat arenas behind the parameters' gradients, in first-seen order, each as ONE 1-D tensor over
    the whole storage.  Grouping is by storage, not by ``._base``: the autograd engine adopts the views a
    backward returns but detaches them, so ``.grad._base`` is None although the storage is still shared."""
    seen, out = set(), []
    for p in params:
        g = p.grad
        if g is None:
            continue
        st = g.untyped_storage()
        key = st.data_ptr()
        if key not in seen:
            seen.add(key)
            out.append(torch.empty(0, dtype=g.dtype, device=g.device).set_(st))
    return out


def allreduce_arenas(arenas: List[torch.Tensor], world_size: int, group=None) -> int:
    """Average the given flat gradient arenas over ranks, in place: one collective per arena (NCCL
    averages inside the collective; gloo sums, then divides)."""
    avg = dist.get_backend(group) == "nccl"
    for a in arenas:
        if avg:
            dist.all_reduce(a, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(a, group=group)
            a.div_(world_size)
    return len(arenas)


def allreduce_param_grads(params: Iterable[torch.nn.Parameter], world_size: int, group=None) -> int:
    """Average gradients over ranks (what DDP does, engine/processor.py:100-105).  Returns the number
    of collectives issued."""
    arenas = grad_arenas(params)
    if len(arenas) <= 4:
        return allreduce_arenas(arenas, world_size, group)
    # the gradients were not delivered as views of a few arenas (e.g. accumulated into existing .grad
    # tensors): one DDP-style bucket -- flatten, one collective, copy back
    from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors
    flat = _flatten_dense_tensors(arenas)
    allreduce_arenas([flat], world_size, group)
    for a, r in zip(arenas, _unflatten_dense_tensors(flat, arenas)):
        a.copy_(r)
    return 1


def shard_batch(n: int, rank: int, world_size: int) -> slice:
    """Contiguous batch shard of rank `rank` (identity-aware samplers hand each rank whole P x K
    groups, data/datasets/sampler_ddp.py:165-175; n must divide evenly)."""
    if n % world_size:
        raise ValueError(f"batch {n} does not divide over {world_size} ranks")
    per = n // world_size
    return slice(rank * per, (rank + 1) * per)
