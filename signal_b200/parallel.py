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


class GradExchange:
    """In-place average of the head's flat fp32 gradient arena over the ranks of one node, as ONE kernel over NVLink
    peer memory (``sig_xchg_allreduce_f32``, csrc/xchg.cu) -- the DDP all-reduce of engine/processor.py:100-105 without
    NCCL's SM footprint, so it can run under the persistent kernels of the backward.

    torch is the plumbing: ``torch.distributed._symmetric_memory`` allocates the symmetric buffer
    ``[arena | flag region]`` and exchanges the peer mappings (and the NVLS multicast mapping when the fabric has one);
    the kernel is ours.  Usage::

        ex = GradExchange(head.grad_numel(), device)          # once, collectively
        head.grad_arena, head.grad_sync = ex.arena, ex.allreduce

    ``FusionHead`` then carves its parameter-gradient views out of ``ex.arena`` (so ``.grad`` lives in symmetric memory:
    overwritten by the next backward, like DDP's ``gradient_as_bucket_view=True``) and calls ``ex.allreduce(piece)``
    on its communication stream as soon as each piece is final.  Every rank must issue the same calls in the same order.
    """

    def __init__(self, numel: int, device, group=None, ctas: int = 0, use_multicast: bool = True):
        import ctypes as C
        import os
        import torch.distributed._symmetric_memory as symm_mem
        from . import lib as L_
        self.lib = L_.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("signal_b200: GradExchange covers the <= 8 GPUs of one node")
        device = torch.device(device)
        self.numel = (int(numel) + 3) // 4 * 4
        flag_floats = int(self.lib.sig_xchg_flag_bytes()) // 4
        with torch.cuda.device(device):
            self.buf = symm_mem.empty(self.numel + flag_floats, dtype=torch.float32, device=device)
            self.buf.zero_()
            torch.cuda.synchronize(device)
            gname = self.group.group_name if hasattr(self.group, "group_name") else self.group
            try:
                symm_mem.enable_symm_mem_for_group(gname)      # (older torch needs it; a no-op / deprecated later)
            except Exception:
                pass
            self.handle = symm_mem.rendezvous(self.buf, gname)
        self.arena = self.buf[:self.numel]
        peers = L_.SigXchgPeers()
        ptrs = list(self.handle.buffer_ptrs)
        for r in range(self.world):
            peers.buf[r] = ptrs[r]
            peers.flags[r] = ptrs[r] + 4 * self.numel
        mc = int(self.handle.multicast_ptr) if use_multicast and os.environ.get("SIG_XCHG_MULTICAST", "1") != "0" else 0
        peers.multicast = mc if mc else None
        peers.rank, peers.world = self.rank, self.world
        self.peers, self.multicast = peers, bool(mc)
        self.ctas = int(os.environ.get("SIG_XCHG_CTAS", str(ctas)))
        self.device = device
        self._C = C
        dist.barrier(self.group)          # every rank's flag region is zero before anyone signals

    def last_call_phases_us(self):
        """(barrier A, data loop, barrier B) of CTA 0 in the last call, microseconds (synchronises)."""
        torch.cuda.synchronize(self.device)
        words = int(self.lib.sig_xchg_flag_bytes()) // 4
        st = self.buf[self.numel + words - 16: self.numel + words].view(torch.int64)[:4].tolist()
        return tuple((st[i + 1] - st[i]) / 1e3 for i in range(3))

    def allreduce(self, piece: torch.Tensor, scale=None):
        """piece: a contiguous fp32 view inside ``self.arena``; averaged over the ranks in place, on the current stream."""
        from . import lib as L_
        off = (piece.data_ptr() - self.arena.data_ptr()) // 4
        n = piece.numel()
        if piece.dtype != torch.float32 or not piece.is_contiguous() or off < 0 or off + n > self.numel or (off | n) % 4:
            raise RuntimeError("signal_b200: GradExchange.allreduce needs a contiguous 16-byte aligned fp32 piece of its own arena")
        s = 1.0 / self.world if scale is None else float(scale)
        with torch.cuda.device(self.device):
            L_.check(self.lib.sig_xchg_allreduce_f32(self._C.byref(self.peers), off, n, s, self.ctas, self.device.index,
                                                     L_.stream_ptr(self.device)), "sig_xchg_allreduce_f32")
        return piece


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
