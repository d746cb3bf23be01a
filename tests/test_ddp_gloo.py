"""CPU, world_size 2 over gloo: the data-parallel host logic (flat gradient arenas, one all-reduce
per arena, batch sharding).  The kernels themselves need a GPU; here gradients are synthetic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from signal_b200 import parallel
from signal_b200.functional import _arena


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = [(6, 4), (6,), (), (3, 1, 2, 2)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes] + [torch.nn.Parameter(torch.zeros(5))]
    class _Views(torch.autograd.Function):   # gradients reach .grad the way the library's backward delivers them
        @staticmethod
        def forward(ctx, *ps):
            return sum(p.sum() for p in ps)

        @staticmethod
        def backward(ctx, g):   # like functional.*.backward: views of flat arenas that only the engine keeps alive
            flat_a, views_a = _arena(shapes[:2], "cpu")
            flat_b, views_b = _arena(shapes[2:], "cpu")
            flat_a.fill_(float(rank + 1))
            flat_b.fill_(float(10 * (rank + 1)))
            return tuple(views_a + views_b)

    _Views.apply(*params[:4]).backward()
    assert all(p.grad._base is None for p in params[:4])   # detached by the engine, storage still shared
    assert params[0].grad.untyped_storage().data_ptr() == params[1].grad.untyped_storage().data_ptr()
    # params[4] has no gradient (like token_selection.W_q/W_k/W_v)
    n = parallel.allreduce_param_grads(params, world)
    ok = n == 2 and all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5)) for p in params[:2]) and \
        all(torch.allclose(p.grad, torch.full_like(p.grad, 15.0)) for p in params[2:4]) and params[4].grad is None
    # gradients that are NOT arena views (six separate tensors): coalesced into one bucket collective
    loose = [torch.nn.Parameter(torch.zeros(3)) for _ in range(6)]
    for p in loose:
        p.grad = torch.full((3,), float(rank + 1))
    n2 = parallel.allreduce_param_grads(loose, world)
    ok = ok and n2 == 1 and all(torch.allclose(p.grad, torch.full((3,), 1.5)) for p in loose)
    sl = parallel.shard_batch(256, rank, world)
    ok = ok and (sl.start, sl.stop) == (rank * 128, rank * 128 + 128)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_arena_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_shard_batch_rejects_ragged():
    import pytest
    with pytest.raises(ValueError):
        parallel.shard_batch(130, 0, 4)


def test_head_arena_layout():
    """One arena for both modules, [SIM late | SIM early | AlignM]: the two exchange pieces tile it once, the late piece
    is exactly in_proj_bias + in_proj_weight rows [0, 2d) (W_q, W_k), every view is 16-byte aligned and the views come
    back in the C structs' field order."""
    from signal_b200 import functional as F_
    d = 64
    flat, pg_s, pg_a, cut, cut2 = F_._head_arena(d, "cpu")
    assert cut2 == min((g.data_ptr() - flat.data_ptr()) // 4 for g in pg_a) and flat.numel() == F_.head_grad_numel(d)
    own = torch.zeros(F_.head_grad_numel(d) + 64)           # a caller-owned arena (parallel.GradExchange) is carved the same way
    flat2, pg_s2, pg_a2, _, _ = F_._head_arena(d, "cpu", own)
    assert flat2.data_ptr() == own.data_ptr() and flat2.numel() == flat.numel()
    assert [(g.data_ptr() - own.data_ptr()) for g in pg_s2 + pg_a2] == [(g.data_ptr() - flat.data_ptr()) for g in pg_s + pg_a]
    sshapes, ashapes = F_._SIM_GRAD_SHAPES(d), F_._align_grad_shapes(d)
    assert [tuple(g.shape) for g in pg_s] == [tuple(s) for s in sshapes]
    assert [tuple(g.shape) for g in pg_a] == [tuple(s) for s in ashapes]
    base = flat.data_ptr()
    off = lambda t: (t.data_ptr() - base) // 4
    assert all(off(g) % 4 == 0 for g in pg_s + pg_a)
    in_w, in_b = pg_s[0], pg_s[1]
    assert off(in_b) == 0 and off(in_w) == (3 * d + 3) // 4 * 4
    assert cut == off(in_w) + 2 * d * d                       # W_v rows are the first thing of the early piece
    assert all(off(g) >= cut for g in pg_s[2:]) and all(off(g) >= cut for g in pg_a)
    assert min(off(g) for g in pg_a) > max(off(g) for g in pg_s)          # AlignM behind SIM
    used = sum(((g.numel() + 3) // 4 * 4) for g in pg_s + pg_a)
    assert used == flat.numel()
    # writes through the views land in the right piece
    flat.zero_()
    in_w[: 2 * d].fill_(1.0); in_b.fill_(1.0)
    assert float(flat[:cut].sum()) == 2 * d * d + 3 * d and float(flat[cut:].sum()) == 0.0


def _worker_pieces(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from signal_b200 import functional as F_
    d = 64
    flat, pg_s, pg_a, cut, _ = F_._head_arena(d, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    flat.copy_(torch.randn(flat.numel(), generator=g))
    mine = flat.clone()
    # the order FusionHead.grad_sync is called in (functional.HeadFunction.backward): early + AlignM piece, then the late piece
    for piece in (flat[cut:], flat[:cut]):
        dist.all_reduce(piece)
        piece.div_(world)
    other = torch.randn(flat.numel(), generator=torch.Generator().manual_seed(100 + (1 - rank)))
    ok = torch.allclose(flat, 0.5 * (mine + other), atol=1e-6)
    ok = ok and torch.allclose(pg_a[1], 0.5 * (mine + other)[(pg_a[1].data_ptr() - flat.data_ptr()) // 4:][: pg_a[1].numel()].view_as(pg_a[1]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_piece_exchange_world2():
    """The two-piece in-backward exchange order on the combined arena averages every gradient exactly once."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pieces, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
