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
