"""Worker of tests/test_gpu_dp.py (one process per GPU under torch.distributed.run).

1. sig_xchg_allreduce_f32 (signal_b200.parallel.GradExchange) against ncclAllReduce on random arenas: several
   offsets/sizes, repeated calls (epoch counters), both the NVLS multicast and the plain peer-load variants, and inside
   a captured CUDA graph that is replayed.
2. Data-parallel semantics of the head (reference: DDP, engine/processor.py:100-105, GAM grid per rank,
   useB.py:76-126): after the in-backward exchange every rank must hold mean_r(grad computed on shard r).  Each rank
   also computes the gradients of EVERY shard locally without any exchange and compares their mean with what the
   exchange delivered -- through the NVLink kernel and through NCCL.
Prints DP_CHECK_OK on rank 0.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def check_exchange(dev, rank, world):
    from signal_b200 import parallel
    n = 3_000_000 + 4 * 37
    for use_mc in (True, False):
        ex = parallel.GradExchange(n, dev, use_multicast=use_mc)
        if rank == 0:
            print(f"exchange: world {world}, multicast {'on' if ex.multicast else 'off (peer loads/stores)'}", flush=True)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        for it, (off, cnt) in enumerate([(0, n - n % 4), (4 * 11, 1_000_000), (2_000_000, 4 * 5), (0, 4), (1024, 2_097_152)] * 2):
            src = torch.randn(ex.numel, generator=g).to(dev)
            ex.arena.copy_(src)
            ref = src.clone()
            dist.all_reduce(ref[off:off + cnt], op=dist.ReduceOp.SUM)
            ref[off:off + cnt] /= world
            torch.cuda.synchronize()
            dist.barrier()
            ex.allreduce(ex.arena[off:off + cnt])
            torch.cuda.synchronize()
            err = float((ex.arena - ref).abs().max())
            assert err <= 2e-6, f"rank {rank} exchange (multicast={ex.multicast}) call {it}: max abs err {err}"
            assert torch.equal(ex.arena[:off], src[:off]) and torch.equal(ex.arena[off + cnt:], src[off + cnt:]), "wrote outside its piece"
        # captured in a CUDA graph and replayed (epochs live in device memory)
        src = torch.randn(ex.numel, generator=g).to(dev)
        stage = src.clone()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            ex.arena.copy_(stage)
            ex.allreduce(ex.arena[:1_000_000])
            ex.allreduce(ex.arena[1_000_000:ex.numel])
        ref = src.clone()
        dist.all_reduce(ref)
        ref /= world
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        err = float((ex.arena - ref).abs().max())
        assert err <= 2e-6, f"rank {rank} exchange in graph replay: {err}"
        dist.barrier()
        del gr


def head_grads(c, dtype, sim_p, al_p, toks, cot, dev, hook=None, arena=None):
    import gpu_harness as gh
    from signal_b200 import modules as M
    sim, al = gh.build_modules(c, sim_p, al_p, device=dev)
    head = M.FusionHead(sim, al, grad_sync=hook)
    head.grad_arena = arena
    tk = [t.to(dev, dtype).requires_grad_(True) for t in toks]
    out, gam, lam = head(*[t[:, 1:] for t in tk], *[t[:, 0] for t in tk])
    torch.autograd.backward([out, gam, lam], [cot.to(dev, dtype), torch.tensor(0.2, device=dev), torch.tensor(0.2, device=dev)])
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in list(sim.named_parameters()) + list(al.named_parameters()) if p.grad is not None}


def check_head(dev, rank, world):
    import golden_util as gu
    from signal_b200 import parallel, synthetic as syn
    for dtype, B, tol in ((torch.float32, 4, 2e-5), (torch.bfloat16, 16, 4e-3)):
        c = dict(d=512, h=16, w=8, B=B, k=80, keep_ratio=None, gain=1.0, structured=False, seed=77)
        sim_p = syn.make_params(syn.sim_param_shapes(c["d"]), c["seed"])
        al_p = syn.make_params(syn.align_param_shapes(c["d"]), c["seed"] + 1)
        toks_all = syn.make_tokens(B * world, c["d"], seed=c["seed"] + 2, dtype=dtype)
        cot_all = syn.make_cotangent(B * world, c["d"], seed=c["seed"] + 3)
        shard = lambda r: ([t[r * B:(r + 1) * B].contiguous() for t in toks_all], cot_all[r * B:(r + 1) * B])
        # what DDP semantics asks for: the mean over ranks of the shard-local gradients (computed here without any exchange)
        local = [head_grads(c, dtype, sim_p, al_p, *shard(r), dev) for r in range(world)]
        want = {n: sum(g[n].double() for g in local) / world for n in local[0]}
        from signal_b200 import functional as F_
        ex = parallel.GradExchange(F_.head_grad_numel(c["d"]), dev)
        toks, cot = shard(rank)
        for name, hook, arena in (("nvlink kernel", ex.allreduce, ex.arena),
                                  ("nccl", lambda a: dist.all_reduce(a, op=dist.ReduceOp.AVG), None)):
            got = head_grads(c, dtype, sim_p, al_p, toks, cot, dev, hook=hook, arena=arena)
            assert got.keys() == want.keys()
            worst = max(float((got[n].double() - want[n]).norm() / want[n].norm().clamp_min(1e-300)) for n in want)
            assert worst <= tol, f"rank {rank} {dtype} via {name}: gradient != mean of shard gradients ({worst:.3e})"
            # and every rank holds the same bits
            for n in sorted(got):
                ref = got[n].clone()
                dist.broadcast(ref, 0)
                assert torch.equal(ref, got[n]), f"{name}: {n} differs between ranks"
            if rank == 0:
                print(f"head dp check {dtype} via {name}: worst rel err {worst:.2e}", flush=True)
        dist.barrier()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    dist.barrier()
    check_exchange(dev, rank, world)
    check_head(dev, rank, world)
    dist.barrier()
    if rank == 0:
        print("DP_CHECK_OK", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
