"""NCCL all-reduce of the head-gradient arenas in isolation (torchrun, one rank per GPU): eager and CUDA-graph replay."""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sizes = [6498816, 3585025]
bufs = [torch.randn(n, device=dev) for n in sizes]
one = torch.randn(sum(sizes), device=dev)

def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def two_sum_div():
    for b in bufs:
        dist.all_reduce(b); b.div_(world)
def two_avg():
    for b in bufs:
        dist.all_reduce(b, op=dist.ReduceOp.AVG)
def one_avg():
    dist.all_reduce(one, op=dist.ReduceOp.AVG)
res = {}
for name, fn in (("two_sum_div", two_sum_div), ("two_avg", two_avg), ("one_avg", one_avg)):
    res[name + "_eager"] = timed(fn)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    res[name + "_graph"] = timed(g.replay)
if rank == 0:
    print("allreduce us:", {k: round(v, 1) for k, v in res.items()}, flush=True)
os._exit(0)
