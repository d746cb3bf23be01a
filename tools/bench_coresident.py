"""Does the exchange kernel run NEXT TO a persistent GEMM (same SMs), or does it wait for SMs to drain?  One GPU:
a CTA-pair tcgen05 GEMM (the LAM offset-net shape, ~55 us, 148 CTAs x ~200 KB smem) on one stream, the exchange kernel
(world = 1: same code path, the 'peers' are this rank) on another, separately and together."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29577")
os.environ.setdefault("RANK", "0")
os.environ.setdefault("WORLD_SIZE", "1")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from signal_b200 import functional as F_, lib, parallel

n = F_.head_grad_numel(768)
ex = parallel.GradExchange(n, dev)
ex.arena.normal_()
M, N, K = 128 * 128 * 3, 768, 768
A = torch.randn(M, K, device=dev).bfloat16()
Bm = torch.randn(N, K, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)


def gemm():
    with torch.cuda.stream(s1):
        for _ in range(4):
            lib.debug_gemm_bf16(A, 0, Bm, 0, M, N, K, out_bf16=True, bn=1024, out=None)


def xchg():
    with torch.cuda.stream(s2):
        ex.allreduce(ex.arena)


def timeit(fns, iters=30):
    for _ in range(3):
        for f in fns:
            f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(iters):
        for f in fns:
            f()
    cur = torch.cuda.current_stream()
    cur.wait_stream(s1); cur.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
big2 = torch.empty_like(big)


def copy():      # HBM-bound neighbour: 4 x (256 MB read + 256 MB write)
    with torch.cuda.stream(s1):
        for _ in range(2):
            big2.copy_(big)


rank = int(os.environ["RANK"])
for mc in ([True, False] if int(os.environ["WORLD_SIZE"]) > 1 else [False]):
    if not mc:
        ex = parallel.GradExchange(n, dev, use_multicast=False)
        ex.arena.normal_()
    for ctas in (74, 148):
        ex.ctas = ctas
        for name, work in (("4 pair GEMMs", gemm), ("2 x 512 MB copies", copy)):
            dist.barrier()
            g, x, both = timeit([work]), timeit([xchg]), timeit([work, xchg])
            if rank == 0:
                print(f"world {dist.get_world_size()} multicast={ex.multicast} ctas={ctas}: {name} alone {g:.1f} us, exchange alone {x:.1f} us, "
                      f"together {both:.1f} us (serial {g + x:.1f}, perfect overlap {max(g, x):.1f})", flush=True)
torch.cuda.synchronize()
os._exit(0)
