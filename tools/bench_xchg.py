"""Stand-alone cost of the gradient exchange (run under torchrun, one rank per GPU):
sig_xchg_allreduce_f32 (NVLS multimem / peer loads-stores, several CTA counts) vs ncclAllReduce on the two pieces of the
head's gradient arena at d = 768 (28.5 MB + 4.7 MB), GPU otherwise idle, CUDA events, max over ranks."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from signal_b200 import functional as F_, parallel
    n = F_.head_grad_numel(768)
    cut = (3 * 768 + 3) // 4 * 4 + 2 * 768 * 768
    pieces = [(cut, n - cut), (0, cut)]

    def timeit(fn, iters=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flat = torch.randn(n, device=dev)
    us = timeit(lambda: [dist.all_reduce(flat[o:o + c], op=dist.ReduceOp.AVG) for o, c in pieces])
    if rank == 0:
        print(f"world {world}: {n * 4 / 1e6:.1f} MB in two pieces; ncclAllReduce(avg): {us:.1f} us", flush=True)
    for mc in (True, False):
        ex = parallel.GradExchange(n, dev, use_multicast=mc)
        ex.arena.normal_()
        for ctas in (32, 148):
            ex.ctas = ctas
            us = timeit(lambda: [ex.allreduce(ex.arena[o:o + c]) for o, c in pieces])
            us1 = timeit(lambda: ex.allreduce(ex.arena[:16]))
            ph1 = ex.last_call_phases_us()
            dist.barrier()
            ex.allreduce(ex.arena[cut:])
            ph = ex.last_call_phases_us()
            if rank == 0:
                print(f"  sig_xchg_allreduce_f32 {'multimem' if ex.multicast else 'peer ld/st'} ctas={ctas}: {us:.1f} us "
                      f"({n * 4 / us / 1e3:.0f} GB/s algorithmic); 64-byte call {us1:.1f} us; CTA 0 phases (barrier A, data, barrier B): "
                      f"28.5 MB call {ph[0]:.1f} / {ph[1]:.1f} / {ph[2]:.1f} us, 64-byte call {ph1[0]:.1f} / {ph1[1]:.1f} / {ph1[2]:.1f} us", flush=True)
        if mc:
            for mb in (0.25, 1, 4, 16):
                cnt = int(mb * (1 << 20) / 4) // 4 * 4
                ex.ctas = 148
                us = timeit(lambda: ex.allreduce(ex.arena[:cnt]))
                ph = ex.last_call_phases_us()
                if rank == 0:
                    print(f"    size sweep {mb} MB: {us:.1f} us per call; phases {ph[0]:.1f} / {ph[1]:.1f} / {ph[2]:.1f}", flush=True)
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
