"""Fixed-overhead probe of the tcgen05 pipeline kernel: tiny GEMMs back to back in a CUDA graph.
Run twice (SIG_PDL=1 / SIG_PDL=0) and with SIG_TC_STAGES to separate launch, prologue and k-loop costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from signal_b200 import lib


def timeit(fn, iters=20, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3


def rnd(*shape):
    return torch.randn(*shape, device="cuda").to(torch.bfloat16)


lib.load()
z1 = torch.empty(1, device="cuda")
print("env PDL", os.environ.get("SIG_PDL"), "STAGES", os.environ.get("SIG_TC_STAGES"))
print("tiny torch kernel: %.2f us" % timeit(lambda: z1.zero_()))
for (M, N, K) in [(128, 128, 64), (128, 128, 128), (128, 128, 768), (128, 768, 128), (384, 768, 768), (384, 768, 64), (16384, 768, 64)]:
    A, B = rnd(M, K), rnd(N, K)
    t = timeit(lambda: lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, bn=128))
    print(f"NT M={M} N={N} K={K}: {t:7.2f} us")

if os.environ.get("SIG_TC_STAMPS"):
    import ctypes as C
    L = lib.load()
    for (M, N, K) in [(128, 128, 64), (128, 128, 768), (384, 768, 768)]:
        A, B = rnd(M, K), rnd(N, K)
        for _ in range(3):
            lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, bn=128)
        torch.cuda.synchronize()
        buf = (C.c_longlong * 16)()
        L.sig_debug_tc_stamps(buf)
        t = list(buf)[:16]
        print(f"stamps M={M} N={N} K={K} (cycles since entry):", [x - t[0] for x in t])
