"""Micro-benchmark of the tcgen05 GEMM core on the shapes the head uses (CUDA events, warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from signal_b200 import lib


def timeit(fn, iters=20, reps=10):
    """GPU time per call: `iters` back-to-back calls captured in a CUDA graph (no host overhead)."""
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
z1 = torch.empty(1, device='cuda')
print("tiny torch kernel:", round(timeit(lambda: z1.zero_()), 2), "us")
for (M, N, K) in [(384, 768, 768), (384, 1536, 768), (384, 768, 1536), (768, 768, 384), (128, 768, 128)]:
    A, B = rnd(M, K), rnd(N, K)
    for bn in (128, 256):
        t = timeit(lambda: lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, bn=bn))
        print(f"NT  M={M} N={N} K={K} bn={bn}: {t:7.2f} us  {2*M*N*K/t/1e6:8.1f} TFLOP/s")
    for ks in (2, 4):
        t = timeit(lambda: lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, ksplit=ks))
        print(f"NT  M={M} N={N} K={K} ksplit={ks}: {t:7.2f} us")
Bs, d = 128, 768
tok = rnd(Bs, 129, d)
X = tok[:, 1:]
W = rnd(d, d)
dH = rnd(Bs * 128, d)
for bn in (128, 256, 512, 1024):
    t = timeit(lambda: lib.debug_gemm_bf16(X, 1, W, 0, Bs * 128, d, d, bn=bn, out_bf16=True))
    print(f"LAM fwd  bn={bn}: {t:7.2f} us  {2*Bs*128*d*d/t/1e6:8.1f} TFLOP/s")
    t = timeit(lambda: lib.debug_gemm_bf16(dH, 0, W, 2, Bs * 128, d, d, bn=bn, out_bf16=True))
    print(f"LAM dX   bn={bn}: {t:7.2f} us  {2*Bs*128*d*d/t/1e6:8.1f} TFLOP/s")
for bn, ks in ((128, 4), (128, 8), (256, 8), (512, 11), (512, 16), (512, 33), (1024, 8), (1024, 16), (1024, 33)):
    t = timeit(lambda: lib.debug_gemm_bf16(dH, 2, X, 3, d, d, Bs * 128, ksplit=ks, bn=bn))
    print(f"LAM dW   bn={bn} ksplit={ks}: {t:7.2f} us  {2*Bs*128*d*d/t/1e6:8.1f} TFLOP/s")
print("stages override:", os.environ.get("SIG_TC_STAGES"), " PDL:", os.environ.get("SIG_PDL"))
