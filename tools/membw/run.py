import ctypes as C, os, torch
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmembw.so"))
lib.membw_launch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
N = 75497472            # bytes of one token pass (B=128, d=768, bf16)
bufs = [torch.empty(N, dtype=torch.uint8, device="cuda").random_() for _ in range(5)]   # rotate: 377 MB > L2
sink = torch.zeros(1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run(mode, unroll, grid, block):
    for b in bufs: lib.membw_launch(mode, unroll, b.data_ptr(), N, grid, block, sink.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(4):
        for b in bufs: lib.membw_launch(mode, unroll, b.data_ptr(), N, grid, block, sink.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    return us, N / us / 1e6
for mode, name in ((0, "stream"), (1, "region")):
    for grid in (148, 296, 384, 592, 1184, 2368, 4736):
        for block in (256, 512, 1024):
            for unroll in (2, 8):
                us, tbs = run(mode, unroll, grid, block)
                print(f"{name:7s} grid={grid:5d} block={block:4d} unroll={unroll}: {us:6.1f} us  {tbs:5.2f} TB/s")
