"""Time line of ONE CUDA-graph replay of the fused step (SIG_PROF_CAPTURE=1 python tools/timeline.py):
the library's phase scopes are recorded as event nodes inside the captured graph."""
import os, sys, faulthandler
faulthandler.enable()
os.environ.setdefault("SIG_PROF_CAPTURE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from signal_b200 import lib, modules as M, synthetic as syn
L_ = lib.load()
d, B = 768, 128
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:      # torchrun --nproc-per-node N tools/timeline.py: with the in-backward gradient exchange
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
sim = M.Select_Interactive_Module(d, k=80); al = M.AlignmentM(d, 16, 8)
sim.load_state_dict(syn.make_params(syn.sim_param_shapes(d), 1234)); al.load_state_dict(syn.make_params(syn.align_param_shapes(d), 1235))
sim, al = sim.to(dev), al.to(dev)
params = list(sim.parameters()) + list(al.parameters())
toks = [t.to(dev).requires_grad_(True) for t in syn.make_tokens(B, d, seed=1, dtype=torch.bfloat16)]
cot = syn.make_cotangent(B, d).to(dev, torch.bfloat16)
wg = torch.tensor(0.2, device=dev); wl = torch.tensor(0.2, device=dev)
head = M.FusionHead(sim, al)
if world > 1:
    marks, t0 = [], [None]

    ex = None
    if os.environ.get("SIG_EXCHANGE", "nvlink") == "nvlink":
        from signal_b200 import parallel
        ex = parallel.GradExchange(head.grad_numel(), dev)
        head.grad_arena = ex.arena
    xch = (lambda f: ex.allreduce(f)) if ex is not None else (lambda f: dist.all_reduce(f, op=dist.ReduceOp.AVG))

    def _sync(flat):      # torch-side external events: laid on the library's time line through the step-start mark
        if not torch.cuda.is_current_stream_capturing():
            xch(flat)
            return
        a, b = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
        a.record()
        xch(flat)
        b.record()
        marks.append((f"{'xchg' if ex is not None else 'allreduce'}_{flat.numel() * 4 / 1e6:.1f}MB", a, b))
    head.grad_sync = _sync
only = sys.argv[1] if len(sys.argv) > 1 else ""
def fwd_bwd():
    patches = [t[:, 1:] for t in toks]; cls = [t[:, 0] for t in toks]
    if only == "sim":
        out = sim(*patches, *cls); torch.autograd.backward([out], [cot])
    elif only == "align":
        gam, lam = al(*patches, stage="together_CLS_Patch"); torch.autograd.backward([gam, lam], [wg, wl])
    else:
        out, gam, lam = head(*patches, *cls, stage="together_CLS_Patch")
        torch.autograd.backward([out, gam, lam], [cot, wg, wl])
for _ in range(3):
    for t in toks: t.grad = None
    for p in params: p.grad = None
    fwd_bwd()
torch.cuda.synchronize()
for t in toks: t.grad = None
for p in params: p.grad = None
g = torch.cuda.CUDAGraph()
lib.profile_enable(True)
with torch.cuda.graph(g):
    if world > 1:
        t0[0] = torch.cuda.Event(enable_timing=True, external=True)
        t0[0].record()
    fwd_bwd()
lib.profile_enable(False)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
n = L_.sig_profile_timeline(buf, len(buf))
rows = [l.split() for l in buf.value.decode().strip().splitlines()]
rows = [(r[0], float(r[1]), float(r[2])) for r in rows]
if world > 1:   # (the first library scope opens at the step start, like t0)
    rows += [(nm, t0[0].elapsed_time(a) * 1e3, t0[0].elapsed_time(b) * 1e3) for nm, a, b in marks]
rows = sorted(rows, key=lambda r: r[1])
if rank != 0:
    os._exit(0)
print(f"{n} scopes; step spans {max(r[2] for r in rows):.1f} us")
for name, a, b in rows:
    print(f"{a:8.1f} {b:8.1f} {b - a:7.1f}  {name}")
if world > 1:
    os._exit(0)
