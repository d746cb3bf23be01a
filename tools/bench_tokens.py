"""Token producer (csrc/tokens.cu) timings: python tools/bench_tokens.py [B]  (3 modalities batched: 3B x 129 rows, 768 -> 512).
Phases are event-timed inside the library (sig_profile_*), the whole fwd+bwd as a CUDA-graph replay; fractions are of the
measured peaks in MEASURED_PEAKS.json (HBM copy GB/s for the LayerNorm passes, burst bf16 TFLOP/s for the GEMMs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from signal_b200 import lib, tokens

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
W, D, L1 = 768, 512, 129
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
dev = torch.device("cuda", 0)
ln = nn.LayerNorm(W).to(dev); p = nn.Parameter(W ** -0.5 * torch.randn(W, D, device=dev))
tp = tokens.TokenProducer(ln, p)
R = 3 * B * L1
xs = [torch.randn(3 * B, L1, W, device=dev).to(torch.bfloat16).requires_grad_(True) for _ in range(4)]   # 4 x 76 MB > L2
cot = torch.randn(3 * B, L1, D, device=dev).to(torch.bfloat16) * 0.01


def step(x):
    t = tp.tokens(x)
    t.backward(cot)


for x in xs:
    step(x)
torch.cuda.synchronize()
lib.profile_enable(True)
for i in range(20):
    step(xs[i % 4])
torch.cuda.synchronize()
lib.profile_enable(False)
prof = {k: v[0] * 1e3 / v[1] for k, v in lib.profile_collect().items() if k.startswith("tokens")}
gs = []
for x in xs:
    x.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step(x)
    gs.append(g)
for g in gs:
    g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(40):
    gs[i % 4].replay()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 40
alg = {  # algorithmic bytes / flops per phase (bf16)
    "tokens_ln_fwd": ("GB/s", 2 * R * W * 2), "tokens_ln_bwd": ("GB/s", 3 * R * W * 2),
    "tokens_proj_fwd": ("TFLOP/s", 2 * R * W * D), "tokens_proj_bwd": ("TFLOP/s", 4 * R * W * D),
    "tokens_patch_mean": ("GB/s", R * D * 2),
}
out = {"rows": R, "W": W, "D": D, "graph_replay_fwd_bwd_ms": round(ms, 4), "phases": {}}
for k, us in sorted(prof.items()):
    unit, work = alg[k]
    ach = work / (us * 1e-6) / (1e9 if unit == "GB/s" else 1e12)
    peak = peaks["hbm_gbs"] if unit == "GB/s" else peaks["bf16_tflops"]
    out["phases"][k] = {"us": round(us, 1), "achieved": round(ach, 1), "unit": unit, "frac": round(ach / peak, 3)}
print(json.dumps(out))
