import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import golden_util as gu
import gpu_harness as gh
from signal_b200 import lib

def al256(n): return (n + 255) // 256 * 256

def layout_tc(B, L, d):
    P = L // 16; BL = B * L; dd = d * d
    items = [("mean", 3*B*d, 4), ("f", 3*B*d, 4), ("nrm", 3*B, 4), ("self4", 4*B, 4), ("lv", B*B, 4), ("la", B*B, 4), ("V", B*B, 4),
             ("rowstat", 2*B, 4), ("colstat", 2*B, 4), ("Wlv", B*B, 4), ("Wla", B*B, 4), ("rowA", B, 4), ("colC", 3*B, 4), ("dtau", 4, 4),
             ("df", 3*B*d, 4), ("dmean", 3*B*d, 4), ("W0b", 3*dd, 2), ("Wqb", 3*dd, 2), ("Wfb", 3*dd, 2), ("dWfb", 3*dd, 2),
             ("bfold", 3*d, 4), ("dWf", 3*dd, 4), ("dbf", 3*d, 4), ("H", 3*BL*d, 2), ("dH", 3*BL*d, 2), ("U", 3*B*P*d, 4),
             ("dU", 3*B*P*d, 4), ("o", 3*B*P, 4), ("dO", 3*B*P, 4), ("S", 3*B*P*d, 4), ("dS", 3*B*P*d, 4), ("part", B*P, 4)]
    off, out = 0, {}
    for n, cnt, es in items:
        out[n] = (off, cnt, es); off += al256(cnt * es)
    return out

def layout_simt(B, L, d):
    P = L // 16; BL = B * L
    items = [("Xf", 3*BL*d), ("mean", 3*B*d), ("f", 3*B*d), ("nrm", 3*B), ("self4", 4*B), ("lv", B*B), ("la", B*B), ("V", B*B),
             ("rowstat", 2*B), ("colstat", 2*B), ("Wlv", B*B), ("Wla", B*B), ("rowA", B), ("colC", 3*B), ("dtau", 4), ("df", 3*B*d), ("dmean", 3*B*d)]
    for m in range(3):
        items += [(f"Q{m}", BL*d), (f"H{m}", BL*d), (f"G{m}", BL*d), (f"U{m}", B*P*d), (f"o{m}", B*P), (f"dO{m}", B*P), (f"dU{m}", B*P*d)]
    items += [("S", 3*B*P*d), ("dS", 3*B*P*d), ("part", B*P), ("dH", BL*d), ("dQ", BL*d), ("dXf", 3*BL*d)]
    off, out = 0, {}
    for n, cnt in items:
        out[n] = (off, cnt, 4); off += al256(cnt * 4)
    return out

def get(buf, lay, name):
    off, cnt, es = lay[name]
    raw = buf[off:off + cnt * es]
    return raw.view(torch.bfloat16 if es == 2 else torch.float32).float().cpu()

name = sys.argv[1] if len(sys.argv) > 1 else "vehicle_d512"
if name == "b128":
    c = dict(d=768, h=16, w=8, B=128, k=80, keep_ratio=None, gain=1.0, structured=False, seed=4242)
else:
    c = gu.CASES[name]
bsel = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sim_p, al_p, toks, cot = gu.case_inputs(c)
B, d, L = c["B"], c["d"], 128
bufs = {}
for tag, flags in (("tc", 0), ("simt", lib.FLAG_FORCE_SIMT)):
    sim, al = gh.build_modules(c, sim_p, al_p)
    al.flags = flags
    tk = [t.to("cuda", torch.bfloat16).requires_grad_(True) for t in toks]
    gam, lam = al(*[t[:, 1:] for t in tk], stage="x")
    lam.backward(retain_graph=True)
    torch.cuda.synchronize()
    bufs[tag] = lam.grad_fn.saved_tensors[-1].clone()
lt, ls = layout_tc(B, L, d), layout_simt(B, L, d)
P = L // 16
o_tc = get(bufs["tc"], lt, "o").view(3, B, P); dO_tc = get(bufs["tc"], lt, "dO").view(3, B, P)
for m in range(3):
    o_s = get(bufs["simt"], ls, f"o{m}").view(B, P); dO_s = get(bufs["simt"], ls, f"dO{m}").view(B, P)
    print("mod", m, "o max abs diff", float((o_tc[m] - o_s).abs().max()), "o rms", float(o_s.pow(2).mean().sqrt()))
    print("   dO rel diff", float((dO_tc[m] - dO_s).norm() / dO_s.norm()))
    print("   o_tc[0]", [round(float(x), 3) for x in o_tc[m, bsel]], "\n   o_s [0]", [round(float(x), 3) for x in o_s[bsel]])
    print("   dO_tc[0]", [float(x) for x in dO_tc[m, bsel]], "\n   dO_s [0]", [float(x) for x in dO_s[bsel]])
S_tc = get(bufs["tc"], lt, "S"); S_s = get(bufs["simt"], ls, "S")
print("S rel diff", float((S_tc - S_s).norm() / S_s.norm()))
dS_tc = get(bufs["tc"], lt, "dS"); dS_s = get(bufs["simt"], ls, "dS")
print("dS rel diff", float((dS_tc - dS_s).norm() / dS_s.norm()))
H_tc = get(bufs["tc"], lt, "H").view(3, B * L, d)
for m in range(3):
    H_s = get(bufs["simt"], ls, f"H{m}").view(B * L, d)
    print("H rel diff", m, float((H_tc[m] - H_s).norm() / H_s.norm()))

import math
for m in range(3):
    o_s = get(bufs["simt"], ls, f"o{m}").view(B, P); dO_s = get(bufs["simt"], ls, f"dO{m}").view(B, P)
    rel = (dO_tc[m] - dO_s).abs() / dO_s.abs().mean()
    idx = torch.topk(rel.flatten(), 5).indices
    for i in idx:
        b, p = int(i) // P, int(i) % P
        ot, os_ = float(o_tc[m, b, p]), float(o_s[b, p])
        Hk, Wk = c["h"] // 4, c["w"] // 4
        py, px = p // Wk, p % Wk
        def pos(o):
            th = math.tanh(o)
            ry = (py + .5) / (Hk - 1) * 2 - 1 + 2 * th / (Hk - 1); rx = (px + .5) / (Wk - 1) * 2 - 1 + 2 * th / (Wk - 1)
            cy = min(max(ry, -1), 1); cx = min(max(rx, -1), 1)
            return (cy + 1) / 2 * (c["h"] - 1), (cx + 1) / 2 * (c["w"] - 1)
        print("mod", m, "b", b, "p", p, "o_tc", ot, "o_s", os_, "dO_tc", float(dO_tc[m, b, p]), "dO_s", float(dO_s[b, p]), "pos_tc", pos(ot), "pos_s", pos(os_))
