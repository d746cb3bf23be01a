import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import golden_util as gu
import gpu_harness as gh
from signal_b200 import lib

name = sys.argv[1] if len(sys.argv) > 1 else "vehicle_d512"
if name == "b128":
    c = dict(d=768, h=16, w=8, B=128, k=80, keep_ratio=None, gain=1.0, structured=False, seed=4242)
else:
    c = gu.CASES[name]
sim_p, al_p, toks, cot = gu.case_inputs(c)
toks = [t.to(torch.bfloat16) for t in toks]
fast = gh.cuda_record(c, torch.bfloat16, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
slow = gh.cuda_record(c, torch.bfloat16, flags=lib.FLAG_FORCE_SIMT, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
for m in range(3):
    a, b = fast["dtok_full_lam"][m], slow["dtok_full_lam"][m]
    err = (a - b).norm(dim=-1)          # [B,129]
    ref = b.norm(dim=-1)
    print("mod", m, "total rel", float((a - b).norm() / b.norm()))
    print("  per-row err (sample 0):", [round(float(x), 4) for x in err[0, :20]])
    print("  per-row ref (sample 0):", [round(float(x), 4) for x in ref[0, :20]])
    print("  rows with err > 0.1*ref:", int((err > 0.1 * ref.clamp_min(1e-9)).sum()), "of", err.numel())
    nz = ref > 0
    print("  ref row-norm stats: max", float(ref.max()), "median nz", float(ref[nz].median()), " err max", float(err.max()), "err median", float(err[nz].median()))
    worst = torch.topk(err.flatten(), 5).indices
    print("  worst rows (b,l):", [(int(i) // 129, int(i) % 129, round(float(err.flatten()[i]), 8), round(float(ref.flatten()[i]), 8)) for i in worst])
for k in sorted(fast):
    if k.startswith("dpar_full_lam/"):
        a, b = fast[k], slow[k]
        print(k, float((a - b).norm() / b.norm().clamp_min(1e-30)))
