"""Run the CUDA drop-in modules on a golden/oracle case and collect the same record the
golden files hold (outputs, masks, projected token gradients, parameter fingerprints)."""
import numpy as np
import torch

import golden_util as gu
from signal_b200 import modules as M
from signal_b200 import synthetic as syn


def build_modules(c, sim_p, al_p, device="cuda"):
    sim = M.Select_Interactive_Module(c["d"], k=c["k"], keep_ratio=c["keep_ratio"])
    al = M.AlignmentM(c["d"], c["h"], c["w"])
    sim.load_state_dict(sim_p)
    al.load_state_dict(al_p)
    return sim.to(device), al.to(device)


def cuda_record(c, dtype=torch.float32, packed=True, flags=0, sim_p=None, al_p=None, toks=None, cot=None):
    if sim_p is None:
        sim_p, al_p, toks, cot = gu.case_inputs(c)
    sim, al = build_modules(c, sim_p, al_p)
    sim.flags = al.flags = flags
    sim.fuse_views = al.fuse_views = packed
    toks = [t.to("cuda", dtype).requires_grad_(True) for t in toks]
    cot = cot.to("cuda")
    patches = [t[:, 1:] for t in toks]
    cls = [t[:, 0] for t in toks]
    out = sim(*patches, *cls)
    masks = sim.token_selection.last_masks
    gam, lam = al(*patches, stage="together_CLS_Patch")
    rec = {
        "sim_out": out.detach().float().cpu().numpy(),
        "masks": np.stack([masks[k][..., 0].cpu().numpy().astype(np.uint8) for k in ("RGB", "NI", "TI")]),
        "gam": float(gam.item()), "lam": float(lam.item()),
    }
    named = [("SIM." + k, p) for k, p in sim.named_parameters()] + [("AlignM." + k, p) for k, p in al.named_parameters()]
    objs = {"sim": (out.float() * cot).sum(), "gam": gam, "lam": lam}
    for oname, J in objs.items():
        grads = torch.autograd.grad(J, toks + [p for _, p in named], retain_graph=True, allow_unused=True)
        gt = [torch.zeros_like(t) if g is None else g for t, g in zip(toks, grads[:3])]
        rec[f"dtok_{oname}"] = gu.project_tokens([g.float() for g in gt], c["d"])
        rec[f"dtok_full_{oname}"] = [g.detach().float().cpu() for g in gt]
        for (key, _), g in zip(named, grads[3:]):
            if g is None or float(g.abs().max()) == 0.0:
                continue
            rec[f"dpar_{oname}/{key}"] = gu.fingerprint_param(key, g)
            rec[f"dpar_full_{oname}/{key}"] = g.detach().float().cpu()
    return rec


def compare_records(got, ref, tol, check_masks=True, ref_masks_key="masks", label=""):
    """Assert got (CUDA) matches ref (golden npz or oracle record) within tol (relative L2)."""
    errs = {}
    if check_masks:
        assert np.array_equal(got["masks"], ref[ref_masks_key]), f"{label}: selected-token masks differ"
    errs["sim_out"] = gu.rel_err(got["sim_out"], ref["sim_out"])
    errs["gam"] = abs(got["gam"] - float(ref["gam"])) / abs(float(ref["gam"]))
    errs["lam"] = abs(got["lam"] - float(ref["lam"])) / abs(float(ref["lam"]))
    for k in ("sim_out", "gam", "lam"):
        assert errs[k] < tol, (label, k, errs[k])
    for oname in ("sim", "gam", "lam"):
        e = gu.rel_err(got[f"dtok_{oname}"], ref[f"dtok_{oname}"])
        errs[f"dtok_{oname}"] = e
        assert e < tol, (label, f"dtok_{oname}", e)
        for key in [k for k in ref if k.startswith(f"dpar_{oname}/")]:
            name = key.split("/", 1)[1]
            if float(np.asarray(ref[key])[0]) < 1e-12:
                assert key not in got or got[key][0] < 1e-6, (label, key)
                continue
            assert key in got, (label, key, "gradient missing")
            e = gu.rel_err(got[key], ref[key])
            errs[key] = e
            assert e < gu.param_tol(name, oname, tol), (label, key, e)
    return errs
