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


def cuda_record(c, dtype=torch.float32, packed=True, flags=0, sim_p=None, al_p=None, toks=None, cot=None, fused=False):
    if sim_p is None:
        sim_p, al_p, toks, cot = gu.case_inputs(c)
    sim, al = build_modules(c, sim_p, al_p)
    sim.flags = al.flags = flags
    sim.fuse_views = al.fuse_views = packed
    toks = [t.to("cuda", dtype).requires_grad_(True) for t in toks]
    cot = cot.to("cuda")
    patches = [t[:, 1:] for t in toks]
    cls = [t[:, 0] for t in toks]
    if fused:
        out, gam, lam = M.FusionHead(sim, al)(*patches, *cls, stage="together_CLS_Patch")
    else:
        out = sim(*patches, *cls)
        gam, lam = al(*patches, stage="together_CLS_Patch")
    masks = sim.token_selection.last_masks
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


def lam_flip_samples(got, ref, tol):
    """(modality, sample) pairs whose LAM token gradient deviates by more than `tol`.

    d(LAM)/d(offset) is discontinuous where a sample point crosses a pixel boundary (the bilinear
    slope switches to another pair of neighbouring tokens).  A reduced-precision run moves the
    predicted offsets by ~1e-3 px, so in a batch of B*8*3 sample points an occasional point lands on
    the other side of a boundary and that one sample's offset-path gradient changes by O(1).  Such
    samples are reported, bounded in number, and excluded from the aggregate comparison.
    """
    a = np.asarray(got["dtok_lam"], dtype=np.float64)
    b = np.asarray(ref["dtok_lam"], dtype=np.float64)
    num = np.linalg.norm((a - b).reshape(a.shape[0], a.shape[1], -1), axis=-1)
    den = np.linalg.norm(b.reshape(b.shape[0], b.shape[1], -1), axis=-1)
    bad = np.argwhere(num > tol * np.maximum(den, 1e-30))
    return [(int(m), int(s)) for m, s in bad]


def compare_records(got, ref, tol, check_masks=True, ref_masks_key="masks", label="", lam_flip_robust=False):
    """Assert got (CUDA) matches ref (golden npz or oracle record) within tol (relative L2).

    lam_flip_robust (reduced-precision runs only): tolerate a few pixel-boundary flips of LAM sample
    points, see lam_flip_samples()."""
    errs = {}
    flips = []
    if lam_flip_robust:
        flips = lam_flip_samples(got, ref, tol)
        nsamp = np.asarray(ref["dtok_lam"]).shape[1] * 3
        assert len(flips) <= max(1, int(0.03 * nsamp)), (label, "too many LAM outlier samples", flips[:10])
        errs["lam_flip_samples"] = float(len(flips))
    if check_masks:
        assert np.array_equal(got["masks"], ref[ref_masks_key]), f"{label}: selected-token masks differ"
    errs["sim_out"] = gu.rel_err(got["sim_out"], ref["sim_out"])
    errs["gam"] = abs(got["gam"] - float(ref["gam"])) / abs(float(ref["gam"]))
    errs["lam"] = abs(got["lam"] - float(ref["lam"])) / abs(float(ref["lam"]))
    for k in ("sim_out", "gam", "lam"):
        assert errs[k] < tol, (label, k, errs[k])
    for oname in ("sim", "gam", "lam"):
        ga, ra = np.array(got[f"dtok_{oname}"], dtype=np.float64), np.array(ref[f"dtok_{oname}"], dtype=np.float64)
        if oname == "lam":
            for m, sidx in flips:
                ga[m, sidx] = 0.0
                ra[m, sidx] = 0.0
        e = gu.rel_err(ga, ra)
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
            ptol = gu.param_tol(name, oname, tol)
            if oname == "lam" and flips:
                mods = {"DAS_r": 0, "DAS_n": 1, "DAS_t": 2}
                hit = [mods[k] for k in mods if k in name]
                if hit and any(m == hit[0] for m, _ in flips):
                    ptol = max(ptol, 0.25)   # one flipped sample carries O(1/B) of this modality's gradient
            assert e < ptol, (label, key, e)
    return errs
