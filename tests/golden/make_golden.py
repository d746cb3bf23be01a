"""Generate golden vectors from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the reference's own modules from /root/reference (read-only) through a
pre-seeded namespace package (so ``modeling/__init__.py``, which pulls timm and
fvcore, is skipped), feeds them seeded synthetic tokens/parameters from
``signal_b200.synthetic`` and stores, per case, in ``tests/golden/<case>.npz``:

* from an fp32 run of the reference: ``masks32`` [3,B,L] uint8, ``sim_out32``, ``gam32``, ``lam32``;
* from an fp64 run of the same reference code (``module.double()``; this removes
  fp32 rounding noise from the pinned values, the semantics are unchanged):
  ``sim_out`` [B,3d], ``masks`` [3,B,L] uint8, ``gam``, ``lam`` and
* for each of the three scalar objectives  J_sim = <sim_out, cot>,
  J_gam = gam, J_lam = lam : the token gradients projected to 4 columns
  (``dtok_<obj>`` [3,B,129,4]) and, per parameter, [norm, 4 probe dots]
  (``dpar_<obj>/<key>``).

* ``dev32/<key>`` for every key above: the relative deviation of the reference's OWN fp32 run from its
  fp64 run (same fingerprints).  The parity tests derive their per-quantity tolerance from it
  (``golden_util.derived_tol``): where fp32 conditioning costs the reference itself more than the flat
  1e-4, the CUDA path is held to a small multiple of what the reference loses, not to a hand-set number.

``bf16dev_<case>.npz`` (cases ``golden_util.BF16_CASES``): the same reference modules run under
``torch.autocast(bfloat16)`` on bf16-rounded tokens against their fp64 run on the same rounded values:
``dev/<key>`` relative deviations, ``mask_flips`` (selection flips of the reference's own reduced-precision
run) and ``lam_flip_samples`` (samples whose LAM token gradient moved by more than 2e-2: bilinear sample
points that crossed a pixel boundary).  These bound what a bf16 implementation can be asked to reproduce.

/root/reference does not exist on the GPU box; only these files travel.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

sys.path.insert(0, os.path.dirname(HERE))
from signal_b200 import synthetic as syn  # noqa: E402
import golden_util as gu  # noqa: E402

CASES = {
    # name: d, h, w, B, topk, keep_ratio, offset_gain, structured, seed
    "rgbnt201_d512": dict(d=512, h=16, w=8, B=8, k=80, keep_ratio=None, gain=1.0, structured=False, seed=101),
    "vehicle_d512": dict(d=512, h=8, w=16, B=8, k=112, keep_ratio=None, gain=40.0, structured=False, seed=202),
    "rgbnt201_d768": dict(d=768, h=16, w=8, B=8, k=80, keep_ratio=None, gain=40.0, structured=False, seed=303),
    "keepratio_d512": dict(d=512, h=16, w=8, B=4, k=40, keep_ratio=0.5, gain=60.0, structured=True, seed=404),
    "msvr_k64_d512": dict(d=512, h=8, w=16, B=6, k=64, keep_ratio=None, gain=25.0, structured=True, seed=505),
}


def import_reference():
    pkg = types.ModuleType("modeling")
    pkg.__path__ = ["/root/reference/modeling"]
    sys.modules["modeling"] = pkg
    sys.path.insert(0, "/root/reference")
    from modeling.AddModule.useA import Select_Interactive_Module
    from modeling.AddModule.useB import AlignmentM
    from modeling.AddModule.DAS import DA_sample
    return Select_Interactive_Module, AlignmentM, DA_sample


def fingerprint_param(key, g):
    g = g.detach().double().reshape(-1)
    pv = syn.probe_vector(key, g.numel())
    return np.array([g.norm().item()] + (pv @ g).tolist(), dtype=np.float64)


def build(c, SIM, ALIGN, DAS, dtype, toks=None, autocast=False):
    d, B = c["d"], c["B"]
    sim = SIM(d, k=c["k"], keep_ratio=c["keep_ratio"])
    al = ALIGN(d, c["h"], c["w"])
    if d != 512:  # useB.py:64 hard-codes 512 channels; rebuild the three DAS at width d
        for n in ("DAS_r", "DAS_n", "DAS_t"):
            setattr(al, n, DAS(1, d, 1, 4, 2, 4))
    sim.load_state_dict(syn.make_params(syn.sim_param_shapes(d), c["seed"]))
    al.load_state_dict(syn.make_params(syn.align_param_shapes(d), c["seed"] + 1, offset_gain=c["gain"]))
    if not autocast:     # (autocast: fp32 master parameters, bf16 tokens -- engine/processor.py:165)
        sim, al = sim.to(dtype), al.to(dtype)
    # the reference's LayerNorm subclass always computes in fp32 (useA.py:420-423) and
    # needs fp32 affine parameters; everything else runs in ``dtype``
    sim.modal_interactive.norm1.float()
    sim.modal_interactive.norm2.float()
    if toks is None:
        toks = syn.make_tokens(B, d, seed=c["seed"] + 2, structured=c["structured"])
    toks = [t.detach().to(dtype).requires_grad_(True) for t in toks]
    patches = [t[:, 1:] for t in toks]
    cls = [t[:, 0] for t in toks]
    # fp64 runs: ``torch.det(G.float())`` (utils/volume.py:57) would keep the Gram determinant in fp32 -- an fp32 island
    # that puts ~3e-4 of rounding noise on d(gam)/d(contra_temp) and would make the pinned "fp64" value noisier than the
    # fp32 parity tolerance.  It is lifted for the fp64 run only (Tensor.float is the identity on fp64 tensors while
    # AlignM runs), so the stored GAM values are true fp64; fp32 and bf16 runs execute the reference unmodified.
    orig_float = torch.Tensor.float
    if dtype == torch.float64:
        torch.Tensor.float = lambda self, *a, **k: self if self.dtype == torch.float64 else orig_float(self, *a, **k)
    try:
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            out = sim(*patches, *cls)
            masks = sim.token_selection.last_masks
            gam, lam = al(*patches, stage="together_CLS_Patch")
    finally:
        torch.Tensor.float = orig_float
    masks = np.stack([masks[k][..., 0].float().numpy().astype(np.uint8) for k in ("RGB", "NI", "TI")])
    return sim, al, toks, out, masks, gam, lam


def grad_record(sim, al, toks, out, gam, lam, cot, d):
    """{dtok_<obj>, dpar_<obj>/<key>} fingerprints of the three objectives' gradients."""
    rec = {}
    proj = syn.token_projection(d)
    objs = {"sim": (out.to(cot.dtype) * cot).sum(), "gam": gam, "lam": lam}
    named = [("SIM." + k, p) for k, p in sim.named_parameters()] + \
            [("AlignM." + k, p) for k, p in al.named_parameters()]
    for oname, J in objs.items():
        grads = torch.autograd.grad(J, toks + [p for _, p in named], retain_graph=True, allow_unused=True)
        gt = [torch.zeros_like(t) if g is None else g for t, g in zip(toks, grads[:3])]
        rec[f"dtok_{oname}"] = torch.stack([g.double() @ proj for g in gt]).numpy()
        for (key, p), g in zip(named, grads[3:]):
            if g is None:
                continue
            rec[f"dpar_{oname}/{key}"] = fingerprint_param(key, g)
    return rec


def deviations(lo, hi):
    """relative deviation of the reduced-precision record `lo` from the fp64 record `hi`, per key"""
    dev = {}
    for k, v in hi.items():
        if k.startswith("masks") or k not in lo:
            continue
        a, b = np.asarray(lo[k], dtype=np.float64), np.asarray(v, dtype=np.float64)
        if k.startswith("dpar_") and b.reshape(-1)[0] < 1e-12:
            continue
        dev[k] = np.float64(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    return dev


def run_case(name, c, SIM, ALIGN, DAS):
    d, B = c["d"], c["B"]
    sim32, al32, toks32, out32, masks32, gam32, lam32 = build(c, SIM, ALIGN, DAS, torch.float32)
    sim, al, toks, out, masks, gam, lam = build(c, SIM, ALIGN, DAS, torch.float64)
    cot = syn.make_cotangent(B, d, seed=c["seed"] + 3).double()
    rec = {
        "sim_out32": out32.detach().numpy(), "masks32": masks32,
        "gam32": np.float64(gam32.item()), "lam32": np.float64(lam32.item()),
        "sim_out": out.detach().numpy().astype(np.float32), "masks": masks,
        "gam": np.float64(gam.item()), "lam": np.float64(lam.item()),
    }
    g64 = grad_record(sim, al, toks, out, gam, lam, cot, d)
    g32 = grad_record(sim32, al32, toks32, out32, gam32, lam32, cot.float(), d)
    hi = dict(g64, sim_out=out.detach().numpy(), gam=rec["gam"], lam=rec["lam"])
    lo = dict(g32, sim_out=rec["sim_out32"], gam=rec["gam32"], lam=rec["lam32"])
    for k, v in g64.items():
        rec[k] = v.astype(np.float32) if k.startswith("dtok_") else v
    for k, v in deviations(lo, hi).items():
        rec["dev32/" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("   masks32==masks64:", bool((masks32 == masks).all()),
          " out32 vs out64:", float(np.linalg.norm(rec["sim_out32"] - rec["sim_out"]) / np.linalg.norm(rec["sim_out"])))
    worst = sorted(((v, k) for k, v in rec.items() if k.startswith("dev32/")), reverse=True)[:4]
    print("   largest fp32-vs-fp64 deviations of the reference:", [(k[6:], float("%.2e" % v)) for v, k in worst])
    kept = rec["masks"].reshape(3, B, -1).sum(-1).mean()
    print(f"{name}: gam={gam.item():.6f} lam={lam.item():.6f} kept={kept:.1f}/128 "
          f"out_rms={out.detach().pow(2).mean().sqrt().item():.4f}")


def run_bf16_case(name, c, SIM, ALIGN, DAS):
    """The reference's own bf16-autocast run vs its fp64 run on the same bf16-rounded tokens."""
    d, B = c["d"], c["B"]
    toks = gu.bf16_case_tokens(c)                      # bf16-rounded (and smoothed where the case says so)
    cot = syn.make_cotangent(B, d, seed=c["seed"] + 3).double()
    sim, al, t64, out, masks, gam, lam = build(c, SIM, ALIGN, DAS, torch.float64, toks=toks)
    hi = dict(grad_record(sim, al, t64, out, gam, lam, cot, d), sim_out=out.detach().numpy(), gam=gam.item(), lam=lam.item())
    simb, alb, tb, outb, masksb, gamb, lamb = build(c, SIM, ALIGN, DAS, torch.bfloat16, toks=toks, autocast=True)
    lo = dict(grad_record(simb, alb, tb, outb, gamb, lamb, cot.float(), d), sim_out=outb.detach().float().numpy(),
              gam=gamb.item(), lam=lamb.item())
    rec = {"dev/" + k: v for k, v in deviations(lo, hi).items()}
    rec["mask_flips"] = np.int64((masks != masksb).sum())
    a, b = lo["dtok_lam"], hi["dtok_lam"]
    num = np.linalg.norm((a - b).reshape(3, B, -1), axis=-1)
    den = np.linalg.norm(b.reshape(3, B, -1), axis=-1)
    rec["lam_flip_samples"] = np.int64((num > 2e-2 * np.maximum(den, 1e-30)).sum())
    np.savez_compressed(os.path.join(HERE, "bf16dev_" + name + ".npz"), **rec)
    worst = sorted(((v, k) for k, v in rec.items() if k.startswith("dev/")), reverse=True)[:5]
    print(f"bf16dev_{name}: reference bf16-autocast vs fp64: mask flips {int(rec['mask_flips'])}, LAM outlier samples "
          f"{int(rec['lam_flip_samples'])}/{3 * B}, largest deviations", [(k[4:], float("%.2e" % v)) for v, k in worst])


if __name__ == "__main__":
    SIM, ALIGN, DAS = import_reference()
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if not only or name in only:
            run_case(name, c, SIM, ALIGN, DAS)
    for name, c in gu.BF16_CASES.items():
        if not only or "bf16dev_" + name in only:
            run_bf16_case(name, c, SIM, ALIGN, DAS)
