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

from signal_b200 import synthetic as syn  # noqa: E402

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


def build(c, SIM, ALIGN, DAS, dtype):
    d, B = c["d"], c["B"]
    sim = SIM(d, k=c["k"], keep_ratio=c["keep_ratio"])
    al = ALIGN(d, c["h"], c["w"])
    if d != 512:  # useB.py:64 hard-codes 512 channels; rebuild the three DAS at width d
        for n in ("DAS_r", "DAS_n", "DAS_t"):
            setattr(al, n, DAS(1, d, 1, 4, 2, 4))
    sim.load_state_dict(syn.make_params(syn.sim_param_shapes(d), c["seed"]))
    al.load_state_dict(syn.make_params(syn.align_param_shapes(d), c["seed"] + 1, offset_gain=c["gain"]))
    sim, al = sim.to(dtype), al.to(dtype)
    # the reference's LayerNorm subclass always computes in fp32 (useA.py:420-423) and
    # needs fp32 affine parameters; everything else runs in ``dtype``
    sim.modal_interactive.norm1.float()
    sim.modal_interactive.norm2.float()
    toks = [t.to(dtype).requires_grad_(True) for t in
            syn.make_tokens(B, d, seed=c["seed"] + 2, structured=c["structured"])]
    patches = [t[:, 1:] for t in toks]
    cls = [t[:, 0] for t in toks]
    out = sim(*patches, *cls)
    masks = sim.token_selection.last_masks
    gam, lam = al(*patches, stage="together_CLS_Patch")
    masks = np.stack([masks[k][..., 0].numpy().astype(np.uint8) for k in ("RGB", "NI", "TI")])
    return sim, al, toks, out, masks, gam, lam


def run_case(name, c, SIM, ALIGN, DAS):
    d, B = c["d"], c["B"]
    _, _, _, out32, masks32, gam32, lam32 = build(c, SIM, ALIGN, DAS, torch.float32)
    sim, al, toks, out, masks, gam, lam = build(c, SIM, ALIGN, DAS, torch.float64)
    cot = syn.make_cotangent(B, d, seed=c["seed"] + 3).double()
    rec = {
        "sim_out32": out32.detach().numpy(), "masks32": masks32,
        "gam32": np.float64(gam32.item()), "lam32": np.float64(lam32.item()),
        "sim_out": out.detach().numpy().astype(np.float32), "masks": masks,
        "gam": np.float64(gam.item()), "lam": np.float64(lam.item()),
    }
    proj = syn.token_projection(d)
    objs = {"sim": (out * cot).sum(), "gam": gam, "lam": lam}
    named = [("SIM." + k, p) for k, p in sim.named_parameters()] + \
            [("AlignM." + k, p) for k, p in al.named_parameters()]
    for oname, J in objs.items():
        grads = torch.autograd.grad(J, toks + [p for _, p in named], retain_graph=True, allow_unused=True)
        gt = [torch.zeros_like(t) if g is None else g for t, g in zip(toks, grads[:3])]
        rec[f"dtok_{oname}"] = torch.stack([g.double() @ proj for g in gt]).numpy().astype(np.float32)
        for (key, p), g in zip(named, grads[3:]):
            if g is None:
                continue
            rec[f"dpar_{oname}/{key}"] = fingerprint_param(key, g)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("   masks32==masks64:", bool((masks32 == masks).all()),
          " out32 vs out64:", float(np.linalg.norm(rec["sim_out32"] - rec["sim_out"]) / np.linalg.norm(rec["sim_out"])))
    kept = rec["masks"].reshape(3, B, -1).sum(-1).mean()
    print(f"{name}: gam={gam.item():.6f} lam={lam.item():.6f} kept={kept:.1f}/128 "
          f"out_rms={out.detach().pow(2).mean().sqrt().item():.4f}")


if __name__ == "__main__":
    SIM, ALIGN, DAS = import_reference()
    torch.set_num_threads(8)
    for name, c in CASES.items():
        run_case(name, c, SIM, ALIGN, DAS)
