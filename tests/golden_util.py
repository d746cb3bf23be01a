"""Shared replay/compare harness for the golden vectors (tests/golden/*.npz)."""
import os

import numpy as np
import torch

from signal_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror tests/golden/make_golden.py::CASES
CASES = {
    "rgbnt201_d512": dict(d=512, h=16, w=8, B=8, k=80, keep_ratio=None, gain=1.0, structured=False, seed=101),
    "vehicle_d512": dict(d=512, h=8, w=16, B=8, k=112, keep_ratio=None, gain=40.0, structured=False, seed=202),
    "rgbnt201_d768": dict(d=768, h=16, w=8, B=8, k=80, keep_ratio=None, gain=40.0, structured=False, seed=303),
    "keepratio_d512": dict(d=512, h=16, w=8, B=4, k=40, keep_ratio=0.5, gain=60.0, structured=True, seed=404),
    "msvr_k64_d512": dict(d=512, h=8, w=16, B=6, k=64, keep_ratio=None, gain=25.0, structured=True, seed=505),
}


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def case_inputs(c, dtype=torch.float32):
    d, B = c["d"], c["B"]
    sim_p = syn.make_params(syn.sim_param_shapes(d), c["seed"])
    al_p = syn.make_params(syn.align_param_shapes(d), c["seed"] + 1, offset_gain=c["gain"])
    toks = syn.make_tokens(B, d, seed=c["seed"] + 2, structured=c["structured"], dtype=dtype)
    cot = syn.make_cotangent(B, d, seed=c["seed"] + 3)
    return sim_p, al_p, toks, cot


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def fingerprint_param(key, g):
    g = g.detach().double().cpu().reshape(-1)
    pv = syn.probe_vector(key, g.numel())
    return np.array([g.norm().item()] + (pv @ g).tolist(), dtype=np.float64)


def project_tokens(grads, d):
    proj = syn.token_projection(d)
    return torch.stack([g.detach().double().cpu() @ proj for g in grads]).numpy()


def param_tol(key, objective, tol):
    """Documented fp32-conditioning exceptions to the flat parity tolerance.

    * d(gam)/d(contra_temp) = sum_ij dZ_ij V_ij / tau^2 with sum_j dZ_ij = 0: on iid
      tokens V ~ 1 everywhere, so the sum cancels to ~1e-3 of its terms and the
      reference's own fp32 ``torch.det`` (volume.py:57, fp32 even in the fp64 golden
      run) shows up as ~3e-4 relative noise; an fp32 evaluation of V (1e-7 relative per
      entry, times 1/tau^2 = 204, over a result of ~0.05) sits at ~1e-3.  Hence 3e-3.
    * LAM parameter gradients with the offsets pushed into tanh saturation
      (offset_gain >= 25) pass through 1 - tanh(o)^2, which loses digits in fp32: the
      fp32 *reference* deviates from its own fp64 run by ~2e-4 there.
    """
    if key.endswith("contra_temp"):
        return max(tol, 3e-3)
    if objective == "lam" and tol >= 1e-2:
        # bf16 runs: the offset-path gradients are sums over channels/positions of terms with random
        # signs; the bf16 rounding of the folded weights and of the stored pre-activation shows up
        # amplified (measured 1-4% on these parameters while every per-token gradient is < 2%)
        return max(tol, 5e-2)
    if objective == "lam" and tol > 1e-5:
        return max(tol, 5e-4)
    return tol
