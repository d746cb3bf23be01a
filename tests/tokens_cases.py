"""Shared helpers for the token-producer tests: the case table and seeded inputs of tests/golden/make_tokens_golden.py
(imported by path so that generator and tests cannot drift apart) and the fingerprint comparison."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_tokens_golden", os.path.join(HERE, "golden", "make_tokens_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)   # (touches /root/reference only inside main())
CASES = gen.CASES


def load_case(name):
    """-> (inputs dict of fp32 torch tensors, golden dict)"""
    c = CASES[name]
    z = dict(np.load(os.path.join(HERE, "golden", f"tokens_{name}.npz")))
    if c["tower"]:
        inp = {k: torch.from_numpy(z[k]) for k in ("x", "ln_w", "ln_b", "proj", "cot")}
    else:
        inp = gen.synthetic_inputs(c)
    return inp, z


def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1) if not torch.is_tensor(a) else a.detach().double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1) if not torch.is_tensor(b) else b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_against_golden(got, z, tol, label):
    """got: dict tokens/dx/d_ln_w/d_ln_b/d_proj (+ patch_mean, cls) of torch tensors.  Full tensors where the golden file has
    them (relative L2 and max-abs), fingerprints (norm + 4 probes, each relative to the norm) otherwise."""
    errs = {}
    for k, v in got.items():
        if "ref/" + k in z:
            ref = torch.from_numpy(z["ref/" + k])
            e = rel(v, ref)
            scale = float(ref.abs().max())
            emax = float((v.detach().double().cpu() - ref).abs().max()) / max(scale, 1e-30)
            errs[k] = max(e, emax / 10.0)   # max-abs error, relative to the largest entry, at 10x the L2 tolerance
        elif "fp/" + k in z:
            fp = gen.fingerprint(k, v.detach().cpu())
            errs[k] = float(np.abs(fp - z["fp/" + k]).max() / z["fp/" + k][0])
        else:
            continue
        t = max(tol, 3.0 * float(z.get("dev32/" + k, 0.0))) if tol <= 1e-3 else tol
        assert errs[k] <= t, f"{label}: {k} err {errs[k]:.3e} > {t:.1e}"
    return errs
