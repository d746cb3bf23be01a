"""Golden vectors for volume_computation4 / volume_computation5 from the LIVE reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_volume_golden.py

Imports utils/volume.py from /root/reference (read-only) and runs the reference's own functions on seeded inputs, forward and
autograd backward, in fp32 (what the reference computes: ``torch.det(G.float())``, volume.py:112,178); for the pin an fp64
evaluation of the same torch ops on the same inputs is stored next to it (``torch.det`` of the fp64 Gram stack) together with
the deviation of the reference's fp32 run from it (`dev32/*`).  Inputs are regenerated from the seed by the tests.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True

CASES = {
    # n modalities, B1 (language), B2 (others), feature dim, correlation of the modalities (0 = independent), seed
    "vol4_indep": dict(n=4, B1=12, B2=9, d=64, corr=0.0, seed=41),
    "vol4_aligned": dict(n=4, B1=16, B2=16, d=128, corr=0.8, seed=42),
    "vol5_indep": dict(n=5, B1=10, B2=14, d=96, corr=0.0, seed=43),
    "vol5_aligned": dict(n=5, B1=16, B2=16, d=128, corr=0.7, seed=44),
}


def inputs(c):
    """L2-normalised features (what Cls_Align feeds volume_computation3, useB.py:99-101); `corr` mixes a shared direction in."""
    g = torch.Generator().manual_seed(c["seed"])
    B = max(c["B1"], c["B2"])
    shared = torch.randn(B, c["d"], generator=g, dtype=torch.float64)
    feats = []
    for k in range(c["n"]):
        rows = c["B1"] if k == 0 else c["B2"]
        x = c["corr"] * shared[:rows] + (1.0 - c["corr"]) * torch.randn(rows, c["d"], generator=g, dtype=torch.float64)
        feats.append(torch.nn.functional.normalize(x, dim=-1).float())
    cot = torch.randn(c["B1"], c["B2"], generator=g).float()
    return feats, cot


def import_reference():
    pkg = types.ModuleType("utils")
    pkg.__path__ = ["/root/reference/utils"]
    sys.modules["utils"] = pkg
    sys.path.insert(0, "/root/reference")
    from utils.volume import volume_computation4, volume_computation5
    return {4: volume_computation4, 5: volume_computation5}


def gram_stack64(feats):
    """the Gram stack the reference builds (volume.py:86-108 / 141-174), in fp64"""
    f = [t.double() for t in feats]
    n = len(f)
    B1, B2 = f[0].shape[0], f[1].shape[0]
    rows = []
    for a in range(n):
        row = []
        for b in range(n):
            if a == 0 and b == 0:
                row.append((f[0] * f[0]).sum(-1)[:, None].expand(B1, B2))
            elif a == 0 or b == 0:
                row.append(f[0] @ f[max(a, b)].T)
            else:
                row.append((f[a] * f[b]).sum(-1)[None, :].expand(B1, B2))
        rows.append(torch.stack(row, dim=-1))
    return torch.stack(rows, dim=-2)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    fns = import_reference()
    for name, c in CASES.items():
        feats, cot = inputs(c)
        xs = [t.clone().requires_grad_(True) for t in feats]
        v32 = fns[c["n"]](*xs)                      # the reference, as it is
        v32.backward(cot)
        g32 = [x.grad.clone() for x in xs]
        xd = [t.double().requires_grad_(True) for t in feats]
        v64 = torch.sqrt(torch.abs(torch.det(gram_stack64(xd))))
        v64.backward(cot.double())
        out = {"ref/vol": v64.detach().numpy(), "dev32/vol": np.float64(rel(v32.detach(), v64.detach())), "vol32": v32.detach().numpy()}
        for k, (a, b) in enumerate(zip(g32, xd)):
            out[f"ref/d{k}"] = b.grad.numpy()
            out[f"dev32/d{k}"] = np.float64(rel(a, b.grad))
        np.savez_compressed(os.path.join(HERE, f"volume_{name}.npz"), **out)
        print(name, "V range", float(v64.min()), float(v64.max()), {k: float(v) for k, v in out.items() if k.startswith("dev32")})


if __name__ == "__main__":
    main()
