"""CPU: the oracle restatement vs golden vectors produced by the live reference."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import signal_oracle as so


# fp64: the restatement must agree with the fp64 run of the reference code to rounding
# (the reference keeps fp32 islands even when run in fp64 -- the offset range tensor,
# DAS.py:144, and its LayerNorm, useA.py:420-423 -- hence 5e-6, not 1e-12; the third island,
# ``torch.det(G.float())`` at volume.py:57, is lifted by the golden generator for its fp64 run).
# fp32: 1e-4, the parity tolerance, or 3x what the reference's OWN fp32 run loses against its
# fp64 run on that quantity where that is more (``dev32/<key>`` in the golden file).
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 5e-6), (torch.float32, 1e-4)])
@pytest.mark.parametrize("name", list(gu.CASES))
def test_oracle_matches_reference_golden(name, dtype, tol):
    c = gu.CASES[name]
    rec = gu.load(name)
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    sim_p = {k: v.to(dtype) for k, v in sim_p.items()}
    al_p = {k: v.to(dtype) for k, v in al_p.items()}
    toks = [t.to(dtype) for t in toks]
    cot = cot.to(dtype)
    for p in list(sim_p.values()) + list(al_p.values()):
        p.requires_grad_(True)
    toks = [t.requires_grad_(True) for t in toks]
    out, gam, lam, masks = so.head_forward(sim_p, al_p, toks, c["k"], c["h"], c["w"], c["keep_ratio"])

    got_masks = np.stack([m[..., 0].numpy().astype(np.uint8) for m in masks])
    ref_masks = rec["masks"] if dtype == torch.float64 else rec["masks32"]
    assert np.array_equal(got_masks, ref_masks), "selected-token masks differ from the reference"
    assert gu.rel_err(out.detach().numpy(), rec["sim_out"]) < max(tol, 2e-7)
    assert abs(gam.item() - rec["gam"]) < tol * abs(rec["gam"])
    assert abs(lam.item() - rec["lam"]) < tol * abs(rec["lam"])

    named = [("SIM." + k, p) for k, p in sim_p.items()] + [("AlignM." + k, p) for k, p in al_p.items()]
    objs = {"sim": (out * cot).sum(), "gam": gam, "lam": lam}
    for oname, J in objs.items():
        grads = torch.autograd.grad(J, toks + [p for _, p in named], retain_graph=True, allow_unused=True)
        gt = [torch.zeros_like(t) if g is None else g for t, g in zip(toks, grads[:3])]
        bound = max(tol, 2e-7) if dtype == torch.float64 else gu.derived_tol(rec, f"dtok_{oname}", tol)
        assert gu.rel_err(gu.project_tokens(gt, c["d"]), rec[f"dtok_{oname}"]) < bound, oname
        for (key, _), g in zip(named, grads[3:]):
            rkey = f"dpar_{oname}/{key}"
            if g is None or float(g.abs().max()) == 0.0:
                # the reference never reaches this parameter either (or gives exact zeros)
                assert rkey not in rec or rec[rkey][0] < 1e-12, key
                continue
            assert rkey in rec, key
            bound = tol if dtype == torch.float64 else gu.derived_tol(rec, rkey, tol)
            assert gu.rel_err(gu.fingerprint_param(key, g), rec[rkey]) < bound, (oname, key)


def test_topk_ties_pick_lowest_index():
    s = torch.tensor([[1., 3., 3., 3., 2., 3., 0.]])
    m = so.topk_mask_lowest_index(s, 2)
    assert m.tolist() == [[False, True, True, False, False, False, False]]
    m = so.topk_mask_lowest_index(torch.zeros(1, 128), 5)
    assert m[0, :5].all() and not m[0, 5:].any()


def test_volume3_matches_det():
    g = torch.Generator().manual_seed(0)
    l, v, a = (torch.nn.functional.normalize(torch.randn(5, 64, generator=g, dtype=torch.float64), dim=-1)
               for _ in range(3))
    V = so.volume3(l, v, a)
    for i in range(5):
        for j in range(5):
            M = torch.stack([l[i], v[j], a[j]])
            assert abs(V[i, j] - torch.sqrt(torch.abs(torch.det(M @ M.T)))) < 1e-10
