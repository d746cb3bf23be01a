"""GPU parity: CUDA drop-in modules (through the C ABI) vs the golden vectors from the live
reference and vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import signal_oracle as so

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4   # BASELINE.json: 1e-4 relative in fp32
BF16_TOL = 2e-2   # BASELINE.json: 2e-2 in bf16


def _harness():
    import gpu_harness
    return gpu_harness


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("name", list(gu.CASES))
def test_fp32_matches_reference_golden(name, packed):
    h = _harness()
    c = gu.CASES[name]
    got = h.cuda_record(c, torch.float32, packed=packed)
    h.compare_records(got, gu.load(name), FP32_TOL, ref_masks_key="masks32", label=name)


def _oracle_record(c, sim_p, al_p, toks, cot):
    sim_p = {k: v.clone().requires_grad_(True) for k, v in sim_p.items()}
    al_p = {k: v.clone().requires_grad_(True) for k, v in al_p.items()}
    toks = [t.clone().float().requires_grad_(True) for t in toks]
    out, gam, lam, masks = so.head_forward(sim_p, al_p, toks, c["k"], c["h"], c["w"], c["keep_ratio"])
    rec = {"sim_out": out.detach().numpy(), "masks": np.stack([m[..., 0].numpy().astype(np.uint8) for m in masks]),
           "gam": gam.item(), "lam": lam.item()}
    named = [("SIM." + k, p) for k, p in sim_p.items()] + [("AlignM." + k, p) for k, p in al_p.items()]
    for oname, J in {"sim": (out * cot).sum(), "gam": gam, "lam": lam}.items():
        grads = torch.autograd.grad(J, toks + [p for _, p in named], retain_graph=True, allow_unused=True)
        gt = [torch.zeros_like(t) if g is None else g for t, g in zip(toks, grads[:3])]
        rec[f"dtok_{oname}"] = gu.project_tokens(gt, c["d"])
        for (key, _), g in zip(named, grads[3:]):
            if g is None:
                continue
            rec[f"dpar_{oname}/{key}"] = gu.fingerprint_param(key, g)
    return rec


@pytest.mark.parametrize("name", ["rgbnt201_d512", "rgbnt201_d768", "vehicle_d512"])
def test_bf16_matches_oracle_on_rounded_inputs(name):
    """bf16 tokens on the GPU vs the fp32 oracle fed the same bf16-rounded values."""
    h = _harness()
    c = gu.CASES[name]
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    toks = [t.to(torch.bfloat16) for t in toks]
    got = h.cuda_record(c, torch.bfloat16, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    ref = _oracle_record(c, sim_p, al_p, [t.float() for t in toks], cot)
    flips = int((got["masks"] != ref["masks"]).sum())
    assert flips <= 2, f"{flips} mask flips vs the oracle on identical (bf16-rounded) inputs"
    h.compare_records(got, ref, BF16_TOL, check_masks=False, label=name)


@pytest.mark.parametrize("d,hw,k", [(512, (16, 8), 80), (768, (16, 8), 80)])
def test_full_batch_b128_matches_oracle(d, hw, k):
    """BASELINE.json config #2 size (B=128) in fp32 against the oracle."""
    h = _harness()
    c = dict(d=d, h=hw[0], w=hw[1], B=128, k=k, keep_ratio=None, gain=30.0, structured=False, seed=900 + d)
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    got = h.cuda_record(c, torch.float32, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    ref = _oracle_record(c, sim_p, al_p, toks, cot)
    flips = int((got["masks"] != ref["masks"]).sum())
    assert flips == 0, f"{flips} mask flips at B=128"
    h.compare_records(got, ref, FP32_TOL, check_masks=False, label=f"B128 d{d}")


@pytest.mark.parametrize("name", ["rgbnt201_d512", "rgbnt201_d768", "vehicle_d512"])
def test_bf16_tensor_core_path_vs_simt_path(name):
    """Same bf16 inputs through the tcgen05 path and through the fp32 SIMT kernels (FORCE_SIMT)."""
    from signal_b200 import lib
    h = _harness()
    c = gu.CASES[name]
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    toks = [t.to(torch.bfloat16) for t in toks]
    fast = h.cuda_record(c, torch.bfloat16, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    slow = h.cuda_record(c, torch.bfloat16, flags=lib.FLAG_FORCE_SIMT, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    h.compare_records(fast, slow, BF16_TOL, check_masks=True, label=name + " tc-vs-simt")


def test_bf16_full_batch_b128_matches_oracle():
    """BASELINE.json config #2 (B=128, bf16, d=768) against the fp32 oracle on the rounded inputs."""
    h = _harness()
    c = dict(d=768, h=16, w=8, B=128, k=80, keep_ratio=None, gain=30.0, structured=False, seed=4242)
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    toks = [t.to(torch.bfloat16) for t in toks]
    got = h.cuda_record(c, torch.bfloat16, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    ref = _oracle_record(c, sim_p, al_p, [t.float() for t in toks], cot)
    flips = int((got["masks"] != ref["masks"]).sum())
    assert flips <= 4, f"{flips} mask flips at B=128 bf16"
    h.compare_records(got, ref, BF16_TOL, check_masks=False, label="B128 d768 bf16")
