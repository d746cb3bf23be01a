"""oracle/tokens_oracle.py against the golden vectors of the live reference's vision-tower tail (clip/model.py:485-487,
meta_arch.py:108-110; tests/golden/make_tokens_golden.py)."""
import pytest
import torch

import tokens_cases as tc
from oracle import tokens_oracle as to


@pytest.mark.parametrize("name", sorted(tc.CASES))
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4)])
def test_tokens_oracle_matches_reference(name, dtype, tol):
    inp, z = tc.load_case(name)
    x, ln_w, ln_b, proj, cot = (inp[k].to(dtype) for k in ("x", "ln_w", "ln_b", "proj", "cot"))
    eps = float(z["eps"])
    tok, mean, xn = to.tokens_fwd(x, ln_w, ln_b, proj, eps)
    dx, dg, db, dp = to.tokens_bwd(x, ln_w, proj, cot, eps, xn=xn)
    x_cash, global_feat = to.split(tok)
    assert x_cash.shape[1] == x.shape[1] - 1 and torch.equal(global_feat, tok[:, 0])
    got = dict(tokens=tok, dx=dx, d_ln_w=dg, d_ln_b=db, d_proj=dp, patch_mean=mean, cls=global_feat)
    tc.check_against_golden(got, z, tol, f"oracle {name} {dtype}")


def test_tokens_oracle_bf16_emulation_is_close_to_fp32():
    """the autocast emulation used as the tight bf16 reference on the GPU stays within bf16 rounding of the fp32 result"""
    inp, z = tc.load_case("small")
    x, ln_w, ln_b, proj = inp["x"], inp["ln_w"], inp["ln_b"], inp["proj"]
    t32, m32, _ = to.tokens_fwd(x, ln_w, ln_b, proj, float(z["eps"]))
    t16, m16, _ = to.tokens_fwd(x, ln_w, ln_b, proj, float(z["eps"]), operand_dtype=torch.bfloat16)
    assert tc.rel(t16, t32) < 1e-2 and tc.rel(m16, m32) < 1e-2


def test_tokens_oracle_explicit_backward_matches_autograd():
    """the oracle's hand-written backward (what the CUDA kernels are compared with) against torch.autograd on its forward"""
    g = torch.Generator().manual_seed(5)
    B, L1, W, D = 3, 9, 48, 24
    x = (torch.randn(B, L1, W, generator=g, dtype=torch.float64) * 1.7 + 0.3).requires_grad_(True)
    ln_w = (1.0 + 0.2 * torch.randn(W, generator=g, dtype=torch.float64)).requires_grad_(True)
    ln_b = (0.1 * torch.randn(W, generator=g, dtype=torch.float64)).requires_grad_(True)
    proj = (torch.randn(W, D, generator=g, dtype=torch.float64) / W ** 0.5).requires_grad_(True)
    cot = torch.randn(B, L1, D, generator=g, dtype=torch.float64)
    tok, mean, xn = to.tokens_fwd(x, ln_w, ln_b, proj, 1e-5)
    tok.backward(cot)
    dx, dg, db, dp = to.tokens_bwd(x.detach(), ln_w.detach(), proj.detach(), cot, 1e-5, xn=xn.detach())
    for got, ref in ((dx, x.grad), (dg, ln_w.grad), (db, ln_b.grad), (dp, proj.grad)):
        assert tc.rel(got, ref) < 1e-12
    # the patch mean is the mean over rows 1.. of the tokens (meta_arch.py:108-110: row 0 is the CLS token)
    assert tc.rel(mean, tok.detach()[:, 1:].mean(dim=1)) < 1e-15
