"""GPU parity of the token producer (signal_b200.tokens.TokenProducer, csrc/tokens.cu: ln_post + proj + CLS/patch split,
clip/model.py:485-487, meta_arch.py:108-110) through the C ABI: fp32 against the golden vectors of the live reference
(1e-4), bf16 / fp16 at the training shape (B=128, 129 tokens, 768 -> 512) against the CPU oracle (2e-2 vs its fp32
evaluation, tighter vs its autocast emulation), the tower's [L,B,W] layout consumed in place, the patch-mean by-product,
and the produced map feeding the fusion head without a copy."""
import pytest
import torch
import torch.nn as nn

import tokens_cases as tc

pytestmark = pytest.mark.gpu


def _mods():
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import tokens
    return tokens


def _producer(tokens, ln_w, ln_b, proj, eps=1e-5, dev="cuda"):
    ln = nn.LayerNorm(ln_w.numel(), eps=eps).to(dev)
    with torch.no_grad():
        ln.weight.copy_(ln_w)
        ln.bias.copy_(ln_b)
    p = nn.Parameter(proj.clone().to(dev))
    return tokens.TokenProducer(ln, p), ln, p


def _run(tokens, inp, eps, dtype, lnd=False):
    """-> dict of results (fp32 on the CPU side of the comparison)"""
    tp, ln, p = _producer(tokens, inp["ln_w"], inp["ln_b"], inp["proj"], eps)
    x = inp["x"].to("cuda", dtype)
    if lnd:   # the tower's layout: [L1, B, W] permuted to [B, L1, W] (clip/model.py:484), consumed in place
        x = x.permute(1, 0, 2).contiguous().permute(1, 0, 2)
        assert not x.is_contiguous()
    x.requires_grad_(True)
    x_cash, global_feat = tp(x)
    tok = torch.cat([global_feat[:, None], x_cash], dim=1)
    assert x_cash.data_ptr() == global_feat.data_ptr() + global_feat.shape[1] * x.element_size()   # two views of one map
    assert tok.dtype == dtype
    torch.autograd.backward([x_cash, global_feat], [inp["cot"][:, 1:].to("cuda", dtype), inp["cot"][:, 0].to("cuda", dtype)])
    assert x.grad.dtype == dtype and x.grad.shape == x.shape
    return dict(tokens=tok.float(), dx=x.grad.float(), d_ln_w=ln.weight.grad, d_ln_b=ln.bias.grad, d_proj=p.grad,
                patch_mean=tp.last_patch_mean, cls=global_feat.float())


@pytest.mark.parametrize("name", sorted(tc.CASES))
@pytest.mark.parametrize("lnd", [False, True])
def test_tokens_fp32_match_reference_golden(name, lnd):
    tokens = _mods()
    inp, z = tc.load_case(name)
    got = _run(tokens, inp, float(z["eps"]), torch.float32, lnd=lnd)
    tc.check_against_golden(got, z, 1e-4, f"cuda fp32 {name}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,grid", [(128, (16, 8)), (64, (8, 16)), (5, (16, 8))])
def test_tokens_half_training_shape_vs_oracle(dtype, B, grid):
    """BASELINE.json configs: 129 tokens, ViT-B/16 width 768 -> 512.  Reference = the oracle in fp32 on the half-rounded
    inputs (2e-2, north_star's bf16 tolerance) and its autocast emulation (same roundings as the kernels: 4e-3)."""
    tokens = _mods()
    from oracle import tokens_oracle as to
    c = dict(h=grid[0], w=grid[1], width=768, out=512, B=B, seed=900 + B)
    inp = tc.gen.synthetic_inputs(c)
    inp["x"] = inp["x"].to(dtype).float()
    inp["cot"] = (inp["cot"] * 0.01).to(dtype).float()
    got = _run(tokens, inp, 1e-5, dtype, lnd=True)
    x, ln_w, ln_b, proj, cot = (inp[k] for k in ("x", "ln_w", "ln_b", "proj", "cot"))
    for emu, tol in ((None, 2e-2), (torch.bfloat16, 4e-3)):
        tok, mean, xn = to.tokens_fwd(x, ln_w, ln_b, proj, 1e-5, operand_dtype=emu)
        dx, dg, db, dp = to.tokens_bwd(x, ln_w, proj, cot, 1e-5, xn=xn, operand_dtype=emu)
        ref = dict(tokens=tok, dx=dx, d_ln_w=dg, d_ln_b=db, d_proj=dp, patch_mean=mean)
        for k, v in ref.items():
            e = tc.rel(got[k], v)
            assert e <= tol, f"{dtype} B={B} emu={emu}: {k} rel err {e:.3e} > {tol}"
    # the by-product is the mean of the ROUNDED patch rows (what AlignM's own pool pass would return)
    own = got["tokens"][:, 1:].mean(dim=1)
    assert tc.rel(got["patch_mean"], own) < 1e-5


@pytest.mark.parametrize("amp_dtype", [torch.bfloat16, torch.float16])
def test_tokens_fp32_input_under_autocast(amp_dtype):
    """torch.autocast keeps the tower's residual stream in fp32, runs layer_norm in fp32 and the matmul in half: fp32 x in,
    half tokens out, fp32 dx back -- compared with stock torch ops under the same autocast and with the oracle."""
    tokens = _mods()
    from oracle import tokens_oracle as to
    c = dict(h=8, w=16, width=768, out=512, B=64, seed=77)
    inp = tc.gen.synthetic_inputs(c)
    tp, ln, p = _producer(tokens, inp["ln_w"], inp["ln_b"], inp["proj"])
    x = inp["x"].cuda().requires_grad_(True)
    cot = (0.01 * inp["cot"]).to("cuda", amp_dtype)
    with torch.autocast("cuda", dtype=amp_dtype):
        tok = tp.tokens(x)
    assert tok.dtype == amp_dtype
    tok.backward(cot)
    got = dict(tokens=tok.float(), dx=x.grad, d_ln_w=ln.weight.grad.clone(), d_ln_b=ln.bias.grad.clone(), d_proj=p.grad.clone())
    assert x.grad.dtype == torch.float32
    # stock torch under the same autocast (what fullstep.ClipViT / the reference tower do)
    x2 = inp["x"].cuda().requires_grad_(True)
    ln.weight.grad = ln.bias.grad = p.grad = None
    with torch.autocast("cuda", dtype=amp_dtype):
        tok2 = ln(x2) @ p
    assert tok2.dtype == amp_dtype
    tok2.backward(cot)
    ref = dict(tokens=tok2.float(), dx=x2.grad, d_ln_w=ln.weight.grad, d_ln_b=ln.bias.grad, d_proj=p.grad)
    for k in ref:
        e = tc.rel(got[k], ref[k])
        assert e < 6e-3, f"autocast {amp_dtype}: {k} vs stock torch {e:.3e}"
    t32, _, xn = to.tokens_fwd(inp["x"], inp["ln_w"], inp["ln_b"], inp["proj"], 1e-5)
    dx32 = to.tokens_bwd(inp["x"], inp["ln_w"], inp["proj"], cot.float().cpu(), 1e-5, xn=xn)[0]
    assert tc.rel(got["tokens"], t32) < 2e-2 and tc.rel(got["dx"], dx32) < 2e-2


def test_tokens_feed_the_head_in_place():
    """the produced [B,129,d] maps are consumed by SIM / AlignM as strided views (no copy), fwd + bwd through both"""
    tokens = _mods()
    from signal_b200 import modules as M, synthetic as syn
    B, W, D = 16, 768, 512
    g = torch.Generator().manual_seed(7)
    ln_w, ln_b = 1.0 + 0.1 * torch.randn(W, generator=g), 0.1 * torch.randn(W, generator=g)
    proj = W ** -0.5 * torch.randn(W, D, generator=g)
    tp, ln, p = _producer(tokens, ln_w, ln_b, proj)
    sim = M.Select_Interactive_Module(D, k=80).cuda()
    al = M.AlignmentM(D, 16, 8).cuda()
    sim.load_state_dict(syn.make_params(syn.sim_param_shapes(D), 11))
    al.load_state_dict(syn.make_params(syn.align_param_shapes(D), 12))
    xs = [torch.randn(B, 129, W, generator=g).to("cuda", torch.bfloat16).requires_grad_(True) for _ in range(3)]
    views = [tp(x) for x in xs]
    patches, cls = [v[0] for v in views], [v[1] for v in views]
    out = sim(*patches, *cls)
    gam, lam = al(*patches, stage="together_CLS_Patch")
    (out.float().square().mean() + 0.2 * gam + 0.2 * lam).backward()
    assert out.shape == (B, 3 * D) and torch.isfinite(out.float()).all()
    for x in xs:
        assert x.grad is not None and torch.isfinite(x.grad.float()).all() and float(x.grad.float().abs().max()) > 0
    assert p.grad is not None and torch.isfinite(p.grad).all()


@pytest.mark.parametrize("stage", ["together_CLS_Patch", "CLS"])
def test_patch_mean_byproduct_replaces_the_gam_pool_pass(stage):
    """AlignM with the producer's patch means deposited in its ctx (SIG_FLAG_PATCH_MEAN) == AlignM pooling the tokens
    itself: same losses and gradients (the two pools add the same 128 bf16 values in a different fp32 order)."""
    tokens = _mods()
    from signal_b200 import lib, modules as M, synthetic as syn
    B, W, D = 32, 768, 512
    g = torch.Generator().manual_seed(17)
    tp, _, _ = _producer(tokens, 1.0 + 0.1 * torch.randn(W, generator=g), 0.1 * torch.randn(W, generator=g),
                         W ** -0.5 * torch.randn(W, D, generator=g))
    al = M.AlignmentM(D, 16, 8).cuda()
    al.load_state_dict(syn.make_params(syn.align_param_shapes(D), 12))
    xs = [torch.randn(B, 129, W, generator=g).to("cuda", torch.bfloat16) for _ in range(3)]
    res = []
    for use_hint in (False, True):
        toks, means = [], []
        for x in xs:
            t = tp.tokens(x).detach().requires_grad_(True)
            toks.append(t)
            means.append(tp.last_patch_mean)
        for p in al.parameters():
            p.grad = None
        n0 = lib.launch_count()
        if use_hint:
            al.patch_mean_hint = means
        r = al(*[t[:, 1:] for t in toks], stage=stage)
        launches = lib.launch_count() - n0
        assert al.patch_mean_hint is None   # consumed
        loss = r if stage == "CLS" else r[0] + 0.5 * r[1]
        loss.backward()
        res.append((launches, [float(v) for v in (r if isinstance(r, tuple) else (r,))], [t.grad.float() for t in toks],
                    float(al.contra_temp.grad)))
    (l0, v0, g0, t0), (l1, v1, g1, t1) = res
    assert l1 == l0 - 1, (l0, l1)          # exactly the pool launch is gone
    for a, b in zip(v0, v1):
        assert abs(a - b) <= 2e-5 * max(abs(a), 1e-3), (v0, v1)
    assert abs(t0 - t1) <= 1e-3 * max(abs(t0), 1e-3)
    for a, b in zip(g0, g1):
        assert tc.rel(b, a) < 1e-3


def test_tokens_rejects_cpu_and_bad_shapes():
    tokens = _mods()
    tp, _, _ = _producer(tokens, torch.ones(64), torch.zeros(64), torch.randn(64, 32))
    with pytest.raises(RuntimeError):
        tp(torch.randn(2, 129, 64))
    with pytest.raises(RuntimeError):
        tp(torch.randn(2, 129, 64, device="cuda")[:, :, ::2])
