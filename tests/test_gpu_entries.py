"""GPU: every exported entry point of the drop-in API is EXECUTED and compared with the oracle on full tensors --
the stand-alone modules and functions a user of the reference can call besides the two top-level modules:

  volume_computation3          utils/volume.py:14-62        sig_volume3_fwd / _bwd
  DA_sample.forward            DAS.py:107-165               sig_das_fwd / _bwd
  TokenSelection.forward       useA.py:223-325              sig_sim_select_fwd (selected != NULL), sig_mask_mul_bwd
  TokenSelection.intra_/inter_modal_token_selection  useA.py:50-221   sig_sim_select_fwd (which = 1 | 2) on production scores
  ModalInteractive.forward     useA.py:364-411              sig_sim_attn_fwd / _bwd (masks NULL)
  AlignmentM.forward(stage="CLS"), Cls_Align, patch_Align   useB.py:76-190
  forward hooks on the sub-modules (zablation/CAM.py:164), instance monkey-patching of patch_Align
  (zablation/offestvisual.py:209-214), torch.no_grad() / eval().
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import signal_oracle as so
from signal_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def _mods():
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import modules as M
    return M


def _err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300)), float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def _close(a, b, tol, what):
    e2, em = _err(a, b)
    assert e2 < tol and em < 4 * tol, (what, "rel-L2 %.3e max-abs %.3e tol %.1e" % (e2, em, tol))
    return e2


# ---------------------------------------------------------------------------------------------
# volume_computation3
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B1,B2,d,corr", [(8, 8, 512, 0.0), (16, 24, 768, 0.0), (24, 16, 512, 0.6), (128, 128, 768, 0.5), (5, 3, 64, 0.8)])
def test_volume3_fwd_bwd(B1, B2, d, corr):
    """B1 != B2, iid and deliberately correlated inputs (small volumes: the determinant cancels).  The bound is derived
    in the test: an fp32 evaluation of the spec (the oracle run in fp32) against its fp64 run on the same inputs."""
    M = _mods()
    g = torch.Generator().manual_seed(B1 * 100 + B2)
    nrm = lambda t: t / t.norm(dim=-1, keepdim=True)
    u = torch.randn(max(B1, B2), d, generator=g)
    l = nrm(corr * u[:B1] + (1 - corr) * torch.randn(B1, d, generator=g))
    v = nrm(corr * u[:B2] + (1 - corr) * torch.randn(B2, d, generator=g))
    a = nrm(corr * u[:B2] + (1 - corr) * torch.randn(B2, d, generator=g))
    dvol = torch.randn(B1, B2, generator=g)

    def run(fn, tensors, dev):
        xs = [t.clone().to(dev).requires_grad_(True) for t in tensors]
        vol = fn(*xs)
        grads = torch.autograd.grad(vol, xs, dvol.to(vol.dtype).to(dev))
        return [vol] + list(grads)

    ref = run(so.volume3, [t.double() for t in (l, v, a)], "cpu")
    ref32 = run(so.volume3, (l, v, a), "cpu")
    got = run(M.volume_computation3, (l, v, a), "cuda")
    assert got[0].dtype == torch.float32 and got[0].shape == (B1, B2)
    for name, x, r, r32 in zip(("vol", "dl", "dv", "da"), got, ref, ref32):
        tol = max(FP32_TOL, 3.0 * _err(r32, r)[0])
        _close(x, r, tol, f"volume3 {name}")


def test_volume3_half_inputs_round_trip():
    """language/video/audio may arrive in half precision under autocast (volume.py:38-39 computes in the input dtype and
    calls G.float()); gradients come back in the input dtype."""
    M = _mods()
    g = torch.Generator().manual_seed(1)
    xs = [torch.nn.functional.normalize(torch.randn(12, 256, generator=g), dim=-1).to(torch.bfloat16).cuda().requires_grad_(True)
          for _ in range(3)]
    vol = M.volume_computation3(*xs)
    vol.sum().backward()
    ref = so.volume3(*[x.detach().float().cpu() for x in xs])
    _close(vol, ref, 1e-4, "volume3 on bf16 inputs")
    assert all(x.grad is not None and x.grad.dtype == torch.bfloat16 for x in xs)


# ---------------------------------------------------------------------------------------------
# DA_sample stand-alone
# ---------------------------------------------------------------------------------------------
def _das_inputs(d, h, w, B, gain, seed, smooth):
    al_p = syn.make_params(syn.align_param_shapes(d), seed, offset_gain=gain)
    toks = syn.make_tokens(B, d, seed=seed + 1)
    if smooth:
        toks = syn.smooth_patches(toks, h, w)
    return al_p, toks[0][:, 1:].contiguous()


@pytest.mark.parametrize("dtype,tol,gain,smooth", [(torch.float32, FP32_TOL, 1.0, False), (torch.float32, FP32_TOL, 30.0, False),
                                                   (torch.bfloat16, BF16_TOL, 1.0, True)])
@pytest.mark.parametrize("d,h,w,B", [(512, 16, 8, 6), (512, 8, 16, 5), (768, 16, 8, 4)])
def test_da_sample_standalone(d, h, w, B, dtype, tol, gain, smooth):
    """DA_sample(1, d, 1, 4, 2, 4)(x [B,C,H,W]) -> [B,C,H/4,W/4] with the reference's NCHW calling convention
    (useB.py:143-159 passes a channels-last view), gradients to x and all seven parameters."""
    M = _mods()
    al_p, x_tok = _das_inputs(d, h, w, B, gain, 1000 + d + h, smooth)
    x_tok = x_tok.to(dtype)
    das = M.DA_sample(1, d, 1, 4, 2, 4)
    das.load_state_dict({k[len("DAS_n."):]: v for k, v in al_p.items() if k.startswith("DAS_n.")})
    das = das.cuda()
    x = x_tok.cuda().reshape(B, h, w, d).permute(0, 3, 1, 2).requires_grad_(True)      # [B,C,H,W] view of channels-last storage
    out = das(x)
    assert out.shape == (B, d, h // 4, w // 4) and out.dtype == dtype
    g = torch.Generator().manual_seed(5)
    cot = torch.randn(B, (h // 4) * (w // 4), d, generator=g)
    names = [n for n, _ in das.named_parameters()]
    grads = torch.autograd.grad(out, [x] + [p for _, p in das.named_parameters()],
                                cot.reshape(B, h // 4, w // 4, d).permute(0, 3, 1, 2).to("cuda", dtype))

    def oracle(dt):
        prm = {k: v.clone().to(dt).requires_grad_(True) for k, v in al_p.items() if k.startswith("DAS_n.")}
        xt = x_tok.to(dt).clone().requires_grad_(True)
        s = so.da_sample(prm, xt, h, w, "DAS_n.")
        gr = torch.autograd.grad(s, [xt] + [prm["DAS_n." + n] for n in names], cot.to(dt))
        return s, gr

    s64, g64 = oracle(torch.float64)
    s32, g32 = oracle(torch.float32)
    got_out = out.permute(0, 2, 3, 1).reshape(B, -1, d)
    _close(got_out, s64, max(tol, 3 * _err(s32, s64)[0]), "DA_sample out")
    gx = grads[0].permute(0, 2, 3, 1).reshape(B, h * w, d)
    if dtype == torch.float32:
        _close(gx, g64[0], max(tol, 3 * _err(g32[0], g64[0])[0]), "DA_sample dx")
        for n, a, r, r32 in zip(names, grads[1:], g64[1:], g32[1:]):
            _close(a, r, max(tol, 3 * _err(r32, r)[0]), "DA_sample d" + n)
    else:
        # per-sample: a sample point that crosses a pixel boundary under bf16 rounding changes that sample's offset-path
        # gradient (gpu_harness.lam_flip_samples); at most one such sample here, everything else within 2e-2
        num = (gx.detach().double().cpu() - g64[0]).reshape(B, -1).norm(dim=1)
        den = g64[0].reshape(B, -1).norm(dim=1)
        bad = int((num > tol * den).sum())
        assert bad <= 1, ("DA_sample dx per-sample", (num / den).tolist())
        if bad == 0:
            for n, a, r in zip(names, grads[1:], g64[1:]):
                e2, _ = _err(a, r)
                assert e2 < 2.5 * tol, ("DA_sample d" + n, e2)


# ---------------------------------------------------------------------------------------------
# TokenSelection
# ---------------------------------------------------------------------------------------------
def _sel_setup(d, k, keep_ratio, B, seed, dtype, structured=False):
    M = _mods()
    sim_p = syn.make_params(syn.sim_param_shapes(d), seed)
    ts = M.TokenSelection(d, k, keep_ratio)
    ts.load_state_dict({k_[len("token_selection."):]: v for k_, v in sim_p.items() if k_.startswith("token_selection.")})
    toks = [t.to(dtype) for t in syn.make_tokens(B, d, seed=seed + 1, structured=structured)]
    return ts.cuda(), sim_p, toks


@pytest.mark.parametrize("d,k,keep_ratio,B,structured", [(512, 80, None, 8, False), (768, 40, None, 5, True), (512, 40, 0.5, 6, True),
                                                         (512, 112, 0.75, 128, False)])
def test_token_selection_forward_selected_and_backward(d, k, keep_ratio, B, structured):
    """TokenSelection.forward returns patches * mask (useA.py:318-320): bit-exact against the oracle (zero rows for
    unselected tokens), masks bit-exact, and the backward is dselected * mask exactly (sig_mask_mul_bwd)."""
    ts, sim_p, toks = _sel_setup(d, k, keep_ratio, B, 77 + d + k, torch.float32, structured)
    tk = [t.cuda().requires_grad_(True) for t in toks]
    patches, cls = [t[:, 1:] for t in tk], [t[:, 0] for t in tk]
    sel = ts(*patches, *cls)
    ref_sel, ref_masks = so.token_selection(sim_p, [t[:, 1:] for t in toks], [t[:, 0] for t in toks], k, keep_ratio)
    g = torch.Generator().manual_seed(9)
    cots = [torch.randn(B, 128, d, generator=g) for _ in range(3)]
    grads = torch.autograd.grad(sel, tk, [c.cuda() for c in cots])
    for m, key in enumerate(("RGB", "NI", "TI")):
        mask = ts.last_masks[key]
        assert mask.shape == (B, 128, 1) and mask.dtype == torch.float32
        assert torch.equal(mask.cpu(), ref_masks[m].float()), f"mask {key} differs"
        assert sel[m].shape == (B, 128, d)
        assert torch.equal(sel[m].detach().cpu(), ref_sel[m]), f"selected {key} not bit-exact"
        dropped = ref_masks[m][..., 0] == 0
        assert dropped.any() and float(sel[m].detach().cpu()[dropped].abs().max()) == 0.0      # zero rows stay in the sequence
        want = torch.zeros(B, 129, d)
        want[:, 1:] = cots[m] * ref_masks[m]
        assert torch.equal(grads[m].cpu(), want), f"d(patches) {key} != dselected * mask"
    if keep_ratio is not None:
        assert all(int(ts.last_masks[key].sum(1).min()) == int(ts.last_masks[key].sum(1).max()) == int(128 * keep_ratio)
                   for key in ("RGB", "NI", "TI"))


def test_token_selection_forward_bf16_selected_is_masked_copy():
    """bf16 tokens: whatever the selection, selected must equal patches * last_masks bit for bit."""
    ts, sim_p, toks = _sel_setup(512, 80, None, 8, 31, torch.bfloat16)
    tk = [t.cuda() for t in toks]
    sel = ts(*[t[:, 1:] for t in tk], *[t[:, 0] for t in tk])
    for m, key in enumerate(("RGB", "NI", "TI")):
        want = tk[m][:, 1:] * ts.last_masks[key].to(torch.bfloat16)
        assert sel[m].dtype == torch.bfloat16 and torch.equal(sel[m], want)


@pytest.mark.parametrize("d,k,B,dtype", [(512, 80, 8, torch.float32), (768, 112, 6, torch.float32), (512, 64, 128, torch.float32),
                                         (512, 80, 8, torch.bfloat16)])
def test_intra_inter_selection_on_production_scores(d, k, B, dtype):
    """The two public selection methods (useA.py:50-96, 98-221) through the production score kernels."""
    ts, sim_p, toks = _sel_setup(d, k, None, B, 55 + d + k, dtype)
    tk = [t.cuda() for t in toks]
    patches, cls = [t[:, 1:] for t in tk], [t[:, 0] for t in tk]
    intra = ts.intra_modal_token_selection(*patches, *cls)
    inter = ts.inter_modal_token_selection(*patches, *cls)
    f32 = [t.float() for t in toks]
    rp, rc = [t[:, 1:] for t in f32], [t[:, 0] for t in f32]
    ref_intra = so.intra_modal_masks(rp, rc, k)
    ref_inter = so.inter_modal_masks_from_scores(so.inter_modal_scores(sim_p, rp, rc), 128, 2 * k)
    flips = 0
    for m in range(3):
        assert intra[m].shape == inter[m].shape == (B, 128, 1) and intra[m].dtype == dtype
        flips += int((intra[m][..., 0].float().cpu() != ref_intra[m].float()).sum())
        flips += int((inter[m][..., 0].float().cpu() != ref_inter[m].float()).sum())
        assert int(intra[m].sum(1).max()) == int(intra[m].sum(1).min()) == min(k, 128)
    if dtype == torch.float32:
        assert flips == 0, f"{flips} mask flips on production scores"
    else:
        # (bf16 tokens, fp32 scores from a split-bf16 fold of W_k^T W_q: the oracle sees the same rounded tokens)
        assert flips <= 2, flips


# ---------------------------------------------------------------------------------------------
# ModalInteractive stand-alone
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,B,dtype,tol", [(512, 8, torch.float32, FP32_TOL), (768, 5, torch.float32, FP32_TOL),
                                           (512, 8, torch.bfloat16, BF16_TOL), (768, 128, torch.bfloat16, BF16_TOL)])
def test_modal_interactive_standalone(d, B, dtype, tol):
    """ModalInteractive.forward on arbitrary (dense, unmasked) K/V maps: output and every gradient, full tensors."""
    M = _mods()
    sim_p = syn.make_params(syn.sim_param_shapes(d), 12 + d)
    mi = M.ModalInteractive(d)
    mi.load_state_dict({k[len("modal_interactive."):]: v for k, v in sim_p.items() if k.startswith("modal_interactive.")})
    mi = mi.cuda()
    toks = [t.to(dtype) for t in syn.make_tokens(B, d, seed=99 + B)]
    cot = syn.make_cotangent(B, d, seed=3)
    tk = [t.cuda().requires_grad_(True) for t in toks]
    out = mi(*[t[:, 1:] for t in tk], *[t[:, 0] for t in tk])
    assert out.shape == (B, 3 * d) and out.dtype == dtype
    names = [n for n, _ in mi.named_parameters()]
    grads = torch.autograd.grad(out, tk + [p for _, p in mi.named_parameters()], cot.to("cuda", dtype))

    def oracle(dt):
        prm = {k: v.clone().to(dt).requires_grad_(True) for k, v in sim_p.items()}
        tt = [t.to(dt).clone().requires_grad_(True) for t in toks]
        o = so.modal_interactive(prm, [t[:, 1:] for t in tt], [t[:, 0] for t in tt])
        gr = torch.autograd.grad(o, tt + [prm["modal_interactive." + n] for n in names], cot.to(dt))
        return o, gr

    o64, g64 = oracle(torch.float64)
    _close(out, o64, tol, "ModalInteractive out")
    for m in range(3):
        _close(grads[m], g64[m], tol, f"ModalInteractive dtokens[{m}]")
    for n, a, r in zip(names, grads[3:], g64[3:]):
        _close(a, r, tol, "ModalInteractive d" + n)


# ---------------------------------------------------------------------------------------------
# AlignmentM: stage == "CLS", Cls_Align / patch_Align, monkey-patching
# ---------------------------------------------------------------------------------------------
def _align_setup(c, dtype=torch.float32):
    import gpu_harness as gh
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    sim, al = gh.build_modules(c, sim_p, al_p)
    return sim, al, sim_p, al_p, [t.to(dtype) for t in toks], cot


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_alignment_stage_cls(dtype, tol, packed):
    """AlignmentM.forward(stage="CLS") (useB.py:181-183): returns the GAM scalar only; gradients reach the tokens and
    contra_temp; the DAS parameters are untouched (grad None, like the reference's unused parameters)."""
    c = dict(d=512, h=16, w=8, B=16, k=80, keep_ratio=None, gain=1.0, structured=True, seed=321)
    sim, al, sim_p, al_p, toks, _ = _align_setup(c, dtype)
    al.fuse_views = packed
    tk = [t.cuda().requires_grad_(True) for t in toks]
    gam = al(*[t[:, 1:] for t in tk], stage="CLS")
    assert isinstance(gam, torch.Tensor) and gam.dim() == 0 and gam.dtype == torch.float32
    gam.backward()
    prm = {k: v.clone().double().requires_grad_(True) for k, v in al_p.items()}
    tt = [t.double().clone().requires_grad_(True) for t in toks]
    ref = so.gam_loss([t[:, 1:] for t in tt], prm["contra_temp"])
    gref = torch.autograd.grad(ref, tt + [prm["contra_temp"]])
    assert abs(float(gam) - float(ref)) < tol * abs(float(ref))
    for m in range(3):
        _close(tk[m].grad, gref[m], tol, f"stage=CLS dtokens[{m}]")
        assert float(tk[m].grad[:, 0].abs().max()) == 0.0          # AlignM never sees the CLS row
    e = abs(float(al.contra_temp.grad) - float(gref[3])) / abs(float(gref[3]))
    # d(gam)/d(tau) on these correlated tokens is well conditioned (V spans [0.03, 0.3]); flat tolerance
    assert e < tol, ("contra_temp", e)
    for n, p in al.named_parameters():
        if n != "contra_temp":
            assert p.grad is None, f"{n} received a gradient under stage='CLS'"
    # the two public methods return the same scalars as forward
    with torch.no_grad():
        g2 = al.Cls_Align(*[t[:, 1:] for t in tk])
        g3, l3 = al(*[t[:, 1:] for t in tk], stage="together_CLS_Patch")
        l2 = al.patch_Align(*[t[:, 1:] for t in tk])
    assert float(g2) == float(gam) == float(g3) and float(l2) == float(l3)


def test_alignment_monkey_patched_patch_align_is_honoured():
    """zablation/offestvisual.py:209-214 replaces model.AlignM.patch_Align on the instance; forward must look it up
    through self at call time.  FusionHead falls back to the two module calls in that case."""
    M = _mods()
    c = gu.CASES["rgbnt201_d512"]
    sim, al, _, _, toks, _ = _align_setup(c)
    tk = [t.cuda() for t in toks]
    patches, cls = [t[:, 1:] for t in tk], [t[:, 0] for t in tk]
    with torch.no_grad():
        g0, l0 = al(*patches, stage="together_CLS_Patch")
    calls = []
    orig = al.patch_Align

    def patched(r, n, t):
        calls.append((r.shape, n.shape, t.shape))
        return orig(r, n, t) * 2.0

    al.patch_Align = patched
    with torch.no_grad():
        g1, l1 = al(*patches, stage="together_CLS_Patch")
        out, g2, l2 = M.FusionHead(sim, al)(*patches, *cls)
    assert len(calls) == 2 and calls[0][0] == (c["B"], 128, c["d"])
    assert float(g1) == float(g0) == float(g2) and float(l1) == float(l2) == 2.0 * float(l0)
    assert out.shape == (c["B"], 3 * c["d"])


def test_forward_hooks_on_submodules_run_them_one_by_one():
    """zablation/CAM.py:164 registers a forward hook on model.SIM.token_selection: the hook must fire with the three
    selected maps, and the result must not change."""
    c = gu.CASES["rgbnt201_d512"]
    sim, al, sim_p, _, toks, cot = _align_setup(c)
    tk = [t.cuda().requires_grad_(True) for t in toks]
    patches, cls = [t[:, 1:] for t in tk], [t[:, 0] for t in tk]
    out0 = sim(*patches, *cls)
    g0 = torch.autograd.grad(out0, tk + list(sim.modal_interactive.parameters()), cot.cuda())
    seen = []
    hnd = sim.token_selection.register_forward_hook(lambda mod, inp, outp: seen.append([o.detach() for o in outp]))
    out1 = sim(*patches, *cls)
    g1 = torch.autograd.grad(out1, tk + list(sim.modal_interactive.parameters()), cot.cuda())
    hnd.remove()
    assert len(seen) == 1 and len(seen[0]) == 3 and seen[0][0].shape == (c["B"], 128, c["d"])
    for m, key in enumerate(("RGB", "NI", "TI")):
        assert torch.equal(seen[0][m], patches[m].detach() * sim.token_selection.last_masks[key])
    _close(out1, out0, 1e-5, "hooked SIM out")
    for a, b in zip(g1, g0):
        _close(a, b, 1e-5, "hooked SIM grads")
    out2 = sim(*patches, *cls)     # hook removed: back on the fused path
    assert torch.equal(out2, out0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_no_grad_and_eval_mode(dtype):
    """Inference (make_model.py:273-290): eval() + torch.no_grad() -- same numbers as the training-mode forward, no
    autograd graph, no saved state."""
    M = _mods()
    c = gu.BF16_CASES["vehicle_d512"]
    sim, al, _, _, _, _ = _align_setup(c)
    toks = gu.bf16_case_tokens(c)
    tk = [t.to("cuda", dtype) for t in toks]
    patches, cls = [t[:, 1:] for t in tk], [t[:, 0] for t in tk]
    out_t = sim(*patches, *cls)
    gam_t, lam_t = al(*patches, stage="together_CLS_Patch")
    sim.eval()
    al.eval()
    with torch.no_grad():
        out_e = sim(*patches, *cls)
        gam_e, lam_e = al(*patches, stage="together_CLS_Patch")
        out_h, gam_h, lam_h = M.FusionHead(sim, al)(*patches, *cls)
    assert not out_e.requires_grad and out_e.grad_fn is None and gam_e.grad_fn is None
    assert torch.equal(out_e, out_t.detach()) and float(gam_e) == float(gam_t) and float(lam_e) == float(lam_t)
    _close(out_h, out_e, 1e-5 if dtype == torch.float32 else 1e-2, "FusionHead no_grad out")
    assert abs(float(gam_h) - float(gam_e)) < 1e-5 * abs(float(gam_e)) + 1e-7
    assert abs(float(lam_h) - float(lam_e)) < (1e-5 if dtype == torch.float32 else 1e-2) * abs(float(lam_e))


# ---------------------------------------------------------------------------------------------
# fp16 tokens (the reference's amp.autocast, engine/processor.py:165)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fused", [False, True])
def test_fp16_autocast_tokens(fused):
    """fp16 token maps under torch.autocast: accepted at the module boundary (one fp16 -> bf16 conversion pass per map),
    outputs and token gradients come back in fp16, everything within the bf16 tolerance of the fp32 oracle fed the
    same fp16-rounded values."""
    import gpu_harness as gh
    M = _mods()
    c = gu.BF16_CASES["rgbnt201_d512"]
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    toks16 = [t.to(torch.float16) for t in toks]
    sim, al = gh.build_modules(c, sim_p, al_p)
    tk = [t.cuda().requires_grad_(True) for t in toks16]
    with torch.autocast("cuda", dtype=torch.float16):
        patches, cls = [t[:, 1:] for t in tk], [t[:, 0] for t in tk]
        if fused:
            out, gam, lam = M.FusionHead(sim, al)(*patches, *cls)
        else:
            out = sim(*patches, *cls)
            gam, lam = al(*patches, stage="together_CLS_Patch")
    assert out.dtype == torch.float16 and gam.dtype == torch.float32 and lam.dtype == torch.float32
    torch.autograd.backward([out, gam, lam], [cot.to("cuda", torch.float16), torch.tensor(0.2, device="cuda"), torch.tensor(0.2, device="cuda")])
    prm_s = {k: v.clone().requires_grad_(True) for k, v in sim_p.items()}
    prm_a = {k: v.clone().requires_grad_(True) for k, v in al_p.items()}
    tt = [t.float().clone().requires_grad_(True) for t in toks16]
    o, g, l, masks = so.head_forward(prm_s, prm_a, tt, c["k"], c["h"], c["w"])
    torch.autograd.backward([o, g, l], [cot.half().float(), torch.tensor(0.2), torch.tensor(0.2)])
    dev = gu.load("bf16dev_rgbnt201_d512")
    got_masks = np.stack([sim.token_selection.last_masks[k][..., 0].cpu().numpy() for k in ("RGB", "NI", "TI")])
    ref_masks = np.stack([m[..., 0].numpy() for m in masks])
    assert int((got_masks != ref_masks).sum()) <= int(dev["mask_flips"])
    _close(out, o, BF16_TOL, "fp16 sim_out")
    assert abs(float(gam) - float(g)) < BF16_TOL * abs(float(g)) and abs(float(lam) - float(l)) < BF16_TOL * abs(float(l))
    for m in range(3):
        assert tk[m].grad.dtype == torch.float16
        # per-sample comparison; a LAM sample point on a pixel boundary may flip (gpu_harness.lam_flip_samples)
        num = (tk[m].grad.float().cpu() - tt[m].grad).reshape(c["B"], -1).norm(dim=1)
        den = tt[m].grad.reshape(c["B"], -1).norm(dim=1)
        assert int((num > BF16_TOL * den).sum()) <= 1, (m, (num / den).tolist())
    e2, _ = _err(sim.modal_interactive.ffn[0].weight.grad, prm_s["modal_interactive.ffn.0.weight"].grad)
    assert e2 < BF16_TOL, e2


def test_convert_half_roundtrip_and_strided_views():
    """sig_convert_half: bit-exact against torch's casts (RNE both ways, saturation at +-65504), strided views."""
    _mods()
    from signal_b200 import functional as F_
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(7, 129, 64, generator=g) * 100).half().cuda()
    x[0, 0, :4] = torch.tensor([65504.0, -65504.0, 6e-8, 0.0]).half()
    for view in (x, x[:, 1:], x[:, 0]):
        assert torch.equal(F_._convert_half(view, torch.bfloat16), view.to(torch.bfloat16))
    y = (torch.randn(7, 129, 64, generator=g) * 1e4).to(torch.bfloat16).cuda()
    y[1, 1, :3] = torch.tensor([1e6, -1e6, 3e-9]).to(torch.bfloat16)
    want = y.float().clamp(-65504, 65504).half()
    for yv, wv in ((y, want), (y[:, 1:], want[:, 1:]), (y[:, 0], want[:, 0])):
        assert torch.equal(F_._convert_half(yv, torch.float16), wv)


# ---------------------------------------------------------------------------------------------
# sizing vs dispatch: accepted shapes outside the tensor-core path must still run
# ---------------------------------------------------------------------------------------------
def test_alignment_bf16_wide_dim_and_mixed_strides_run():
    """bf16 with d in (768, 1024] runs on the exact SIMT path with a ctx the library sized for it; three patch maps
    with different strides are handed over as contiguous copies (the tensor-core path needs one geometry)."""
    M = _mods()
    for d, mixed in ((1024, False), (512, True)):
        al = M.AlignmentM(d, 16, 8).cuda()
        sim_like = syn.make_tokens(3, d, seed=5, dtype=torch.bfloat16)
        tk = [t.cuda().requires_grad_(True) for t in sim_like]
        patches = [t[:, 1:] for t in tk]
        if mixed:
            patches[1] = patches[1].contiguous()
        gam, lam = al(*patches, stage="together_CLS_Patch")
        (gam + lam).backward()
        ref = so.gam_loss([t[:, 1:].float().cpu() for t in tk], torch.tensor(0.07))
        assert abs(float(gam) - float(ref)) < BF16_TOL * abs(float(ref))
        assert all(t.grad is not None and bool(torch.isfinite(t.grad).all()) for t in tk)


def test_fusion_head_make_graphed_matches_eager():
    """FusionHead.make_graphed: CUDA-graph replay of forward and backward through the ordinary autograd API -- same
    outputs and gradients as the eager call, on fresh inputs, repeatedly."""
    import gpu_harness as gh
    M = _mods()
    c = gu.BF16_CASES["rgbnt201_d512"]
    sim_p, al_p, _, cot = gu.case_inputs(c)
    sim, al = gh.build_modules(c, sim_p, al_p)
    head = M.FusionHead(sim, al)
    sample = [t.cuda().requires_grad_(True) for t in gu.bf16_case_tokens(c)]
    graphed = head.make_graphed(*sample)
    params = list(sim.parameters()) + list(al.parameters())
    w = [cot.to("cuda", torch.bfloat16), torch.tensor(0.2, device="cuda"), torch.tensor(0.2, device="cuda")]
    for seed in (1, 2):
        toks = [t.to(torch.bfloat16) for t in syn.make_tokens(c["B"], c["d"], seed=900 + seed)]
        res = {}
        for name, fn in (("graphed", lambda a, b, c_: graphed(a, b, c_)),
                         ("eager", lambda a, b, c_: head(a[:, 1:], b[:, 1:], c_[:, 1:], a[:, 0], b[:, 0], c_[:, 0]))):
            tk = [t.cuda().requires_grad_(True) for t in toks]
            for p in params:
                p.grad = None
            out, gam, lam = fn(*tk)
            torch.autograd.backward([out, gam, lam], w)
            torch.cuda.synchronize()
            res[name] = ([out.detach().clone(), gam.detach().clone(), lam.detach().clone()] + [t.grad.clone() for t in tk]
                         + [p.grad.clone() for p in params if p.grad is not None],
                         {k: v.clone() for k, v in sim.token_selection.last_masks.items()})
        assert len(res["graphed"][0]) == len(res["eager"][0])
        for a, b in zip(res["graphed"][0], res["eager"][0]):
            _close(a.float(), b.float(), 2e-3, "graphed vs eager")       # (split-K atomics: not bit-reproducible)
        for k in ("RGB", "NI", "TI"):
            assert torch.equal(res["graphed"][1][k], res["eager"][1][k])


# ---- volume_computation4 / volume_computation5 (utils/volume.py:65-182) -----------------------------------------------------
@pytest.mark.parametrize("name", ["vol4_indep", "vol4_aligned", "vol5_indep", "vol5_aligned"])
def test_volume_computation_4_5_match_reference_golden(name):
    import __graft_entry__ as entry
    entry.build()
    import volume_cases as vc
    from signal_b200 import modules as M
    c = vc.CASES[name]
    z = vc.load(name)
    feats, cot = vc.gen.inputs(c)
    xs = [t.cuda().requires_grad_(True) for t in feats]
    fn = M.volume_computation4 if c["n"] == 4 else M.volume_computation5
    V = fn(*xs)
    assert V.shape == (c["B1"], c["B2"]) and V.dtype == torch.float32
    V.backward(cot.cuda())

    def rel(a, b):
        a, b = a.detach().double().cpu(), torch.from_numpy(b).double()
        return float((a - b).norm() / b.norm().clamp_min(1e-30))

    tol = lambda key: max(1e-4, 3.0 * float(z["dev32/" + key]))     # never below what the reference itself loses in fp32
    assert rel(V, z["ref/vol"]) <= tol("vol"), rel(V, z["ref/vol"])
    for k, x in enumerate(xs):
        assert rel(x.grad, z[f"ref/d{k}"]) <= tol(f"d{k}"), (name, k, rel(x.grad, z[f"ref/d{k}"]))


def test_volume_n_entry_with_three_modalities_equals_volume3():
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import functional as F_, modules as M
    g = torch.Generator().manual_seed(9)
    l, v, a = [torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=-1).cuda().requires_grad_(True) for n in (20, 33, 33)]
    cot = torch.randn(20, 33, generator=g).cuda()
    V3 = M.volume_computation3(l, v, a)
    V3.backward(cot)
    g3 = [t.grad.clone() for t in (l, v, a)]
    for t in (l, v, a):
        t.grad = None
    Vn = F_.VolumeNFunction.apply(l, v, a)
    Vn.backward(cot)
    assert float((Vn - V3).abs().max()) < 1e-5
    for x, y in zip((l, v, a), g3):
        assert float((x.grad - y).norm() / y.norm()) < 1e-4
