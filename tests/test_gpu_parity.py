"""GPU parity: CUDA drop-in modules (through the C ABI) vs the golden vectors from the live
reference and vs the CPU oracle on the same seeded inputs."""
import numpy as np
import os

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


# How the exceptions to the flat tolerances are bounded (no hand-set numbers):
#  * fp32: every golden file stores ``dev32/<quantity>``, the deviation of the reference's own fp32 run from its fp64 run
#    (make_golden.py); a quantity's bound is max(1e-4, FP32_DEV_C x that) -- gpu_harness.compare_records.  Live-oracle
#    comparisons (B=128) derive the same record on the spot from an fp32 and an fp64 run of the oracle.
#  * bf16: ``bf16dev_<case>.npz`` stores what the reference's own bf16-autocast run loses against its fp64 run on the same
#    bf16-rounded tokens: per-quantity deviations, selection flips, LAM outlier samples.  The CUDA path must stay within
#    max(2e-2, BF16_DEV_C x the deviation of the reference's own bf16 run) per quantity -- both are samples of the same
#    rounding noise, so a factor 2 either way is chance -- while the MEDIAN of err / reference-deviation over all those
#    quantities must be <= 1: on the whole at least as close to the fp64 truth as the reference's reduced-precision run
#    (gpu_harness.compare_records).  It must flip no more mask entries than the reference does, and show at most BF16_FLIP_FRAC of the reference's count of LAM outlier samples (the reference's bf16 run puts
#    EVERY sample over 2e-2; see the generator's log lines in tests/golden/make_golden.py).
FP32_DEV_C = 3.0
BF16_DEV_C = 2.0
BF16_FLIP_FRAC = 0.25
BF16_CASES = gu.BF16_CASES
SMALL_BF16 = [n for n, c in BF16_CASES.items() if c["B"] <= 8]
FULL_BF16 = [n for n, c in BF16_CASES.items() if c["B"] == 128]


def _report(label, rep):
    """one line per comparison that used more than half of its bound (visible with pytest -s / in the log on failure)"""
    tight = sorted(((e / b, k, e, b) for k, (e, b) in rep.items() if b > 0 and e / b > 0.5), reverse=True)[:6]
    if tight:
        print(f"[{label}] closest to their bounds:", [(k, "%.2e/%.2e" % (e, b)) for _, k, e, b in tight])


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("name", list(gu.CASES))
def test_fp32_matches_reference_golden(name, packed):
    h = _harness()
    c = gu.CASES[name]
    got = h.cuda_record(c, torch.float32, packed=packed)
    ref = gu.load(name)
    rep = {}
    h.compare_records(got, ref, FP32_TOL, ref_masks_key="masks32", label=name, dev=ref, dev_prefix="dev32/", dev_c=FP32_DEV_C, report=rep)
    _report(name, rep)


def _oracle_record(c, sim_p, al_p, toks, cot, dtype=torch.float32):
    return _harness().oracle_record(c, sim_p, al_p, toks, cot, dtype)


def _live_dev32(lo, hi):
    """dev32 record derived on the spot: the oracle's fp32 run against its fp64 run, on the full tensors"""
    dev = {}
    for k, v in hi.items():
        if k in ("sim_out",):
            dev["dev32/" + k] = gu.rel_err(lo[k], v)
        elif k in ("gam", "lam"):
            dev["dev32/" + k] = abs(lo[k] - v) / abs(v)
        elif k.startswith("dtok_full_"):
            a, b = torch.stack(lo[k]).double(), torch.stack(v).double()
            dev["dev32/" + k.replace("_full", "")] = float((a - b).norm() / b.norm().clamp_min(1e-300))
        elif k.startswith("dpar_full_") and float(v.abs().max()) > 0:
            dev["dev32/" + k.replace("_full", "")] = float((lo[k].double() - v.double()).norm() / v.double().norm())
    return dev


def _bf16_inputs(c):
    sim_p, al_p, _, cot = gu.case_inputs(c)
    return sim_p, al_p, gu.bf16_case_tokens(c), cot


def _bf16_budget(name):
    dev = gu.load("bf16dev_" + name)
    return dev, int(dev["mask_flips"]), int(BF16_FLIP_FRAC * int(dev["lam_flip_samples"]))


@pytest.mark.parametrize("name", SMALL_BF16 + FULL_BF16)
def test_bf16_matches_oracle_on_rounded_inputs(name):
    """bf16 tokens on the GPU vs the fp32 oracle fed the same bf16-rounded values (full gradient tensors); includes
    BASELINE.json configs #2 and #3 at B=128 (grid 16x8 TOPK 80 at d=768 and d=512; grid 8x16 TOPK 112 with keep_ratio)."""
    h = _harness()
    c = BF16_CASES[name]
    sim_p, al_p, toks, cot = _bf16_inputs(c)
    dev, flip_budget, lam_budget = _bf16_budget(name)
    got = h.cuda_record(c, torch.bfloat16, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    ref = _oracle_record(c, sim_p, al_p, [t.float() for t in toks], cot)
    flips = int((got["masks"] != ref["masks"]).sum())
    assert flips <= flip_budget, f"{flips} mask flips vs the oracle on identical (bf16-rounded) inputs; the reference's own bf16 run flips {flip_budget}"
    rep = {}
    h.compare_records(got, ref, BF16_TOL, check_masks=False, label=name, dev=dev, dev_prefix="dev/", dev_c=BF16_DEV_C,
                      lam_flip_budget=lam_budget, report=rep)
    _report(name, rep)


@pytest.mark.parametrize("d,hw,k,keep", [(512, (16, 8), 80, None), (768, (16, 8), 80, None), (512, (8, 16), 112, 0.75)])
def test_full_batch_b128_matches_oracle(d, hw, k, keep):
    """BASELINE.json config #2 / #3 sizes (B=128) in fp32 against the oracle, full gradient tensors; the per-quantity
    bound is derived from the oracle's own fp32-vs-fp64 deviation on these inputs."""
    h = _harness()
    c = dict(d=d, h=hw[0], w=hw[1], B=128, k=k, keep_ratio=keep, gain=30.0, structured=False, seed=900 + d + hw[0])
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    got = h.cuda_record(c, torch.float32, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    ref32 = _oracle_record(c, sim_p, al_p, toks, cot)
    ref = _oracle_record(c, sim_p, al_p, toks, cot, torch.float64)
    flips = int((got["masks"] != ref["masks"]).sum())
    assert flips == 0, f"{flips} mask flips at B=128"
    rep = {}
    h.compare_records(got, ref, FP32_TOL, check_masks=False, label=f"B128 d{d}", dev=_live_dev32(ref32, ref), dev_prefix="dev32/",
                      dev_c=FP32_DEV_C, report=rep)
    _report(f"B128 d{d} {hw}", rep)


@pytest.mark.parametrize("B", [256, 512])
def test_large_batch_matches_its_shards_and_the_oracle(B):
    """BASELINE.json config #5 read as strong scaling: 512 / 256 samples per GPU at 2 / 4 GPUs.  SIM and LAM are per-sample
    independent, so a B-sample call must reproduce the 128-sample calls on its shards (SIM rows and masks; LAM loss = mean of
    the shard losses; SIM's token gradient rows); GAM couples the batch through the B x B grid and is checked against the
    oracle's Cls_Align on the full batch."""
    _harness()
    import gpu_harness
    from oracle import signal_oracle as so
    from signal_b200 import modules as M
    c = dict(d=512, h=16, w=8, B=B, k=80, keep_ratio=None, gain=1.0, structured=False, seed=5000 + B)
    sim_p, al_p, toks, cot = gu.case_inputs(c)
    sim, al = gpu_harness.build_modules(c, sim_p, al_p)
    head = M.FusionHead(sim, al)
    tk = [t.to("cuda", torch.bfloat16).requires_grad_(True) for t in toks]
    cotd = cot.to("cuda", torch.bfloat16)

    def run(rows):
        xs = [t.detach()[rows].clone().requires_grad_(True) for t in tk]
        out, gam, lam = head(*[x[:, 1:] for x in xs], *[x[:, 0] for x in xs])
        masks = torch.stack([sim.token_selection.last_masks[k] for k in ("RGB", "NI", "TI")]).clone()
        out.backward(cotd[rows])                      # SIM's part of the token gradient only (per-sample independent)
        return out.detach().float(), float(gam), float(lam), masks, [x.grad.float() for x in xs]

    out, gam, lam, masks, dt = run(slice(0, B))
    lams = []
    for s0 in range(0, B, 128):
        o, _, l, mk, g = run(slice(s0, s0 + 128))
        lams.append(l)
        assert torch.equal(mk, masks[:, s0:s0 + 128]), f"masks of shard {s0 // 128} differ from the B={B} call"
        assert float((o - out[s0:s0 + 128]).norm() / o.norm()) < 2e-3
        for m in range(3):
            assert float((g[m] - dt[m][s0:s0 + 128]).norm() / g[m].norm()) < 5e-3
    assert abs(lam - sum(lams) / len(lams)) < 2e-3 * abs(lam), (lam, lams)
    ref_gam = float(so.gam_loss([t.detach().float().cpu()[:, 1:] for t in tk], al_p["contra_temp"].float()))
    assert abs(gam - ref_gam) < BF16_TOL * abs(ref_gam), (gam, ref_gam)


@pytest.mark.parametrize("name", ["rgbnt201_d512", "rgbnt201_d768", "vehicle_d512"])
def test_fused_pool_in_score_pass_matches_own_pool(name):
    """SIG_FUSE_POOL (N2 forward): GAM's mean pool delivered by SIM's score pass == AlignM pooling the tokens itself -- GAM
    loss, every gradient and the masks (the score part of the kernel must be unchanged: bit-equal masks and SIM output)."""
    _harness()
    import gpu_harness
    from signal_b200 import lib, modules as M
    c = BF16_CASES[name]
    sim_p, al_p, toks, cot = _bf16_inputs(c)
    res = {}
    for mode in ("0", "1"):
        os.environ["SIG_FUSE_POOL"] = mode
        try:
            sim, al = gpu_harness.build_modules(c, sim_p, al_p)
            tk = [t.to("cuda", torch.bfloat16).requires_grad_(True) for t in toks]
            n0 = lib.launch_count()
            out, gam, lam = M.FusionHead(sim, al)(*[t[:, 1:] for t in tk], *[t[:, 0] for t in tk])
            launches = lib.launch_count() - n0
            torch.autograd.backward([out, gam, lam], [cot.to("cuda", torch.bfloat16), torch.tensor(0.2, device="cuda"), torch.tensor(0.2, device="cuda")])
            masks = torch.stack([sim.token_selection.last_masks[k] for k in ("RGB", "NI", "TI")]).clone()
            res[mode] = (launches, out.detach().float(), float(gam), float(lam), masks, [t.grad.float() for t in tk],
                         float(al.contra_temp.grad))
        finally:
            os.environ.pop("SIG_FUSE_POOL", None)
    a, b = res["0"], res["1"]
    assert b[0] == a[0] - 1, (a[0], b[0])                       # AlignM's pool launch is gone
    assert torch.equal(a[4], b[4]) and torch.equal(a[1], b[1])  # selection and SIM output untouched
    assert abs(a[2] - b[2]) <= 2e-5 * abs(a[2]) and abs(a[3] - b[3]) <= 1e-6 * abs(a[3])
    assert abs(a[6] - b[6]) <= 1e-3 * max(abs(a[6]), 1e-3)
    for x, y in zip(a[5], b[5]):
        assert float((x - y).norm() / x.norm()) < 1e-3


@pytest.mark.parametrize("name", SMALL_BF16)
def test_bf16_tensor_core_path_vs_simt_path(name):
    """Same bf16 inputs through the tcgen05 path and through the fp32 SIMT kernels (FORCE_SIMT)."""
    from signal_b200 import lib
    h = _harness()
    c = BF16_CASES[name]
    sim_p, al_p, toks, cot = _bf16_inputs(c)
    dev, _, lam_budget = _bf16_budget(name)
    fast = h.cuda_record(c, torch.bfloat16, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    slow = h.cuda_record(c, torch.bfloat16, flags=lib.FLAG_FORCE_SIMT, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot)
    h.compare_records(fast, slow, BF16_TOL, check_masks=True, label=name + " tc-vs-simt", dev=dev, dev_prefix="dev/", dev_c=BF16_DEV_C,
                      lam_flip_budget=lam_budget)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("name", ["rgbnt201_d512", "vehicle_d512"])
def test_fusion_head_matches_the_two_modules(name, dtype, tol):
    """FusionHead (AlignM on a side stream, one token-gradient writer) vs calling SIM and AlignM separately."""
    h = _harness()
    c = BF16_CASES[name]
    sim_p, al_p, toks, cot = _bf16_inputs(c)
    toks = [t.to(dtype) for t in toks]
    a = h.cuda_record(c, dtype, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot, fused=True)
    b = h.cuda_record(c, dtype, sim_p=sim_p, al_p=al_p, toks=toks, cot=cot, fused=False)
    h.compare_records(a, b, tol, check_masks=True, label=name + " fused-vs-separate")
    # all three objectives at once: the shared gradient map must hold the SUM of the three contributions
    import gpu_harness
    from signal_b200 import modules as M
    sim, al = gpu_harness.build_modules(c, sim_p, al_p)
    tk = [t.to("cuda", dtype).requires_grad_(True) for t in toks]
    out, gam, lam = M.FusionHead(sim, al)(*[t[:, 1:] for t in tk], *[t[:, 0] for t in tk])
    torch.autograd.backward([out, gam, lam], [cot.to("cuda", dtype), torch.tensor(0.2, device="cuda"), torch.tensor(0.2, device="cuda")])
    for m in range(3):
        want = a["dtok_full_sim"][m] + 0.2 * a["dtok_full_gam"][m] + 0.2 * a["dtok_full_lam"][m]
        got = tk[m].grad.float().cpu()
        assert float((got - want).norm() / want.norm()) < (2e-2 if dtype == torch.bfloat16 else 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fusion_head_grad_sync_pieces(dtype):
    """FusionHead.grad_sync (the data-parallel exchange hook) is called with contiguous pieces of the two
    gradient arenas on a communication stream, each after the event the library records when that piece is
    final (sig_sim_param_grads.early_event, sig_align_param_grads.done_event).  A hook that doubles its piece in
    place must therefore yield exactly 2x the gradients of a run without it, and the pieces must tile both
    arenas once -- a piece handed over before its producer finished would come back undoubled."""
    _harness()
    import gpu_harness
    from signal_b200 import modules as M
    from signal_b200 import parallel
    c = BF16_CASES["rgbnt201_d512"]
    sim_p, al_p, toks, cot = _bf16_inputs(c)

    def run(hook):
        sim, al = gpu_harness.build_modules(c, sim_p, al_p)
        head = M.FusionHead(sim, al, grad_sync=hook)
        tk = [t.to("cuda", dtype).requires_grad_(True) for t in toks]
        for _ in range(3):      # repeated steps re-use the same events and streams
            for p in list(sim.parameters()) + list(al.parameters()):
                p.grad = None
            out, gam, lam = head(*[t[:, 1:] for t in tk], *[t[:, 0] for t in tk])
            torch.autograd.backward([out, gam, lam], [cot.to("cuda", dtype), torch.tensor(0.2, device="cuda"), torch.tensor(0.2, device="cuda")])
        torch.cuda.synchronize()
        params = list(sim.named_parameters()) + list(al.named_parameters())
        return {n: p.grad.clone() for n, p in params if p.grad is not None}, parallel.grad_arenas(p for _, p in params)

    pieces = []

    def hook(flat):
        pieces.append((flat.data_ptr(), flat.numel()))
        flat.mul_(2.0)

    # (with or without a hook the step takes the same code path: AlignM's eager backward chain is on in both)
    base, _ = run(None)
    got, arenas = run(hook)
    assert len(pieces) == 6 and len(arenas) == 1          # two pieces per step, one arena for both modules
    total = sum(n for _, n in pieces[-2:])
    assert total == sum(a.numel() for a in arenas)
    assert base.keys() == got.keys()
    # not bit-equal: the split-K weight-gradient GEMM adds its partial tiles with fp32 atomics, and on the bf16 path
    # the un-fold GEMMs read a bf16 rounding of that sum; an undoubled piece would be off by 0.5
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    errs = {n: float((got[n] - 2.0 * base[n]).norm()) / (float(base[n].norm()) + 1e-30) for n in base}
    bad = {n: e for n, e in errs.items() if not e <= tol}
    assert not bad, bad
