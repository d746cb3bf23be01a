"""GPU: the inference tail (feature concat + L2 norm, euclidean distance matrix, market1501 CMC / mAP) through the C ABI
against the oracle and the golden vectors of the live reference's utils/metrics.py."""
import numpy as np
import pytest
import torch

import golden_util as gu
import metrics_cases as mc
from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu


def _ev():
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import evaluation
    return evaluation


@pytest.mark.parametrize("name", list(mc.CASES))
def test_distmat_and_rank_eval_match_reference_golden(name):
    ev = _ev()
    c = mc.CASES[name]
    rec = gu.load("metrics_" + name)
    feats, pids, camids = mc.make_case(c)
    nq = c["nq"]
    f = torch.from_numpy(feats).cuda()
    if c["norm"]:
        f = torch.nn.functional.normalize(f, dim=1, p=2)
    dist = ev.euclidean_distance(f[:nq], f[nq:])
    assert dist.is_cuda and dist.dtype == torch.float32
    assert float((dist.cpu() - torch.from_numpy(rec["distmat"])).abs().max()) <= 1e-5 * float(np.abs(rec["distmat"]).max())
    # identical fp32 distances in -> CMC bit-exact, mAP to fp64 rounding (the ranking is integer work)
    cmc, mAP = ev.eval_func(torch.from_numpy(rec["distmat"]).cuda(), pids[:nq], pids[nq:], camids[:nq], camids[nq:], max_rank=c["max_rank"])
    assert np.array_equal(cmc, rec["cmc"]) and abs(mAP - float(rec["mAP"])) < 1e-12
    # and end to end on the device-computed distances against the oracle on the same matrix
    cmc2, mAP2 = ev.eval_func(dist, pids[:nq], pids[nq:], camids[:nq], camids[nq:], max_rank=c["max_rank"])
    ocmc, omAP = mo.eval_func(dist.cpu().numpy(), pids[:nq], pids[nq:], camids[:nq], camids[nq:], max_rank=c["max_rank"])
    assert np.array_equal(cmc2, ocmc) and abs(mAP2 - omAP) < 1e-12


def test_rank_eval_ties_and_large_gallery():
    """quantised distances (ties everywhere) and a gallery of RGBNT100 size; stable order = lowest index first"""
    ev = _ev()
    g = np.random.default_rng(5)
    nq, ng, ids = 64, 8575, 50
    dist = (g.integers(0, 40, size=(nq, ng)) / 40.0).astype(np.float32)
    q_pids, g_pids = g.integers(0, ids, nq), g.integers(0, ids, ng)
    q_cams, g_cams = g.integers(0, 8, nq), g.integers(0, 8, ng)
    cmc, mAP = ev.eval_func(torch.from_numpy(dist).cuda(), q_pids, g_pids, q_cams, g_cams, max_rank=50)
    ocmc, omAP = mo.eval_func(dist, q_pids, g_pids, q_cams, g_cams, max_rank=50)
    assert np.array_equal(cmc, ocmc) and abs(mAP - omAP) < 1e-12


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("normalize", [False, True])
def test_inference_features(dtype, normalize):
    ev = _ev()
    g = torch.Generator().manual_seed(3)
    B, d = 37, 512
    toks = [torch.randn(B, 129, d, generator=g).to(dtype).cuda() for _ in range(3)]
    vars_total = torch.randn(B, 3 * d, generator=g).to(dtype).cuda()
    f = ev.inference_features(toks[0][:, 0], toks[1][:, 0], toks[2][:, 0], vars_total, normalize=normalize)     # strided CLS views
    want = mo.inference_features([t[:, 0].float().cpu().numpy() for t in toks], vars_total.float().cpu().numpy(), normalize)
    assert f.shape == (B, 6 * d) and f.dtype == torch.float32
    assert np.abs(f.cpu().numpy() - want).max() <= 2e-6 * np.abs(want).max()


def test_r1_map_eval_class_mirrors_the_reference_flow():
    ev = _ev()
    c = mc.CASES["rgbnt201_like"]
    rec = gu.load("metrics_rgbnt201_like")
    feats, pids, camids = mc.make_case(c)
    e = ev.R1_mAP_eval(c["nq"], max_rank=c["max_rank"], feat_norm=True)
    for lo in range(0, feats.shape[0], 64):       # batches, as engine/processor.py feeds them
        e.update((torch.from_numpy(feats[lo:lo + 64]).cuda(), pids[lo:lo + 64], camids[lo:lo + 64], ["x"] * len(pids[lo:lo + 64])))
    cmc, mAP, distmat, _, _, qf, gf = e.compute()
    assert np.abs(cmc - rec["cmc"]).max() <= 1.0 / c["nq"] + 1e-6 and abs(mAP - float(rec["mAP"])) < 5e-3
    assert distmat.shape == (c["nq"], c["ng"]) and qf.shape[0] == c["nq"]


def test_alignm_skip_in_eval_launches_nothing():
    """make_model.py:277-281 computes AlignM at inference and discards it: the opt-in switch returns zeros without a launch"""
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import lib, modules as M
    al = M.AlignmentM(512, 16, 8).cuda()
    toks = [torch.randn(4, 129, 512, device="cuda").to(torch.bfloat16) for _ in range(3)]
    patches = [t[:, 1:] for t in toks]
    al.eval()
    ref = al(*patches, stage="together_CLS_Patch")
    assert float(ref[0]) != 0.0 and float(ref[1]) != 0.0        # default: computed like the reference does
    al.skip_in_eval = True
    n0 = lib.launch_count()
    gam, lam = al(*patches, stage="together_CLS_Patch")
    cls_only = al(*patches, stage="CLS")
    assert lib.launch_count() == n0
    assert float(gam) == 0.0 and float(lam) == 0.0 and float(cls_only) == 0.0 and gam.dtype == torch.float32 and gam.dim() == 0
    al.train()
    assert float(al(*patches, stage="together_CLS_Patch")[0]) != 0.0   # training mode is never skipped
