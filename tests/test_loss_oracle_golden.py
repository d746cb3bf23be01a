"""CPU: the loss oracle (oracle/loss_oracle.py) against golden vectors produced by the live reference classes
(tests/golden/make_loss_golden.py: layers/softmax_loss.py CrossEntropyLabelSmooth, layers/triplet_loss.py TripletLoss)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import loss_oracle as lo  # noqa: E402
from make_loss_golden import CASES  # noqa: E402

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "losses_*.npz")))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def test_golden_files_exist():
    assert len(FILES) == len(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("dtype,tag,tol", [(torch.float64, "", 1e-9), (torch.float32, "32", 1e-4)])
def test_loss_oracle_matches_reference(name, dtype, tag, tol):
    c = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"losses_{name}.npz"))
    labels = torch.from_numpy(g["labels"])
    z = torch.from_numpy(g["logits"]).to(dtype).requires_grad_(True)
    xent = lo.xent_label_smooth(z, labels, c["C"], 0.1)
    xent.backward()
    # (the reference builds its smoothed targets in fp32 even for fp64 logits -- torch.zeros default dtype,
    #  softmax_loss.py:30-32 -- so its fp64 run carries 6e-8 of target rounding)
    assert rel(xent.item(), g["xent" + tag]) < max(tol, 1e-6)
    assert rel(z.grad.numpy(), g["dlogits" + tag]) < max(tol, 1e-6)
    x = torch.from_numpy(g["feat"]).to(dtype).requires_grad_(True)
    tl, ap, an = lo.triplet_loss(x, labels, c["margin"], c["hf"])
    tl.backward()
    assert float(g["tri"]) > 0                       # the case exercises the loss (violated triplets exist)
    assert rel(tl.item(), g["tri" + tag]) < tol
    assert rel(ap.detach().numpy(), g["dist_ap" + tag]) < tol
    assert rel(an.detach().numpy(), g["dist_an" + tag]) < tol
    assert rel(x.grad.numpy(), g["dfeat" + tag]) < max(tol, 1e-8) * (10 if tag else 1)


def test_hard_mining_ties_and_self_positive():
    # two identical positives: the lowest index wins; the anchor itself is a positive (distance sqrt(1e-12))
    x = torch.tensor([[0.0, 0.0], [1.0, 0.0], [1.0, 0.0], [5.0, 0.0], [5.0, 1.0]], dtype=torch.float64)
    y = torch.tensor([0, 0, 0, 1, 1])
    d = lo.pairwise_euclid(x)
    ap, an, pi, ni = lo.hard_mining(d, y)
    assert pi.tolist()[0] == 1 and ni.tolist()[0] == 3
    assert abs(float(d[0, 0]) - 1e-6) < 1e-12
    assert pi.tolist()[1] == 0 and pi.tolist()[3] == 4


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_bnneck_oracle_matches_torch_modules(name, mode):
    """BNNeck + classifier restatement vs nn.BatchNorm1d -> nn.Linear(bias=False) (what make_model.py:128-131 builds)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", f"losses_{name}.npz"))
    t = lambda k: torch.from_numpy(g[k])
    x = t("feat").clone().requires_grad_(True)
    gamma, beta, W = (t(k).clone().requires_grad_(True) for k in ("bn_w", "bn_b", "cls_w"))
    fb, sc, rm, rv = lo.bnneck_classifier(x, gamma, beta, W, 1e-5, (t("bn_rm0"), t("bn_rv0")), 0.1, mode == "train")
    ((sc * t("cot_s")).sum() + (fb * t("cot_f")).sum()).backward()
    for got, key in ((fb, "feat"), (sc, "score"), (x.grad, "dx"), (gamma.grad, "dbn_w"), (beta.grad, "dbn_b"), (W.grad, "dcls_w"),
                     (rm, "rm"), (rv, "rv")):
        assert rel(got.detach().numpy(), g[f"nk_{mode}_{key}"]) < 1e-9, key
