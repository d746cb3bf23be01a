"""GPU parity of the ID / metric loss drop-ins (signal_b200.losses, csrc/losses.cu) through the C ABI:
against the golden vectors of the live reference (fp32 1e-4), against the CPU oracle at training sizes
(B = 128, D = 3 * 768, C = 171), in bf16 (2e-2), and the API surface of the reference classes."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

pytestmark = pytest.mark.gpu


def _mods():
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import losses
    return losses


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).norm() / (b.norm() + 1e-300))


@pytest.mark.parametrize("name", ["rgbnt201", "msvr310_margin", "hardfactor"])
def test_losses_match_reference_golden_fp32(name):
    losses = _mods()
    from make_loss_golden import CASES
    c = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"losses_{name}.npz"))
    labels = torch.from_numpy(g["labels"]).cuda()
    z = torch.from_numpy(g["logits"]).float().cuda().requires_grad_(True)
    xent = losses.CrossEntropyLabelSmooth(c["C"], epsilon=0.1)(z, labels)
    xent.backward()
    assert xent.dtype == torch.float32 and xent.dim() == 0
    assert rel(xent, g["xent"]) < 1e-5 and rel(z.grad, g["dlogits"]) < 1e-4
    x = torch.from_numpy(g["feat"]).float().cuda().requires_grad_(True)
    tri = losses.TripletLoss(margin=c["margin"], hard_factor=c["hf"])
    tl, ap, an = tri(x, labels)
    tl.backward()
    assert rel(ap, g["dist_ap"]) < 1e-4 and rel(an, g["dist_an"]) < 1e-4
    assert rel(tl, g["tri"]) < 1e-4 and rel(x.grad, g["dfeat"]) < 1e-4
    assert tri.last_indices.shape == (2, c["B"])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("margin", [None, 0.3])
def test_losses_training_size_vs_oracle(dtype, tol, margin):
    losses = _mods()
    from oracle import loss_oracle as lo
    B, K, D, C = 128, 8, 3 * 768, 171
    g = torch.Generator().manual_seed(5)
    labels = torch.randperm(C, generator=g)[: B // K].repeat_interleave(K)[torch.randperm(B, generator=g)]
    centers = 0.02 * torch.randn(C, D, generator=g)
    feat = (centers[labels] + 0.05 * torch.randn(B, D, generator=g)).to(dtype)
    logits = (3.0 * torch.randn(B, C, generator=g)).to(dtype)
    cot_ap = torch.randn(B, generator=g) * 1e-3
    # oracle on the same (rounded) values, fp32 -> fp64 arithmetic
    zo = logits.double().requires_grad_(True)
    xo = feat.double().requires_grad_(True)
    lx = lo.xent_label_smooth(zo, labels, C, 0.1)
    lt, apo, ano = lo.triplet_loss(xo, labels, margin, 0.0)
    (0.25 * lx + lt + (apo * cot_ap.double()).sum()).backward()
    z = logits.cuda().requires_grad_(True)
    x = feat.cuda().requires_grad_(True)
    lxg = losses.CrossEntropyLabelSmooth(C)(z, labels.cuda())
    ltg, ap, an = losses.TripletLoss(margin=margin)(x, labels.cuda())
    (0.25 * lxg + ltg + (ap * cot_ap.cuda()).sum()).backward()
    assert float(lt.detach()) > 0.05
    assert rel(lxg, lx.detach()) < tol and rel(ltg, lt.detach()) < tol
    assert rel(ap, apo.detach()) < tol and rel(an, ano.detach()) < tol
    assert z.grad.dtype == dtype and x.grad.dtype == dtype
    assert rel(z.grad, zo.grad) < tol and rel(x.grad, xo.grad) < tol


def test_losses_no_cpu_path_and_strided_rows():
    losses = _mods()
    with pytest.raises(RuntimeError):
        losses.CrossEntropyLabelSmooth(10)(torch.randn(4, 10), torch.zeros(4, dtype=torch.long))
    # rows of a wider buffer (leading dimension > C): consumed without a copy
    buf = torch.randn(8, 64, device="cuda")
    z = buf[:, :40].detach().requires_grad_(True)
    y = torch.arange(8, device="cuda") % 40
    a = losses.CrossEntropyLabelSmooth(40)(z, y)
    b = losses.CrossEntropyLabelSmooth(40)(z.detach().contiguous(), y)
    assert torch.equal(a, b)
    # LabelSmoothingCrossEntropy (softmax_loss.py:36-55) is the same smoothed target with C from the logits
    c = losses.LabelSmoothingCrossEntropy(0.1)(z, y)
    ref = torch.nn.functional.cross_entropy(z.detach(), y, label_smoothing=0.1)
    assert float((c.detach() - ref).abs()) < 1e-5 * float(ref.abs()) and torch.equal(a, c)


@pytest.mark.parametrize("name", ["rgbnt201", "hardfactor"])
@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_bnneck_classifier_matches_torch_modules(name, mode, dtype, tol):
    """BNNeckClassifier(bottleneck, classifier) vs the fp64 golden run of nn.BatchNorm1d -> nn.Linear(bias=False):
    both outputs, every gradient, and the in-place running-statistics update."""
    losses = _mods()
    from make_loss_golden import CASES
    c = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"losses_{name}.npz"))
    t = lambda k: torch.from_numpy(g[k]).float().cuda()
    bn = torch.nn.BatchNorm1d(c["D"]).cuda()
    cls = torch.nn.Linear(c["D"], c["C"], bias=False).cuda()
    with torch.no_grad():
        bn.weight.copy_(t("bn_w")); bn.bias.copy_(t("bn_b")); cls.weight.copy_(t("cls_w"))
        bn.running_mean.copy_(t("bn_rm0")); bn.running_var.copy_(t("bn_rv0"))
    bn.bias.requires_grad_(False)                       # make_model.py:129
    bn.train(mode == "train")
    x = t("feat").to(dtype).requires_grad_(True)
    if dtype == torch.bfloat16:                         # compare against the golden run only where rounding of x is negligible
        tol = 3e-2
    fb, sc = losses.BNNeckClassifier(bn, cls)(x)
    assert fb.dtype == dtype and sc.dtype == dtype
    ((sc.float() * t("cot_s")).sum() + (fb.float() * t("cot_f")).sum()).backward()
    assert rel(fb, g[f"nk_{mode}_feat"]) < tol and rel(sc, g[f"nk_{mode}_score"]) < tol
    assert rel(x.grad, g[f"nk_{mode}_dx"]) < tol
    assert rel(bn.weight.grad, g[f"nk_{mode}_dbn_w"]) < tol and rel(cls.weight.grad, g[f"nk_{mode}_dcls_w"]) < tol
    assert bn.bias.grad is None
    assert rel(bn.running_mean, g[f"nk_{mode}_rm"]) < tol and rel(bn.running_var, g[f"nk_{mode}_rv"]) < tol
    assert int(bn.num_batches_tracked) == (1 if mode == "train" else 0)
