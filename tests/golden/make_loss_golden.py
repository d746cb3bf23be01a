"""Golden vectors for the ID / metric losses from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_loss_golden.py

Imports layers/softmax_loss.py and layers/triplet_loss.py from /root/reference (read-only; the package __init__ is
skipped through a pre-seeded namespace package), runs CrossEntropyLabelSmooth and TripletLoss in fp64 (pins the values
without fp32 noise) and fp32 on seeded inputs and stores inputs, outputs and input gradients in
tests/golden/losses_<case>.npz.  /root/reference does not exist on the GPU box; only these files travel.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True

CASES = {
    # name: B (= ids x instances), instances per id, feature dim, classes, margin (None = soft), hard_factor, seed
    "rgbnt201": dict(B=32, K=4, D=192, C=101, margin=None, hf=0.0, seed=11),
    "msvr310_margin": dict(B=24, K=4, D=96, C=75, margin=0.3, hf=0.0, seed=12, noise=0.45),
    "hardfactor": dict(B=16, K=2, D=128, C=40, margin=None, hf=0.2, seed=13),
}


def import_reference():
    pkg = types.ModuleType("layers")
    pkg.__path__ = ["/root/reference/layers"]
    sys.modules["layers"] = pkg
    sys.path.insert(0, "/root/reference")
    from layers.softmax_loss import CrossEntropyLabelSmooth
    from layers.triplet_loss import TripletLoss
    return CrossEntropyLabelSmooth, TripletLoss


def case_inputs(c):
    g = torch.Generator().manual_seed(c["seed"])
    ids = torch.randperm(c["C"], generator=g)[: c["B"] // c["K"]]
    labels = ids.repeat_interleave(c["K"])[torch.randperm(c["B"], generator=g)]
    # overlapping identities: hardest negatives closer than hardest positives for a good part of the anchors
    centers = 0.25 * torch.randn(c["C"], c["D"], generator=g, dtype=torch.float64)
    feat = centers[labels] + c.get("noise", 0.2) * torch.randn(c["B"], c["D"], generator=g, dtype=torch.float64)
    logits = 3.0 * torch.randn(c["B"], c["C"], generator=g, dtype=torch.float64)
    return feat, logits, labels


def main():
    Xent, Triplet = import_reference()
    for name, c in CASES.items():
        feat, logits, labels = case_inputs(c)
        out = {"feat": feat.numpy(), "logits": logits.numpy(), "labels": labels.numpy()}
        for tag, dt in (("", torch.float64), ("32", torch.float32)):
            z = logits.detach().clone().to(dt).requires_grad_(True)
            xent = Xent(c["C"], epsilon=0.1, use_gpu=False)(z, labels)
            xent.backward()
            x = feat.detach().clone().to(dt).requires_grad_(True)
            tl, ap, an = Triplet(margin=c["margin"], hard_factor=c["hf"])(x, labels)
            tl.backward()
            out.update({"xent" + tag: xent.detach().double().numpy(), "dlogits" + tag: z.grad.double().numpy(),
                        "tri" + tag: tl.detach().double().numpy(), "dist_ap" + tag: ap.detach().double().numpy(),
                        "dist_an" + tag: an.detach().double().numpy(), "dfeat" + tag: x.grad.double().numpy()})
        # BNNeck + classifier: the reference uses torch's own modules (make_model.py:128-131), run here in fp64
        g = torch.Generator().manual_seed(c["seed"] + 1000)
        D, C = c["D"], c["C"]
        bn = torch.nn.BatchNorm1d(D).double()
        cls = torch.nn.Linear(D, C, bias=False).double()
        with torch.no_grad():
            bn.weight.copy_(1.0 + 0.3 * torch.randn(D, generator=g, dtype=torch.float64))
            bn.bias.copy_(0.1 * torch.randn(D, generator=g, dtype=torch.float64))
            bn.running_mean.copy_(0.2 * torch.randn(D, generator=g, dtype=torch.float64))
            bn.running_var.copy_(0.5 + torch.rand(D, generator=g, dtype=torch.float64))
            cls.weight.copy_(0.05 * torch.randn(C, D, generator=g, dtype=torch.float64))
        cot_s = torch.randn(c["B"], C, generator=g, dtype=torch.float64)
        cot_f = torch.randn(c["B"], D, generator=g, dtype=torch.float64)
        out.update({"bn_w": bn.weight.detach().numpy().copy(), "bn_b": bn.bias.detach().numpy().copy(), "cls_w": cls.weight.detach().numpy().copy(),
                    "bn_rm0": bn.running_mean.numpy().copy(), "bn_rv0": bn.running_var.numpy().copy(),
                    "cot_s": cot_s.numpy(), "cot_f": cot_f.numpy()})
        for mode in ("train", "eval"):
            bn.train(mode == "train")
            bn.running_mean.copy_(torch.from_numpy(out["bn_rm0"])); bn.running_var.copy_(torch.from_numpy(out["bn_rv0"]))
            for p_ in list(bn.parameters()) + list(cls.parameters()):
                p_.grad = None
            x = feat.detach().clone().requires_grad_(True)
            fb = bn(x)
            sc = cls(fb)
            ((sc * cot_s).sum() + (fb * cot_f).sum()).backward()
            out.update({f"nk_{mode}_feat": fb.detach().numpy(), f"nk_{mode}_score": sc.detach().numpy(), f"nk_{mode}_dx": x.grad.numpy(),
                        f"nk_{mode}_dbn_w": bn.weight.grad.numpy().copy(), f"nk_{mode}_dbn_b": bn.bias.grad.numpy().copy(),
                        f"nk_{mode}_dcls_w": cls.weight.grad.numpy().copy(),
                        f"nk_{mode}_rm": bn.running_mean.numpy().copy(), f"nk_{mode}_rv": bn.running_var.numpy().copy()})
        np.savez_compressed(os.path.join(HERE, f"losses_{name}.npz"), **out)
        print(name, float(out["xent"]), float(out["tri"]))


if __name__ == "__main__":
    main()
