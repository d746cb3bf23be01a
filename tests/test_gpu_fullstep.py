"""GPU: the head inside a complete training iteration (BASELINE.json configs[3], signal_b200/fullstep.py): backbone ->
tokens -> FusionHead -> four BNNeck heads -> ID + triplet + GAM + LAM -> backward -> Adam, against the same iteration with
the head and the losses in stock torch ops (the oracle port bench.py uses as its comparison arm)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _build(th, seed=7):
    from signal_b200 import fullstep as fs
    torch.manual_seed(seed)
    return fs.SignalTrainStep((8, 16), 64, 31, 4, 512, 0.25, 1.0, 0.2, 0.01, torch_head=th, layers=2, width=128, heads=4).cuda()


@pytest.mark.parametrize("autocast", [False, True])
def test_full_train_step_matches_stock_torch_head(autocast):
    import __graft_entry__ as entry
    entry.build()
    import bench
    B = 16
    g = torch.Generator().manual_seed(3)
    imgs = [torch.randn(B, 3, 128, 256, generator=g).cuda() for _ in range(3)]
    target = (torch.arange(B) // 4).cuda()
    cam = torch.randint(0, 4, (B,), generator=g).cuda()
    res = {}
    for arm, th in (("b200", None), ("torch", bench._torch_head)):
        step = _build(th)
        opt = torch.optim.Adam([p for p in step.parameters() if p.requires_grad], lr=1e-3)
        losses = []
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                loss = step(imgs, target, cam)
            loss.backward()
            if not losses:
                grads = {n: p.grad.detach().float().clone() for n, p in step.named_parameters() if p.grad is not None}
            opt.step()
            losses.append(float(loss.detach()))
        res[arm] = (losses, grads)
    tol = 2e-2 if autocast else 2e-4
    la, lb = res["b200"][0], res["torch"][0]
    assert abs(la[0] - lb[0]) < tol * abs(lb[0]), (la, lb)
    # after two Adam updates (lr 1e-3: every parameter moves by ~lr whatever its gradient's size, so round-off-level gradient
    # differences on near-zero entries change the trajectory at the percent level): same trajectory, loosely
    print("losses b200 / torch:", la, lb)
    assert all(abs(a - b) < 5e-2 * abs(b) for a, b in zip(la, lb)), (la, lb)
    assert la[-1] < 0.6 * la[0]                                                     # and it trains
    ga, gb = res["b200"][1], res["torch"][1]
    assert ga.keys() == gb.keys()
    # first-iteration gradients: head parameters and the backbone's (which receive the head's token gradients)
    for n in ("SIM.modal_interactive.ffn.0.weight", "AlignM.contra_temp", "classifier_var.weight", "bottleneck_r.weight",
              "backbone.proj", "backbone.blocks.0.mlp.c_fc.weight", "cv_embed"):
        e = float((ga[n] - gb[n]).norm() / gb[n].norm().clamp_min(1e-30))
        assert e < (8e-2 if autocast else 1e-3), (n, e)
