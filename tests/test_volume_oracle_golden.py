"""oracle.signal_oracle.volume_n (Laplace expansion, fp64) against the golden vectors of the live reference's
volume_computation4 / volume_computation5 (utils/volume.py:65-182; tests/golden/make_volume_golden.py)."""
import numpy as np
import pytest
import torch

import volume_cases as vc
from oracle import signal_oracle as so


def _rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", sorted(vc.CASES))
def test_volume_n_oracle_matches_reference(name):
    c = vc.CASES[name]
    z = vc.load(name)
    feats, cot = vc.gen.inputs(c)
    xs = [t.double().requires_grad_(True) for t in feats]
    V = so.volume_n(*xs)
    assert V.shape == (c["B1"], c["B2"])
    assert _rel(V.detach(), z["ref/vol"]) < 1e-10
    # the reference's own fp32 result is within its measured deviation of the pin
    assert _rel(z["vol32"], z["ref/vol"]) <= 3.0 * float(z["dev32/vol"]) + 1e-12
    V.backward(cot.double())
    for k, x in enumerate(xs):
        assert _rel(x.grad, z[f"ref/d{k}"]) < 1e-8, (name, k)


def test_volume_n_with_three_modalities_is_volume3():
    c = dict(n=3, B1=7, B2=5, d=32, corr=0.5, seed=3)
    feats, _ = vc.gen.inputs(c)
    assert _rel(so.volume_n(*[f.double() for f in feats]), so.volume3(*[f.double() for f in feats])) < 1e-12


@pytest.mark.parametrize("n", [3, 4, 5])
def test_volume_n_oracle_equals_torch_det(n):
    """the Laplace-expansion restatement against torch.det of the same Gram stack (fp64), values and gradients"""
    c = dict(n=n, B1=6, B2=5, d=40, corr=0.6, seed=100 + n)
    feats, cot = vc.gen.inputs(c)
    xs = [t.double().requires_grad_(True) for t in feats]
    V = so.volume_n(*xs)
    V.backward(cot.double())
    ys = [t.double().requires_grad_(True) for t in feats]
    Vd = torch.sqrt(torch.abs(torch.det(vc.gen.gram_stack64(ys))))
    Vd.backward(cot.double())
    assert _rel(V.detach(), Vd.detach()) < 1e-10
    for a, b in zip(xs, ys):
        assert _rel(a.grad, b.grad) < 1e-8
