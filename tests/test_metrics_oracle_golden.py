"""CPU: the inference-tail oracle (oracle/metrics_oracle.py) against golden vectors from the live reference's
utils/metrics.py (eval_func, euclidean_distance)."""
import numpy as np
import pytest
import torch

import golden_util as gu
import metrics_cases as mc
from oracle import metrics_oracle as mo


@pytest.mark.parametrize("name", list(mc.CASES))
def test_metrics_oracle_matches_reference(name):
    c = mc.CASES[name]
    rec = gu.load("metrics_" + name)
    feats, pids, camids = mc.make_case(c)
    nq = c["nq"]
    f = torch.nn.functional.normalize(torch.from_numpy(feats), dim=1, p=2).numpy() if c["norm"] else feats
    dist = mo.euclidean_distance(f[:nq], f[nq:])
    assert np.abs(dist - rec["distmat"]).max() <= 2e-5 * np.abs(rec["distmat"]).max()
    # ranking on the reference's own fp32 distance matrix: identical CMC curve, mAP to rounding
    cmc, mAP = mo.eval_func(rec["distmat"], pids[:nq], pids[nq:], camids[:nq], camids[nq:], max_rank=c["max_rank"])
    assert cmc.dtype == np.float32 and cmc.shape == rec["cmc"].shape and np.array_equal(cmc, rec["cmc"])
    assert abs(mAP - float(rec["mAP"])) < 1e-12


def test_metrics_oracle_ties_take_the_lowest_index():
    dist = np.zeros((1, 4), dtype=np.float32)        # all distances equal: stable order 0,1,2,3
    cmc, mAP = mo.eval_func(dist, [7], [1, 7, 2, 7], [0], [1, 1, 1, 1], max_rank=4)
    assert cmc.tolist() == [0.0, 1.0, 1.0, 1.0] and abs(mAP - (1 / 2 + 2 / 4) / 2) < 1e-12


def test_inference_features_concat_and_norm():
    g = np.random.default_rng(0)
    cls = [g.standard_normal((5, 8)) for _ in range(3)]
    v = g.standard_normal((5, 24))
    f = mo.inference_features(cls, v, normalize=True)
    want = torch.nn.functional.normalize(torch.cat([torch.from_numpy(x) for x in cls + [v]], dim=-1), dim=1, p=2).numpy()
    assert f.shape == (5, 48) and np.abs(f - want).max() < 1e-15
