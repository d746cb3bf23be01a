"""GPU: token selection from IDENTICAL fp32 scores must be bit-exact against the oracle, ties broken by the lowest
index (BASELINE north_star; useA.py:50-96, 136-221, 254-314) -- through the C-ABI test seam sig_sim_select_from_scores."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _scores(B, L, levels, seed):
    """Scores quantised to a few levels (ties everywhere), one all-equal row and one strictly increasing row."""
    g = torch.Generator().manual_seed(seed)
    q = lambda *shape: torch.randint(0, levels, shape, generator=g).float() / levels
    intra, inter, raw = q(3, B, L), q(3, B, 2 * L), q(3, B, L) * 7.0 - 3.0
    intra[0, 0] = 0.25
    inter[1, 0] = 0.5
    raw[2, 0] = 1.0
    intra[1, 1] = torch.arange(L).float()
    inter[2, 1] = -torch.arange(2 * L).float()
    return intra, inter, raw


def _oracle_masks(intra, inter, raw, which, k1, k2, max_keep):
    from oracle import signal_oracle as so
    L = intra.shape[2]
    out = []
    intra_m = [so.topk_mask_lowest_index(intra[m], k1) for m in range(3)]
    sel = [so.topk_mask_lowest_index(inter[m], k2) for m in range(3)]       # D_rgb = [NIR | TIR], D_nir = [RGB | TIR], D_tir = [RGB | NIR]
    inter_m = [sel[1][:, :L] | sel[2][:, :L], sel[0][:, :L] | sel[2][:, L:], sel[0][:, L:] | sel[1][:, L:]]
    if which == 1:
        out = intra_m
    elif which == 2:
        out = inter_m
    else:
        out = [a | b for a, b in zip(inter_m, intra_m)]
        if max_keep >= 0:
            out = [so.keep_ratio_adjust(m, raw[i], max_keep) for i, m in enumerate(out)]
    return torch.stack(out).float()


@pytest.mark.parametrize("which", [1, 2, 3])
@pytest.mark.parametrize("levels,k,max_keep", [(4, 80, -1), (16, 112, -1), (3, 40, 64), (1000, 64, 96), (2, 200, -1)])
def test_selection_from_identical_scores_is_bit_exact(which, levels, k, max_keep):
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import functional as F_
    B, L = 5, 128
    if which != 3 and max_keep >= 0:
        pytest.skip("keep_ratio only applies to the union")
    intra, inter, raw = _scores(B, L, levels, seed=levels * 31 + k)
    got = F_.select_from_scores(intra.cuda(), inter.cuda(), raw.cuda(), which, k, 2 * k, max_keep).cpu()
    want = _oracle_masks(intra, inter, raw, which, k, 2 * k, max_keep)
    assert got.shape == want.shape == (3, B, L)
    assert torch.equal(got, want), f"{int((got != want).sum())} mask entries differ"
