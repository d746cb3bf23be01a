"""Seeded inputs of the inference-tail tests (shared by tests/golden/make_metrics_golden.py and the tests)."""
import numpy as np

CASES = {
    # clustered features (identity centre + noise): meaningful rankings, distinct distances
    "rgbnt201_like": dict(nq=60, ng=240, ids=30, cams=4, D=96, noise=2.5, norm=True, max_rank=50, seed=1, junk=True),
    # gallery smaller than max_rank (the reference clips max_rank to ng); query cameras are disjoint from the gallery's, so
    # nothing is discarded (the reference's eval_func cannot stack CMC rows shorter than max_rank)
    "no_norm_small_gallery": dict(nq=25, ng=45, ids=10, cams=3, D=64, noise=1.2, norm=False, max_rank=50, seed=2, junk=False),
    # some query identities never appear in the gallery (skipped queries), one camera only for a few ids (all junk)
    "absent_ids": dict(nq=40, ng=150, ids=25, cams=2, D=48, noise=1.8, norm=True, max_rank=20, seed=3, junk=True, absent=5),
}


def make_case(c):
    g = np.random.default_rng(c["seed"])
    n = c["nq"] + c["ng"]
    centres = g.standard_normal((c["ids"], c["D"])).astype(np.float32)
    pids = g.integers(0, c["ids"], size=n)
    if c.get("absent"):
        # the first `absent` identities only occur among the queries
        gal = pids[c["nq"]:]
        gal[gal < c["absent"]] = c["absent"] + (gal[gal < c["absent"]] % (c["ids"] - c["absent"]))
        pids[:c["absent"]] = np.arange(c["absent"])
    camids = g.integers(0, c["cams"], size=n)
    if not c["junk"]:
        camids[:c["nq"]] += 100
    feats = centres[pids] + c["noise"] * g.standard_normal((n, c["D"])).astype(np.float32)
    return feats.astype(np.float32), pids.astype(np.int64), camids.astype(np.int64)
