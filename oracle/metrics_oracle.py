"""CPU oracle of the inference tail.  TEST INFRASTRUCTURE ONLY -- nothing in ``signal_b200/`` imports it.

numpy restatement of utils/metrics.py:111-170 (``eval_func``), :494-501 (``euclidean_distance``) and of the feature
concat / normalisation (modeling/make_model.py:284-290, metrics.py:266-268), pinned against the live reference functions
by tests/golden/make_metrics_golden.py -> tests/golden/metrics_*.npz.  One deliberate difference: the ranking is a STABLE
ascending sort (ties -> lowest gallery index); numpy's default ``argsort`` in the reference leaves ties unspecified.
"""
import numpy as np


def inference_features(cls3, vars_total, normalize=False):
    """make_model.py:284-290 (+ F.normalize(dim=1, p=2, eps=1e-12), metrics.py:266-268)."""
    f = np.concatenate([np.asarray(c, dtype=np.float64) for c in cls3] + [np.asarray(vars_total, dtype=np.float64)], axis=1)
    if normalize:
        f = f / np.maximum(np.linalg.norm(f, axis=1, keepdims=True), 1e-12)
    return f


def euclidean_distance(qf, gf):
    """|q|^2 + |g|^2 - 2 q.g^T -- metrics.py:494-501 (squared distances)."""
    qf, gf = np.asarray(qf, dtype=np.float64), np.asarray(gf, dtype=np.float64)
    return (qf ** 2).sum(1, keepdims=True) + (gf ** 2).sum(1, keepdims=True).T - 2.0 * qf @ gf.T


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50):
    """metrics.py:111-170: for each query drop the gallery entries with its pid AND camid, rank the rest by distance,
    CMC = cumulative first-match indicator, AP = mean precision at the match positions; queries whose identity is not in
    the (valid) gallery are skipped."""
    distmat = np.asarray(distmat)
    q_pids, g_pids, q_camids, g_camids = (np.asarray(x) for x in (q_pids, g_pids, q_camids, g_camids))
    num_q, num_g = distmat.shape
    max_rank = min(max_rank, num_g)
    indices = np.argsort(distmat, axis=1, kind="stable")
    all_cmc, all_ap = [], []
    for q in range(num_q):
        order = indices[q]
        keep = ~((g_pids[order] == q_pids[q]) & (g_camids[order] == q_camids[q]))
        hits = (g_pids[order] == q_pids[q])[keep].astype(np.int64)
        if not hits.any():
            continue
        cmc = np.minimum(hits.cumsum(), 1)
        all_cmc.append(cmc[:max_rank])
        prec = hits.cumsum() / np.arange(1, hits.shape[0] + 1, dtype=np.float64)
        all_ap.append(float((prec * hits).sum() / hits.sum()))
    assert all_cmc, "Error: all query identities do not appear in gallery"
    return np.asarray(all_cmc).astype(np.float32).sum(0) / np.float32(len(all_cmc)), float(np.mean(all_ap))
