"""Inference tail of the Signal model on the device: feature concat + L2 norm, euclidean distance
matrix, market1501 CMC / mAP -- drop-ins for

* modeling/make_model.py:284-290   ``torch.cat([ori, vars_total], dim=-1)``  -> :func:`inference_features`
* utils/metrics.py:494-501          ``euclidean_distance(qf, gf)``             -> :func:`euclidean_distance`
* utils/metrics.py:111-170          ``eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50)`` -> :func:`eval_func`
* utils/metrics.py:222-301          ``R1_mAP_eval`` (reset / update / compute; the plotting helpers are out of scope)

The reference copies every batch of features to the host (metrics.py:245) and ranks in numpy; here features, distance
matrix and per-query statistics stay in HBM, all arithmetic runs in libsignal_b200.so, and only the final CMC curve / mAP
are read back.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import lib as L_

__all__ = ["inference_features", "euclidean_distance", "eval_func", "R1_mAP_eval"]


def inference_features(rgb_global, ni_global, ti_global, vars_total, normalize: bool = False) -> torch.Tensor:
    """[B, 6d] fp32 = cat([RGB_global, NI_global, TI_global, vars_total], -1) (make_model.py:284-290), optionally with the
    L2 normalisation R1_mAP_eval.compute applies (metrics.py:266-268) fused in."""
    lib = L_.load()
    cls = [rgb_global, ni_global, ti_global]
    for t in cls + [vars_total]:
        L_._require_cuda(t, "features")
    if vars_total.dtype == torch.float16:
        cls = [c.float() for c in cls]
        vars_total = vars_total.float()
    B, d = cls[0].shape
    if any(c.shape != (B, d) or c.stride(1) != 1 or c.dtype != cls[0].dtype for c in cls) or vars_total.shape != (B, 3 * d) \
            or vars_total.stride(1) != 1 or vars_total.dtype != cls[0].dtype:
        raise RuntimeError("signal_b200: inference_features needs three [B,d] CLS views and vars_total [B,3d] of one dtype")
    dev = cls[0].device
    out = torch.empty(B, 6 * d, dtype=torch.float32, device=dev)
    ptrs = (C.c_void_p * 3)(*[c.data_ptr() for c in cls])
    strides = (C.c_int64 * 3)(*[c.stride(0) for c in cls])
    with torch.cuda.device(dev):
        L_.check(lib.sig_infer_features(ptrs, strides, vars_total.data_ptr(), vars_total.stride(0), L_.dtype_enum(cls[0]), B, d,
                                        int(normalize), out.data_ptr(), dev.index, L_.stream_ptr(dev)), "sig_infer_features")
    return out


def euclidean_distance(qf: torch.Tensor, gf: torch.Tensor, as_numpy: bool = False):
    """Squared euclidean distances [nq, ng] fp32 (utils/metrics.py:494-501).  Returns a CUDA tensor (``as_numpy=True``: the
    reference's numpy array, which forces the device-to-host copy the reference always makes)."""
    lib = L_.load()
    L_._require_cuda(qf, "qf")
    L_._require_cuda(gf, "gf")
    qf, gf = qf.float().contiguous(), gf.float().contiguous()
    nq, D = qf.shape
    ng = gf.shape[0]
    dev = qf.device
    dist = torch.empty(nq, ng, dtype=torch.float32, device=dev)
    ws = torch.empty(nq + ng, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L_.check(lib.sig_euclidean_distmat(qf.data_ptr(), gf.data_ptr(), nq, ng, D, dist.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                           dev.index, L_.stream_ptr(dev)), "sig_euclidean_distmat")
    return dist.cpu().numpy() if as_numpy else dist


def _ids(x, dev) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(dev, torch.int64).contiguous()
    return torch.as_tensor(np.asarray(x), dtype=torch.int64).to(dev)


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank: int = 50):
    """market1501 CMC / mAP (utils/metrics.py:111-170) -> (all_cmc float32 [max_rank], mAP float)."""
    lib = L_.load()
    if not isinstance(distmat, torch.Tensor) or not distmat.is_cuda:
        raise RuntimeError("signal_b200: eval_func needs the CUDA distance matrix of euclidean_distance (there is no CPU path)")
    dist = distmat.float().contiguous()
    nq, ng = dist.shape
    dev = dist.device
    if ng < max_rank:
        max_rank = ng
        print("Note: number of gallery samples is quite small, got {}".format(ng))
    qp, gp, qc, gc = (_ids(x, dev) for x in (q_pids, g_pids, q_camids, g_camids))
    cmc = torch.empty(max_rank, dtype=torch.float32, device=dev)
    mp = torch.empty(2, dtype=torch.float64, device=dev)
    stats = torch.empty(3 * nq, dtype=torch.float64, device=dev)
    ovf = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L_.check(lib.sig_rank_eval(dist.data_ptr(), dist.stride(0), qp.data_ptr(), gp.data_ptr(), qc.data_ptr(), gc.data_ptr(), nq, ng,
                                   max_rank, cmc.data_ptr(), mp.data_ptr(), stats.data_ptr(), ovf.data_ptr(), dev.index,
                                   L_.stream_ptr(dev)), "sig_rank_eval")
    mp_h = mp.cpu()
    assert float(mp_h[1]) > 0, "Error: all query identities do not appear in gallery"
    if int(ovf.item()):
        raise RuntimeError("signal_b200: a query has more than 2048 gallery matches")
    return cmc.cpu().numpy(), float(mp_h[0])


class R1_mAP_eval:
    """utils/metrics.py:222-301 without the host round trips: ``update`` keeps the features on the device."""

    def __init__(self, num_query, max_rank=50, feat_norm=True, reranking=False):
        if reranking:
            raise NotImplementedError("signal_b200: k-reciprocal re-ranking (utils/reranking.py) is outside the hot path")
        self.num_query = num_query
        self.max_rank = max_rank
        self.feat_norm = feat_norm
        self.reranking = reranking
        self.reset()

    def reset(self):
        self.feats = []
        self.pids = []
        self.camids = []
        self.img_paths = []

    def update(self, output):
        feat, pid, camid = output[0], output[1], output[2]
        self.feats.append(feat.detach())
        self.pids.extend(np.asarray(pid))
        self.camids.extend(np.asarray(camid))
        if len(output) > 3:
            self.img_paths.extend(output[3])

    def compute(self):
        feats = torch.cat(self.feats, dim=0).float()
        if self.feat_norm and self.feat_norm != "no":
            feats = torch.nn.functional.normalize(feats, dim=1, p=2)
        qf, gf = feats[:self.num_query], feats[self.num_query:]
        q_pids, g_pids = np.asarray(self.pids[:self.num_query]), np.asarray(self.pids[self.num_query:])
        q_camids, g_camids = np.asarray(self.camids[:self.num_query]), np.asarray(self.camids[self.num_query:])
        distmat = euclidean_distance(qf, gf)
        cmc, mAP = eval_func(distmat, q_pids, g_pids, q_camids, g_camids, self.max_rank)
        return cmc, mAP, distmat, self.pids, self.camids, qf, gf
