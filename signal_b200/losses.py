"""Drop-ins for the reference's ID / metric losses: same class names, constructor
signatures and return values as layers/softmax_loss.py and layers/triplet_loss.py, computed by the CUDA
kernels of csrc/losses.cu through the C ABI (sig_xent_ls_*, sig_triplet_*).  No host synchronisation
(the reference copies the targets to the host on every call, softmax_loss.py:30), no CPU fallback."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as L_


def _rows(t: torch.Tensor, what: str) -> torch.Tensor:
    L_._require_cuda(t, what)
    if t.dim() != 2:
        raise RuntimeError(f"signal_b200: {what} must be [B, N]")
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()
    return t if t.stride(1) == 1 else t.contiguous()


def _labels(t: torch.Tensor, B: int, dev) -> torch.Tensor:
    if t.dim() != 1 or t.numel() != B:
        raise RuntimeError("signal_b200: labels must be [B]")
    return t.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()


class _XentLS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, eps):
        lib = L_.load()
        z = _rows(logits, "logits")
        B, C_ = z.shape
        dev = z.device
        y = _labels(targets, B, dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        lse = torch.empty(B, dtype=torch.float32, device=dev)
        nws = lib.sig_loss_ws_bytes(B)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L_.check(lib.sig_xent_ls_fwd(z.data_ptr(), L_.dtype_enum(z), z.stride(0), y.data_ptr(), B, C_, float(eps), loss.data_ptr(),
                                         lse.data_ptr(), ws.data_ptr(), nws, dev.index, L_.stream_ptr(dev)), "sig_xent_ls_fwd")
        ctx.save_for_backward(z, y, lse)
        ctx.eps, ctx.in_dtype = float(eps), logits.dtype
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lib = L_.load()
        z, y, lse = ctx.saved_tensors
        B, C_ = z.shape
        dev = z.device
        g = dloss.float().contiguous()
        dz = torch.empty_like(z, memory_format=torch.contiguous_format)
        with torch.cuda.device(dev):
            L_.check(lib.sig_xent_ls_bwd(z.data_ptr(), L_.dtype_enum(z), z.stride(0), y.data_ptr(), B, C_, ctx.eps, lse.data_ptr(),
                                         g.data_ptr(), dz.data_ptr(), dz.stride(0), dev.index, L_.stream_ptr(dev)), "sig_xent_ls_bwd")
        return dz.to(ctx.in_dtype), None, None


class CrossEntropyLabelSmooth(nn.Module):
    """layers/softmax_loss.py:4-34: ``(-t * log_softmax(inputs)).mean(0).sum()`` with
    ``t = (1 - epsilon) * onehot(targets) + epsilon / num_classes``.  fp32 0-dim result."""

    def __init__(self, num_classes, epsilon=0.1, use_gpu=True):
        super().__init__()
        self.num_classes = num_classes
        self.epsilon = epsilon
        self.use_gpu = use_gpu
        self.logsoftmax = nn.LogSoftmax(dim=1)     # (attribute kept for state_dict / attribute compatibility)

    def forward(self, inputs, targets):
        if inputs.size(1) != self.num_classes:
            raise RuntimeError(f"signal_b200: logits have {inputs.size(1)} classes, num_classes={self.num_classes}")
        return _XentLS.apply(inputs, targets, self.epsilon)


class LabelSmoothingCrossEntropy(nn.Module):
    """layers/softmax_loss.py:36-55: ``confidence * nll + smoothing * (-mean_k log p)``, averaged over the batch --
    the same smoothed target ``(1 - s) onehot + s / C`` as CrossEntropyLabelSmooth with C taken from the logits."""

    def __init__(self, smoothing=0.1):
        super().__init__()
        assert smoothing < 1.0
        self.smoothing = smoothing
        self.confidence = 1. - smoothing

    def forward(self, x, target):
        return _XentLS.apply(x, target, self.smoothing)


class _Triplet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, labels, margin, hard_factor):
        lib = L_.load()
        x = _rows(feat, "global_feat")
        B, D = x.shape
        dev = x.device
        y = _labels(labels, B, dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ap = torch.empty(B, dtype=torch.float32, device=dev)
        an = torch.empty(B, dtype=torch.float32, device=dev)
        idx = torch.empty(2, B, dtype=torch.int32, device=dev)
        nws = lib.sig_loss_ws_bytes(B)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        soft = margin is None
        m = 0.0 if soft else float(margin)
        with torch.cuda.device(dev):
            L_.check(lib.sig_triplet_fwd(x.data_ptr(), L_.dtype_enum(x), x.stride(0), y.data_ptr(), B, D, m, int(soft), float(hard_factor),
                                         loss.data_ptr(), ap.data_ptr(), an.data_ptr(), idx[0].data_ptr(), idx[1].data_ptr(),
                                         ws.data_ptr(), nws, dev.index, L_.stream_ptr(dev)), "sig_triplet_fwd")
        ctx.save_for_backward(x, ap, an, idx)
        ctx.cfg = (m, int(soft), float(hard_factor), feat.dtype)
        ctx.mark_non_differentiable(idx)
        return loss, ap, an, idx

    @staticmethod
    def backward(ctx, dloss, dap, dan, _didx):
        lib = L_.load()
        x, ap, an, idx = ctx.saved_tensors
        m, soft, hf, in_dtype = ctx.cfg
        B, D = x.shape
        dev = x.device
        g = None if dloss is None else dloss.float().contiguous()
        gap = None if dap is None else dap.float().contiguous()
        gan = None if dan is None else dan.float().contiguous()
        dx = torch.empty_like(x, memory_format=torch.contiguous_format)
        ptr = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            L_.check(lib.sig_triplet_bwd(x.data_ptr(), L_.dtype_enum(x), x.stride(0), B, D, m, soft, hf, ap.data_ptr(), an.data_ptr(),
                                         idx[0].data_ptr(), idx[1].data_ptr(), ptr(g), ptr(gap), ptr(gan), dx.data_ptr(), dx.stride(0),
                                         dev.index, L_.stream_ptr(dev)), "sig_triplet_bwd")
        return dx.to(in_dtype), None, None, None


def normalize(x, axis=-1):
    """layers/triplet_loss.py:5-13."""
    return 1. * x / (torch.norm(x, 2, axis, keepdim=True).expand_as(x) + 1e-12)


class TripletLoss(object):
    """layers/triplet_loss.py:106-135: triplet loss with hard example mining on the euclidean distance
    matrix of the batch; ``margin=None`` -> SoftMarginLoss, else MarginRankingLoss(margin).
    ``__call__`` returns ``(loss, dist_ap, dist_an)`` like the reference."""

    def __init__(self, margin=None, hard_factor=0.0):
        self.margin = margin
        self.hard_factor = hard_factor
        self.last_indices = None     # int32 [2, B]: hardest positive / negative of every anchor (diagnostic)

    def __call__(self, global_feat, labels, normalize_feature=False):
        if normalize_feature:
            global_feat = normalize(global_feat, axis=-1)
        loss, ap, an, idx = _Triplet.apply(global_feat, labels, self.margin, self.hard_factor)
        self.last_indices = idx
        return loss, ap, an


class _BNNeckCls(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, bn_w, bn_b, cls_w, run_mean, run_var, momentum, eps, training):
        lib = L_.load()
        x = _rows(feat, "feat")
        B, D = x.shape
        C_ = cls_w.shape[0]
        dev = x.device
        for t in (bn_w, bn_b, cls_w):
            L_._f32c(t)
        out = torch.empty(B, D, dtype=x.dtype, device=dev)
        logits = torch.empty(B, C_, dtype=x.dtype, device=dev)
        stats = torch.empty(2, D, dtype=torch.float32, device=dev)
        y32 = torch.empty(B, D, dtype=torch.float32, device=dev)
        nws = lib.sig_bnneck_ws_bytes(B, D, C_)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            L_.check(lib.sig_bnneck_cls_fwd(x.data_ptr(), L_.dtype_enum(x), x.stride(0), B, D, C_, bn_w.data_ptr(), bn_b.data_ptr(),
                                            ptr(run_mean), ptr(run_var), float(momentum), float(eps), int(training), cls_w.data_ptr(),
                                            out.data_ptr(), D, logits.data_ptr(), C_, stats[0].data_ptr(), stats[1].data_ptr(),
                                            y32.data_ptr(), ws.data_ptr(), nws, dev.index, L_.stream_ptr(dev)), "sig_bnneck_cls_fwd")
        ctx.save_for_backward(x, bn_w, cls_w, stats, y32)
        ctx.cfg = (int(training), feat.dtype)
        return out, logits

    @staticmethod
    def backward(ctx, dout, dlogits):
        lib = L_.load()
        x, bn_w, cls_w, stats, y32 = ctx.saved_tensors
        training, in_dtype = ctx.cfg
        B, D = x.shape
        C_ = cls_w.shape[0]
        dev = x.device
        cast = lambda t: None if t is None else t.to(x.dtype).contiguous()
        dout, dlogits = cast(dout), cast(dlogits)
        dx = torch.empty(B, D, dtype=x.dtype, device=dev)
        dg = torch.empty(D, dtype=torch.float32, device=dev)
        db = torch.empty(D, dtype=torch.float32, device=dev)
        dW = torch.empty(C_, D, dtype=torch.float32, device=dev)
        nws = lib.sig_bnneck_ws_bytes(B, D, C_)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            L_.check(lib.sig_bnneck_cls_bwd(x.data_ptr(), L_.dtype_enum(x), x.stride(0), B, D, C_, bn_w.data_ptr(), cls_w.data_ptr(),
                                            stats[0].data_ptr(), stats[1].data_ptr(), training, y32.data_ptr(), ptr(dlogits), C_,
                                            ptr(dout), D, dx.data_ptr(), D, dg.data_ptr(), db.data_ptr(), dW.data_ptr(), ws.data_ptr(),
                                            nws, dev.index, L_.stream_ptr(dev)), "sig_bnneck_cls_bwd")
        return dx.to(in_dtype), dg, db, dW, None, None, None, None, None


class BNNeckClassifier:
    """The BNNeck of the reference as ONE call (modeling/make_model.py:128-131, 194-195, 212-214):

        feat_bn = self.bottleneck(feat); score = self.classifier(feat_bn)      # nn.BatchNorm1d, nn.Linear(bias=False)

    becomes ``feat_bn, score = BNNeckClassifier(model.bottleneck, model.classifier)(feat)``.  Owns no parameters
    (checkpoints unchanged); honours ``bottleneck.training`` (batch statistics + in-place running-statistics update
    vs running statistics), ``momentum``, ``eps`` and a frozen ``bias`` (``requires_grad_(False)``)."""

    def __init__(self, bottleneck: nn.BatchNorm1d, classifier: nn.Linear):
        if classifier.bias is not None:
            raise RuntimeError("signal_b200: the reference's classifier has no bias (make_model.py:130)")
        self.bottleneck, self.classifier = bottleneck, classifier

    def __call__(self, feat):
        bn = self.bottleneck
        training = bn.training or bn.running_mean is None
        if training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        momentum = 0.1 if bn.momentum is None else bn.momentum
        return _BNNeckCls.apply(feat, bn.weight, bn.bias, self.classifier.weight, bn.running_mean, bn.running_var,
                                momentum, bn.eps, training)
