"""Event-timed fwd+bwd of the ID / triplet loss drop-ins (signal_b200.losses) at the training shape
(B=128, D=3*768, C=171), eager and as a CUDA-graph replay, next to the same math written with stock
torch CUDA ops the way the reference does (layers/softmax_loss.py: one-hot built on the HOST every call;
layers/triplet_loss.py: boolean-mask reshapes).  python tools/bench_losses.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from signal_b200 import losses

B, K, D, C = 128, 8, 3 * 768, 171
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(5)
labels = torch.randperm(C, generator=g)[: B // K].repeat_interleave(K)[torch.randperm(B, generator=g)].to(dev)
feat = (0.02 * torch.randn(C, D, generator=g)[labels.cpu()] + 0.05 * torch.randn(B, D, generator=g)).to(dev).requires_grad_(True)
logits = (3.0 * torch.randn(B, C, generator=g)).to(dev).requires_grad_(True)
xent, tri = losses.CrossEntropyLabelSmooth(C), losses.TripletLoss()


def ours():
    feat.grad = None; logits.grad = None
    (0.25 * xent(logits, labels) + tri(feat, labels)[0]).backward()


def stock():   # the reference's formulation with torch CUDA ops
    feat.grad = None; logits.grad = None
    logp = torch.log_softmax(logits, 1)
    t = torch.zeros(logp.size()).scatter_(1, labels.unsqueeze(1).data.cpu(), 1).to(dev)     # softmax_loss.py:30-31
    t = 0.9 * t + 0.1 / C
    lx = (-t * logp).mean(0).sum()
    xx = feat.pow(2).sum(1, keepdim=True).expand(B, B)
    dist = (xx + xx.t() - 2 * feat @ feat.t()).clamp(min=1e-12).sqrt()
    pos = labels.expand(B, B).eq(labels.expand(B, B).t())
    ap = dist[pos].view(B, -1).max(1)[0]
    an = dist[~pos].view(B, -1).min(1)[0]
    lt = torch.nn.functional.soft_margin_loss(an - ap, torch.ones_like(an))
    (0.25 * lx + lt).backward()


def time(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


print(f"ours  eager : {time(ours):8.1f} us per fwd+bwd (xent + triplet)")
print(f"stock eager : {time(stock):8.1f} us  (includes the reference's host round trip for the one-hot targets)")
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ours(); torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        ours()
torch.cuda.synchronize()
print(f"ours  graph : {time(gr.replay):8.1f} us  (6 kernels; the stock version cannot be captured: .cpu())")
