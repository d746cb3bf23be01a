"""One warm LAM-forward-shaped GEMM launch sequence for ncu (tools/prof_one_gemm.py [bn])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from signal_b200 import lib
bn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Bs, d = 128, 768
tok = torch.randn(3, Bs, 129, d, device="cuda").to(torch.bfloat16)
W = torch.randn(d, d, device="cuda").to(torch.bfloat16)
bias = torch.randn(d, device="cuda")
out = torch.empty(Bs * 128, d, dtype=torch.bfloat16, device="cuda")
for i in range(6):
    lib.debug_gemm_bf16(tok[i % 3][:, 1:], 1, W, 0, Bs * 128, d, d, bn=bn, out_bf16=True, bias=bias)
torch.cuda.synchronize()
print("done")
