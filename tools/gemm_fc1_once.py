"""Dev tool for `ncu --set full`: the roofline GEMM of bench.py (FFN fc1, M = 131 584, N = 2048, K = 384, +bias, bf16 out) and
its fused-SwiGLU training form (hidden + pre-activation outputs), three launches each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vtp_b200 import lib

BF = torch.bfloat16
M, N, K = 131584, 2048, 384
A = (torch.randn(M, K, device="cuda") * 0.1).to(BF)
W = (torch.randn(N, K, device="cuda") * 0.1).to(BF)
b = torch.zeros(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=BF)
hid = torch.empty(M, N // 2, device="cuda", dtype=BF)
for _ in range(3):
    lib.gemm(A, W, out, M=M, N=N, K=K, bias=b)
for _ in range(3):
    lib.gemm(A, W, hid, M=M, N=N, K=K, bias=b, act=lib.ACT_SWIGLU8, ldo=N // 2, out2=out)
torch.cuda.synchronize()
print("done")
