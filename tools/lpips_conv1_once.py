"""Dev tool for ncu: VGG conv1_2 (64 -> 64 channels at 256 x 256, 32 images) through the implicit-conv GEMM, 3 launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vtp_b200 import lib

BF = torch.bfloat16
B, hw, ci, co = 32, 256, 64, 64
M = B * hw * hw
x = (torch.randn(B, hw, hw, ci, device="cuda") * 0.5).to(BF)
w = (torch.randn(co, 9 * ci, device="cuda") * 0.05).to(BF)
bias = torch.zeros(co, device="cuda")
y = torch.empty(B, hw, hw, co, device="cuda", dtype=BF)
for _ in range(3):
    lib.gemm(x, w, y, M=M, N=co, K=9 * ci, lda=ci, ldb=9 * ci, bias=bias, act=lib.ACT_RELU, ldo=co, conv=(ci, hw, hw))
torch.cuda.synchronize()
print("done")
