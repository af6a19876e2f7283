"""Dev tool: one launch set of three recurring short-K GEMM shapes for `ncu --set full --import-source on`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtp_b200 import lib
BF = torch.bfloat16; dev = "cuda"; M = 2 * 256 * 257; D = 384; Hs = 1024
x = (torch.randn(M, D, device=dev) * 0.1).to(BF); xh = (torch.randn(M, Hs, device=dev) * 0.1).to(BF)
W12 = (torch.randn(2 * Hs, D, device=dev) * 0.1).to(BF); Wp = (torch.randn(D, D, device=dev) * 0.1).to(BF)
W3 = (torch.randn(D, Hs, device=dev) * 0.1).to(BF)
b2h = torch.zeros(2 * Hs, device=dev); bd = torch.zeros(D, device=dev)
o2h = torch.empty(M, 2 * Hs, device=dev, dtype=BF); stream = torch.randn(M, D, device=dev)
for _ in range(2):
    lib.gemm(x, W12, o2h, M=M, N=2 * Hs, K=D, bias=b2h)
    lib.gemm(x, Wp, stream, M=M, N=D, K=D, bias=bd, resid=stream)
    lib.gemm(xh, W3, stream, M=M, N=D, K=Hs, bias=bd, resid=stream)
torch.cuda.synchronize()
print("done")
