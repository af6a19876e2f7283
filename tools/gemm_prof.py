"""Dev tool: launch the bench's dominant GEMM (FFN fc1, M = 2*256*257) a few times for ncu --set full."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtp_b200 import lib
BF = torch.bfloat16; dev = "cuda"; M, N, K = 2 * 256 * 257, 2048, 384
A = (torch.randn(M, K, device=dev) * 0.1).to(BF); W = (torch.randn(N, K, device=dev) * 0.1).to(BF)
b = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=BF)
for _ in range(4):
    lib.gemm(A, W, out, M=M, N=N, K=K, bias=b)
torch.cuda.synchronize()
print("done", M, N, K)
