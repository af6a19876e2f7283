"""Dev tool: launch a few representative GEMM shapes once each (for ncu --set full)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtp_b200 import lib
BF = torch.bfloat16; dev = "cuda"; M = 65792; D, Hs = 384, 1024
mk = lambda *s: (torch.randn(*s, device=dev) * 0.1).to(BF)
x, x2h = mk(M, D), mk(M, 2 * Hs)
Wqkv, Wproj, W12 = mk(3 * D, D), mk(D, D), mk(2 * Hs, D)
b3d, bd, b2h = torch.zeros(3 * D, device=dev), torch.zeros(D, device=dev), torch.zeros(2 * Hs, device=dev)
stream = torch.randn(M, D, device=dev)
o_qkv = torch.empty(M, 3 * D, device=dev, dtype=BF); o_d = torch.empty(M, D, device=dev, dtype=BF)
o_h = torch.empty(M, Hs, device=dev, dtype=BF); o_2h = torch.empty(M, 2 * Hs, device=dev, dtype=BF)
torch.cuda.synchronize()
lib.gemm(x, Wqkv, o_qkv, M=M, N=3 * D, K=D, bias=b3d)                                                    # 0 qkv plain
lib.gemm(x, Wproj, stream, M=M, N=D, K=D, bias=bd, resid=stream)                                          # 1 proj + resid
lib.gemm(x, W12, o_h, M=M, N=2 * Hs, K=D, bias=b2h, act=lib.ACT_SWIGLU8, ldo=Hs, out2=o_2h)               # 2 fc1 swiglu + pre
lib.gemm(x2h, W12, o_d, M=M, N=D, K=2 * Hs, b_mn=True, ldb=D, round_bf16=False)                            # 3 dgrad fc1
torch.cuda.synchronize()
print("done")
