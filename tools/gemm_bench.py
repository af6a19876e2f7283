"""Dev tool: time the tcgen05 GEMM on the shapes of the VTP-Small training step (GPU only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vtp_b200 import lib

BF = torch.bfloat16
dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131584


def t(fn, reps=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def mk(*s):
    return (torch.randn(*s, device=dev) * 0.1).to(BF)


D, Hs = 384, 1024
x = mk(M, D); xh = mk(M, Hs); x2h = mk(M, 2 * Hs); xq = mk(M, 3 * D)
Wqkv, Wproj, W12, W3 = mk(3 * D, D), mk(D, D), mk(2 * Hs, D), mk(D, Hs)
b3d, bd, b2h = torch.zeros(3 * D, device=dev), torch.zeros(D, device=dev), torch.zeros(2 * Hs, device=dev)
stream = torch.randn(M, D, device=dev)
sin = torch.zeros(256, 64, device=dev, dtype=BF); cos = torch.ones(256, 64, device=dev, dtype=BF)
o_qkv = torch.empty(M, 3 * D, device=dev, dtype=BF); o_d = torch.empty(M, D, device=dev, dtype=BF)
o_h = torch.empty(M, Hs, device=dev, dtype=BF); o_2h = torch.empty(M, 2 * Hs, device=dev, dtype=BF)
gW = torch.zeros(2 * Hs, D, device=dev); gW3 = torch.zeros(D, Hs, device=dev); gWq = torch.zeros(3 * D, D, device=dev); gWp = torch.zeros(D, D, device=dev)
cases = [
    ("fwd qkv+rope      ", 2 * M * 3 * D * D, lambda: lib.gemm(x, Wqkv, o_qkv, M=M, N=3 * D, K=D, bias=b3d, act=lib.ACT_ROPE, rope=(sin, cos, 257, 1, 2 * D))),
    ("fwd qkv plain     ", 2 * M * 3 * D * D, lambda: lib.gemm(x, Wqkv, o_qkv, M=M, N=3 * D, K=D, bias=b3d)),
    ("fwd proj+resid f32", 2 * M * D * D, lambda: lib.gemm(x, Wproj, stream, M=M, N=D, K=D, bias=bd, resid=stream)),
    ("fwd fc1 swiglu+pre", 2 * M * 2 * Hs * D, lambda: lib.gemm(x, W12, o_h, M=M, N=2 * Hs, K=D, bias=b2h, act=lib.ACT_SWIGLU8, ldo=Hs, out2=o_2h)),
    ("fwd fc1 swiglu    ", 2 * M * 2 * Hs * D, lambda: lib.gemm(x, W12, o_h, M=M, N=2 * Hs, K=D, bias=b2h, act=lib.ACT_SWIGLU8, ldo=Hs)),
    ("fwd fc1 plain     ", 2 * M * 2 * Hs * D, lambda: lib.gemm(x, W12, o_2h, M=M, N=2 * Hs, K=D, bias=b2h)),
    ("fwd fc2+resid f32 ", 2 * M * D * Hs, lambda: lib.gemm(xh, W3, stream, M=M, N=D, K=Hs, bias=bd, resid=stream)),
    ("dgrad fc2 (N=Hs)  ", 2 * M * D * Hs, lambda: lib.gemm(x, W3, o_h, M=M, N=Hs, K=D, b_mn=True, ldb=Hs, round_bf16=False)),
    ("dgrad fc1 (K=2Hs) ", 2 * M * 2 * Hs * D, lambda: lib.gemm(x2h, W12, o_d, M=M, N=D, K=2 * Hs, b_mn=True, ldb=D, round_bf16=False)),
    ("dgrad qkv (K=3D)  ", 2 * M * 3 * D * D, lambda: lib.gemm(xq, Wqkv, o_d, M=M, N=D, K=3 * D, b_mn=True, ldb=D, round_bf16=False)),
    ("wgrad fc1 splitK  ", 2 * M * 2 * Hs * D, lambda: lib.gemm(x2h, x, gW, M=2 * Hs, N=D, K=M, a_mn=True, b_mn=True, lda=2 * Hs, ldb=D, ldo=D, accumulate=True, split_k=8, round_bf16=False)),
    ("wgrad fc1 auto    ", 2 * M * 2 * Hs * D, lambda: lib.gemm(x2h, x, gW, M=2 * Hs, N=D, K=M, a_mn=True, b_mn=True, lda=2 * Hs, ldb=D, ldo=D, accumulate=True, split_k=-1, round_bf16=False)),
    ("wgrad fc2 auto    ", 2 * M * Hs * D, lambda: lib.gemm(x, xh, gW3, M=D, N=Hs, K=M, a_mn=True, b_mn=True, lda=D, ldb=Hs, ldo=Hs, accumulate=True, split_k=-1, round_bf16=False)),
    ("wgrad fc2 splitK8 ", 2 * M * Hs * D, lambda: lib.gemm(x, xh, gW3, M=D, N=Hs, K=M, a_mn=True, b_mn=True, lda=D, ldb=Hs, ldo=Hs, accumulate=True, split_k=8, round_bf16=False)),
    ("wgrad qkv auto    ", 2 * M * 3 * D * D, lambda: lib.gemm(xq, x, gWq, M=3 * D, N=D, K=M, a_mn=True, b_mn=True, lda=3 * D, ldb=D, ldo=D, accumulate=True, split_k=-1, round_bf16=False)),
    ("wgrad proj auto   ", 2 * M * D * D, lambda: lib.gemm(x, x, gWp, M=D, N=D, K=M, a_mn=True, b_mn=True, lda=D, ldb=D, ldo=D, accumulate=True, split_k=-1, round_bf16=False)),
    ("wgrad fc1 splitK32", 2 * M * 2 * Hs * D, lambda: lib.gemm(x2h, x, gW, M=2 * Hs, N=D, K=M, a_mn=True, b_mn=True, lda=2 * Hs, ldb=D, ldo=D, accumulate=True, split_k=32, round_bf16=False)),
]
print(f"M = {M}")
for name, fl, fn in cases:
    ms = t(fn)
    print(f"{name}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s")
