"""Dev tool: the 13 VGG16 convolutions of the LPIPS term as the step runs them (chunks of 32 images at 256 x 256, implicit-conv
GEMM, bias + ReLU epilogue; conv1_1 through the 27 -> 32 im2col) — us per launch, TFLOP/s, and bytes moved, per layer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vtp_b200 import lib

BF = torch.bfloat16
B = 32
cfg = [(3, 64, 256), (64, 64, 256), (64, 128, 128), (128, 128, 128), (128, 256, 64), (256, 256, 64), (256, 256, 64),
       (256, 512, 32), (512, 512, 32), (512, 512, 32), (512, 512, 16), (512, 512, 16), (512, 512, 16)]


def t(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tot_f = tot_b = 0.0
for i, (ci, co, hw) in enumerate(cfg):
    M = B * hw * hw
    bias = torch.zeros(co, device="cuda")
    y = torch.empty(B, hw, hw, co, device="cuda", dtype=BF)
    if i == 0:
        col = (torch.randn(M, 32, device="cuda") * 0.5).to(BF)
        w = (torch.randn(co, 32, device="cuda") * 0.1).to(BF)
        fwd = lambda: lib.gemm(col, w, y, M=M, N=co, K=32, bias=bias, act=lib.ACT_RELU, ldo=co)
        K = 27
        bwd = None
    else:
        x = (torch.randn(B, hw, hw, ci, device="cuda") * 0.5).to(BF)
        w = (torch.randn(co, 9 * ci, device="cuda") * 0.05).to(BF)
        fwd = lambda: lib.gemm(x, w, y, M=M, N=co, K=9 * ci, lda=ci, ldb=9 * ci, bias=bias, act=lib.ACT_RELU, ldo=co, conv=(ci, hw, hw))
        K = 9 * ci
        # dgrad: conv of dY [.., co] with the rotated kernel [ci, 9 co], masked by the layer input's ReLU
        dy = (torch.randn(B, hw, hw, co, device="cuda") * 0.5).to(BF)
        wb = (torch.randn(ci, 9 * co, device="cuda") * 0.05).to(BF)
        dx = torch.empty(B, hw, hw, ci, device="cuda", dtype=BF)
        bwd = lambda: lib.gemm(dy, wb, dx, M=M, N=ci, K=9 * co, lda=co, ldb=9 * co, ldo=ci, conv=(co, hw, hw), round_bf16=False, mask_pos=x)
    uf = t(fwd)
    fl = 2.0 * M * co * K
    line = f"conv{i:2d} {ci:3d}->{co:3d} @{hw:3d}: fwd {uf:7.1f} us {fl / uf / 1e6:7.0f} TFLOP/s ({(M * (ci if i else 32) + M * co) * 2 / uf / 1e3:6.0f} GB/s in+out)"
    tot_f += uf
    if bwd is not None:
        ub = t(bwd)
        tot_b += ub
        line += f"   dgrad {ub:7.1f} us {fl / ub / 1e6:7.0f} TFLOP/s"
    print(line, flush=True)
print(f"per 32-image chunk: fwd {tot_f / 1e3:.2f} ms, dgrad {tot_b / 1e3:.2f} ms  ->  per 256-image step (2 fwd + 1 dgrad): {(16 * tot_f + 8 * tot_b) / 1e3:.1f} ms")
