"""Dev tool: where does the short-K GEMM lose its time?  Times the step's recurring shapes under the kernel's
diagnostic switches (VTP_GEMM_DBG bits, VTP_GEMM_NO_CLUSTER, VTP_GEMM_NO_2PERSM) and prints torch.matmul (cuBLAS) on
the same shapes as an "achievable on this box" yardstick (information only: nothing in the product calls it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vtp_b200 import lib

BF = torch.bfloat16
dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131584


def t(fn, reps=8):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def mk(*s):
    return (torch.randn(*s, device=dev) * 0.1).to(BF)


D, Hs = 384, 1024
x = mk(M, D); xh = mk(M, Hs)
Wqkv, Wproj, W12, W3 = mk(3 * D, D), mk(D, D), mk(2 * Hs, D), mk(D, Hs)
b3d, bd, b2h = torch.zeros(3 * D, device=dev), torch.zeros(D, device=dev), torch.zeros(2 * Hs, device=dev)
stream = torch.randn(M, D, device=dev)
o_qkv = torch.empty(M, 3 * D, device=dev, dtype=BF); o_d = torch.empty(M, D, device=dev, dtype=BF)
o_2h = torch.empty(M, 2 * Hs, device=dev, dtype=BF)
cases = [
    ("fc1 plain  N2048 K384 bf16", 2 * M * 2 * Hs * D, lambda: lib.gemm(x, W12, o_2h, M=M, N=2 * Hs, K=D, bias=b2h),
     lambda: torch.addmm(b2h.to(BF), x, W12.t(), out=o_2h)),
    ("qkv plain  N1152 K384 bf16", 2 * M * 3 * D * D, lambda: lib.gemm(x, Wqkv, o_qkv, M=M, N=3 * D, K=D, bias=b3d),
     lambda: torch.addmm(b3d.to(BF), x, Wqkv.t(), out=o_qkv)),
    ("proj+res   N384  K384 f32 ", 2 * M * D * D, lambda: lib.gemm(x, Wproj, stream, M=M, N=D, K=D, bias=bd, resid=stream),
     lambda: torch.mm(x, Wproj.t(), out=o_d)),
    ("proj bf16  N384  K384 bf16", 2 * M * D * D, lambda: lib.gemm(x, Wproj, o_d, M=M, N=D, K=D, bias=bd),
     lambda: torch.mm(x, Wproj.t(), out=o_d)),
    ("fc2+res    N384  K1024 f32", 2 * M * D * Hs, lambda: lib.gemm(xh, W3, stream, M=M, N=D, K=Hs, bias=bd, resid=stream),
     lambda: torch.mm(xh, W3.t(), out=o_d)),
]
variants = [("default (TMA store)", {}), ("dbg16 LSU store", {"VTP_GEMM_DBG": "16"}), ("dbg1 no store", {"VTP_GEMM_DBG": "1"})]
KEYS = ["VTP_GEMM_DBG", "VTP_GEMM_NO_CLUSTER", "VTP_GEMM_NO_2PERSM", "VTP_GEMM_NO_FAST", "VTP_GEMM_NO_G2"]
print(f"M = {M}   (us per launch; TFLOP/s in brackets for the full-work variants)")
print(f"{'variant':18s}" + "".join(f"{c[0]:>30s}" for c in cases))
for vname, env in variants:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    row = f"{vname:18s}"
    for name, fl, fn, _ in cases:
        us = t(fn)
        row += f"{us:18.1f} [{fl / us / 1e6:7.1f}]  "
    print(row, flush=True)
for k in KEYS:
    os.environ.pop(k, None)
row = f"{'cuBLAS (info)':18s}"
for name, fl, _, ref in cases:
    us = t(ref)
    row += f"{us:18.1f} [{fl / us / 1e6:7.1f}]  "
print(row)
# plain device copy of the fc1 output size, for the HBM yardstick
src = torch.empty(M, 2 * Hs, device=dev, dtype=BF)
us = t(lambda: o_2h.copy_(src))
print(f"copy {src.numel() * 2 / 1e6:.0f} MB: {us:.1f} us = {2 * src.numel() * 2 / us / 1e3:.0f} GB/s (r+w)")
