"""Dev tool: attention fwd/bwd at the training-step shape (B=512 global crops, T=257, H=6); times them, and with
`--once` launches each twice for `ncu --set full --import-source on`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtp_b200 import lib, rope as rope_mod
BF = torch.bfloat16; dev = "cuda"
B, T, H = 512, 257, 6
if len(sys.argv) > 2:
    B, T = int(sys.argv[1]), int(sys.argv[2])
D = 64 * H; M = B * T
qkv = (torch.randn(M, 3 * D, device=dev) * 0.5).to(BF)
o = torch.empty(M, D, device=dev, dtype=BF); do = (torch.randn(M, D, device=dev) * 0.1).to(BF)
lse = torch.empty(B, H, T, device=dev); dqkv = torch.empty(M, 3 * D, device=dev, dtype=BF)
prefix = 1 if T in (257, 37) else 0
fwd = lambda: lib.attention_fwd(qkv, o, B, T, H, prefix=prefix, lse=lse)
bwd = lambda: lib.attention_bwd(qkv, o, do, lse, dqkv, B, T, H, prefix=prefix)
if "--once" in sys.argv:
    for _ in range(2):
        fwd(); bwd()
    os.environ["VTP_ATTN_FWD_PIPE"] = "1"   # persistent ping-pong forward (attention_pipe.cu)
    fwd(); fwd()
    os.environ["VTP_ATTN_FWD_PIPE"] = "0"
    torch.cuda.synchronize(); print("done"); sys.exit(0)
def t(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
HW = T - prefix
fl = 4.0 * T * T * 64 * B * H
uf, ub = t(fwd), t(bwd)
os.environ["VTP_ATTN_FWD8"] = "1"     # opt-in variant: two row threads per query row (attn_fwd8_kernel)
o_ref = o.clone()
uf8 = t(fwd)
os.environ["VTP_ATTN_FWD8"] = "0"
if os.environ.get("VTP_TEST_UNVALIDATED") == "1" and 128 < T - prefix <= 256:
    os.environ["VTP_ATTN_FWD_PIPE"] = "1"   # persistent ping-pong kernel (attention_pipe.cu)
    ufp = t(fwd)
    os.environ["VTP_ATTN_FWD_PIPE"] = "0"
    print(f"fwd pipe variant: {ufp:.1f} us (x{uf / ufp:.2f} vs rows4), max |diff| vs rows4 = {(o.float() - o_ref.float()).abs().max().item():.3e}")
print(f"fwd rows8 variant: {uf8:.1f} us (x{uf / uf8:.2f} vs rows4), max |diff| vs rows4 = {(o.float() - o_ref.float()).abs().max().item():.3e}")
print(f"B={B} T={T} H={H}: fwd {uf:.1f} us ({fl / uf / 1e6:.0f} TFLOP/s, {(M * 4 * D * 2) / uf / 1e3:.0f} GB/s)   "
      f"bwd {ub:.1f} us ({2.5 * fl / ub / 1e6:.0f} TFLOP/s, {(M * 8 * D * 2) / ub / 1e3:.0f} GB/s)")
