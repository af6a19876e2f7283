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
    os.environ["VTP_ATTN_FWD_PIPE"] = "0"   # the one-tile-per-CTA forward for comparison
    fwd(); fwd()
    del os.environ["VTP_ATTN_FWD_PIPE"]
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
uf, ub = t(fwd), t(bwd)            # defaults: persistent ping-pong forward for 128 < HW <= 256, FULL paths at HW == 256
o_def = o.clone()
os.environ["VTP_ATTN_FWD_PIPE"] = "0"     # one-tile-per-CTA kernel, one thread per row
u4 = t(fwd)
d4 = (o.float() - o_def.float()).abs().max().item()
os.environ["VTP_ATTN_FWD8"] = "1"         # two row threads per query row (attn_fwd8_kernel)
u8 = t(fwd)
os.environ["VTP_ATTN_FWD8"] = "0"
del os.environ["VTP_ATTN_FWD_PIPE"]
# TMA-store epilogues (round 2): same bits expected, time each setting
def with_env(k, v, fn):
    old = os.environ.get(k)
    os.environ[k] = v
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[k]
        else:
            os.environ[k] = old
for v in ("0", "1"):
    uf_v = with_env("VTP_ATTN_PIPE_TSO", v, lambda: t(fwd))
    o_v, lse_v = o.clone(), lse.clone()
    ub_v = with_env("VTP_ATTN_BWD_TSO", v, lambda: t(bwd))
    dq_v = dqkv.clone()
    ub_np = with_env("VTP_ATTN_BWD_TSO", v, lambda: with_env("VTP_ATTN_BWD_NO_PREFETCH", "1", lambda: t(bwd)))
    if v == "0":
        o_0, lse_0, dq_0 = o_v, lse_v, dq_v
    print(f"TSO={v}: fwd {uf_v:.1f} us  bwd {ub_v:.1f} us (without the O-row prefetch {ub_np:.1f} us)"
          + ("" if v == "0" else f"   max|d o| {(o_v.float() - o_0.float()).abs().max().item():.3e}  max|d lse| {(lse_v - lse_0).abs().max().item():.3e}"
             f"  max|d dqkv| {(dq_v.float() - dq_0.float()).abs().max().item():.3e}"), flush=True)
print(f"fwd rows4 (one tile per CTA): {u4:.1f} us, rows8: {u8:.1f} us; default is x{u4 / uf:.2f} of rows4, max |default - rows4| = {d4:.3e}")
print(f"B={B} T={T} H={H}: fwd {uf:.1f} us ({fl / uf / 1e6:.0f} TFLOP/s, {(M * 4 * D * 2) / uf / 1e3:.0f} GB/s)   "
      f"bwd {ub:.1f} us ({2.5 * fl / ub / 1e6:.0f} TFLOP/s, {(M * 8 * D * 2) / ub / 1e3:.0f} GB/s)")
