"""Host orchestration of the sm_100a kernels: weight packing and tower forwards.

Mirrors, stage by stage, the reference's L1/L2 code (vtp/models/layers/*, encoders/*, decoders/*) — every numeric op
is a C-ABI kernel call from `lib`; torch is used for device memory (torch.empty) and integer index glue only.

Two precision modes, selected by the caller (VTPModel maps them from the autocast state like the reference):
  "bf16" — equals the reference under torch.autocast(bfloat16): bf16 GEMM/attention operands, fp32 accumulation,
           fp32 norms, fp32 residual stream in the encoder/text tower, bf16 stream in the decoder.
  "fp32" — equals the reference in fp32: every GEMM runs as a bf16x3 split (hi·hi + hi·lo + lo·hi, K-concatenated so
           the same tcgen05 kernel is used; error ~2^-16), activations fp32, attention on fp32 CUDA cores.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib
from .rope import rope_sincos

BF = torch.bfloat16
F32 = torch.float32
# bf16 hot path: run the RoPE rotation and the SwiGLU gate as stand-alone full-occupancy kernels after a plain GEMM instead
# of inside the GEMM epilogue (measured faster at K = 384: see csrc/elementwise.cu).  The fused epilogues stay available
# (fp32 mode uses them; set False to use them in bf16 mode too — identical numerics, tests cover both).
SPLIT_EPILOGUES = True
# SwiGLU gate: the lean TMA-store epilogue variant of round 2 (csrc/gemm.cu fast_swiglu_tile) writes the hidden tensor
# (and, for training, the pre-activation) straight from the fc1 GEMM — no stand-alone gate pass re-reading [M, 2Hs].
# VTP_FUSED_SWIGLU=0 restores the stand-alone swiglu_fwd kernel.
import os as _os
FUSED_SWIGLU = _os.environ.get("VTP_FUSED_SWIGLU", "1") != "0"


def _e(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


# ------------------------------------------------------------------------------------------------------ packing
@dataclass
class Lin:
    """One packed linear: w bf16 [N, Kp] (Kp = K in bf16 mode, 3K split hi|lo|hi in fp32 mode), b fp32 [N] | None."""
    w: torch.Tensor
    b: Optional[torch.Tensor]
    N: int
    K: int

    @property
    def Kp(self) -> int:
        return self.w.shape[1]


def pack_lin(w: torch.Tensor, b: Optional[torch.Tensor], mode: str) -> Lin:
    """w fp32 [N, K] (already on the device).  bf16 mode rounds the bias to bf16 as autocast does."""
    N, K = w.shape
    w = w.detach().to(F32).contiguous()
    if mode == "bf16":
        wp = w.to(BF).contiguous()
        bp = None if b is None else b.detach().to(BF).to(F32).contiguous()
    else:
        Kpad = (K + 7) // 8 * 8
        if Kpad != K:
            w = torch.nn.functional.pad(w, (0, Kpad - K))
        wp = _e((N, 3 * Kpad), BF, w.device)
        lib.split3(w, wp, N, Kpad, b_side=True)
        bp = None if b is None else b.detach().to(F32).contiguous()
    return Lin(wp, bp, N, K)


def interleave8(w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """rows [16g,16g+8) = w1[8g:8g+8], rows [16g+8,16g+16) = w2[8g:8g+8]  (SwiGLU gate epilogue layout)."""
    Hs = w1.shape[0]
    rest = w1.shape[1:]
    return torch.stack([w1.reshape(Hs // 8, 8, *rest), w2.reshape(Hs // 8, 8, *rest)], dim=1).reshape(2 * Hs, *rest)


@dataclass
class BlockW:
    n1_w: torch.Tensor
    n1_b: Optional[torch.Tensor]
    qkv: Lin
    proj: Lin
    n2_w: torch.Tensor
    n2_b: Optional[torch.Tensor]
    fc1: Lin          # SwiGLU: 8-interleaved w1|w2 (N = 2*Hs);  text MLP: c_fc (N = 4*D)
    fc2: Lin          # w3 / c_proj
    hidden: int


@dataclass
class TowerW:
    """A ViT-style stack (vision trunk, pixel decoder or text transformer) in packed form."""
    D: int
    heads: int
    norm: str                 # "rms" | "ln"
    eps: float
    stream_bf16: bool         # residual stream dtype under autocast (decoder: bf16)
    prefix: int               # cls tokens
    ffn: str                  # "swiglu" | "gelu"
    blocks: List[BlockW] = field(default_factory=list)
    norm_w: Optional[torch.Tensor] = None
    norm_b: Optional[torch.Tensor] = None
    periods: Optional[torch.Tensor] = None
    extra: Dict[str, object] = field(default_factory=dict)
    _rope: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = field(default_factory=dict)

    def rope(self, H: int, W: int, dev):
        key = (H, W)
        if key not in self._rope:
            sin, cos = rope_sincos(H, W, self.periods.to(BF))
            self._rope[key] = (sin.to(dev).contiguous(), cos.to(dev).contiguous())
        return self._rope[key]


def _f(t):
    return t.detach().to(F32).contiguous()


def pack_vit_blocks(sd: Dict[str, torch.Tensor], pre: str, depth: int, mode: str, ln: bool) -> List[BlockW]:
    blocks = []
    for i in range(depth):
        p = f"{pre}blocks.{i}."
        w1, w2 = sd[p + "mlp.w1.weight"], sd[p + "mlp.w2.weight"]
        Hs = w1.shape[0]
        blocks.append(BlockW(
            n1_w=_f(sd[p + "norm1.weight"]), n1_b=_f(sd[p + "norm1.bias"]) if ln else None,
            qkv=pack_lin(sd[p + "attn.qkv.weight"], sd.get(p + "attn.qkv.bias"), mode),
            proj=pack_lin(sd[p + "attn.proj.weight"], sd.get(p + "attn.proj.bias"), mode),
            n2_w=_f(sd[p + "norm2.weight"]), n2_b=_f(sd[p + "norm2.bias"]) if ln else None,
            fc1=pack_lin(interleave8(w1, w2), interleave8(sd[p + "mlp.w1.bias"], sd[p + "mlp.w2.bias"]), mode),
            fc2=pack_lin(sd[p + "mlp.w3.weight"], sd.get(p + "mlp.w3.bias"), mode),
            hidden=Hs))
    return blocks


def pack_trunk(sd, cfg, mode: str, pre: str = "trunk.") -> TowerW:
    """encoders/vision_transformer.py:58-187 + vision_transformer_bottleneck.py:11-46 parameters."""
    ln = cfg.vision_norm_layer != "rmsnorm"
    eps = {"rmsnorm": 1e-5, "layernorm": 1e-6, "layernormbf16": 1e-5}[cfg.vision_norm_layer]
    W = TowerW(D=cfg.vision_embed_dim, heads=cfg.vision_num_heads, norm="ln" if ln else "rms", eps=eps,
               stream_bf16=False, prefix=1, ffn="swiglu")
    W.blocks = pack_vit_blocks(sd, pre, cfg.vision_depth, mode, ln)
    W.norm_w = _f(sd[pre + "norm.weight"])
    W.norm_b = _f(sd[pre + "norm.bias"]) if ln else None
    W.periods = sd[pre + "rope_embed.periods"].detach().cpu()
    pw = sd[pre + "patch_embed.proj.weight"]
    W.extra["patch"] = pack_lin(pw.flatten(1), sd[pre + "patch_embed.proj.bias"], mode)
    W.extra["patch_size"] = pw.shape[-1]
    cls = _f(sd[pre + "cls_token"]).reshape(-1) + 0 * _f(sd[pre + "mask_token"]).reshape(-1)
    W.extra["cls"] = cls.contiguous()
    mt = _f(sd[pre + "mask_token"]).reshape(-1)
    W.extra["mask_token"] = (mt.to(BF).to(F32) if mode == "bf16" else mt).contiguous()
    if (pre + "feature_bottleneck.weight") in sd:
        W.extra["bneck"] = pack_lin(sd[pre + "feature_bottleneck.weight"], None, mode)
    return W


def pack_decoder(sd, cfg, mode: str, pre: str = "pixel_decoder.") -> TowerW:
    """decoders/pixel_decoder.py:15-132 parameters."""
    ln = cfg.decoder_norm_layer != "rmsnorm"
    eps = {"rmsnorm": 1e-5, "layernorm": 1e-6, "layernormbf16": 1e-5}[cfg.decoder_norm_layer]
    W = TowerW(D=cfg.decoder_embed_dim, heads=cfg.decoder_num_heads, norm="ln" if ln else "rms", eps=eps,
               stream_bf16=(mode == "bf16"), prefix=0, ffn="swiglu")
    W.blocks = pack_vit_blocks(sd, pre, cfg.decoder_depth, mode, ln)
    W.norm_w = _f(sd[pre + "norm.weight"])
    W.norm_b = _f(sd[pre + "norm.bias"]) if ln else None
    W.periods = sd[pre + "rope_embed.periods"].detach().cpu()
    W.extra["proj_in"] = pack_lin(sd[pre + "proj_in.weight"].flatten(1), sd.get(pre + "proj_in.bias"), mode)
    W.extra["proj_out"] = pack_lin(sd[pre + "proj_out.weight"].flatten(1), sd.get(pre + "proj_out.bias"), mode)
    return W


def pack_text(sd, cfg, mode: str, pre: str = "text_transformer.") -> TowerW:
    """encoders/text_transformer.py:231-332 + layers/block.py:370-427 parameters."""
    W = TowerW(D=cfg.text_embed_dim, heads=cfg.text_num_heads, norm="ln", eps=1e-5, stream_bf16=False, prefix=0,
               ffn="gelu")
    for i in range(cfg.text_depth):
        p = f"{pre}resblocks.{i}."
        W.blocks.append(BlockW(
            n1_w=_f(sd[p + "ln_1.weight"]), n1_b=_f(sd[p + "ln_1.bias"]),
            qkv=pack_lin(sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"], mode),
            proj=pack_lin(sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"], mode),
            n2_w=_f(sd[p + "ln_2.weight"]), n2_b=_f(sd[p + "ln_2.bias"]),
            fc1=pack_lin(sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"], mode),
            fc2=pack_lin(sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"], mode),
            hidden=sd[p + "mlp.c_fc.weight"].shape[0]))
    W.norm_w, W.norm_b = _f(sd["ln_final.weight"]), _f(sd["ln_final.bias"])
    W.extra["tok_emb"] = _f(sd["token_embedding.weight"])
    W.extra["pos"] = _f(sd["positional_embedding"])
    W.extra["proj"] = pack_lin(sd["text_projection"].t(), None, mode)  # x @ P  ==  linear(x, Pᵀ)
    return W


# ------------------------------------------------------------------------------------------------------ primitives
def operand(x: torch.Tensor, M: int, K: int, mode: str) -> torch.Tensor:
    """GEMM A operand of an activation: bf16 mode -> x itself (bf16); fp32 mode -> bf16x3 split [M, 3K]."""
    if mode == "bf16":
        assert x.dtype == BF
        return x
    assert x.dtype == F32
    Kpad = (K + 7) // 8 * 8
    assert Kpad == K, "fp32 mode needs K % 8 == 0"
    out = _e((M, 3 * K), BF, x.device)
    lib.split3(x, out, M, K, b_side=False)
    return out


def linear(a_op: torch.Tensor, lin: Lin, out: torch.Tensor, M: int, mode: str, **epi) -> None:
    """out = epi(a_op · lin.wᵀ + lin.b).  a_op from `operand()` / norm(y_mode=split)."""
    lib.gemm(a_op, lin.w, out, M=M, N=lin.N, K=lin.Kp, lda=lin.Kp, ldb=lin.Kp, bias=lin.b,
             round_bf16=(mode == "bf16"), **epi)


def norm(x, M, D, w, b, eps, mode, *, want: str, tape=None):
    """want: 'op' -> GEMM operand (bf16 | split3), 'f32' -> fp32 values.  Returns tensor."""
    dev = x.device
    rstd = mean = None
    if tape is not None:
        rstd = _e((M,), F32, dev)
        mean = _e((M,), F32, dev) if b is not None else None
    if want == "f32":
        y = _e((M, D), F32, dev)
        lib.norm_fwd(x, y, w, b, eps, M, D, y_mode=lib.OUT_F32, rstd=rstd, mean=mean)
    elif mode == "bf16":
        y = _e((M, D), BF, dev)
        lib.norm_fwd(x, y, w, b, eps, M, D, y_mode=lib.OUT_BF16, rstd=rstd, mean=mean)
    else:
        y = _e((M, 3 * D), BF, dev)
        lib.norm_fwd(x, y, w, b, eps, M, D, y_mode=lib.OUT_SPLIT3, rstd=rstd, mean=mean)
    if tape is not None:
        tape["rstd"], tape["mean"] = rstd, mean
    return y


class DropPlan:
    """Batch-subset stochastic depth for one tower pass (layers/block.py:20-118 `get_branges_scales`, :201-298): every
    sub-layer (attention, FFN) of every block runs on a fresh random subset of the images and its output is added back
    scaled by residual_scale_factor.

      single process:  keep = max(int(b (1 - ratio)), 1),  scale = b / keep                     (block.py:33-38)
      data parallel:   global_keep = max(int(b W (1 - ratio)), W) dealt out evenly over the W ranks (the first
                       global_keep % W ranks get one more, capped at b), scale = b W / Σ allocation  (block.py:40-67)
    The reference has rank 0 compute that allocation and broadcast it (block.py:94); it is a pure function of
    (b, ratio, W), so every rank derives it locally here — no collective.  The subset itself is
    `torch.randperm(b, device)[:keep]` like the reference (device-side, graph-capturable).  `preset` (a sequence of index
    tensors) replaces the random draws — used by the parity tests."""

    def __init__(self, ratio: float, world: int = 1, rank: int = 0, preset: Optional[Sequence[torch.Tensor]] = None):
        self.ratio, self.world, self.rank = float(ratio), int(world), int(rank)
        self.preset = list(preset) if preset is not None else None
        self.calls = 0

    def keep_and_scale(self, b: int) -> Tuple[int, float]:
        if self.world <= 1:
            keep = max(int(b * (1 - self.ratio)), 1)
            return keep, b / keep
        gb = b * self.world
        global_keep = max(int(gb * (1 - self.ratio)), self.world)
        base, extra = divmod(global_keep, self.world)
        alloc = [min(base + (1 if i < extra else 0), b) for i in range(self.world)]
        return alloc[self.rank], gb / max(sum(alloc), 1)

    def next(self, b: int, dev) -> Tuple[torch.Tensor, float]:
        keep, scale = self.keep_and_scale(b)
        if self.preset is not None:
            idx = self.preset[self.calls].to(dev, torch.long).contiguous()
            assert idx.numel() == keep, f"preset subset {self.calls} has {idx.numel()} images, the plan keeps {keep}"
        else:
            idx = torch.randperm(b, device=dev)[:keep].contiguous()
        self.calls += 1
        return idx, scale


def _block_drop(W: TowerW, bw: BlockW, x: torch.Tensor, B: int, T: int, rope, drop: DropPlan, t: Optional[dict]):
    """One block of the fp32-stream trunk under batch-subset stochastic depth, bf16 (autocast) mode
    (layers/block.py:201-233 / :238-289): gather the kept images, run the sub-layer on them, add the bf16 residual back
    into a copy of the stream with alpha = residual_scale_factor."""
    dev, D, H = x.device, W.D, W.heads
    mode = "bf16"
    # ---- attention sub-layer on subset 1
    idx1, a1 = drop.next(B, dev)
    n1 = idx1.numel()
    M1 = n1 * T
    xs1 = _e((M1, D), F32, dev)
    lib.gather_images(x, xs1, idx1, T, D)
    nt = {} if t is not None else None
    h = norm(xs1, M1, D, bw.n1_w, bw.n1_b, W.eps, mode, want="op", tape=nt)
    qkv = _e((M1, 3 * D), BF, dev)
    linear(h, bw.qkv, qkv, M1, mode)
    if rope is not None:
        lib.rope_fwd(qkv, rope[0], rope[1], M1, T, W.prefix, D)
    o = _e((M1, D), BF, dev)
    lse = _e((n1, H, T), F32, dev) if t is not None else None
    lib.attention_fwd(qkv, o, n1, T, H, prefix=W.prefix, lse=lse)
    res1 = _e((M1, D), BF, dev)
    linear(o, bw.proj, res1, M1, mode)
    x_mid = x.clone()
    lib.scatter_add_images(res1, x_mid, idx1, T, D, a1)
    # ---- FFN sub-layer on subset 2 (drawn independently, block.py:219)
    idx2, a2 = drop.next(B, dev)
    n2 = idx2.numel()
    M2 = n2 * T
    xs2 = _e((M2, D), F32, dev)
    lib.gather_images(x_mid, xs2, idx2, T, D)
    nt2 = {} if t is not None else None
    h2 = norm(xs2, M2, D, bw.n2_w, bw.n2_b, W.eps, mode, want="op", tape=nt2)
    Hd = bw.hidden
    pre = _e((M2, 2 * Hd), BF, dev)
    hid = _e((M2, Hd), BF, dev)
    if FUSED_SWIGLU:
        linear(h2, bw.fc1, hid, M2, mode, act=lib.ACT_SWIGLU8, ldo=Hd, out2=pre)
    else:
        linear(h2, bw.fc1, pre, M2, mode)
        lib.swiglu_fwd(pre, hid, M2, Hd)
    res2 = _e((M2, D), BF, dev)
    linear(hid, bw.fc2, res2, M2, mode)
    x_out = x_mid.clone()
    lib.scatter_add_images(res2, x_out, idx2, T, D, a2)
    if t is not None:
        t.update(drop1=(idx1, a1, xs1), drop2=(idx2, a2, xs2), h1=h, n1=nt, qkv=qkv, o=o, lse=lse, h2=h2, n2=nt2, pre=pre,
                 hid=hid)
    return x_out


def tower_blocks(W: TowerW, x: torch.Tensor, B: int, T: int, rope, mode: str, *, causal: bool = False,
                 tape: Optional[list] = None, taps: Optional[Dict[int, torch.Tensor]] = None,
                 drop: Optional[DropPlan] = None) -> torch.Tensor:
    """The block loop (encoders/vision_transformer.py:228-233, decoders/pixel_decoder.py:147-148,
    encoders/text_transformer.py:100-104): x [B*T, D] residual stream (fp32, or bf16 for the autocast decoder).
    Inference updates x in place; with a tape every sub-layer writes a fresh stream buffer and saves what backward
    needs.  `taps` {block index: None} is filled with copies of the stream after those blocks."""
    dev = x.device
    M, D, H = B * T, W.D, W.heads
    act = BF if mode == "bf16" else F32
    if drop is not None and drop.ratio > 0.0:
        if mode != "bf16" or W.stream_bf16 or W.ffn != "swiglu" or causal:
            raise NotImplementedError("batch-subset stochastic depth is implemented for the vision trunk in bf16 mode "
                                      "(the only tower the reference applies drop_ratio to: vtp.py:275-293,452-463,487-500)")
        for li, bw in enumerate(W.blocks):
            t = {} if tape is not None else None
            x = _block_drop(W, bw, x, B, T, rope, drop, t)
            if tape is not None:
                tape.append(t)
            if taps is not None and li in taps:
                taps[li] = x.clone()
        return x
    for li, bw in enumerate(W.blocks):
        t = {} if tape is not None else None
        # ---- attention sub-layer: x + proj(attn(rope(qkv(norm1 x))))      layers/block.py:293, attention.py:91-126
        nt = {} if t is not None else None
        h = norm(x, M, D, bw.n1_w, bw.n1_b, W.eps, mode, want="op", tape=nt)
        qkv = _e((M, 3 * D), act, dev)
        if rope is not None and mode == "bf16" and SPLIT_EPILOGUES:
            linear(h, bw.qkv, qkv, M, mode)                 # rounding to bf16 == q.to(bf16) of the reference
            lib.rope_fwd(qkv, rope[0], rope[1], M, T, W.prefix, D)
        elif rope is not None:
            linear(h, bw.qkv, qkv, M, mode, act=lib.ACT_ROPE, rope=(rope[0], rope[1], T, W.prefix, 2 * D))
        else:
            linear(h, bw.qkv, qkv, M, mode)
        o = _e((M, D), act, dev)
        lse = _e((B, H, T), F32, dev) if t is not None else None
        if mode == "bf16":
            lib.attention_fwd(qkv, o, B, T, H, prefix=W.prefix, causal=causal, lse=lse)
        else:
            lib.attention_fwd_f32(qkv, o, B, T, H, causal=causal)
        x_mid = x if t is None else torch.empty_like(x)
        linear(operand(o, M, D, mode), bw.proj, x_mid, M, mode, resid=x)
        if t is not None:
            t.update(x_in=x, h1=h, n1=nt, qkv=qkv, o=o, lse=lse)
        # ---- FFN sub-layer: x + w3(silu(w1 x) * w2 x)   /   x + c_proj(gelu(c_fc x))       layers/block.py:294
        nt2 = {} if t is not None else None
        h2 = norm(x_mid, M, D, bw.n2_w, bw.n2_b, W.eps, mode, want="op", tape=nt2)
        Hd = bw.hidden
        hid = _e((M, Hd), act, dev)
        pre = None
        if W.ffn == "swiglu" and mode == "bf16" and SPLIT_EPILOGUES and not FUSED_SWIGLU:
            pre = _e((M, 2 * Hd), BF, dev)
            linear(h2, bw.fc1, pre, M, mode)
            lib.swiglu_fwd(pre, hid, M, Hd)
            if t is None:
                pre = None
        elif W.ffn == "swiglu":
            pre = _e((M, 2 * Hd), BF, dev) if t is not None else None
            linear(h2, bw.fc1, hid, M, mode, act=lib.ACT_SWIGLU8, ldo=Hd, out2=pre)
        else:
            pre = _e((M, Hd), BF, dev) if t is not None else None
            linear(h2, bw.fc1, hid, M, mode, act=lib.ACT_GELU, out2=pre)
        x_out = x_mid if t is None else torch.empty_like(x)
        linear(operand(hid, M, Hd, mode), bw.fc2, x_out, M, mode, resid=x_mid)
        if t is not None:
            t.update(x_mid=x_mid, h2=h2, n2=nt2, pre=pre, hid=hid)
            tape.append(t)
        x = x_out
        if taps is not None and li in taps:
            taps[li] = x.clone()
    return x


# ------------------------------------------------------------------------------------------------------ vision trunk
def trunk_tokens(W: TowerW, img: torch.Tensor, mode: str, mask_idx: Optional[torch.Tensor] = None):
    """layers/embeddings.py:61-70 + encoders/vision_transformer.py:189-219: image -> token stream [B*(1+HW), D] fp32."""
    dev = img.device
    B, C, Hi, Wi = img.shape
    ps = W.extra["patch_size"]
    gh, gw = Hi // ps, Wi // ps
    HW, T, D = gh * gw, gh * gw + 1, W.D
    img = img.to(F32).contiguous()
    patch: Lin = W.extra["patch"]
    if mode == "bf16":
        a = _e((B * HW, patch.K), BF, dev)
        lib.patchify(img, a, ps)
    else:
        a32 = _e((B * HW, patch.K), F32, dev)
        lib.patchify(img, a32, ps)
        a = operand(a32, B * HW, patch.K, mode)
    x = _e((B * T, D), F32, dev)
    linear(a, patch, x, B * HW, mode, rr_group=HW, rr_skip=1)
    lib.fill_prefix_tokens(x, W.extra["cls"], B, T, 1, D)
    if mask_idx is not None and mask_idx.numel() > 0:
        lib.apply_mask_tokens(x, W.extra["mask_token"], mask_idx, HW, T, 1, D)
    return x, (B, T, gh, gw), a


def trunk_forward(W: TowerW, img: torch.Tensor, mode: str, *, mask_idx=None, tape: Optional[dict] = None,
                  taps=None, drop: Optional[DropPlan] = None):
    """encoders/vision_transformer.py:221-258 for one resolution group.  Returns (x_prenorm [B*T,D] fp32, meta)."""
    x, (B, T, gh, gw), a = trunk_tokens(W, img, mode, mask_idx)
    if tape is not None:
        tape["patch_a"] = a
    rope = W.rope(gh, gw, img.device)
    blk_tape = [] if tape is not None else None
    x = tower_blocks(W, x, B, T, rope, mode, tape=blk_tape, taps=taps, drop=drop)
    if tape is not None:
        tape["blocks"] = blk_tape
        tape["meta"] = (B, T, gh, gw)
    return x, (B, T, gh, gw)


def trunk_outputs(W: TowerW, x: torch.Tensor, meta, mode: str, *, use_bottleneck: bool):
    """final norm + cls/patch split + optional bottleneck (encoders/vision_transformer.py:246-258,
    vision_transformer_bottleneck.py:66-79).  Returns dict of [B, ...] tensors in the reference's dtypes."""
    B, T, gh, gw = meta
    M, D = B * T, W.D
    dev = x.device
    if use_bottleneck and "bneck" in W.extra:
        bn: Lin = W.extra["bneck"]
        xn = norm(x, M, D, W.norm_w, W.norm_b, W.eps, mode, want="op")
        out = _e((M, bn.N), BF if mode == "bf16" else F32, dev)
        linear(xn, bn, out, M, mode)
        out = out.view(B, T, bn.N)
    else:
        out = norm(x, M, D, W.norm_w, W.norm_b, W.eps, mode, want="f32").view(B, T, D)
    return {"x_norm_clstoken": out[:, 0], "x_norm_patchtokens": out[:, 1:], "x_prenorm": x.view(B, T, D)}


def latents_nchw(patch_tokens: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """vtp_hf/modeling_vtp.py:379-395: (B, N, C) -> (B, C, h, w); patch_tokens is a strided view [B, HW, C] of the
    [B, T, C] bottleneck output (cls row skipped via the batch stride)."""
    B, HW, C = patch_tokens.shape
    out = _e((B, C, gh, gw), patch_tokens.dtype, patch_tokens.device)
    lib.transpose_batched(patch_tokens, out, B, HW, C, in_bstride=patch_tokens.stride(0))
    return out


# ------------------------------------------------------------------------------------------------------ pixel decoder
def decoder_forward(W: TowerW, lat: torch.Tensor, mode: str, *, tape: Optional[dict] = None) -> torch.Tensor:
    """decoders/pixel_decoder.py:134-162: latents [B, C, h, w] -> image [B, 3, 16h, 16w]."""
    dev = lat.device
    B, C, gh, gw = lat.shape
    HW, D = gh * gw, W.D
    M = B * HW
    act = BF if mode == "bf16" else F32
    lat = lat.contiguous()
    if lat.dtype not in (BF, F32):
        lat = lat.to(F32)
    tok = _e((M, C), act, dev)  # (B,C,HW) -> (B,HW,C): flatten(2).transpose(1,2)
    lib.transpose_batched(lat, tok, B, C, HW)
    pin: Lin = W.extra["proj_in"]
    x = _e((M, D), BF if W.stream_bf16 else F32, dev)
    linear(operand(tok, M, C, mode), pin, x, M, mode)
    rope = W.rope(gh, gw, dev)
    blk_tape = [] if tape is not None else None
    x = tower_blocks(W, x, B, HW, rope, mode, tape=blk_tape)
    nt = {} if tape is not None else None
    xn = norm(x, M, D, W.norm_w, W.norm_b, W.eps, mode, want="op", tape=nt)
    pout: Lin = W.extra["proj_out"]
    r = int(round((pout.N // 3) ** 0.5))
    img = _e((B, 3, gh * r, gw * r), act, dev)
    linear(xn, pout, img, M, mode, pixel_shuffle=(r, gh, gw, 3), ldo=gw * r)
    if tape is not None:
        tape.update(blocks=blk_tape, meta=(B, HW, gh, gw), tok=tok, x_final=x, xn=xn, nf=nt)
    return img


# ------------------------------------------------------------------------------------------------------ text tower
def text_forward(W: TowerW, ids: torch.Tensor, mode: str, *, tape: Optional[dict] = None) -> torch.Tensor:
    """vtp_hf/modeling_vtp.py:278-310 up to (not including) the final normalize: ids int64 [B, L] -> [B, E]."""
    dev = ids.device
    B, L = ids.shape
    D = W.D
    M = B * L
    ids = ids.contiguous()
    x = _e((M, D), F32, dev)
    lib.embed_tokens(ids, W.extra["tok_emb"], W.extra["pos"], x)
    blk_tape = [] if tape is not None else None
    x = tower_blocks(W, x, B, L, None, mode, causal=True, tape=blk_tape)
    nt = {} if tape is not None else None
    xn = norm(x, M, D, W.norm_w, W.norm_b, W.eps, mode, want="f32", tape=nt)
    # text_global_pool 'argmax' (encoders/text_transformer.py:222-224): integer index glue, bit-exact
    eot = ids.argmax(dim=-1) + torch.arange(B, device=dev) * L
    act = BF if mode == "bf16" else F32
    pooled = _e((B, D), act, dev)
    lib.gather_rows(xn, pooled, eot, D)
    proj: Lin = W.extra["proj"]
    f = _e((B, proj.N), act, dev)
    linear(operand(pooled, B, D, mode), proj, f, B, mode)
    if tape is not None:
        tape.update(blocks=blk_tape, meta=(B, L), x_final=x, nf=nt, eot=eot, pooled=pooled)
    return f


def l2_normalize(f: torch.Tensor, eps: float = 1e-12, norm_out=None) -> torch.Tensor:
    out = torch.empty_like(f)
    lib.l2norm_fwd(f, out, f.shape[0], f.shape[1], eps, norm_out=norm_out)
    return out
