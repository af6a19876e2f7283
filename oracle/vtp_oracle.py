"""TEST INFRASTRUCTURE ONLY — CPU oracle for the VTP hot path.

A plain, functional PyTorch-CPU restatement of the reference's forward algorithm (MiniMax-AI/VTP @ 5ce1eb6), written
from the reference's definitions and citing them file:line (paths relative to the reference root).  It consumes a
reference-format state dict (same keys as `VTPModel.state_dict()` / legacy `VTP.state_dict()`).  Autograd on these
functions is the gradient oracle for the training step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product path (vtp_b200/) never does.

PINNING: the model-side functions are pinned against the real reference imported in the dev container
(tests/test_oracle_golden.py::test_oracle_matches_live_reference, runs where /root/reference exists) and against
committed golden vectors generated from the real reference (tests/golden/, scripts oracle/make_golden.py for the HF
inference API and oracle/make_golden_legacy.py for the legacy training meta-arch: `VTP.forward_ssl_learning`,
`update_teacher`, `DINOHead`, `LPIPS`).  The three LOSS FUNCTIONS do not exist in the reference — they
are restated from OpenCLIP ClipLoss and DINOv2 DINOLoss/iBOTPatchLoss definitions: **loss parity is unpinned**
(checked only against an independent fp64 restatement in tests).

`mode`:
  "fp32"  — the reference run in fp32 (what `tools/test_reconstruction_hf.py` does for the decoder / on CPU).
  "bf16"  — the reference under `torch.autocast(bfloat16)`: nn.Linear/conv/matmul/SDPA inputs+outputs are bf16,
            norms compute in fp32, the encoder/text residual streams stay fp32, the decoder stream is bf16
            (probe in oracle/ref_harness.py docstring).  Rounding points are restated explicitly.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BF = torch.bfloat16


def _r(x: Tensor, mode: str) -> Tensor:
    """bf16 rounding point (value kept in fp32 so that autograd/fp64 checks stay simple)."""
    return x.to(BF).to(torch.float32) if mode == "bf16" else x


def linear(x: Tensor, w: Tensor, b: Optional[Tensor], mode: str) -> Tensor:
    """nn.Linear; under autocast input, weight and bias are cast to bf16 and the output is bf16."""
    if mode == "bf16":
        y = F.linear(_r(x, mode), _r(w, mode), None if b is None else _r(b, mode))
        return _r(y, mode)
    return F.linear(x, w, b)


# ----------------------------------------------------------------------------------------------- RoPE
def rope_periods(head_dim: int = 64, base: float = 100.0) -> Tensor:
    """layers/embeddings.py:182-195 — periods = base ** (2*arange(hd/4)/(hd/2)), computed AND stored in bf16."""
    return base ** (2 * torch.arange(head_dim // 4, dtype=BF) / (head_dim // 2))


def rope_table(H: int, W: int, periods: Tensor) -> Tuple[Tensor, Tensor]:
    """layers/embeddings.py:131-180, normalize_coords='separate', eval (no shift/jitter/rescale). All ops in the
    dtype of `periods` (bf16, vision_transformer.py:74,136).  Returns (sin, cos), each [H*W, head_dim]."""
    dd = {"dtype": periods.dtype}
    coords_h = torch.arange(0.5, H, **dd) / H
    coords_w = torch.arange(0.5, W, **dd) / W
    coords = torch.stack(torch.meshgrid(coords_h, coords_w, indexing="ij"), dim=-1).flatten(0, 1)
    coords = 2.0 * coords - 1.0
    angles = 2 * math.pi * coords[:, :, None] / periods[None, None, :]
    angles = angles.flatten(1, 2).tile(2)
    return torch.sin(angles), torch.cos(angles)


def _rot_half(x: Tensor) -> Tensor:
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat([-x2, x1], dim=-1)


def apply_rope(q: Tensor, k: Tensor, sin: Tensor, cos: Tensor) -> Tuple[Tensor, Tensor]:
    """layers/attention.py:70-89 — q,k [B,H,N,64] are cast to the table dtype (bf16), rotated on the last HW tokens
    with every elementwise op rounded to bf16, then cast back.  This happens in fp32 mode too."""
    prefix = q.shape[-2] - sin.shape[-2]
    assert prefix >= 0
    out = []
    for t in (q, k):
        tb = t.to(sin.dtype)
        rot = (tb[:, :, prefix:] * cos) + (_rot_half(tb[:, :, prefix:]) * sin)
        out.append(torch.cat([tb[:, :, :prefix], rot], dim=-2).to(t.dtype))
    return out[0], out[1]


class _RopeSTE(torch.autograd.Function):
    """RoPE with the reference's bf16 forward values and the exact linear (fp32) backward — autograd through the
    bf16 casts gives the same thing; this is only here so the oracle can be run in float64 for loss checks."""

    @staticmethod
    def forward(ctx, q, k, sin, cos):
        ctx.save_for_backward(sin, cos)
        return apply_rope(q, k, sin, cos)

    @staticmethod
    def backward(ctx, gq, gk):  # pragma: no cover - helper
        sin, cos = ctx.saved_tensors
        sin, cos = sin.to(gq.dtype), cos.to(gq.dtype)
        prefix = gq.shape[-2] - sin.shape[-2]

        def bw(g):
            gr = g[:, :, prefix:]
            gx = gr * cos - _rot_half(gr * sin)
            return torch.cat([g[:, :, :prefix], gx], dim=-2)

        return bw(gq), bw(gk), None, None


# ----------------------------------------------------------------------------------------------- norms
def rmsnorm(x: Tensor, w: Tensor, eps: float = 1e-5) -> Tensor:
    """layers/normalization.py:17-22 — fp32 internal, cast back to x.dtype BEFORE the weight multiply."""
    xf = x.float()
    y = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x)
    return y * w


def layernorm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# ----------------------------------------------------------------------------------------------- blocks
def sdpa(q: Tensor, k: Tensor, v: Tensor, mode: str, causal: bool = False) -> Tensor:
    """F.scaled_dot_product_attention(q,k,v) (layers/attention.py:124): softmax(q kᵀ / sqrt(hd)) v, no dropout.
    bf16 mode: q,k,v are bf16 values; P is rounded to bf16 before P·V (tensor-core semantics), output rounded."""
    scale = q.shape[-1] ** -0.5
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        n = s.shape[-1]
        s = s + torch.full((n, n), float("-inf"), dtype=s.dtype).triu(1)
    m = s.amax(-1, keepdim=True)
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    o = (_r(p, mode) @ v) / l
    return _r(o, mode)


def self_attention(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int, rope, mode: str) -> Tensor:
    """layers/attention.py:91-126 (SelfAttention.forward + compute_attention), qkv packing [q|k|v] x [H,hd]."""
    B, N, C = x.shape
    qkv = linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias"), mode)
    qkv = qkv.reshape(B, N, 3, heads, C // heads)
    q, k, v = [t.transpose(1, 2) for t in torch.unbind(qkv, 2)]
    if rope is not None:
        q, k = _RopeSTE.apply(q, k, rope[0], rope[1]) if q.dtype == torch.float64 else apply_rope(q, k, *rope)
    o = sdpa(q, k, v, mode).transpose(1, 2).reshape(B, N, C)
    return linear(o, sd[pre + "proj.weight"], sd.get(pre + "proj.bias"), mode)


def swiglu(x: Tensor, sd: Dict[str, Tensor], pre: str, mode: str) -> Tensor:
    """layers/ffn.py:77-81."""
    x1 = linear(x, sd[pre + "w1.weight"], sd.get(pre + "w1.bias"), mode)
    x2 = linear(x, sd[pre + "w2.weight"], sd.get(pre + "w2.bias"), mode)
    h = _r(_r(F.silu(x1), mode) * x2, mode)
    return linear(h, sd[pre + "w3.weight"], sd.get(pre + "w3.bias"), mode)


def _norm(x: Tensor, sd, pre: str, kind: str) -> Tensor:
    if kind == "rmsnorm":
        return rmsnorm(x, sd[pre + "weight"])
    eps = 1e-6 if kind == "layernorm" else 1e-5
    return layernorm(x, sd[pre + "weight"], sd[pre + "bias"], eps)


def block(x: Tensor, sd, pre: str, heads: int, rope, norm_kind: str, mode: str, stream_bf16: bool, drop=None) -> Tensor:
    """layers/block.py:290-296 (eval / drop_ratio=0 branch; LayerScale is Identity when init_values is None).
    drop = (idx1, scale1, idx2, scale2): the training branch with batch-subset stochastic depth, layers/block.py:201-233 —
    `x[indices]` -> sub-layer -> `torch.index_add(x, 0, residual, indices, alpha=residual_scale_factor)`; the subsets are
    arguments here (the reference draws them with torch.randperm inside get_branges_scales, block.py:20-118)."""
    if drop is not None:
        idx1, s1, idx2, s2 = drop
        r1 = self_attention(_norm(x[idx1], sd, pre + "norm1.", norm_kind), sd, pre + "attn.", heads, rope, mode)
        x = torch.index_add(x, 0, idx1, r1.to(x.dtype), alpha=s1)
        r2 = swiglu(_norm(x[idx2], sd, pre + "norm2.", norm_kind), sd, pre + "mlp.", mode)
        return torch.index_add(x, 0, idx2, r2.to(x.dtype), alpha=s2)
    a = self_attention(_norm(x, sd, pre + "norm1.", norm_kind), sd, pre + "attn.", heads, rope, mode)
    x = x + a
    if stream_bf16:
        x = _r(x, mode)
    m = swiglu(_norm(x, sd, pre + "norm2.", norm_kind), sd, pre + "mlp.", mode)
    x = x + m
    if stream_bf16:
        x = _r(x, mode)
    return x


# ----------------------------------------------------------------------------------------------- trunk
def patch_embed(img: Tensor, sd, pre: str, mode: str) -> Tensor:
    """layers/embeddings.py:61-70 — Conv2d(3,D,16,16) == im2col GEMM; returns [B, HW, D]."""
    w, b = sd[pre + "proj.weight"], sd[pre + "proj.bias"]
    ps = w.shape[-1]
    y = F.conv2d(_r(img, mode), _r(w, mode), _r(b, mode), stride=ps)
    return _r(y, mode).flatten(2).transpose(1, 2)


def trunk_forward(img_list: Sequence[Tensor], masks_list: Sequence[Optional[Tensor]], sd, *, pre: str = "trunk.",
                  depth: int, heads: int, norm_kind: str = "rmsnorm", mode: str = "fp32",
                  use_bottleneck: bool = True, drops=None) -> List[Dict[str, Tensor]]:
    """encoders/vision_transformer.py:189-258 (prepare_tokens_with_masks + forward_features_list) and
    encoders/vision_transformer_bottleneck.py:48-79.  Encoder residual stream is fp32 in both modes."""
    periods = sd[pre + "rope_embed.periods"].to(BF)
    xs, ropes = [], []
    for img, masks in zip(img_list, masks_list):
        x = patch_embed(img, sd, pre + "patch_embed.", mode)
        B, HW, D = x.shape
        h, w = img.shape[-2] // 16, img.shape[-1] // 16
        cls = sd[pre + "cls_token"]
        if masks is not None:
            x = torch.where(masks.unsqueeze(-1), sd[pre + "mask_token"].to(x.dtype).unsqueeze(0), x)
        else:
            cls = cls + 0 * sd[pre + "mask_token"]
        xs.append(torch.cat([cls.expand(B, -1, -1).to(x.dtype), x], dim=1))
        ropes.append(rope_table(h, w, periods))
    # drops[j][i] = (idx1, scale1, idx2, scale2) of list element j in block i (stochastic depth), or None
    for i in range(depth):
        xs = [block(x, sd, f"{pre}blocks.{i}.", heads, r, norm_kind, mode, False,
                    drop=None if drops is None or drops[j] is None else drops[j][i])
              for j, (x, r) in enumerate(zip(xs, ropes))]
    outs = []
    for x, masks in zip(xs, masks_list):
        xn = _norm(x, sd, pre + "norm.", norm_kind)
        cls_t, patch_t = xn[:, 0], xn[:, 1:]
        if use_bottleneck and (pre + "feature_bottleneck.weight") in sd:
            wb = sd[pre + "feature_bottleneck.weight"]
            cls_t, patch_t = linear(cls_t, wb, None, mode), linear(patch_t, wb, None, mode)
        outs.append({"x_norm_clstoken": cls_t, "x_norm_patchtokens": patch_t, "x_prenorm": x, "masks": masks})
    return outs


def reconstruction_latents(img: Tensor, sd, *, depth: int, heads: int, mode: str = "fp32") -> Tensor:
    """vtp_hf/modeling_vtp.py:337-360,379-395."""
    out = trunk_forward([img], [None], sd, depth=depth, heads=heads, mode=mode)[0]
    pt = out["x_norm_patchtokens"]
    B, N, C = pt.shape
    return pt.transpose(1, 2).reshape(B, C, img.shape[-2] // 16, img.shape[-1] // 16)


def decode_latents(lat: Tensor, sd, *, pre: str = "pixel_decoder.", depth: int, heads: int, mode: str = "fp32",
                   norm_kind: str = "layernorm") -> Tensor:
    """decoders/pixel_decoder.py:134-162.  Under autocast the decoder residual stream is bf16."""
    B, _, H, W = lat.shape
    w_in, b_in = sd[pre + "proj_in.weight"], sd[pre + "proj_in.bias"]
    x = linear(lat.flatten(2).transpose(1, 2), w_in.flatten(1), b_in, mode)
    rope = rope_table(H, W, sd[pre + "rope_embed.periods"].to(BF))
    for i in range(depth):
        x = block(x, sd, f"{pre}blocks.{i}.", heads, rope, norm_kind, mode, True)
    x = _r(_norm(x, sd, pre + "norm.", norm_kind), mode)
    w_out, b_out = sd[pre + "proj_out.weight"], sd[pre + "proj_out.bias"]
    y = linear(x, w_out.flatten(1), b_out, mode)  # [B, HW, 3*r*r]
    r = int(math.isqrt(w_out.shape[0] // 3))
    y = y.transpose(1, 2).reshape(B, -1, H, W)
    return F.pixel_shuffle(y, r)


def clip_image_feature(img: Tensor, sd, *, depth: int, heads: int, mode: str = "fp32", normalize: bool = True,
                       trunk_pre: str = "trunk.", proj_key: str = "visual_proj.weight") -> Tensor:
    """vtp_hf/modeling_vtp.py:244-276 (bottleneck_ae_only=True, clip_feat='cls')."""
    out = trunk_forward([img], [None], sd, pre=trunk_pre, depth=depth, heads=heads, mode=mode, use_bottleneck=False)[0]
    f = linear(out["x_norm_clstoken"], sd[proj_key], None, mode)
    return F.normalize(f, dim=-1) if normalize else f


# ----------------------------------------------------------------------------------------------- text tower
def text_feature(ids: Tensor, sd, *, layers: int, heads: int, mode: str = "fp32", normalize: bool = True,
                 pre: str = "text_transformer.") -> Tensor:
    """vtp_hf/modeling_vtp.py:278-310; layers/block.py:370-427 (ResidualAttentionBlock with nn.MultiheadAttention,
    additive causal mask encoders/text_transformer.py:334-338); text_global_pool argmax text_transformer.py:213-228."""
    x = sd["token_embedding.weight"][ids] + sd["positional_embedding"]
    B, L, D = x.shape
    hd = D // heads
    for i in range(layers):
        p = f"{pre}resblocks.{i}."
        h = layernorm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"], mode)
        q, k, v = [t.reshape(B, L, heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        o = sdpa(q, k, v, mode, causal=True).transpose(1, 2).reshape(B, L, D)
        x = x + linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"], mode)
        h = layernorm(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"], mode)
        h = _r(F.gelu(h), mode)
        x = x + linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"], mode)
    x = layernorm(x, sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    pooled = x[torch.arange(B), ids.argmax(dim=-1)]
    f = _r(_r(pooled, mode) @ _r(sd["text_projection"], mode), mode)
    return F.normalize(f, dim=-1) if normalize else f


# ----------------------------------------------------------------------------------------------- DINO head
def dino_head(x: Tensor, sd, pre: str, mode: str = "fp32") -> Tensor:
    """heads/dino_head.py:65-89 with nlayers=3, weight-normed last layer (weight_g * v/||v||, :48-49)."""
    h = _r(F.gelu(linear(x, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"], mode)), mode)
    h = _r(F.gelu(linear(h, sd[pre + "mlp.2.weight"], sd[pre + "mlp.2.bias"], mode)), mode)
    h = linear(h, sd[pre + "mlp.4.weight"], sd[pre + "mlp.4.bias"], mode)
    h = F.normalize(h, dim=-1, p=2, eps=1e-12)
    g = sd.get(pre + "last_layer.weight_g", sd.get(pre + "last_layer.parametrizations.weight.original0"))
    v = sd.get(pre + "last_layer.weight_v", sd.get(pre + "last_layer.parametrizations.weight.original1"))
    w = g * v / v.norm(dim=1, keepdim=True)
    return linear(h, w, None, mode)


# ----------------------------------------------------------------------------------------------- LPIPS
VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
LPIPS_TAPS = (1, 3, 6, 9, 12)  # conv index (0-based) whose ReLU output is tapped: relu1_2, 2_2, 3_3, 4_3, 5_3
LPIPS_SHIFT = (-0.030, -0.088, -0.188)
LPIPS_SCALE = (0.458, 0.448, 0.450)


def lpips(x: Tensor, y: Tensor, vgg_w: Sequence[Tensor], vgg_b: Sequence[Tensor], lin_w: Sequence[Tensor],
          mode: str = "fp32") -> Tensor:
    """utils/lpips.py:84-100,103-114,127-171 — ScalingLayer, VGG16 features (13 conv3x3+ReLU, 4 maxpool), channel
    unit-normalise (eps 1e-10), squared diff, 1x1 lin (no bias, dropout is identity in eval), spatial mean, sum of
    the 5 taps.  Returns [B,1,1,1].  Weights are caller-supplied (the pretrained ones need network access)."""
    shift = torch.tensor(LPIPS_SHIFT, dtype=x.dtype)[None, :, None, None]
    scale = torch.tensor(LPIPS_SCALE, dtype=x.dtype)[None, :, None, None]

    def feats(t):
        t = (t - shift) / scale
        outs, ci = [], 0
        for c in VGG_CFG:
            if c == "M":
                t = F.max_pool2d(t, 2, 2)
            else:
                t = _r(F.relu(F.conv2d(_r(t, mode), _r(vgg_w[ci], mode), _r(vgg_b[ci], mode), padding=1)), mode)
                if ci in LPIPS_TAPS:
                    outs.append(t)
                ci += 1
        return outs

    fx, fy = feats(x), feats(y)
    val = 0
    for k in range(5):
        nx = fx[k] / (torch.sqrt(torch.sum(fx[k] ** 2, dim=1, keepdim=True)) + 1e-10)
        ny = fy[k] / (torch.sqrt(torch.sum(fy[k] ** 2, dim=1, keepdim=True)) + 1e-10)
        d = (nx - ny) ** 2
        val = val + F.conv2d(d, lin_w[k]).mean([2, 3], keepdim=True)
    return val


# ----------------------------------------------------------------------------------------------- losses (RESTATED)
def clip_loss(img_f: Tensor, txt_f: Tensor, logit_scale_exp: Tensor) -> Tensor:
    """OpenCLIP `ClipLoss` (open_clip/loss.py): logits = scale * I Tᵀ ; 0.5*(CE(logits, arange)+CE(logitsᵀ, arange)).
    Multi-rank: features are all-gathered first, so this is evaluated on the GLOBAL batch."""
    logits = logit_scale_exp * img_f @ txt_f.t()
    labels = torch.arange(logits.shape[0])
    return 0.5 * (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels))


def teacher_probs(t_logits: Tensor, center: Tensor, temp: float) -> Tensor:
    """DINOv2 `softmax_center_teacher`: softmax((t - center) / teacher_temp)."""
    return F.softmax((t_logits - center) / temp, dim=-1)


def dino_ibot_loss(student_local: Tensor, student_global: Tensor, student_masked: Tensor, teacher_cls: Tensor,
                   teacher_masked: Tensor, masks_weight: Tensor, *, n_local: int, n_images: int,
                   student_temp: float = 0.1) -> Dict[str, Tensor]:
    """DINOv2 ssl_meta_arch.forward_backward loss terms (dinov2/train/ssl_meta_arch.py) with centering teacher:
      teacher_cls    [2B, K] softmaxed+centred teacher cls probs ALREADY SWAPPED (vtp.py:425-426)
      student_global [2B, K] logits;  student_local [n_local*B, K] logits (crop-major);  student_masked [n_m, K]
      teacher_masked [n_m, K] probs;  masks_weight [n_m] = 1 / (#masked patches in that image)
    dino_local  = sum_{l,g} mean_b( -t_g · logsoftmax(s_l/τ) ) / (n_g_terms + n_l_terms)
    dino_global = mean_rows( -t_swapped · logsoftmax(s_g/τ) ) * 2 / (n_g_terms + n_l_terms)
    ibot        = -sum_i w_i (t_i · logsoftmax(s_i/τ)) / n_images * 2        (loss_scales = 2, ibot_loss_weight = 1)
    with n_g_terms = 2, n_l_terms = 2*n_local."""
    n_g_terms, n_l_terms = 2, max(2 * n_local, 1)
    B = teacher_cls.shape[0] // 2
    lsm = lambda s: F.log_softmax(s / student_temp, dim=-1)
    t_views = teacher_cls.chunk(2)
    loc = 0
    for sl in student_local.chunk(n_local):
        ls = lsm(sl)
        for tv in t_views:
            loc = loc - (tv * ls).sum(-1).mean()
    loc = loc / (n_g_terms + n_l_terms)
    glo = -(teacher_cls * lsm(student_global)).sum(-1).mean() * 2 / (n_g_terms + n_l_terms)
    ib = -((teacher_masked * lsm(student_masked)).sum(-1) * masks_weight).sum() / n_images * 2
    return {"dino_local": loc, "dino_global": glo, "ibot": ib}


def recon_loss(rec: Tensor, target: Tensor, lpips_val: Optional[Tensor], lpips_weight: float = 1.0) -> Tensor:
    """pixel L1 + λ·LPIPS (north_star 'pixel/LPIPS reconstruction'; weights fixed by SURVEY.md §8d)."""
    l = (rec - target).abs().mean()
    if lpips_val is not None:
        l = l + lpips_weight * lpips_val.mean()
    return l


# ----------------------------------------------------------------------------------------------- legacy meta-arch
def legacy_to_hf_keys(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Legacy `VTP.state_dict()` (vtp/models/vtp.py: `proj`, `transformer.resblocks.*`, `teacher_*`, `dino_head.*`)
    -> the HF `VTPModel` key names the rest of this file reads (`visual_proj`, `text_transformer.resblocks.*`); teacher
    and head keys pass through unchanged."""
    out = {}
    for k, v in sd.items():
        if k == "proj.weight":
            out["visual_proj.weight"] = v
        elif k.startswith("transformer.resblocks."):
            out["text_" + k] = v
        else:
            out[k] = v
    return out


def ssl_forward(sd, global_crops: Tensor, local_crops: Tensor, masks: Tensor, mask_indices: Tensor, *, depth: int,
                heads: int, mode: str = "fp32", upperbound: Optional[int] = None) -> Dict[str, Tensor]:
    """`VTP.forward_ssl_learning` (vtp/models/vtp.py:365-386): teacher side `get_teacher_forward_outputs` :410-450
    (no-grad teacher trunk on the global crops, the two views' cls tokens swapped :425-426, masked patch tokens
    index_select'ed behind them into an `upperbound`-row buffer :432-439, teacher head on the buffer) and student side
    `get_student_ssl_outputs` :452-484 (list forward [global with masks, local], three head calls; the masked patches
    go through the head inside a zero-padded `upperbound` buffer :470-477).  sd uses legacy keys (`teacher_trunk.*`,
    `dino_head.*`, `teacher_dino_head.*`)."""
    n_m = int(mask_indices.numel())
    ub = n_m if upperbound is None else upperbound
    with torch.no_grad():
        t = trunk_forward([global_crops], [None], sd, pre="teacher_trunk.", depth=depth, heads=heads, mode=mode,
                          use_bottleneck=False)[0]
        tcls = t["x_norm_clstoken"]
        half = tcls.shape[0] // 2
        tcls = torch.cat([tcls[half:], tcls[:half]])
        tpatch = t["x_norm_patchtokens"].flatten(0, 1)
        buf = tpatch.new_zeros(ub + tcls.shape[0], tpatch.shape[-1])
        buf[:tcls.shape[0]] = tcls
        buf[tcls.shape[0]:tcls.shape[0] + n_m] = tpatch[mask_indices]
        t_after = dino_head(buf, sd, "teacher_dino_head.", mode=mode)
    sg, sl = trunk_forward([global_crops, local_crops], [masks, None], sd, depth=depth, heads=heads, mode=mode,
                           use_bottleneck=False)
    sp = sg["x_norm_patchtokens"].flatten(0, 1)
    sbuf = sp.new_zeros(ub, sp.shape[-1])
    sbuf[:n_m] = sp[mask_indices]
    return {"teacher_cls": t_after[:tcls.shape[0]], "teacher_masked": t_after[tcls.shape[0]:tcls.shape[0] + n_m],
            "student_local": dino_head(sl["x_norm_clstoken"], sd, "dino_head.", mode=mode),
            "student_global": dino_head(sg["x_norm_clstoken"], sd, "dino_head.", mode=mode),
            "student_masked": dino_head(sbuf, sd, "dino_head.", mode=mode)[:n_m]}


TEACHER_PAIRS = (("trunk.", "teacher_trunk."), ("proj.", "teacher_proj."), ("visual_proj.", "teacher_proj."),
                 ("dino_head.", "teacher_dino_head."))


def update_teacher(sd: Dict[str, Tensor], momentum: float) -> None:
    """`VTP.update_teacher` (vtp/models/vtp.py:388-401), in place on a legacy-keyed dict: every PARAMETER of the trunk,
    the clip projection and the DINO head (buffers such as rope_embed.periods are not parameters and are skipped)."""
    with torch.no_grad():
        for k in list(sd.keys()):
            for s_pre, t_pre in TEACHER_PAIRS:
                if k.startswith(s_pre) and not k.endswith("rope_embed.periods"):
                    tk = t_pre + k[len(s_pre):]
                    if tk in sd:
                        sd[tk] = momentum * sd[tk] + (1 - momentum) * sd[k]
