"""Algorithmic FLOP model of the path (2·M·N·K per GEMM), used for the roofline fractions bench.py
reports.  Backward of a trainable pass is counted as 2x its forward (dgrad + wgrad), the EMA teacher as forward only."""
from __future__ import annotations

from .config import VTPConfig


def swiglu_hidden(D: int, ratio: float = 4.0) -> int:
    d = int(int(D * ratio) * 2 / 3)
    return d + (-d % 8)


def vit_block_flops(N: int, D: int, Hs: int) -> float:
    """one block forward for one sequence of N tokens: qkv + QK/PV + proj + SwiGLU (w1,w2,w3)."""
    return 2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D + 6 * N * D * Hs


def trunk_pass_flops(cfg: VTPConfig, image: int) -> float:
    D, hw = cfg.vision_embed_dim, (image // cfg.vision_patch_size) ** 2
    N = hw + 1
    Hs = swiglu_hidden(D, cfg.vision_mlp_ratio)
    patch = 2 * hw * (3 * cfg.vision_patch_size ** 2) * D
    return patch + cfg.vision_depth * vit_block_flops(N, D, Hs)


def bottleneck_flops(cfg: VTPConfig, image: int) -> float:
    return 2 * ((image // cfg.vision_patch_size) ** 2 + 1) * cfg.vision_embed_dim * cfg.vision_feature_bottleneck


def decoder_pass_flops(cfg: VTPConfig, image: int) -> float:
    D, hw = cfg.decoder_embed_dim, (image // 16) ** 2
    Hs = swiglu_hidden(D, 4.0)
    return 2 * hw * cfg.vision_feature_bottleneck * D + cfg.decoder_depth * vit_block_flops(hw, D, Hs) + 2 * hw * D * 768


def text_pass_flops(cfg: VTPConfig) -> float:
    D, L = cfg.text_embed_dim, cfg.text_context_length
    blk = 2 * L * D * 3 * D + 4 * L * L * D + 2 * L * D * D + 4 * L * D * int(D * cfg.text_mlp_ratio)
    return cfg.text_depth * blk + 2 * D * D


def head_flops(tokens: float, D: int, hidden: int, bottleneck: int, K: int) -> float:
    return 2 * tokens * (D * hidden + hidden * hidden + hidden * bottleneck + bottleneck * K)


def encode_decode_flops(cfg: VTPConfig, image: int = 256) -> float:
    """forward only: get_reconstruction_latents + get_latents_decoded_images (configs 1, 5)."""
    return trunk_pass_flops(cfg, image) + bottleneck_flops(cfg, image) + decoder_pass_flops(cfg, image)


def train_step_flops_per_image(cfg: VTPConfig, *, image: int = 256, local: int = 96, n_local: int = 8,
                               mask_ratio: float = 0.3, mask_prob: float = 0.5, head_hidden: int = 2048,
                               head_bottleneck: int = 256, K: int = 65536, lpips: bool = False) -> dict:
    """per SOURCE image of the 3-objective step defined in SURVEY.md §8(d)."""
    Pg, Pl = trunk_pass_flops(cfg, image), trunk_pass_flops(cfg, local)
    Pd, Pt = decoder_pass_flops(cfg, image), text_pass_flops(cfg)
    hw = (image // 16) ** 2
    n_mask = 2 * mask_prob * round(mask_ratio * hw)          # masked patches per source image (2 global crops)
    D = cfg.vision_embed_dim
    t_teacher, t_student = 2 + n_mask, 2 + n_local + n_mask
    clip = 3 * Pg + 3 * Pt + 3 * 2 * D * cfg.text_embed_dim
    ssl_teacher = 2 * Pg + head_flops(t_teacher, D, head_hidden, head_bottleneck, K)
    ssl_student = 3 * (2 * Pg + n_local * Pl) + 3 * head_flops(t_student, D, head_hidden, head_bottleneck, K)
    rec = 3 * (Pg + bottleneck_flops(cfg, image)) + 3 * Pd
    out = dict(clip=clip, ssl_teacher=ssl_teacher, ssl_student=ssl_student, rec=rec)
    if lpips:
        out["lpips"] = 3 * 40.1e9 * (image / 256) ** 2
    out["total"] = sum(out.values())
    return out
