"""TEST INFRASTRUCTURE ONLY — deterministic, reference-independent weights and inputs.

`seeded_state_dict` fills a {key: shape} spec with per-key seeded values (CPU torch generator, seeded from a CRC of
the key), so the dev container (where the real reference is importable) and the GPU box (where it is not) construct
bit-identical weights without sharing a file.  Distributions are chosen to exercise the kernels harder than the
reference's own init: non-zero biases, non-unit norm weights, and q/k weights large enough for peaky attention."""
from __future__ import annotations

import zlib
from typing import Dict, Sequence

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2**31 - 1))


def seeded_state_dict(spec: Dict[str, Sequence[int]], seed: int = 0, dtypes: Dict[str, torch.dtype] | None = None,
                      qkv_std: float = 0.06):
    """qkv_std: std of the q/k/v projection weights (0.06 gives attention logits of std ~1.4 at D = 384; scale it by
    sqrt(384 / D) for wider models to keep the same peakiness)."""
    sd = {}
    for k in sorted(spec):
        shape = tuple(spec[k])
        g = _gen(k, seed)
        leaf = k.rsplit(".", 1)[-1]
        if k.endswith("rope_embed.periods"):
            hd4 = shape[0]
            sd[k] = (100.0 ** (2 * torch.arange(hd4, dtype=torch.bfloat16) / (2 * hd4)))
            continue
        if k == "logit_scale":
            sd[k] = torch.full(shape, 2.659260036932778)
            continue
        r = torch.randn(shape, generator=g)
        if "norm" in k or ".ln_" in k or k.startswith("ln_final") or k.endswith("weight_g"):
            sd[k] = (1.0 + 0.1 * r) if leaf in ("weight", "weight_g") else 0.05 * r
        elif leaf in ("bias", "in_proj_bias"):
            sd[k] = 0.02 * r
        elif "qkv.weight" in k or "in_proj_weight" in k:
            sd[k] = qkv_std * r
        elif k in ("trunk.cls_token", "trunk.mask_token", "positional_embedding"):
            sd[k] = 0.05 * r
        elif k == "text_projection" or "visual_proj" in k or "proj.weight" in k:
            sd[k] = r * (shape[-1] if k != "text_projection" else shape[0]) ** -0.5
        elif "patch_embed.proj.weight" in k:
            sd[k] = r * 0.03
        else:
            sd[k] = 0.03 * r
    if dtypes:
        for k, dt in dtypes.items():
            if k in sd:
                sd[k] = sd[k].to(dt)
    return sd


def seeded_images(B: int, H: int, W: int, seed: int = 1234) -> torch.Tensor:
    return torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed))


def seeded_captions(B: int, L: int = 77, vocab: int = 49408, seed: int = 4321) -> torch.Tensor:
    """SOT, U{4..40} random ids, EOT (= vocab-1, the largest id so argmax finds it), zero padding."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(B, L, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(4, 41, (1,), generator=g))
        ids[b, 0] = vocab - 2
        ids[b, 1:1 + n] = torch.randint(1, vocab - 2, (n,), generator=g)
        ids[b, 1 + n] = vocab - 1
    return ids
