"""Host-side construction of the axial RoPE sin/cos table.

The table is input-independent in eval (reference recomputes it in every block, encoders/vision_transformer.py:228-233);
it is built ONCE per (H, W) with the same torch ops and dtype (bf16) as layers/embeddings.py:131-195 and handed to the
QKV-GEMM epilogue as a constant."""
from __future__ import annotations

import math

import torch


def rope_periods(head_dim: int = 64, base: float = 100.0, dtype=torch.bfloat16) -> torch.Tensor:
    """layers/embeddings.py:182-188 (base parametrisation)."""
    return base ** (2 * torch.arange(head_dim // 4, dtype=dtype) / (head_dim // 2))


def rope_sincos(H: int, W: int, periods: torch.Tensor):
    """layers/embeddings.py:131-180, normalize_coords='separate', eval mode. Returns (sin, cos) [H*W, head_dim]."""
    periods = periods.detach().to("cpu")
    dd = {"dtype": periods.dtype}
    ch = torch.arange(0.5, H, **dd) / H
    cw = torch.arange(0.5, W, **dd) / W
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)
    coords = 2.0 * coords - 1.0
    angles = 2 * math.pi * coords[:, :, None] / periods[None, None, :]
    angles = angles.flatten(1, 2).tile(2)
    return torch.sin(angles).contiguous(), torch.cos(angles).contiguous()
