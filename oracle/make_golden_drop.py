"""TEST INFRASTRUCTURE ONLY — golden for batch-subset stochastic depth (layers/block.py:20-118,201-298), from the REAL
reference run in the dev container:

    python -m oracle.make_golden_drop

The reference draws its subsets with torch.randperm inside `get_branges_scales`; to make the run reproducible on the
GPU box that function is replaced by one that hands out preset subsets (stored in the golden) with the reference's own
single-process scale b / keep.  Everything else — `x[indices]`, the sub-layer, `torch.index_add(..., alpha)` — is the
reference's training branch of `SelfAttentionBlock._forward_list`."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from tests.util import golden_inputs, load_golden  # noqa: E402

RATIO = 0.3


def main():
    rh.import_reference()
    import vtp.models.layers.block as blk
    from vtp.models.vtp_hf import VTPConfig, VTPModel

    meta, _ = load_golden("tiny")
    sd, x, _ = golden_inputs(meta)
    m = VTPModel(VTPConfig(**meta["config"]))
    m.load_state_dict(sd)
    m.train()
    B = x.shape[0]
    keep = max(int(B * (1 - RATIO)), 1)
    g = torch.Generator().manual_seed(123)
    presets = [torch.randperm(B, generator=g)[:keep] for _ in range(2 * meta["config"]["vision_depth"])]
    calls = []

    def preset_branges(xx, ratio=0.0):
        calls.append(1)
        return presets[len(calls) - 1].to(xx.device), xx.shape[0] / keep

    orig, blk.get_branges_scales = blk.get_branges_scales, preset_branges
    try:
        with torch.no_grad():
            out = m.trunk(x, is_training=True, use_bottleneck=True, drop_ratio=RATIO)
    finally:
        blk.get_branges_scales = orig
    assert len(calls) == len(presets)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tiny_drop.npz"), patch=out["x_norm_patchtokens"].numpy(),
                        cls=out["x_norm_clstoken"].numpy(), presets=torch.stack(presets).numpy(), ratio=np.array([RATIO]))
    print("tiny_drop: keep", keep, "of", B, "presets", [p.tolist() for p in presets])


if __name__ == "__main__":
    main()
