"""TEST INFRASTRUCTURE ONLY — golden vectors of `VTPModel.get_intermediate_layers_feature` (the linear-probe feature
path, tools/test_linear_probing_hf.py:109-152) from the REAL reference on the `tiny` seeded model:

    python -m oracle.make_golden_layers      ->  tests/golden/tiny_layers.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.make_golden import CONFIGS  # noqa: E402
from oracle.seeded import seeded_images, seeded_state_dict  # noqa: E402


def main():
    rh.import_reference()
    from vtp.models.vtp_hf import VTPConfig, VTPModel

    kw, B, size, _ = CONFIGS["tiny"][:4]
    m = VTPModel(VTPConfig(**kw)).eval()
    m.load_state_dict(seeded_state_dict({k: list(v.shape) for k, v in m.state_dict().items()}, seed=0))
    x = seeded_images(B, size, size)
    out = {}
    with torch.no_grad():
        a = m.get_intermediate_layers_feature(x, n=2, return_class_token=True, norm=True)
        for i, (patch, cls) in enumerate(a):
            out[f"last2_patch{i}"], out[f"last2_cls{i}"] = patch.numpy(), cls.numpy()
        (raw,) = m.get_intermediate_layers_feature(x, n=[0], reshape=True, norm=False)
        out["block0_raw_nchw"] = raw.numpy()
        b = m.get_intermediate_layers_feature(x, n=[1, 0], norm=True)        # unsorted request -> ascending block order
        out["order_patch0"], out["order_patch1"] = b[0].numpy(), b[1].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tiny_layers.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
