import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        meta = json.load(f)
    data = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    return meta, {k: torch.from_numpy(v) for k, v in data.items()}


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def golden_inputs(meta):
    from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict

    sd = seeded_state_dict(meta["spec"], seed=0, **meta.get("seed_opts", {}))
    x = seeded_images(meta["batch"], meta["image_size"], meta["image_size"])
    ids = seeded_captions(meta["batch"], 77, meta["config"]["text_vocab_size"])
    return sd, x, ids
