"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the REAL reference (imported from /root/reference,
see oracle/ref_harness.py) on seeded weights/inputs (oracle/seeded.py).  Run in the dev container:

    python -m oracle.make_golden [name ...]

Each file holds the reference outputs in fp32 and under torch.autocast("cpu", bfloat16), plus the state-dict
key->shape spec so the GPU box can rebuild the same weights without the reference."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    # name: (VTPConfig kwargs, B, image size, store_text)
    "tiny": (dict(vision_embed_dim=128, vision_depth=2, vision_num_heads=2, text_embed_dim=128, text_num_heads=2,
                  text_depth=2, decoder_embed_dim=128, decoder_num_heads=2, decoder_depth=2, text_vocab_size=1000),
             3, 64, True),
    "tiny96": (dict(vision_embed_dim=128, vision_depth=2, vision_num_heads=2, text_embed_dim=128, text_num_heads=2,
                    text_depth=2, decoder_embed_dim=128, decoder_num_heads=2, decoder_depth=2, text_vocab_size=1000),
               2, 96, False),
    "small": (dict(vision_embed_dim=384, vision_depth=12, vision_num_heads=6, text_embed_dim=384, text_num_heads=6,
                   text_depth=12, decoder_embed_dim=384, decoder_num_heads=6, decoder_depth=12, text_vocab_size=2048),
              2, 256, True),
    # VTP-Large geometry (BASELINE configs 4/5) at depth 2: D = 1024, 16 heads, SwiGLU hidden 2736 (= 16·171, the ragged
    # N/K tile case), text tower 768/12 heads
    "large2": (dict(vision_embed_dim=1024, vision_depth=2, vision_num_heads=16, text_embed_dim=768, text_num_heads=12,
                    text_depth=2, decoder_embed_dim=1024, decoder_num_heads=16, decoder_depth=2, text_vocab_size=2048),
               2, 256, True, dict(qkv_std=0.0367)),   # 0.06·sqrt(384/1024): same attention peakiness as "small"
}


def main():
    rh.import_reference()
    from vtp.models.vtp_hf import VTPConfig, VTPModel

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])
    for name, entry in CONFIGS.items():
        kw, B, size, with_text = entry[:4]
        seed_opts = entry[4] if len(entry) > 4 else {}
        if only and name not in only:
            continue
        cfg = VTPConfig(**kw)
        m = VTPModel(cfg).eval()
        spec = {k: list(v.shape) for k, v in m.state_dict().items()}
        sd = seeded_state_dict(spec, seed=0, **seed_opts)
        m.load_state_dict(sd)
        x = seeded_images(B, size, size)
        ids = seeded_captions(B, 77, kw["text_vocab_size"])
        out = {}
        with torch.no_grad():
            for tag, ctx in (("fp32", torch.autocast("cpu", enabled=False)),
                             ("bf16", torch.autocast("cpu", dtype=torch.bfloat16))):
                with ctx:
                    lat = m.get_reconstruction_latents(x)
                    rec = m.get_latents_decoded_images(lat)
                    out[f"latents_{tag}"] = lat.float().numpy()
                    out[f"recon_{tag}"] = rec.float().numpy()
                    fi = m.get_clip_image_feature(x)
                    out[f"img_feat_{tag}"] = fi.float().numpy()
                    feats = m.get_last_layer_feature(x)
                    out[f"cls_{tag}"] = feats["cls_token"].float().numpy()
                    if name not in ("small", "large2"):
                        out[f"patch_{tag}"] = feats["patch_tokens"].float().numpy()
                    if with_text:
                        ft = m.get_clip_text_feature(ids)
                        out[f"txt_feat_{tag}"] = ft.float().numpy()
                        lg, _ = m.get_clip_logits(x, ids)
                        out[f"logits_{tag}"] = lg.float().numpy()
        # conditioning of the REFERENCE ITSELF: response of each fp32 output to a 1e-6 relative input perturbation.
        # (bf16 roundings inside the fp32 path — RoPE, layers/attention.py:74-89 — make some outputs, notably the cls
        # token, discontinuous; an independent implementation cannot agree better than this floor.)
        with torch.no_grad():
            xp = x * (1 + 1e-6)
            relf = lambda a, b: float(((a.float() - torch.from_numpy(b)).norm() / torch.from_numpy(b).norm()))
            latp = m.get_reconstruction_latents(xp)
            sens = {"latents": relf(latp, out["latents_fp32"]),
                    "recon": relf(m.get_latents_decoded_images(latp), out["recon_fp32"]),
                    "img_feat": relf(m.get_clip_image_feature(xp), out["img_feat_fp32"]),
                    "cls": relf(m.get_last_layer_feature(xp)["cls_token"], out["cls_fp32"])}
            if with_text:
                sens["logits"] = relf(m.get_clip_logits(xp, ids)[0], out["logits_fp32"])
        out["ids"] = ids.numpy()
        out["x_checksum"] = np.array([x.double().sum().item(), x.double().abs().sum().item()])
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        with open(os.path.join(OUT, f"{name}.json"), "w") as f:
            json.dump({"config": kw, "batch": B, "image_size": size, "spec": spec,
                       "reference_commit": "5ce1eb6", "torch": torch.__version__, "seed_opts": seed_opts,
                       "ref_sensitivity_1e-6": sens}, f)
        print(name, "reference sensitivity to 1e-6 input perturbation:", sens)


if __name__ == "__main__":
    main()
