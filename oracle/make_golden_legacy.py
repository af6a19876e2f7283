"""TEST INFRASTRUCTURE ONLY — pins the LEGACY / training side of the oracle to the REAL reference (run in the dev
container, where /root/reference is importable; see oracle/ref_harness.py):

    python -m oracle.make_golden_legacy

Writes tests/golden/legacy_tiny.{npz,json} with the outputs of the reference's own modules on seeded weights/inputs:
  * `VTP.forward(forward_type="ssl")`  (vtp/models/vtp.py:365-386,410-484): teacher cls swap, iBOT masked-patch gather
    into the `upperbound` buffer, list forward [global(masked), local] of the student, the three DINO-head calls;
  * `VTP.forward(forward_type="clip")` / `"rec"` of the legacy meta-arch (vtp.py:340-363,487-512);
  * `VTP.update_teacher(momentum)`     (vtp.py:388-401): the teacher outputs after one EMA update;
  * `DINOHead.forward` stand-alone     (heads/dino_head.py:65-89), fp32 and CPU-autocast-bf16;
  * `LPIPS.forward`                    (utils/lpips.py:84-100) with the seeded-random VGG16 / lin weights of
    vtp_b200.lpips.random_weights (the constructor's weight download, lpips.py:68,78,130, is bypassed — no network).
The weights are NOT stored: the state-dict spec is, and oracle/seeded.py regenerates them bit-identically."""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

LEGACY_TINY = dict(embed_dim=128, depth=2, heads=2, text_dim=128, text_heads=2, text_layers=2, dec_dim=128, dec_depth=2,
                   dec_heads=2, image_size=64, head_out_dim=512, head_hidden=256, head_bottleneck=64, vocab=1000)
B, N_LOCAL, GLOBAL, LOCAL, MOMENTUM = 3, 2, 64, 32, 0.994


def ssl_inputs():
    """Seeded SSL batch: 2 global crops (view-major [2B]) + N_LOCAL local crops (crop-major), iBOT masks on every
    other global image (5 of 16 patches), `upperbound` larger than the number of masked patches."""
    gc = seeded_images(2 * B, GLOBAL, GLOBAL, seed=5)
    lc = seeded_images(N_LOCAL * B, LOCAL, LOCAL, seed=6)
    HW = (GLOBAL // 16) ** 2
    masks = torch.zeros(2 * B, HW, dtype=torch.bool)
    g = torch.Generator().manual_seed(7)
    for b in range(0, 2 * B, 2):
        masks[b, torch.randperm(HW, generator=g)[:5]] = True
    idx = masks.flatten().nonzero().flatten()
    return gc, lc, masks, idx


def build_lpips():
    """The reference's LPIPS module with its network-dependent constructor bypassed and seeded weights loaded."""
    from torch import nn
    from vtp.utils import lpips as L
    from vtp_b200.lpips import random_weights

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = L.LPIPS.__new__(L.LPIPS)
        nn.Module.__init__(m)
        m.scaling_layer = L.ScalingLayer()
        m.chns = [64, 128, 256, 512, 512]
        m.net = L.vgg16(pretrained=False, requires_grad=False)
        for i, c in enumerate(m.chns):
            setattr(m, f"lin{i}", L.NetLinLayer(c, use_dropout=True))
    vw, vb, lw = random_weights(0)
    convs = [mod for mod in m.net.modules() if isinstance(mod, nn.Conv2d)]
    assert len(convs) == 13
    with torch.no_grad():
        for c, w, b in zip(convs, vw, vb):
            c.weight.copy_(w), c.bias.copy_(b)
        for i, w in enumerate(lw):
            getattr(m, f"lin{i}").model[-1].weight.copy_(w)
    return m.eval()


def main():
    rh.import_reference()
    from vtp.models.heads.dino_head import DINOHead
    from vtp.models.vtp import VTP

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = VTP(vtp_config=rh.legacy_vtp_config(**LEGACY_TINY)).eval()
    spec = {k: list(v.shape) for k, v in m.state_dict().items()}
    sd = seeded_state_dict(spec, seed=0)
    m.load_state_dict(sd)
    out = {}
    gc, lc, masks, idx = ssl_inputs()
    n_m = int(idx.numel())
    ssl = dict(global_crops=gc, n_global_crops=2, mask_indices_list=idx, n_masked_patches=n_m, upperbound=n_m + 3,
               local_crops=lc, masks=masks)
    x = seeded_images(B, GLOBAL, GLOBAL)
    ids = seeded_captions(B, 77, LEGACY_TINY["vocab"])
    with torch.no_grad():
        for tag, ctx in (("fp32", torch.autocast("cpu", enabled=False)), ("bf16", torch.autocast("cpu", dtype=torch.bfloat16))):
            with ctx:
                t, s = m(ssl_dict=ssl, forward_type="ssl")
                out[f"t_cls_{tag}"] = t["teacher_cls_tokens_after_head"].float().numpy()
                out[f"t_masked_{tag}"] = t["masked_teacher_patch_tokens_after_head"].float().numpy()
                out[f"s_local_{tag}"] = s["student_local_cls_tokens_after_head"].float().numpy()
                out[f"s_global_{tag}"] = s["student_global_cls_tokens_after_head"].float().numpy()
                out[f"s_masked_{tag}"] = s["student_global_masked_patch_tokens_after_head"].float().numpy()
                c = m(image=x, text=ids, forward_type="clip")
                out[f"clip_img_{tag}"] = c["image_features"].float().numpy()
                out[f"clip_txt_{tag}"] = c["text_features"].float().numpy()
                r = m(reconstruction_image=x, forward_type="rec")
                out[f"rec_{tag}"] = r["reconstructed_image"].float().numpy()
                # stand-alone DINOHead on seeded tokens (the student's head)
                tok = torch.randn(11, LEGACY_TINY["embed_dim"], generator=torch.Generator().manual_seed(9))
                out[f"head_{tag}"] = m.dino_head(tok).float().numpy()
        # EMA: one update, then the teacher's outputs again (vtp.py:388-401 covers trunk, proj, dino_head)
        m.update_teacher(MOMENTUM)
        t2, _ = m(ssl_dict=ssl, forward_type="ssl")
        out["t_cls_after_ema_fp32"] = t2["teacher_cls_tokens_after_head"].float().numpy()
        out["teacher_qkv0_after_ema"] = m.teacher_trunk.blocks[0].attn.qkv.weight.numpy().copy()
        out["teacher_proj_after_ema"] = m.teacher_proj.weight.numpy().copy()
        out["teacher_head_v_after_ema"] = m.teacher_dino_head.last_layer.weight_v.numpy().copy()
        # LPIPS (inputs in [-1, 1]-ish), fp32 and CPU autocast
        lp = build_lpips()
        g = torch.Generator().manual_seed(3)
        a = torch.randn(2, 3, 64, 64, generator=g) * 0.5
        b = torch.randn(2, 3, 64, 64, generator=g) * 0.5
        out["lpips_fp32"] = lp(a, b).float().numpy()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out["lpips_bf16"] = lp(a, b).float().numpy()
    # a DINOHead built by its own constructor must expose the keys the oracle reads
    hk = sorted(DINOHead(8, 16, hidden_dim=8, bottleneck_dim=4).state_dict().keys())
    np.savez_compressed(os.path.join(OUT, "legacy_tiny.npz"), **out)
    with open(os.path.join(OUT, "legacy_tiny.json"), "w") as f:
        json.dump({"legacy_config": LEGACY_TINY, "batch": B, "n_local": N_LOCAL, "global": GLOBAL, "local": LOCAL,
                   "momentum": MOMENTUM, "upperbound": n_m + 3, "spec": spec, "dino_head_keys": hk,
                   "reference_commit": "5ce1eb6", "torch": torch.__version__}, f)
    print("legacy_tiny:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
