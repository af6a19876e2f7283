"""GPU parity of the drop-in API (VTPModel.get_*) against golden vectors produced by the REAL reference, in both
precision modes, through the C-ABI kernels only.

Tolerances (relative L2, written here as the contract):
  fp32 mode  vs reference fp32          : max(1e-3, 3 x the REFERENCE's own response to a 1e-6 input perturbation,
                                          recorded in the golden file; the reference's fp32 path contains bf16
                                          roundings (RoPE) that make e.g. the cls token discontinuous)
  bf16 mode  vs reference bf16-autocast : 2e-2  and not worse than 2x the reference's own |bf16 - fp32| deviation
"""
import pytest
import torch

from tests.util import golden_inputs, load_golden, rel
from vtp_b200.config import VTPConfig
from vtp_b200.model import VTPModel

pytestmark = pytest.mark.gpu


def _build(name):
    meta, g = load_golden(name)
    sd, x, ids = golden_inputs(meta)
    m = VTPModel(VTPConfig(**meta["config"]))
    m.load_state_dict(sd)
    return m.cuda(), g, x.cuda(), ids.cuda(), meta


@pytest.mark.parametrize("name", ["tiny", "tiny96", "small", "large2"])
def test_fp32_mode_matches_reference_fp32(name):
    m, g, x, ids, meta = _build(name)
    lat = m.get_reconstruction_latents(x)
    assert lat.dtype == torch.float32 and tuple(lat.shape) == tuple(g["latents_fp32"].shape)
    e = {"latents": rel(lat, g["latents_fp32"])}
    rec = m.get_latents_decoded_images(lat)
    e["recon"] = rel(rec, g["recon_fp32"])
    e["recon_from_golden_latents"] = rel(m.get_latents_decoded_images(g["latents_fp32"].cuda()), g["recon_fp32"])
    e["img_feat"] = rel(m.get_clip_image_feature(x), g["img_feat_fp32"])
    f = m.get_last_layer_feature(x)
    e["cls"] = rel(f["cls_token"], g["cls_fp32"])
    if "patch_fp32" in g:
        e["patch"] = rel(f["patch_tokens"], g["patch_fp32"])
    if "txt_feat_fp32" in g:
        e["txt_feat"] = rel(m.get_clip_text_feature(ids), g["txt_feat_fp32"])
        lg, lgt = m.get_clip_logits(x, ids)
        e["logits"] = rel(lg, g["logits_fp32"])
        assert torch.equal(lg.T, lgt)
    sens = meta["ref_sensitivity_1e-6"]
    floor = {"latents": sens["latents"], "recon": sens["recon"], "recon_from_golden_latents": sens["recon"],
             "img_feat": sens["img_feat"], "cls": sens["cls"], "patch": sens["latents"], "txt_feat": 0.0,
             "logits": sens.get("logits", 0.0)}
    print(name, "fp32-mode rel errors:", {k: f"{v:.2e}" for k, v in e.items()}, "ref floor:", sens)
    for k, v in e.items():
        tol = max(1e-3, 3 * floor[k])
        if k == "logits":
            # bilinear in two features that are each held to the tolerance above; with B*B = 4 entries and a
            # cancellation-prone dot product the relative error of the matrix can reach the SUM of the feature errors
            # (large2 on hardware: img_feat 6.2e-4, txt_feat 1.4e-5 -> logits 1.12e-3)
            tol = max(tol, 2 * (e["img_feat"] + e["txt_feat"]))
        assert v < tol, (k, v, floor[k])


@pytest.mark.parametrize("name", ["tiny", "tiny96", "small", "large2"])
def test_bf16_mode_matches_reference_autocast(name):
    m, g, x, ids, meta = _build(name)
    e, dev = {}, {}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lat = m.get_reconstruction_latents(x)
        assert lat.dtype == torch.bfloat16
        rec = m.get_latents_decoded_images(g["latents_bf16"].cuda().to(torch.bfloat16))
        fi = m.get_clip_image_feature(x)
        cls = m.get_last_layer_feature(x)["cls_token"]
        if "txt_feat_bf16" in g:
            ft = m.get_clip_text_feature(ids)
    for key, val in (("latents", lat), ("recon", rec), ("img_feat", fi), ("cls", cls)):
        e[key] = rel(val, g[f"{key}_bf16"])
        dev[key] = (rel(val, g[f"{key}_fp32"]), rel(g[f"{key}_bf16"], g[f"{key}_fp32"]))
    if "txt_feat_bf16" in g:
        e["txt_feat"] = rel(ft, g["txt_feat_bf16"])
        dev["txt_feat"] = (rel(ft, g["txt_feat_fp32"]), rel(g["txt_feat_bf16"], g["txt_feat_fp32"]))
    print(name, "bf16-mode rel vs ref-autocast:", {k: f"{v:.2e}" for k, v in e.items()})
    print(name, "  (ours vs fp32, ref-bf16 vs fp32):", {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in dev.items()})
    # Derived bounds, per output, with the REFERENCE's own bf16 sensitivity d = |ref autocast-bf16 - ref fp32|
    # as the yardstick (7e-3 at tiny ... 2.4e-2 for the chaotic cls token at depth 12):
    #   (i)  our distance to the reference's bf16 output <= 1.25 d  (two independent bf16 evaluations cannot be closer to
    #        each other than each is to fp32; measured 0.27-1.03 d), capped at 2e-2 absolute;
    #   (ii) our distance to the reference's FP32 output <= 1.15 d + 2e-4  (measured 0.97-1.03 d: our bf16 path is exactly
    #        as far from fp32 as the reference's own autocast path).
    for k, (ours, theirs) in dev.items():
        assert e[k] < min(2e-2, 1.25 * theirs + 2e-4), (k, e[k], theirs)
        if k == "recon":
            continue  # decoded from the reference's bf16 latents: compared with recon_bf16 above only
        assert ours < 1.15 * theirs + 2e-4, (k, ours, theirs)


def test_text_argmax_pool_index_is_exact():
    m, g, x, ids, meta = _build("tiny")
    assert torch.equal(ids.cpu(), g["ids"])
    assert torch.equal(ids.argmax(-1).cpu(), g["ids"].argmax(-1))


def test_errors_mirror_reference():
    from vtp_b200.config import preset

    m = VTPModel(preset("tiny", train_reconstruction=False, train_clip=False)).cuda()
    with pytest.raises(RuntimeError):
        m.get_latents_decoded_images(torch.zeros(1, 64, 4, 4, device="cuda"))
    with pytest.raises(RuntimeError):
        m.get_clip_image_feature(torch.zeros(1, 3, 64, 64, device="cuda"))
    with pytest.raises(ValueError):
        m.forward(torch.zeros(1, 3, 64, 64, device="cuda"), forward_type="bogus")


def test_split_and_fused_epilogues_agree():
    """bf16 mode: stand-alone RoPE / SwiGLU kernels (default) vs the fused GEMM epilogues — same rounding sequence."""
    from vtp_b200 import engine

    m, g, x, ids, meta = _build("tiny")
    outs = {}
    for flag in (True, False):
        engine.SPLIT_EPILOGUES = flag
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                lat = m.get_reconstruction_latents(x)
                outs[flag] = (lat.float().clone(), m.get_latents_decoded_images(lat).float().clone())
        finally:
            engine.SPLIT_EPILOGUES = True
    assert rel(outs[False][0], outs[True][0]) < 2e-3 and rel(outs[False][1], outs[True][1]) < 2e-3


def test_intermediate_layer_features_match_reference():
    """Linear-probe feature path (tools/test_linear_probing_hf.py:109-152): tests/golden/tiny_layers.npz comes from the
    real reference (oracle/make_golden_layers.py).  First hardware run: round 2 (green)."""
    import os

    import numpy as np

    m, _, x, _, meta = _build("tiny")
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_layers.npz")).items()}
    tol = max(1e-3, 3 * meta["ref_sensitivity_1e-6"]["cls"])
    a = m.get_intermediate_layers_feature(x, n=2, return_class_token=True, norm=True)
    assert len(a) == 2
    for i, (patch, cls) in enumerate(a):
        assert patch.dtype == torch.float32 and tuple(patch.shape) == tuple(g[f"last2_patch{i}"].shape)
        assert rel(patch, g[f"last2_patch{i}"]) < tol and rel(cls, g[f"last2_cls{i}"]) < tol
    (raw,) = m.get_intermediate_layers_feature(x, n=[0], reshape=True, norm=False)
    assert tuple(raw.shape) == tuple(g["block0_raw_nchw"].shape) and rel(raw, g["block0_raw_nchw"]) < tol
    b = m.get_intermediate_layers_feature(x, n=[1, 0], norm=True)       # ascending block order, like the reference
    assert rel(b[0], g["order_patch0"]) < tol and rel(b[1], g["order_patch1"]) < tol
    with pytest.raises(AssertionError):
        m.get_intermediate_layers_feature(x, n=[0, 0])
