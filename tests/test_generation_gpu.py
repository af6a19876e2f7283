"""Consumer side of the path: uint8 image production and latent shards / statistics.
Integer outputs are bit-exact against the reference's torch expressions."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("norm", ["imagenet", "half"])
def test_image_to_u8_bit_exact(dtype, norm):
    """generation/tokenizer/vtp_tokenizer.py:106-119: Normalize(inv_mean, inv_std) -> *255 -> clamp -> uint8 -> NHWC."""
    from vtp_b200 import lib
    from vtp_b200.generation import NORMALIZE_HALF, NORMALIZE_IMAGENET

    cfg = NORMALIZE_IMAGENET if norm == "imagenet" else NORMALIZE_HALF
    inv_mean = [-m / s for m, s in zip(cfg["mean"], cfg["std"])]
    inv_std = [1.0 / s for s in cfg["std"]]
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(3, 3, 40, 64, generator=g) * 1.5).to(dtype)
    x[0, :, 0, :8] = torch.tensor([-10.0, 10.0, 0.0, float("inf"), -float("inf"), 2.64, -2.1179, 1e-3]).to(dtype)
    xc = x.cuda()
    sub = torch.tensor(inv_mean, dtype=torch.float32, device="cuda")
    div = torch.tensor(inv_std, dtype=torch.float32, device="cuda")
    out = torch.empty(3, 40, 64, 3, dtype=torch.uint8, device="cuda")
    lib.image_to_u8(xc, sub, div, out)
    # the reference expression, evaluated by torch on the same device in fp32
    t = (xc.float() - sub.view(1, 3, 1, 1)) / div.view(1, 3, 1, 1)
    ref = torch.clamp(t * 255, 0, 255).permute(0, 2, 3, 1).to(torch.uint8)
    assert torch.equal(out, ref)
    # and the evaluation program's order (tools/test_reconstruction_hf.py:371-372,401): clamp to [0,1] first
    ref2 = (torch.clamp(t, 0, 1).permute(0, 2, 3, 1) * 255.0).to(torch.uint8)
    assert torch.equal(out, ref2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_latent_stats_match_torch(dtype):
    from vtp_b200 import lib

    lat = (torch.randn(7, 64, 16, 16, generator=torch.Generator().manual_seed(1)) * 2 + 0.3).to(dtype).cuda()
    s = torch.zeros(64, dtype=torch.float64, device="cuda")
    ss = torch.zeros(64, dtype=torch.float64, device="cuda")
    lib.latent_stats(lat, s, ss)
    lib.latent_stats(lat[:3].contiguous(), s, ss)            # accumulates
    both = torch.cat([lat, lat[:3]]).double()
    assert torch.allclose(s, both.sum(dim=(0, 2, 3)), rtol=1e-12, atol=1e-9)
    assert torch.allclose(ss, (both * both).sum(dim=(0, 2, 3)), rtol=1e-12, atol=1e-9)


def test_tokenizer_roundtrip_and_shards(tmp_path):
    """VTP_Tokenizer surface (encode_images / decode_to_images) on the tiny golden model + the shard writer."""
    from safetensors import safe_open
    from tests.util import golden_inputs, load_golden
    from vtp_b200.config import VTPConfig
    from vtp_b200.generation import LatentShardWriter, VTP_Tokenizer
    from vtp_b200.model import VTPModel

    meta, g = load_golden("tiny")
    sd, x, _ = golden_inputs(meta)
    m = VTPModel(VTPConfig(**meta["config"]))
    m.load_state_dict(sd)
    tok = VTP_Tokenizer(model=m, img_size=64, normalize_type="imagenet")
    assert (tok.patch_size, tok.embed_dim, tok.latent_size) == (16, 64, 4)
    z = tok.encode_images(x)
    assert z.device.type == "cpu" and tuple(z.shape) == tuple(g["latents_fp32"].shape)
    assert ((z - g["latents_fp32"]).norm() / g["latents_fp32"].norm()).item() < 1e-3
    img = tok.decode_to_images(z)
    assert isinstance(img, np.ndarray) and img.dtype == np.uint8 and img.shape == (x.shape[0], 64, 64, 3)
    dec = m.get_latents_decoded_images(z.cuda())
    ref = torch.clamp(tok.transform_inv(dec) * 255, 0, 255).permute(0, 2, 3, 1).to("cpu", dtype=torch.uint8).numpy()
    assert np.array_equal(img, ref)
    # shards: 3 batches of B images, shard_size 2B -> one full shard + one partial, plus the statistics file
    B = x.shape[0]
    w = LatentShardWriter(str(tmp_path), rank=1, shard_size=2 * B)
    zs = []
    for i in range(3):
        xi = x.cuda() * (1.0 + 0.1 * i)
        zi, zf = tok.encode_images_device(xi), tok.encode_images_device(torch.flip(xi, dims=[3]))
        zs.append(zi.cpu())
        w.add(zi, zf, torch.arange(B) + 10 * i)
    w.close()                         # default: only rank 0 writes the statistics (multi-rank extraction shares the directory)
    assert not os.path.exists(os.path.join(tmp_path, "latents_stats.pt"))
    torch.save(w.stats(), os.path.join(tmp_path, "latents_stats.pt"))   # what close(write_stats=True) / rank 0 writes
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith(".safetensors"))
    assert files == ["latents_rank01_shard000.safetensors", "latents_rank01_shard001.safetensors"]
    with safe_open(os.path.join(tmp_path, files[0]), "pt") as f:
        assert set(f.keys()) == {"latents", "latents_flip", "labels"} and f.metadata()["total_size"] == str(2 * B)
        assert torch.equal(f.get_tensor("latents"), torch.cat(zs[:2]))
        assert torch.equal(f.get_tensor("labels"), torch.cat([torch.arange(B), torch.arange(B) + 10]))
    st = torch.load(os.path.join(tmp_path, "latents_stats.pt"))
    allz = torch.cat(zs).double()
    assert st["mean"].shape == (1, 64, 1, 1)
    assert torch.allclose(st["mean"].double(), allz.mean(dim=(0, 2, 3), keepdim=True), atol=1e-5)
    assert torch.allclose(st["std"].double(), allz.std(dim=(0, 2, 3), keepdim=True), rtol=1e-4, atol=1e-6)
