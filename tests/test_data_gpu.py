"""(f)4: device-side input pipeline of the training step (vtp_b200/data.py, csrc/data.cu): crop + bilinear resize + flip +
normalise against torch's own F.interpolate, iBOT masks, crop-box sampling, and the batch layout train_step expects."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_crop_resize_norm_matches_interpolate():
    from vtp_b200 import lib
    from vtp_b200.data import IMAGENET_MEAN, IMAGENET_STD

    g = torch.Generator().manual_seed(0)
    B, H, W = 3, 120, 200
    src = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    boxes = torch.tensor([[0, 0, 200, 120], [10, 20, 64, 64], [37, 5, 150, 90], [100, 60, 17, 33], [0, 0, 1, 1]], dtype=torch.float32)
    idx = torch.tensor([0, 1, 2, 1, 0], dtype=torch.int32)
    flips = torch.tensor([0, 1, 0, 1, 0], dtype=torch.uint8)
    for S in (96, 256):
        out = torch.empty(5, 3, S, S, device="cuda")
        lib.crop_resize_norm(src.cuda(), idx.cuda(), boxes.cuda(), flips.cuda(), out, mean=IMAGENET_MEAN, std=IMAGENET_STD)
        torch.cuda.synchronize()
        mean = torch.tensor(IMAGENET_MEAN)[None, :, None, None]
        std = torch.tensor(IMAGENET_STD)[None, :, None, None]
        for n in range(5):
            x0, y0, w, h = [int(v) for v in boxes[n]]
            crop = src[idx[n], y0:y0 + h, x0:x0 + w].permute(2, 0, 1)[None].float()
            ref = F.interpolate(crop, size=(S, S), mode="bilinear", align_corners=False, antialias=False)
            if flips[n]:
                ref = ref.flip(-1)
            ref = (ref / 255.0 - mean) / std
            assert (out[n].cpu() - ref[0]).abs().max().item() < 2e-4, (S, n)


def test_ibot_masks_and_boxes():
    from vtp_b200.data import ibot_masks, random_resized_crop_boxes

    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    idx, w = ibot_masks(64, 256, 0.3, 0.5, "cuda", gen)
    assert idx.dtype == torch.int64 and idx.numel() == 32 * 77 and w.numel() == idx.numel()
    assert bool((idx[1:] > idx[:-1]).all())                       # ascending, no duplicates
    per_img = torch.bincount(idx // 256, minlength=64)
    assert set(per_img.tolist()) == {0, 77} and int((per_img > 0).sum()) == 32
    assert torch.allclose(w, torch.full_like(w, 1 / 77))
    rng = np.random.default_rng(0)
    for scale in ((0.32, 1.0), (0.05, 0.32)):
        b = random_resized_crop_boxes(rng, 4000, 300, 400, scale)
        assert (b[:, 0] >= 0).all() and (b[:, 1] >= 0).all() and (b[:, 0] + b[:, 2] <= 400).all() and (b[:, 1] + b[:, 3] <= 300).all()
        area = b[:, 2] * b[:, 3] / (300 * 400)
        assert scale[0] * 0.9 <= area.min() and area.max() <= scale[1] * 1.05
        ar = b[:, 2] / b[:, 3]
        assert 0.70 < ar.min() and ar.max() < 1.40


def test_pipeline_feeds_the_training_step():
    from oracle.seeded import seeded_state_dict
    from tests.util import load_golden
    from vtp_b200.config import VTPConfig
    from vtp_b200.data import TrainBatchPipeline
    from vtp_b200.train import TrainConfig, VTPTrainer

    meta, _ = load_golden("tiny")
    cfg = VTPConfig(**meta["config"])
    tr = VTPTrainer(cfg, TrainConfig(head_out_dim=512, head_hidden=256, head_bottleneck=64, n_local_crops=2))
    tr.import_state_dict(seeded_state_dict(meta["spec"], seed=0))
    vocab = cfg.text_vocab_size

    def toy_tokenizer(texts):                        # stands in for the reference's BPE (vtp/tokenizers): str -> ids [B, 77]
        ids = torch.zeros(len(texts), 77, dtype=torch.long)
        for i, t in enumerate(texts):
            body = [1 + (ord(ch) % (vocab - 3)) for ch in t][:60]
            ids[i, 0] = vocab - 2
            ids[i, 1:1 + len(body)] = torch.tensor(body)
            ids[i, 1 + len(body)] = vocab - 1
        return ids

    pipe = TrainBatchPipeline("cuda", image_size=64, local_size=32, n_local=2, tokenizer=toy_tokenizer, seed=3)
    g = torch.Generator().manual_seed(0)
    B = 4
    imgs = torch.randint(0, 256, (B, 80, 100, 3), generator=g, dtype=torch.uint8).pin_memory()
    caps = [f"a photo number {i} of something" for i in range(B)]
    pipe.submit(imgs, caps)
    losses = []
    for _ in range(3):
        batch = pipe.get()
        pipe.submit(imgs, caps)
        assert batch["image"].shape == (B, 3, 64, 64) and batch["global_crops"].shape == (2 * B, 3, 64, 64)
        assert batch["local_crops"].shape == (2 * B, 3, 32, 32) and batch["rec_image"].shape == (B, 3, 64, 64)
        assert batch["text"].shape == (B, 77) and batch["text"].dtype == torch.int64
        assert batch["mask_indices"].numel() == B * 5 and float(batch["image"].abs().max()) < 4.0
        losses.append(tr.train_step(batch).cpu().clone())
    pipe.close()
    assert all(torch.isfinite(l).all() for l in losses)
    # the reconstruction view is the whole image, un-flipped: same pixels every step -> its loss goes down
    assert losses[-1][4] < losses[0][4]
