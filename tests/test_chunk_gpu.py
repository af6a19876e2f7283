"""TrainConfig.ssl_chunk / rec_chunk (activation-memory control for VTP-Base/Large at 256 images per GPU): processing
the per-GPU batch in image groups gives the same losses, centre statistics and parameter gradients as one pass — the
per-token arithmetic is identical, only fp32 accumulation order (atomics, split-K) differs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(**kw):
    from oracle.seeded import seeded_state_dict
    from tests.util import load_golden
    from vtp_b200.config import VTPConfig
    from vtp_b200.train import TrainConfig, VTPTrainer

    meta, _ = load_golden("tiny")
    cfg = VTPConfig(**meta["config"])
    tc = TrainConfig(head_out_dim=512, head_hidden=256, head_bottleneck=64, n_local_crops=2, **kw)
    tr = VTPTrainer(cfg, tc)
    tr.import_state_dict(seeded_state_dict(meta["spec"], seed=0))
    return tr, cfg


def _batch(B, vocab):
    from vtp_b200.synthetic import make_batch, to_device

    return to_device(make_batch(B, image_size=64, local_size=32, n_local=2, vocab=vocab, seed=7), "cuda", non_blocking=False)


@pytest.mark.parametrize("chunk", [1, 2, 3])
def test_ssl_chunks_equal_whole_batch(chunk):
    ref, cfg = _trainer()
    b = _batch(4, cfg.text_vocab_size)
    assert b["mask_indices"].numel() > 0
    ref.ssl_fwd_bwd(b["global_crops"], b["local_crops"], b["mask_indices"], b["masks_weight"], 1.0)
    tr, _ = _trainer(ssl_chunk=chunk)
    tr.ssl_fwd_bwd(b["global_crops"], b["local_crops"], b["mask_indices"], b["masks_weight"], 1.0)
    torch.cuda.synchronize()
    l0, l1 = ref.loss_acc.cpu(), tr.loss_acc.cpu()
    assert torch.allclose(l0, l1, rtol=1e-4, atol=1e-6), (l0, l1)
    assert torch.allclose(ref.center_dino, tr.center_dino, rtol=1e-4, atol=1e-6)
    assert torch.allclose(ref.center_ibot, tr.center_ibot, rtol=1e-4, atol=1e-6)
    g0, g1 = ref.store.g, tr.store.g
    assert ((g0 - g1).norm() / g0.norm()).item() < 2e-3


def test_rec_chunks_and_full_step_equal_whole_batch():
    ref, cfg = _trainer()
    b = _batch(4, cfg.text_vocab_size)
    tr, _ = _trainer(ssl_chunk=2, rec_chunk=3)
    for t in (ref, tr):
        t.enable_lpips(seed=0, chunk=2)
    # gradients of one whole step (all three objectives), before the optimiser consumes them
    for t in (ref, tr):
        t.loss_acc.zero_()
        t.clip_fwd_bwd(b["image"], b["text"], 1.0)
        t.ssl_fwd_bwd(b["global_crops"], b["local_crops"], b["mask_indices"], b["masks_weight"], 1.0)
    ref.rec_fwd_bwd(b["rec_image"], 1.0)
    for b0 in range(0, 4, 3):
        tr.rec_fwd_bwd(b["rec_image"][b0:b0 + 3], 1.0, norm_B=4)
    torch.cuda.synchronize()
    assert torch.allclose(ref.loss_acc.cpu(), tr.loss_acc.cpu(), rtol=2e-4, atol=1e-6), (ref.loss_acc, tr.loss_acc)
    g0, g1 = ref.store.g, tr.store.g
    assert ((g0 - g1).norm() / g0.norm()).item() < 2e-3
    # and train_step drives the same chunking (parameters after one step agree)
    ref2, _ = _trainer()
    tr2, _ = _trainer(ssl_chunk=2, rec_chunk=3)
    ref2.train_step(b)
    tr2.train_step(b)
    torch.cuda.synchronize()
    p0, p1 = ref2.store.p, tr2.store.p
    assert ((p0 - p1).norm() / p0.norm()).item() < 1e-4


def test_chunked_step_is_graph_capturable():
    """ssl_chunk / rec_chunk inside capture_step: the per-group mask lists are split once per batch outside the graph
    (VTPTrainer.split_masks), so the captured step has no data-dependent shape; replay == eager steps."""
    te, cfg = _trainer(ssl_chunk=2, rec_chunk=3)
    tg, _ = _trainer(ssl_chunk=2, rec_chunk=3)
    b = _batch(4, cfg.text_vocab_size)
    b2 = _batch(4, cfg.text_vocab_size)
    b2["mask_indices"], b2["masks_weight"] = b["mask_indices"], b["masks_weight"]      # same number of masked patches
    b2["global_crops"] = b["global_crops"].flip(0).contiguous()
    seq = [b, b2, b]
    le = [te.train_step(x).cpu().clone() for x in seq]
    tg.capture_step(b, warmup=1)
    lg = [tg.replay_step(x).cpu().clone() for x in seq[1:]]
    for a, c in zip(le[1:], lg):
        assert torch.isfinite(c).all() and torch.allclose(a, c, rtol=2e-3, atol=1e-5), (a, c)
    assert ((tg.store.p - te.store.p).norm() / te.store.p.norm()).item() < 1e-4
