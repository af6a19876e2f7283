"""HBM budget model of the training step (vtp_b200/memory.py) — host logic, no GPU."""
from vtp_b200 import memory as m
from vtp_b200.config import preset


def test_small_matches_the_measured_peak():
    """profiles/bench_n2_r1.log: VTP-Small, 256 images/GPU, K = 65536 -> 41.7 GiB peak (torch.cuda.max_memory_allocated)."""
    est = m.train_step_bytes(preset("small"), 256)
    assert abs(est["peak"] / m.GIB - 41.7) < 0.15 * 41.7
    assert est["ssl"] > est["rec"] > est["clip"]
    assert abs(est["params"] / 1e6 - 106.2) < 1.0          # 83.8 M model + ~22.3 M DINO head


def test_large_needs_chunks_and_chunks_fit():
    cfg = preset("large")
    est = m.train_step_bytes(cfg, 256)
    assert est["peak"] > 180e9                               # config 4 does not fit a B200 in one piece
    ssl_chunk, rec_chunk = m.suggest_chunks(cfg, 256, budget_bytes=150 * m.GIB)
    assert 0 < ssl_chunk < 256
    fit = m.train_step_bytes(cfg, 256, ssl_chunk=ssl_chunk, rec_chunk=rec_chunk)
    assert fit["peak"] <= 150 * m.GIB
    assert m.suggest_chunks(preset("small"), 256) == (0, 0) and m.suggest_chunks(preset("base"), 256) == (0, 0)


def test_tape_bytes_formula():
    # trunk block, fp32 stream, SwiGLU: 20 D + 6 Hs (+ lse / rstd floats)
    D, Hs, H = 384, 1024, 6
    assert m.block_tape_bytes(D, Hs, "swiglu", False, H) == 20 * D + 6 * Hs + 4 * H + 8
    assert m.block_tape_bytes(D, Hs, "swiglu", True, H) == 16 * D + 6 * Hs + 4 * H + 8
    assert m.block_tape_bytes(D, 4 * D, "gelu", False, H) == 20 * D + 4 * 4 * D + 4 * H + 8


def test_suggest_chunks_rejects_what_cannot_fit():
    import pytest

    with pytest.raises(ValueError):
        m.suggest_chunks(preset("large"), 256, budget_bytes=40 * m.GIB)     # the contrastive pass alone is larger
    # a tighter but feasible budget only makes the groups smaller
    a = m.suggest_chunks(preset("large"), 256, budget_bytes=150 * m.GIB)
    b = m.suggest_chunks(preset("large"), 256, budget_bytes=110 * m.GIB)
    assert 0 < b[0] <= a[0]


def test_latent_statistics_formula():
    """LatentShardWriter.stats(): mean / unbiased std from fp64 sums (host arithmetic only)."""
    import torch

    from vtp_b200.generation import LatentShardWriter

    z = torch.randn(5, 4, 3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 2 + 1
    w = LatentShardWriter.__new__(LatentShardWriter)
    w._sum, w._sumsq, w._n = z.sum(dim=(0, 2, 3)), (z * z).sum(dim=(0, 2, 3)), z.numel() // z.shape[1]
    st = w.stats()
    assert st["mean"].shape == (1, 4, 1, 1) and st["std"].dtype == torch.float32
    assert torch.allclose(st["mean"].double(), z.mean(dim=(0, 2, 3), keepdim=True), atol=1e-6)
    assert torch.allclose(st["std"].double(), z.std(dim=(0, 2, 3), keepdim=True), atol=1e-6)


def test_gradient_buckets_partition_the_flat_buffer():
    """train.grad_buckets: the per-tower all-reduce ranges cover [0, n) exactly once, and every tensor lies inside the
    bucket of its tower (text + clip projection + logit scale | DINO head | pixel decoder | trunk)."""
    from vtp_b200.train import ParamStore, _vit_specs, grad_buckets

    st = ParamStore("cpu")
    st.add("trunk.patch.w", (128, 768), True, True); st.add("trunk.cls", (128,), False, True)
    _vit_specs(st, "trunk.", 128, 2, 344, False, True, 688)
    st.add("visual_proj.w", (128, 128), True, True)
    st.add("head.mlp0.w", (256, 128), True, True); st.add("head.last_g", (512,), False, True)
    st.add("decoder.proj_in.w", (128, 64), True)
    _vit_specs(st, "decoder.", 128, 2, 344, True, False, 688)
    st.add("text.tok_emb", (1000, 128), True); st.add("text.pos", (77, 128), False)
    _vit_specs(st, "text.", 128, 2, 512, True, False, 512)
    st.add("logit_scale", (1,), False)
    st.finalize()
    b = grad_buckets(st.offset, st.n)
    ranges = sorted(r for rs in b.values() for r in rs)
    assert ranges[0][0] == 0 and ranges[-1][1] == st.n
    assert all(a[1] == c[0] for a, c in zip(ranges, ranges[1:]))          # no gap, no overlap
    inside = lambda name, which: any(lo <= st.offset[name] < hi for lo, hi in b[which])
    assert inside("text.blocks.1.fc1.w", "text") and inside("logit_scale", "text") and inside("visual_proj.w", "text")
    assert inside("head.last_g", "head") and inside("decoder.blocks.0.qkv.b", "decoder") and inside("trunk.cls", "trunk")
    assert sum(len(v) for v in b.values()) <= 16                          # a handful of NCCL calls per step
