"""2-GPU NCCL tests of the data-parallel training step (skipped on a 1-GPU box): the sharded gradients of ALL THREE
objectives after the bucketed all-reduce equal the single-process global-batch gradients; parameters stay bit-identical
across ranks; the captured CUDA graph of the step (NCCL collectives inside) trains like the eager launches."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ssl_batch(Bg, HW=16, n_loc=2):
    """Global SSL batch (view-major global crops, crop-major local crops, iBOT masks on every other global crop)."""
    from oracle.seeded import seeded_images

    masks = torch.zeros(2 * Bg, HW, dtype=torch.bool)
    g = torch.Generator().manual_seed(5)
    for i in range(0, 2 * Bg, 2):
        masks[i, torch.randperm(HW, generator=g)[:5]] = True
    return dict(global_crops=seeded_images(2 * Bg, 64, 64, seed=21).cuda(), local_crops=seeded_images(n_loc * Bg, 32, 32, seed=22).cuda(),
                masks=masks, mask_indices=masks.flatten().nonzero().flatten().cuda(),
                masks_weight=(1.0 / masks.sum(-1).clamp(min=1).float())[:, None].expand_as(masks)[masks].cuda(), n_loc=n_loc)


def _ssl_slice(gb, rank, world):
    """This rank's share: images [rank*B, (rank+1)*B) of BOTH views and of every local crop."""
    Bg = gb["global_crops"].shape[0] // 2
    B = Bg // world
    sl = slice(rank * B, (rank + 1) * B)
    gc = gb["global_crops"].view(2, Bg, *gb["global_crops"].shape[1:])[:, sl].reshape(2 * B, *gb["global_crops"].shape[1:])
    lc = gb["local_crops"].view(gb["n_loc"], Bg, *gb["local_crops"].shape[1:])[:, sl].reshape(gb["n_loc"] * B, *gb["local_crops"].shape[1:])
    m = gb["masks"].view(2, Bg, -1)[:, sl].reshape(2 * B, -1)
    return dict(global_crops=gc.contiguous(), local_crops=lc.contiguous(), mask_indices=m.flatten().nonzero().flatten().cuda(),
                masks_weight=(1.0 / m.sum(-1).clamp(min=1).float())[:, None].expand_as(m)[m].cuda())


def _worker(rank, world, port, out, exchange="nccl"):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict
    from tests.util import load_golden
    from vtp_b200.config import VTPConfig
    from vtp_b200.train import TrainConfig, VTPTrainer

    meta, _ = load_golden("tiny")
    cfg = VTPConfig(**meta["config"])
    sd = seeded_state_dict(meta["spec"], seed=0)
    tc = TrainConfig(head_out_dim=512, head_hidden=256, head_bottleneck=64, n_local_crops=2, clip_exchange=exchange)
    Bg = 8
    B = Bg // world
    x = seeded_images(Bg, 64, 64).cuda()
    ids = seeded_captions(Bg, 77, 1000).cuda()
    tr = VTPTrainer(cfg, tc, device=f"cuda:{rank}")
    tr.import_state_dict(sd)
    sl = slice(rank * B, (rank + 1) * B)
    gb = _ssl_batch(Bg)
    rb = _ssl_slice(gb, rank, world)
    tr.clip_fwd_bwd(x[sl].contiguous(), ids[sl].contiguous(), 1.0)
    tr.ssl_fwd_bwd(rb["global_crops"], rb["local_crops"], rb["mask_indices"], rb["masks_weight"], 1.0)
    tr.rec_fwd_bwd(x[sl].contiguous(), 1.0, final_group=True)
    tr.allreduce_grads()
    if exchange == "p2p":
        tr.peer.check()                      # the flag barriers did not time out
    g = (tr.store.g / world).cpu()
    loss = tr.loss_acc.clone()
    dist.all_reduce(loss)
    loss = (loss / world).cpu()
    # single-process global-batch reference on the same device
    ref = VTPTrainer(cfg, tc, device=f"cuda:{rank}")
    ref.world, ref.rank = 1, 0              # (its contrastive exchange, if peer-memory, is built for world = 1 too)
    ref.import_state_dict(sd)
    ref.clip_fwd_bwd(x, ids, 1.0)
    ref.ssl_fwd_bwd(gb["global_crops"], gb["local_crops"], gb["mask_indices"], gb["masks_weight"], 1.0)
    ref.rec_fwd_bwd(x, 1.0)
    centre = (tr.center_dino - ref.center_dino).abs().max().item() / max(ref.center_dino.abs().max().item(), 1e-12)
    gr = ref.store.g.cpu()
    rel = ((g - gr).norm() / gr.norm()).item()
    # one optimiser step on both ranks: parameters must remain identical across ranks
    tr.optimizer_step()
    p = tr.store.p.clone()
    p0 = p.clone()
    dist.broadcast(p0, src=0)
    out[rank] = (rel, float((p - p0).abs().max()), loss.tolist(), ref.loss_acc.cpu().tolist(), centre)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_two_rank_step_equals_global_batch(exchange):
    """nccl: all-gather + all-reduced cross terms; p2p: peer-memory gather fused with the logits (csrc/clip.cu)."""
    import torch.multiprocessing as mp

    if exchange == "p2p" and os.environ.get("VTP_TEST_UNVALIDATED") != "1":
        pytest.skip("the peer-memory exchange has only run on one GPU so far (set VTP_TEST_UNVALIDATED=1 on a 2-GPU box)")
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, exchange), nprocs=world, join=True)
    for r in range(world):
        rel, pdiff, loss, loss_ref, centre = out[r]
        assert rel < 2e-2, rel                 # bf16 noise: different batch tiling of the same math
        assert pdiff == 0.0
        for j in (0, 1, 2, 3, 4):              # clip, dino_local, dino_global, ibot, rec: mean over ranks == global batch
            assert abs(loss[j] - loss_ref[j]) < 2e-2 * abs(loss_ref[j]), (j, loss, loss_ref)
        assert centre < 1e-3, centre           # teacher centre: all-reduced sums == global-batch mean


def _graph_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict
    from tests.util import load_golden
    from vtp_b200.config import VTPConfig
    from vtp_b200.train import TrainConfig, VTPTrainer

    meta, _ = load_golden("tiny")
    cfg = VTPConfig(**meta["config"])
    sd = seeded_state_dict(meta["spec"], seed=0)
    Bg = 8
    B = Bg // world
    sl = slice(rank * B, (rank + 1) * B)
    gb = _ssl_batch(Bg)
    batch = dict(_ssl_slice(gb, rank, world), image=seeded_images(Bg, 64, 64)[sl].cuda(), text=seeded_captions(Bg, 77, 1000)[sl].cuda(),
                 rec_image=seeded_images(Bg, 64, 64, seed=31)[sl].cuda())
    res = {}
    for mode in ("eager", "graph"):
        tr = VTPTrainer(cfg, TrainConfig(head_out_dim=512, head_hidden=256, head_bottleneck=64, n_local_crops=2), device=f"cuda:{rank}")
        tr.import_state_dict(sd)
        tr.enable_lpips(seed=0, chunk=2)
        if mode == "eager":
            for _ in range(4):
                loss = tr.train_step(batch)
        else:
            tr.capture_step(batch, warmup=2)
            for _ in range(2):
                loss = tr.replay_step()
        torch.cuda.synchronize()
        res[mode] = (tr.store.p.clone(), loss.clone())
        tr.release_graph()       # a live graph that captured NCCL collectives blocks destroy_process_group()
    p_e, p_g = res["eager"][0], res["graph"][0]
    p0 = p_g.clone()
    dist.broadcast(p0, src=0)
    out[rank] = (((p_e - p_g).norm() / p_e.norm()).item(), float((p_g - p0).abs().max()), res["eager"][1].cpu().tolist(),
                 res["graph"][1].cpu().tolist())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_graph_step_equals_eager():
    """capture_step with NCCL inside (bucketed gradient all-reduce on NCCL's stream, feature all-gather, centre sums):
    two replays after two warm-up steps == four eager steps; parameters identical across ranks."""
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_graph_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        rel, pdiff, le, lg = out[r]
        assert rel < 1e-4, rel
        assert pdiff == 0.0
        assert all(abs(a - b) <= 2e-3 * abs(a) + 1e-5 for a, b in zip(le, lg)), (le, lg)
