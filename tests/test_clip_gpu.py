"""Contrastive exchange kernels (csrc/clip.cu) on one GPU: the per-rank feature buffers of a multi-rank job are
emulated by separately allocated tensors ("virtual ranks"), which exercises the pointer-table gather, the ragged tile
edges and both softmax directions exactly as the multi-GPU run does; the IPC / flag-barrier plumbing is exercised with
world = 1.  The 2-GPU equivalence test lives in test_dist_gpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _features(world, B, E, seed):
    g = torch.Generator().manual_seed(seed)
    fi = [torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF).cuda() for _ in range(world)]
    ft = [torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF).cuda() for _ in range(world)]
    return fi, ft


@pytest.mark.parametrize("world,B,E,rank", [(1, 8, 64, 0), (3, 5, 72, 1), (4, 96, 384, 3), (2, 256, 768, 0)])
def test_gather_logits_softmax_grad(world, B, E, rank):
    from vtp_b200 import lib

    fi, ft = _features(world, B, E, seed=world * 100 + B)
    Bg = world * B
    Bgp = (Bg + 7) // 8 * 8
    S = torch.full((Bg, Bgp), float("nan"), device="cuda")
    St = torch.full((Bg, Bgp), float("nan"), device="cuda")
    fi_all = torch.zeros(Bgp, E, dtype=BF, device="cuda")
    ft_all = torch.zeros(Bgp, E, dtype=BF, device="cuda")
    lib.clip_gather_logits([t.data_ptr() for t in fi], [t.data_ptr() for t in ft], B, E, S, St, fi_all, ft_all)
    I, T = torch.cat(fi), torch.cat(ft)
    assert torch.equal(fi_all[:Bg], I) and torch.equal(ft_all[:Bg], T)          # the gather is a bit-exact copy
    ref = I.double() @ T.double().t()
    assert (S[:, :Bg].double() - ref).abs().max().item() < 5e-5                  # fp32 accumulation of bf16 products
    assert torch.equal(St[:, :Bg], S[:, :Bg].t())
    # ---- both softmax directions, loss, d(log_scale), gradient matrices
    ls = torch.tensor([2.3], device="cuda")
    coef = 0.5 / B
    lse = torch.empty(2, Bg, device="cuda")
    acc = torch.zeros(2, device="cuda")
    row0 = rank * B
    lib.clip_lse(S, St, Bg, row0, B, ls, coef, lse, acc[0:1], acc[1:2])
    x = (ls.double().exp() * S[:, :Bg].double())
    lse_i, lse_t = torch.logsumexp(x, 1), torch.logsumexp(x, 0)
    assert (lse[0].double() - lse_i).abs().max().item() < 1e-4 and (lse[1].double() - lse_t).abs().max().item() < 1e-4
    own = slice(row0, row0 + B)
    eye = torch.eye(Bg, dtype=torch.float64, device="cuda")
    p_row, p_col = torch.exp(x - lse_i[:, None]), torch.exp(x - lse_t[None, :])
    d = x.diagonal()
    loss = coef * ((lse_i[own] - d[own]).sum() + (lse_t[own] - d[own]).sum())
    dls = coef * (((p_row - eye)[own, :] * x[own, :]).sum() + ((p_col - eye)[:, own] * x[:, own]).sum())
    assert abs(acc[0].item() - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
    assert abs(acc[1].item() - dls.item()) < 2e-4 * max(1.0, abs(dls.item()))
    dMi = torch.full((B, Bgp), float("nan"), dtype=BF, device="cuda")
    dMt = torch.full((B, Bgp), float("nan"), dtype=BF, device="cuda")
    lib.clip_grad(S, St, Bg, row0, B, ls, coef, lse, dMi, dMt)
    dM = coef * ls.double().exp() * (p_row + p_col - 2 * eye)
    scale = dM.abs().max().item()
    assert (dMi[:, :Bg].double() - dM[own, :]).abs().max().item() < 6e-3 * scale      # bf16 output rounding
    assert (dMt[:, :Bg].double() - dM[:, own].t()).abs().max().item() < 6e-3 * scale
    if Bgp > Bg:
        assert (dMi[:, Bg:] == 0).all() and (dMt[:, Bg:] == 0).all()


def test_comm_buffer_alias_and_barrier_world1():
    from vtp_b200 import lib
    from vtp_b200.comm import PeerFeatures

    pf = PeerFeatures(6, 64, "cuda")
    assert pf.world == 1 and pf.img.shape == (6, 64) and pf.img.dtype == BF
    assert pf.img.data_ptr() == pf.img_ptrs[0] and pf.txt.data_ptr() == pf.txt_ptrs[0]
    assert len(lib.comm_get_handle(pf.base)) == 64
    pf.img.copy_(torch.ones(6, 64, dtype=BF, device="cuda"))
    pf.txt.fill_(2.0)
    for _ in range(3):
        pf.barrier()
    pf.check()
    assert pf.epoch == 3 and float(pf.img.float().sum()) == 6 * 64 and float(pf.txt.float().sum()) == 2 * 6 * 64
    pf.close()


def test_trainer_p2p_exchange_equals_nccl_formulation():
    """World = 1: the peer-memory path (full logits, fused gather) and the default path give the same loss and the same
    parameter gradients up to bf16 rounding of the gradient matrices."""
    from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict
    from tests.util import load_golden
    from vtp_b200.config import VTPConfig
    from vtp_b200.train import TrainConfig, VTPTrainer

    meta, _ = load_golden("tiny")
    cfg = VTPConfig(**meta["config"])
    sd = seeded_state_dict(meta["spec"], seed=0)
    x = seeded_images(6, 64, 64).cuda()
    ids = seeded_captions(6, 77, 1000).cuda()
    res = {}
    for mode in ("nccl", "p2p"):
        tc = TrainConfig(head_out_dim=512, head_hidden=256, head_bottleneck=64, n_local_crops=2, clip_exchange=mode)
        tr = VTPTrainer(cfg, tc)
        tr.import_state_dict(sd)
        tr.clip_fwd_bwd(x, ids, 1.0)
        if mode == "p2p":
            tr.peer.check()
        res[mode] = (tr.store.g.clone(), tr.loss_acc.clone())
    g0, l0 = res["nccl"]
    g1, l1 = res["p2p"]
    assert abs(l0[0].item() - l1[0].item()) < 1e-3 * abs(l0[0].item())
    rel = ((g0 - g1).norm() / g0.norm()).item()
    assert rel < 2e-2, rel                 # bf16 rounding of the gradient matrices (one combined vs two separate)
