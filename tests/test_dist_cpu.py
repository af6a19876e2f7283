"""CPU, world_size 2 over gloo: the data-parallel formulation used by VTPTrainer (feature all-gather + local-row
logits + cross-term all-reduce; flat-buffer gradient all-reduce averaged in the optimiser) equals the single-process
global-batch computation."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, formulation="sharded"):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dist_clip

    sharded_clip_feature_grads = getattr(dist_clip, f"{formulation}_clip_feature_grads")

    torch.manual_seed(0)
    Bg, E = 8, 16
    B = Bg // world
    fi_g = torch.nn.functional.normalize(torch.randn(Bg, E, dtype=torch.float64), dim=-1)
    ft_g = torch.nn.functional.normalize(torch.randn(Bg, E, dtype=torch.float64), dim=-1)
    ls = torch.tensor(2.0, dtype=torch.float64)
    loss, dfi, dft, dls = sharded_clip_feature_grads(fi_g[rank * B:(rank + 1) * B], ft_g[rank * B:(rank + 1) * B], ls,
                                                     rank, world)
    # parameter-gradient convention: sum over ranks then * 1/world in the optimiser (flat all-reduce)
    flat = torch.cat([dls.reshape(1), loss.reshape(1)])
    dist.all_reduce(flat)
    flat /= world
    out[rank] = (dfi / world, dft / world, flat[0].item(), flat[1].item())
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("formulation", ["sharded", "replicated"])
def test_sharded_clip_equals_global_batch(formulation):
    """sharded = NCCL path (local-row logits + all-reduced cross terms); replicated = peer-memory path (full logits on
    every rank, forward gather only — csrc/clip.cu)."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, formulation), nprocs=world, join=True)
    from oracle import vtp_oracle as vo

    torch.manual_seed(0)
    Bg, E = 8, 16
    fi = torch.nn.functional.normalize(torch.randn(Bg, E, dtype=torch.float64), dim=-1).requires_grad_(True)
    ft = torch.nn.functional.normalize(torch.randn(Bg, E, dtype=torch.float64), dim=-1).requires_grad_(True)
    ls = torch.tensor(2.0, dtype=torch.float64, requires_grad=True)
    loss = vo.clip_loss(fi, ft, ls.exp())     # single process, global batch (OpenCLIP ClipLoss)
    loss.backward()
    B = Bg // world
    for r in range(world):
        dfi, dft, dls, l = out[r]
        assert torch.allclose(dfi, fi.grad[r * B:(r + 1) * B], atol=1e-12)
        assert torch.allclose(dft, ft.grad[r * B:(r + 1) * B], atol=1e-12)
        assert abs(dls - ls.grad.item()) < 1e-12 and abs(l - loss.item()) < 1e-12


def test_param_store_regions_cover_everything():
    """flat buffer: every tensor 128B-aligned, regions disjoint, teacher regions first (EMA slices line up)."""
    from vtp_b200.train import ParamStore

    st = ParamStore.__new__(ParamStore)
    st.device, st.specs, st.offset, st.shape, st.regions = "cpu", [], {}, {}, []
    st.add("a.w", (5, 7), True, True), st.add("a.b", (7,), False, True), st.add("c.w", (3, 3), True, False)
    st.add("c.b", (3,), False, False), st.add("d.w", (64,), True, True)
    st.finalize()
    offs = sorted((st.offset[n], n) for n in st.offset)
    assert all(o % 64 == 0 for o, _ in offs)
    ends = [e for _, e, _, _ in st.regions]
    starts = [s for s, _, _, _ in st.regions]
    assert starts == [0] + ends[:-1] and ends[-1] == st.n
    assert [t for *_, t in st.regions] == sorted([t for *_, t in st.regions], reverse=True)
    assert st.n_teacher == max(e for s, e, d, t in st.regions if t)
    assert st.f32("a.w").shape == (5, 7) and st.grad("c.b").shape == (3,)
