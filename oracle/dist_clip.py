"""TEST INFRASTRUCTURE ONLY — CPU mirror (plain torch ops) of the trainer's data-parallel contrastive step
(vtp_b200/train.py::clip_fwd_bwd): all-gather of normalised features (collective C2), local-rows x global-columns
logits, softmax-CE with labels offset by rank, feature gradients = local term + all-reduced cross term (local slice).
Used by the world_size-2 gloo test to prove the sharded formulation equals the single-process global-batch gradient."""
from __future__ import annotations

import torch
import torch.distributed as dist


def sharded_clip_feature_grads(fi: torch.Tensor, ft: torch.Tensor, log_scale: torch.Tensor, rank: int, world: int):
    """fi, ft: this rank's L2-normalised features [B, E].  Returns (loss_local, dfi, dft, dlog_scale) where the
    gradients are those of  L = mean_r L_local(r)  AFTER the (mean) parameter all-reduce, i.e. d(sum_r L_local)/d f_local
    here and 1/world applied by the optimiser — exactly the trainer's convention."""
    B, E = fi.shape
    fi_all = [torch.empty_like(fi) for _ in range(world)]
    ft_all = [torch.empty_like(ft) for _ in range(world)]
    dist.all_gather(fi_all, fi)
    dist.all_gather(ft_all, ft)
    fi_all, ft_all = torch.cat(fi_all), torch.cat(ft_all)
    s = log_scale.exp()
    sim_i, sim_t = fi @ ft_all.t(), ft @ fi_all.t()
    labels = rank * B + torch.arange(B)
    coef = 0.5 / B

    def ce(sim):
        x = s * sim
        lse = torch.logsumexp(x, dim=1)
        loss = coef * (lse - x[torch.arange(B), labels]).sum()
        g = coef * (torch.softmax(x, dim=1) - torch.nn.functional.one_hot(labels, x.shape[1]).to(x.dtype))
        return loss, s * g, (g * x).sum()

    li, Gi, dsi = ce(sim_i)
    lt, Gt, dst = ce(sim_t)
    dfi = Gi @ ft_all
    dft = Gt @ fi_all
    cross_i = Gt.t() @ ft          # [B_g, E]: contribution of MY text rows to every image feature
    cross_t = Gi.t() @ fi
    dist.all_reduce(cross_i)
    dist.all_reduce(cross_t)
    dfi = dfi + cross_i[rank * B:(rank + 1) * B]
    dft = dft + cross_t[rank * B:(rank + 1) * B]
    return li + lt, dfi, dft, dsi + dst


def replicated_clip_feature_grads(fi: torch.Tensor, ft: torch.Tensor, log_scale: torch.Tensor, rank: int, world: int):
    """CPU mirror of the peer-memory formulation (vtp_b200/csrc/clip.cu, train.py::_clip_loss_p2p): ONE forward gather,
    the full Bg x Bg similarity matrix on every rank, both softmax directions over all rows, and the feature gradients
    of the rank's own rows from  dM = coef·e^s·(softmax_row + softmax_col − 2·I)  — no backward collective.
    Same return convention as `sharded_clip_feature_grads`."""
    B, E = fi.shape
    fi_all = [torch.empty_like(fi) for _ in range(world)]
    ft_all = [torch.empty_like(ft) for _ in range(world)]
    dist.all_gather(fi_all, fi)
    dist.all_gather(ft_all, ft)
    fi_all, ft_all = torch.cat(fi_all), torch.cat(ft_all)
    Bg = world * B
    s = log_scale.exp()
    S = fi_all @ ft_all.t()                     # [Bg images, Bg captions]
    x = s * S
    lse_i = torch.logsumexp(x, dim=1)           # image -> text direction, one per image row
    lse_t = torch.logsumexp(x, dim=0)           # text -> image direction, one per caption
    coef = 0.5 / B
    own = slice(rank * B, (rank + 1) * B)
    eye = torch.eye(Bg, dtype=x.dtype)
    p_row = torch.exp(x - lse_i[:, None])       # softmax over captions, per image
    p_col = torch.exp(x - lse_t[None, :])       # softmax over images, per caption
    dM = coef * s * (p_row + p_col - 2 * eye)   # d(Σ_ranks L_local)/dS
    dfi = dM[own, :] @ ft_all
    dft = dM[:, own].t() @ fi_all
    diag = x.diagonal()
    loss = coef * ((lse_i[own] - diag[own]).sum() + (lse_t[own] - diag[own]).sum())
    g_row = coef * (p_row - eye)
    g_col = coef * (p_col - eye)
    dls = (g_row[own, :] * x[own, :]).sum() + (g_col[:, own] * x[:, own]).sum()
    return loss, dfi, dft, dls
