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
