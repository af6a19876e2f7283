"""TEST INFRASTRUCTURE / CPU BASELINE ONLY — the 3-objective training step restated on the CPU oracle
(oracle/vtp_oracle.py forward + torch.autograd backward + torch.optim.AdamW + EMA teacher), i.e. what running the
reference's towers (vtp/models/vtp.py forward_clip / forward_ssl_learning / forward_reconstruction, update_teacher) plus
the restated losses costs on host cores.  Used by bench.py's `cpu_baseline` leg and `--impl reference`, and by tests.
Never imported by vtp_b200/."""
from __future__ import annotations

import copy
from typing import Dict

import torch

from . import vtp_oracle as vo


class OracleTrainer:
    def __init__(self, sd: Dict[str, torch.Tensor], head_sd: Dict[str, torch.Tensor], dims: dict, *, n_local: int,
                 lr=1e-4, betas=(0.9, 0.95), wd=0.05, teacher_momentum=0.994, mode="bf16", lpips=None, lpips_weight=1.0):
        """lpips: (vgg_w, vgg_b, lin_w) of the frozen perceptual network (utils/lpips.py) or None = L1 only."""
        self.dims, self.n_local, self.mode, self.mom = dims, n_local, mode, teacher_momentum
        self.lpips, self.lpips_weight = lpips, lpips_weight
        self.p = {k: (v.clone().float().requires_grad_(True) if v.is_floating_point() and "periods" not in k else v.clone())
                  for k, v in sd.items()}
        self.h = {"h." + k: v.clone().float().requires_grad_(True) for k, v in head_sd.items()}
        self.teacher = {k: v.detach().clone() for k, v in self.p.items()}
        self.teacher_h = {k: v.detach().clone() for k, v in self.h.items()}
        params = [v for v in list(self.p.values()) + list(self.h.values()) if torch.is_tensor(v) and v.requires_grad]
        self.opt = torch.optim.AdamW(params, lr=lr, betas=betas, weight_decay=wd)
        K = head_sd["last_layer.weight_v"].shape[0]
        self.center_dino, self.center_ibot = torch.zeros(K), torch.zeros(K)

    def step(self, batch) -> Dict[str, float]:
        d, m = self.dims, self.mode
        dv, hv = d["vision_depth"], d["vision_num_heads"]
        p, h = self.p, self.h
        losses = {}
        # CLIP
        fi = vo.clip_image_feature(batch["image"], p, depth=dv, heads=hv, mode=m)
        ft = vo.text_feature(batch["text"], p, layers=d["text_depth"], heads=d["text_num_heads"], mode=m)
        losses["clip"] = vo.clip_loss(fi, ft, p["logit_scale"].exp())
        # SSL
        gc, lc, mask_idx, mw = batch["global_crops"], batch["local_crops"], batch["mask_indices"], batch["masks_weight"]
        B2 = gc.shape[0]
        B = B2 // 2
        HW = (gc.shape[-1] // 16) * (gc.shape[-2] // 16)
        masks = torch.zeros(B2 * HW, dtype=torch.bool)
        masks[mask_idx] = True
        masks = masks.view(B2, HW)
        with torch.no_grad():
            t_out = vo.trunk_forward([gc], [None], self.teacher, depth=dv, heads=hv, mode=m, use_bottleneck=False)[0]
            tcls = t_out["x_norm_clstoken"]
            tcls = torch.cat([tcls[B:], tcls[:B]])
            tpatch = t_out["x_norm_patchtokens"].flatten(0, 1)[mask_idx]
            tlog = vo.dino_head(torch.cat([tcls, tpatch]), self.teacher_h, "h.", mode=m)
            tp_cls = vo.teacher_probs(tlog[:B2], self.center_dino, 0.07)
            tp_m = vo.teacher_probs(tlog[B2:], self.center_ibot, 0.07)
            self.center_dino = 0.9 * self.center_dino + 0.1 * tlog[:B2].mean(0)
            if mask_idx.numel():
                self.center_ibot = 0.9 * self.center_ibot + 0.1 * tlog[B2:].mean(0)
        sg, sl = vo.trunk_forward([gc, lc], [masks, None], p, depth=dv, heads=hv, mode=m, use_bottleneck=False)
        s_in = torch.cat([sl["x_norm_clstoken"], sg["x_norm_clstoken"], sg["x_norm_patchtokens"].flatten(0, 1)[mask_idx]])
        slog = vo.dino_head(s_in, h, "h.", mode=m)
        nl = lc.shape[0]
        terms = vo.dino_ibot_loss(slog[:nl], slog[nl:nl + B2], slog[nl + B2:], tp_cls, tp_m, mw, n_local=self.n_local,
                                  n_images=B2)
        losses.update(terms)
        # REC
        lat = vo.reconstruction_latents(batch["rec_image"], p, depth=dv, heads=hv, mode=m)
        rec = vo.decode_latents(lat, p, depth=d["decoder_depth"], heads=d["decoder_num_heads"], mode=m)
        lp = None
        if self.lpips is not None:
            lp = vo.lpips(rec, batch["rec_image"], *self.lpips, mode=m)
        losses["rec"] = vo.recon_loss(rec, batch["rec_image"], lp, self.lpips_weight)
        total = sum(losses.values())
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        self.opt.step()
        with torch.no_grad():  # EMA teacher, vtp.py:388-401
            for k, v in self.p.items():
                if torch.is_tensor(v) and v.requires_grad and (k.startswith("trunk.") or k.startswith("visual_proj")):
                    self.teacher[k].mul_(self.mom).add_(v.detach(), alpha=1 - self.mom)
            for k, v in self.h.items():
                self.teacher_h[k].mul_(self.mom).add_(v.detach(), alpha=1 - self.mom)
        return {k: float(v.detach()) for k, v in losses.items()}
