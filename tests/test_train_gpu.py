"""GPU parity of the hand-written 3-objective training step (vtp_b200/train.py) against autograd on the CPU oracle
(oracle/vtp_oracle.py, bf16-emulation mode) with identical seeded weights and inputs, objective by objective.

Loss VALUES are compared at 2e-2 relative, parameter GRADIENTS by relative L2 per tensor at 6e-2 (bf16 GEMM operands on
both sides, different accumulation order).  The loss definitions themselves are restated (reference ships none)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import vtp_oracle as vo
from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict
from tests.util import load_golden, rel
from vtp_b200.config import VTPConfig
from vtp_b200.engine import interleave8
from vtp_b200.train import TrainConfig, VTPTrainer

pytestmark = pytest.mark.gpu

TOL_G, TOL_L = 6e-2, 2e-2
K, HH, HB = 512, 256, 64

# Geometries the gradient-parity cases run at.  "tiny" is the fast default; "small2" / "large2" are the BENCHED shapes at
# depth 2: VTP-Small (D = 384, 6 heads, T = 257 in-step attention backward, K = 65 536 smem-resident loss rows, 96-pixel local
# crops = packed T = 37 tiles) and VTP-Large (D = 1024, 16 heads, SwiGLU hidden 2736 = ragged N / K tiles, text tower 768 / 12).
GEOS = {
    "tiny": dict(golden="tiny", img=64, local=32, K=512, HH=256, HB=64, heads=2, theads=2, dheads=2),
    "small2": dict(cfg=dict(vision_embed_dim=384, vision_depth=2, vision_num_heads=6, text_embed_dim=384, text_num_heads=6,
                            text_depth=2, decoder_embed_dim=384, decoder_num_heads=6, decoder_depth=2, text_vocab_size=2048),
                   img=256, local=96, K=65536, HH=2048, HB=256, heads=6, theads=6, dheads=6),
    "large2": dict(golden="large2", img=256, local=96, K=65536, HH=2048, HB=256, heads=16, theads=12, dheads=16),
}


class _Ctx:
    pass


def _setup(geo="tiny"):
    g = GEOS[geo]
    c = _Ctx()
    c.geo, c.g = geo, g
    seed_opts = {}
    if "golden" in g:
        meta, _ = load_golden(g["golden"])
        cfg = VTPConfig(**meta["config"])
        spec, seed_opts = meta["spec"], meta.get("seed_opts", {})
    else:
        from vtp_b200.model import VTPModel
        cfg = VTPConfig(**g["cfg"])
        spec = {k: list(v.shape) for k, v in VTPModel(cfg).state_dict().items()}
    sd = seeded_state_dict(spec, seed=0, **seed_opts)
    D = cfg.vision_embed_dim
    Kp, hh, hb = g["K"], g["HH"], g["HB"]
    head_spec = {"mlp.0.weight": [hh, D], "mlp.0.bias": [hh], "mlp.2.weight": [hh, hh], "mlp.2.bias": [hh],
                 "mlp.4.weight": [hb, hh], "mlp.4.bias": [hb], "last_layer.weight_g": [Kp, 1], "last_layer.weight_v": [Kp, hb]}
    hsd = seeded_state_dict(head_spec, seed=3)
    hsd["last_layer.weight_v"] = torch.randn(Kp, hb, generator=torch.Generator().manual_seed(9)) * 0.5
    tc = TrainConfig(head_out_dim=Kp, head_hidden=hh, head_bottleneck=hb, n_local_crops=2)
    tr = VTPTrainer(cfg, tc)
    tr.import_state_dict(sd, hsd)
    c.cfg, c.sd, c.hsd, c.tr = cfg, sd, hsd, tr
    c.img, c.local, c.K = g["img"], g["local"], Kp
    c.HW = (g["img"] // 16) ** 2
    c.heads, c.theads, c.dheads = g["heads"], g["theads"], g["dheads"]
    c.vocab = cfg.text_vocab_size
    return c


def _setup_tiny():
    c = _setup("tiny")
    return c.cfg, c.sd, c.hsd, c.tr


def _leafs(sd):
    return {k: (v.clone().float().requires_grad_(True) if v.is_floating_point() and "periods" not in k else v) for k, v in sd.items()}


FLOORS = {}      # name -> (error, tolerance used) of the current test, for the printed summary


def _check(tr, name, ref, ref32=None):
    """Gradient parity bound.  Both sides compute in bf16 with different accumulation orders, so the yardstick is the
    reference algorithm's OWN sensitivity to that precision: ref32 = the same gradient from the oracle in fp32 mode;
    tolerance = max(4 %, 1.5 x |oracle bf16 - oracle fp32|) for this tensor (measured: ~3.5 % at the tiny geometry, ~10 %
    at VTP-Large width / depth 2 — two bf16 implementations cannot agree better than each agrees with fp32).  Without a
    noise floor the flat 6 % bound of round 1 applies."""
    got = tr.store.grad(name).float().cpu()
    e = rel(got, ref.reshape(got.shape))
    tol = TOL_G if ref32 is None else max(4e-2, 1.5 * rel(ref, ref32))
    FLOORS[name] = (e, tol)
    assert e < tol, (name, e, tol)
    return e


def _vit_checks(tr, p, pre_ref, pre, blocks, ln, p32=None):
    g32 = (lambda k: p32[k].grad) if p32 is not None else (lambda k: None)
    errs = {}
    for i in blocks:
        r, q = f"{pre_ref}blocks.{i}.", f"{pre}blocks.{i}."
        errs[q + "qkv.w"] = _check(tr, q + "qkv.w", p[r + "attn.qkv.weight"].grad, g32(r + "attn.qkv.weight"))
        errs[q + "qkv.b"] = _check(tr, q + "qkv.b", p[r + "attn.qkv.bias"].grad, g32(r + "attn.qkv.bias"))
        errs[q + "proj.w"] = _check(tr, q + "proj.w", p[r + "attn.proj.weight"].grad, g32(r + "attn.proj.weight"))
        errs[q + "fc1.w"] = _check(tr, q + "fc1.w", interleave8(p[r + "mlp.w1.weight"].grad, p[r + "mlp.w2.weight"].grad),
                                  None if p32 is None else interleave8(p32[r + "mlp.w1.weight"].grad, p32[r + "mlp.w2.weight"].grad))
        errs[q + "fc1.b"] = _check(tr, q + "fc1.b", interleave8(p[r + "mlp.w1.bias"].grad, p[r + "mlp.w2.bias"].grad),
                                  None if p32 is None else interleave8(p32[r + "mlp.w1.bias"].grad, p32[r + "mlp.w2.bias"].grad))
        errs[q + "fc2.w"] = _check(tr, q + "fc2.w", p[r + "mlp.w3.weight"].grad, g32(r + "mlp.w3.weight"))
        errs[q + "fc2.b"] = _check(tr, q + "fc2.b", p[r + "mlp.w3.bias"].grad, g32(r + "mlp.w3.bias"))
        errs[q + "n1_w"] = _check(tr, q + "n1_w", p[r + "norm1.weight"].grad, g32(r + "norm1.weight"))
        errs[q + "n2_w"] = _check(tr, q + "n2_w", p[r + "norm2.weight"].grad, g32(r + "norm2.weight"))
        if ln:
            errs[q + "n1_b"] = _check(tr, q + "n1_b", p[r + "norm1.bias"].grad, g32(r + "norm1.bias"))
    return errs


def _summary(geo, what):
    worst = max(FLOORS.items(), key=lambda kv: kv[1][0] / kv[1][1])
    print(f"[{geo}] {what}: {len(FLOORS)} gradient tensors, max rel error {max(e for e, _ in FLOORS.values()):.3f}, median tolerance "
          f"{sorted(t for _, t in FLOORS.values())[len(FLOORS) // 2]:.3f}, tightest = {worst[0]} ({worst[1][0]:.3f} of {worst[1][1]:.3f})")


GEO_PARAMS = ["tiny", "small2", "large2"]


@pytest.mark.parametrize("geo", GEO_PARAMS)
def test_rec_objective_gradients(geo):
    c = _setup(geo)
    tr = c.tr
    x = seeded_images(3 if geo == "tiny" else 2, c.img, c.img)
    def oracle(mode):
        q = _leafs(c.sd)
        lat = vo.reconstruction_latents(x, q, depth=2, heads=c.heads, mode=mode)
        rec_ = vo.decode_latents(lat, q, depth=2, heads=c.dheads, mode=mode)
        l = vo.recon_loss(rec_, x, None)
        l.backward()
        return q, rec_.detach(), l

    p, rec, loss = oracle("bf16")
    p32, _, _ = oracle("fp32")          # the algorithm's own bf16 sensitivity = the yardstick of _check
    out = tr.rec_fwd_bwd(x.cuda(), 1.0, return_image=True)
    torch.cuda.synchronize()
    assert rel(out, rec) < 2e-2
    assert abs(tr.loss_acc[4].item() - loss.item()) < TOL_L * loss.item()
    FLOORS.clear()
    g, g32 = (lambda k: p[k].grad), (lambda k: p32[k].grad)
    errs = _vit_checks(tr, p, "trunk.", "trunk.", [0, 1], False, p32)
    errs.update(_vit_checks(tr, p, "pixel_decoder.", "decoder.", [0, 1], True, p32))
    for ours, key, flat in (("trunk.patch.w", "trunk.patch_embed.proj.weight", True), ("trunk.patch.b", "trunk.patch_embed.proj.bias", False),
                            ("trunk.cls", "trunk.cls_token", False), ("trunk.norm_w", "trunk.norm.weight", False),
                            ("trunk.bneck.w", "trunk.feature_bottleneck.weight", False),
                            ("decoder.proj_in.w", "pixel_decoder.proj_in.weight", True), ("decoder.proj_in.b", "pixel_decoder.proj_in.bias", False),
                            ("decoder.proj_out.w", "pixel_decoder.proj_out.weight", True), ("decoder.proj_out.b", "pixel_decoder.proj_out.bias", False),
                            ("decoder.norm_w", "pixel_decoder.norm.weight", False), ("decoder.norm_b", "pixel_decoder.norm.bias", False)):
        errs[ours] = _check(tr, ours, g(key).flatten(1) if flat else g(key), g32(key).flatten(1) if flat else g32(key))
    _summary(geo, "rec")


@pytest.mark.parametrize("geo", GEO_PARAMS)
def test_clip_objective_gradients(geo):
    c = _setup(geo)
    tr = c.tr
    B = 6 if geo == "tiny" else 4
    x = seeded_images(B, c.img, c.img)
    ids = seeded_captions(B, 77, c.vocab)
    def oracle(mode):
        q = _leafs(c.sd)
        fi = vo.clip_image_feature(x, q, depth=2, heads=c.heads, mode=mode)
        ft = vo.text_feature(ids, q, layers=2, heads=c.theads, mode=mode)
        l = vo.clip_loss(vo._r(fi, mode), vo._r(ft, mode), q["logit_scale"].exp())
        l.backward()
        return q, l

    p, loss = oracle("bf16")
    p32, _ = oracle("fp32")
    tr.clip_fwd_bwd(x.cuda(), ids.cuda(), 1.0)
    torch.cuda.synchronize()
    assert abs(tr.loss_acc[0].item() - loss.item()) < TOL_L * abs(loss.item()), (tr.loss_acc[0].item(), loss.item())
    FLOORS.clear()
    g, g32 = (lambda k: p[k].grad), (lambda k: p32[k].grad)
    errs = _vit_checks(tr, p, "trunk.", "trunk.", [0, 1], False, p32)
    errs["visual_proj"] = _check(tr, "visual_proj.w", g("visual_proj.weight"), g32("visual_proj.weight"))
    errs["patch.w"] = _check(tr, "trunk.patch.w", g("trunk.patch_embed.proj.weight").flatten(1), g32("trunk.patch_embed.proj.weight").flatten(1))
    errs["cls"] = _check(tr, "trunk.cls", g("trunk.cls_token"), g32("trunk.cls_token"))
    errs["logit_scale"] = _check(tr, "logit_scale", g("logit_scale"), g32("logit_scale"))
    errs["text.proj"] = _check(tr, "text.proj.w", g("text_projection").t(), g32("text_projection").t())
    errs["text.tok_emb"] = _check(tr, "text.tok_emb", g("token_embedding.weight"), g32("token_embedding.weight"))
    errs["text.pos"] = _check(tr, "text.pos", g("positional_embedding"), g32("positional_embedding"))
    errs["text.norm_w"] = _check(tr, "text.norm_w", g("ln_final.weight"), g32("ln_final.weight"))
    for i in (0, 1):
        r, q = f"text_transformer.resblocks.{i}.", f"text.blocks.{i}."
        for ours, key in (("qkv.w", "attn.in_proj_weight"), ("qkv.b", "attn.in_proj_bias"), ("proj.w", "attn.out_proj.weight"),
                          ("fc1.w", "mlp.c_fc.weight"), ("fc2.w", "mlp.c_proj.weight"), ("n1_b", "ln_1.bias")):
            errs[q + ours] = _check(tr, q + ours, g(r + key), g32(r + key))
    _summary(geo, "clip")


@pytest.mark.parametrize("geo", GEO_PARAMS)
def test_ssl_objective_gradients(geo):
    c = _setup(geo)
    tr, sd, hsd, K = c.tr, c.sd, c.hsd, c.K
    B, n_loc = (3, 2) if geo == "tiny" else (2, 2)
    gc = seeded_images(2 * B, c.img, c.img, seed=11)
    lc = seeded_images(n_loc * B, c.local, c.local, seed=12)
    HW = c.HW
    # mask 30 % of the patches on some of the global crops (5 of 16 at the tiny geometry)
    gsel = torch.Generator().manual_seed(5)
    masks = torch.zeros(2 * B, HW, dtype=torch.bool)
    for img in ((0, 2, 5) if geo == "tiny" else (0, 3)):
        masks[img, torch.randperm(HW, generator=gsel)[:max(5, int(0.3 * HW))]] = True
    mask_idx = masks.flatten().nonzero().flatten()
    mw = (1.0 / masks.sum(-1).clamp(min=1).float())[:, None].expand_as(masks)[masks]
    n_m = mask_idx.numel()
    p = _leafs(sd)
    hp = _leafs(hsd)
    hp_full = {"h." + k: v for k, v in hp.items()}
    with torch.no_grad():
        t_out = vo.trunk_forward([gc], [None], sd, depth=2, heads=c.heads, mode="bf16", use_bottleneck=False)[0]
        tcls = t_out["x_norm_clstoken"]
        tcls = torch.cat([tcls[B:], tcls[:B]])
        tpatch = t_out["x_norm_patchtokens"].flatten(0, 1)[mask_idx]
        th = {"h." + k: v for k, v in hsd.items()}
        tlog = vo.dino_head(vo._r(torch.cat([tcls, tpatch]), "bf16"), th, "h.", mode="bf16")
        tp_cls = vo.teacher_probs(tlog[:2 * B], torch.zeros(K), 0.07)
        tp_m = vo.teacher_probs(tlog[2 * B:], torch.zeros(K), 0.07)
    sg, sl = vo.trunk_forward([gc, lc], [masks, None], p, depth=2, heads=c.heads, mode="bf16", use_bottleneck=False)
    s_in = torch.cat([sl["x_norm_clstoken"], sg["x_norm_clstoken"], sg["x_norm_patchtokens"].flatten(0, 1)[mask_idx]])
    slog = vo.dino_head(vo._r(s_in, "bf16"), hp_full, "h.", mode="bf16")
    nl = n_loc * B
    terms = vo.dino_ibot_loss(slog[:nl], slog[nl:nl + 2 * B], slog[nl + 2 * B:], tp_cls, tp_m, mw, n_local=n_loc,
                              n_images=2 * B)
    loss = terms["dino_local"] + terms["dino_global"] + terms["ibot"]
    loss.backward()
    tr.ssl_fwd_bwd(gc.cuda(), lc.cuda(), mask_idx.cuda(), mw.cuda(), 1.0)
    torch.cuda.synchronize()
    got = tr.loss_acc[1:4].cpu()
    for j, k in enumerate(("dino_local", "dino_global", "ibot")):
        assert abs(got[j].item() - terms[k].item()) < 3e-2 * abs(terms[k].item()), (k, got[j].item(), terms[k].item())
    FLOORS.clear()                  # flat 6 % bound here (measured max 1.0 / 1.5 / 2.1 % at tiny / small2 / large2)
    errs = _vit_checks(tr, p, "trunk.", "trunk.", [0, 1], False)
    errs["patch.w"] = _check(tr, "trunk.patch.w", p["trunk.patch_embed.proj.weight"].grad.flatten(1))
    errs["cls"] = _check(tr, "trunk.cls", p["trunk.cls_token"].grad)
    errs["mask_token"] = _check(tr, "trunk.mask_token", p["trunk.mask_token"].grad)
    errs["norm"] = _check(tr, "trunk.norm_w", p["trunk.norm.weight"].grad)
    for j in (0, 2, 4):
        errs[f"head.mlp{j}.w"] = _check(tr, f"head.mlp{j}.w", hp[f"mlp.{j}.weight"].grad)
        errs[f"head.mlp{j}.b"] = _check(tr, f"head.mlp{j}.b", hp[f"mlp.{j}.bias"].grad)
    errs["head.last_v"] = _check(tr, "head.last_v", hp["last_layer.weight_v"].grad)
    errs["head.last_g"] = _check(tr, "head.last_g", hp["last_layer.weight_g"].grad)
    _summary(geo, "ssl")


def test_full_step_runs_and_learns():
    cfg, sd, hsd, tr = _setup_tiny()
    tr.tc.lr = 2e-4
    B, n_loc, HW = 4, 2, 16
    masks = torch.zeros(2 * B, HW, dtype=torch.bool)
    masks[::2, :5] = True
    batch = dict(image=seeded_images(B, 64, 64).cuda(), text=seeded_captions(B, 77, 1000).cuda(),
                 global_crops=seeded_images(2 * B, 64, 64, seed=21).cuda(), local_crops=seeded_images(n_loc * B, 32, 32, seed=22).cuda(),
                 mask_indices=masks.flatten().nonzero().flatten().cuda(),
                 masks_weight=(1.0 / masks.sum(-1).clamp(min=1).float())[:, None].expand_as(masks)[masks].cuda(),
                 rec_image=seeded_images(B, 64, 64).cuda())
    hist = []
    for _ in range(6):
        hist.append(tr.train_step(batch).cpu().clone())
    assert all(torch.isfinite(h).all() for h in hist)
    assert hist[-1][4] < hist[0][4]            # reconstruction L1 goes down on a fixed batch
    assert hist[-1][0] < hist[0][0] + 1e-3     # contrastive loss does not blow up
    # EMA teacher moved towards the student, grads were zeroed by the fused optimiser
    assert float(tr.store.g.abs().max()) == 0.0
    d = (tr.store.tp - tr.store.p[:tr.store.n_teacher]).abs().max().item()
    assert d > 0
    # exported weights load into the inference model and reproduce the trainer's reconstruction
    from vtp_b200.model import VTPModel
    m = VTPModel(cfg).cuda()
    m.load_state_dict(tr.export_state_dict(), strict=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        rec_m = m.get_latents_decoded_images(m.get_reconstruction_latents(batch["rec_image"]))
    rec_t = tr.rec_fwd_bwd(batch["rec_image"], 1.0, return_image=True)
    assert rel(rec_m, rec_t) < 2e-2


def test_rec_objective_with_lpips_gradients():
    """recon = L1 + 1.0 * LPIPS (frozen seeded-random VGG16): loss terms and trunk/decoder gradients vs oracle autograd."""
    from vtp_b200.lpips import LPIPSLoss, random_weights

    cfg, sd, hsd, tr = _setup_tiny()
    vw, vb, lw = random_weights(0)
    tr.enable_lpips(LPIPSLoss(vw, vb, lw, device="cuda", chunk=2))
    x = seeded_images(3, 64, 64) * 0.5
    p = _leafs(sd)
    lat = vo.reconstruction_latents(x, p, depth=2, heads=2, mode="bf16")
    rec = vo.decode_latents(lat, p, depth=2, heads=2, mode="bf16")
    lp = vo.lpips(rec, x, vw, vb, lw, mode="bf16")
    loss = vo.recon_loss(rec, x, lp, 1.0)
    loss.backward()
    tr.rec_fwd_bwd(x.cuda(), 1.0)
    torch.cuda.synchronize()
    l1, lpv = tr.loss_acc[4].item(), tr.loss_acc[5].item()
    assert abs(lpv - lp.mean().item()) < 3e-2 * abs(lp.mean().item()), (lpv, lp.mean().item())
    assert abs(l1 + lpv - loss.item()) < TOL_L * loss.item()
    errs = _vit_checks(tr, p, "pixel_decoder.", "decoder.", [0, 1], True)
    errs.update(_vit_checks(tr, p, "trunk.", "trunk.", [0], False))
    errs["proj_out.w"] = _check(tr, "decoder.proj_out.w", p["pixel_decoder.proj_out.weight"].grad.flatten(1))
    errs["bneck"] = _check(tr, "trunk.bneck.w", p["trunk.feature_bottleneck.weight"].grad)
    print("rec+lpips grad rel errors: max", max(errs.values()))


def _tiny_batch(B=4, n_loc=2, HW=16):
    masks = torch.zeros(2 * B, HW, dtype=torch.bool)
    masks[::2, :5] = True
    return dict(image=seeded_images(B, 64, 64).cuda(), text=seeded_captions(B, 77, 1000).cuda(),
                global_crops=seeded_images(2 * B, 64, 64, seed=21).cuda(), local_crops=seeded_images(n_loc * B, 32, 32, seed=22).cuda(),
                mask_indices=masks.flatten().nonzero().flatten().cuda(),
                masks_weight=(1.0 / masks.sum(-1).clamp(min=1).float())[:, None].expand_as(masks)[masks].cuda(),
                rec_image=seeded_images(B, 64, 64).cuda())


def test_graph_step_equals_eager_steps():
    """VTPTrainer.capture_step / replay_step: the CUDA graph of the whole step (3 objectives + LPIPS + optimiser + EMA,
    device-side step counter and bias corrections) must train exactly like the eager launches — same kernels, same order;
    only the split-K fp32 atomics of the weight gradients may reorder, hence a tolerance instead of bit equality."""
    from vtp_b200 import lib

    batch = _tiny_batch()
    b2 = dict(batch)
    b2["rec_image"] = seeded_images(4, 64, 64, seed=77).cuda()     # a second batch of the same shapes
    seq = [batch, batch, b2, batch, b2]                             # the capture's two warm-up steps train on `batch`

    def trainer():
        _, _, _, t = _setup_tiny()
        t.hyper[3] = 2e-4
        t.enable_lpips(seed=0, chunk=2)
        return t

    te = trainer()
    le = [te.train_step(b).cpu().clone() for b in seq]
    tg = trainer()
    tg.capture_step(batch, warmup=2)
    lg = [tg.replay_step(b).cpu().clone() for b in seq[2:]]
    assert tg.step_count == te.step_count == 5
    assert int(tg.hyper[0].item()) == 5 and int(te.hyper[0].item()) == 5
    for a, b in zip(le[2:], lg):
        assert torch.isfinite(b).all()
        assert torch.allclose(a, b, rtol=2e-3, atol=1e-5), (a, b)
    assert rel(tg.store.p, te.store.p) < 1e-4
    assert rel(tg.store.tp, te.store.tp) < 1e-5
    assert float(tg.store.g.abs().max()) == 0.0
    assert tg.graph_launches > 100
    with pytest.raises(ValueError):
        bad = dict(batch)
        bad["rec_image"] = seeded_images(2, 64, 64).cuda()
        tg.replay_step(bad)


def test_device_schedules_drive_the_optimizer():
    """set_schedules(): lr / weight decay / teacher momentum follow the device tables by the optimiser's own step counter
    (eagerly and inside the captured graph), holding the last value past the end."""
    from vtp_b200.schedules import CosineSchedule

    _, _, _, tr = _setup_tiny()
    lr = CosineSchedule(1e-3, 1e-5, total_iters=6, warmup_iters=2, start_warmup_value=1e-6)
    mom = CosineSchedule(0.99, 1.0, total_iters=4)
    tr.set_schedules(lr=lr, teacher_momentum=mom)
    batch = _tiny_batch()
    seen = []
    for _ in range(3):
        tr.train_step(batch)
        seen.append(tr.scheduled_values())
    tr.capture_step(batch, warmup=1)                 # step 4 eagerly
    seen.append(tr.scheduled_values())
    for _ in range(4):                               # steps 5..8 by replay: beyond both tables
        tr.replay_step()
        seen.append(tr.scheduled_values())
    for i, sv in enumerate(seen):
        assert sv["step"] == i + 1
        assert abs(sv["lr"] - lr[i]) <= 1e-6 * abs(lr[i]) + 1e-12, (i, sv, lr[i])
        assert abs(sv["teacher_momentum"] - mom[i]) < 1e-6, (i, sv, mom[i])
        assert abs(sv["weight_decay"] - tr.tc.weight_decay) < 1e-9
    # momentum 1.0 from step 5 on: the teacher stops moving
    t0 = tr.store.tp.clone()
    tr.replay_step()
    assert torch.equal(t0, tr.store.tp)


def test_stochastic_depth_forward_matches_reference_golden():
    """a8: batch-subset stochastic depth (layers/block.py:201-298) on the kernels vs the REAL reference's training branch
    with the same preset subsets (tests/golden/tiny_drop.npz, oracle/make_golden_drop.py), bf16 mode vs the fp32 reference."""
    import os

    import numpy as np

    from vtp_b200 import engine as E

    c = _setup("tiny")
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_drop.npz")).items()}
    x = seeded_images(3, 64, 64)
    W = c.tr.towers[("trunk", "param")]
    plan = E.DropPlan(float(g["ratio"][0]), preset=[p for p in g["presets"]])
    xo, meta = E.trunk_forward(W, x.cuda(), "bf16", drop=plan)
    out = E.trunk_outputs(W, xo, meta, "bf16", use_bottleneck=True)
    assert plan.calls == 4
    # bf16 kernels vs fp32 reference: the usual bf16-mode distance of this geometry (2e-2 bound)
    e_drop = rel(out["x_norm_patchtokens"].float(), g["patch"])
    assert e_drop < 2e-2
    # and it is NOT the plain path: without the subsets the distance to the reference's stochastic-depth output is several
    # times larger (2.7e-2 measured: random-init residual branches are small next to the stream)
    xo0, meta0 = E.trunk_forward(W, x.cuda(), "bf16")
    out0 = E.trunk_outputs(W, xo0, meta0, "bf16", use_bottleneck=True)
    e_plain = rel(out0["x_norm_patchtokens"].float(), g["patch"])
    print(f"stochastic depth forward: rel {e_drop:.2e} (plain path {e_plain:.2e})")
    assert e_plain > 2 * e_drop


def test_stochastic_depth_gradients_match_oracle():
    """Reconstruction objective with rec_drop_rate > 0 and preset subsets: loss and every trunk / decoder gradient vs autograd
    through the oracle's restatement of the reference's index / index_add branch."""
    c = _setup("tiny")
    tr = c.tr
    B = 4
    x = seeded_images(B, c.img, c.img)
    ratio = 0.5
    keep = max(int(B * (1 - ratio)), 1)
    gen = torch.Generator().manual_seed(77)
    presets = [torch.randperm(B, generator=gen)[:keep] for _ in range(4)]
    sc = B / keep
    p = _leafs(c.sd)
    drops = [[(presets[0], sc, presets[1], sc), (presets[2], sc, presets[3], sc)]]
    o = vo.trunk_forward([x], [None], p, depth=2, heads=c.heads, mode="bf16", drops=drops)[0]
    pt = o["x_norm_patchtokens"]
    lat = pt.transpose(1, 2).reshape(B, pt.shape[-1], c.img // 16, c.img // 16)
    rec = vo.decode_latents(lat, p, depth=2, heads=c.dheads, mode="bf16")
    loss = vo.recon_loss(rec, x, None)
    loss.backward()
    tr.tc.rec_drop_rate = ratio
    tr.drop_presets = [presets]
    out = tr.rec_fwd_bwd(x.cuda(), 1.0, return_image=True)
    torch.cuda.synchronize()
    assert rel(out, rec.detach()) < 2e-2
    assert abs(tr.loss_acc[4].item() - loss.item()) < TOL_L * loss.item()
    errs = _vit_checks(tr, p, "trunk.", "trunk.", [0, 1], False)
    errs.update(_vit_checks(tr, p, "pixel_decoder.", "decoder.", [0, 1], True))
    errs["patch.w"] = _check(tr, "trunk.patch.w", p["trunk.patch_embed.proj.weight"].grad.flatten(1))
    errs["cls"] = _check(tr, "trunk.cls", p["trunk.cls_token"].grad)
    errs["bneck"] = _check(tr, "trunk.bneck.w", p["trunk.feature_bottleneck.weight"].grad)
    print("stochastic-depth rec grad rel errors: max", max(errs.values()), {k: f"{v:.1e}" for k, v in errs.items()})


def test_stochastic_depth_full_step_random_subsets():
    """All three objectives with drop rates > 0 and the random (torch.randperm, device-side) subsets: the step runs, learns,
    and is graph-capturable (the permutation is drawn inside the graph by the graph-safe CUDA generator)."""
    c = _setup("tiny")
    tr = c.tr
    tr.tc.clip_drop_rate = tr.tc.ssl_drop_rate = tr.tc.rec_drop_rate = 0.25
    tr.hyper[3] = 2e-4
    batch = _tiny_batch()
    l0 = tr.train_step(batch).cpu().clone()
    for _ in range(4):
        l1 = tr.train_step(batch).cpu().clone()
    assert torch.isfinite(l0).all() and torch.isfinite(l1).all() and l1[4] < l0[4]
    tr.capture_step(batch, warmup=1)
    a = tr.replay_step().cpu().clone()
    b = tr.replay_step().cpu().clone()
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    assert not torch.equal(a, b)          # different subsets (and parameters) every replay
