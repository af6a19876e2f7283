"""The 3-objective training step (CLIP contrastive + DINO/iBOT self-distillation + pixel reconstruction) on the sm_100a
kernels, with hand-written backward (no torch.autograd): the B200-native counterpart of the reference's legacy
meta-arch `VTP` (vtp/models/vtp.py:88-512: forward_clip / forward_ssl_learning / forward_reconstruction,
update_teacher) plus the loss / optimiser layer that the reference does not release.

  * parameters live in ONE flat fp32 master buffer with matching bf16 compute copy, fp32 gradient and Adam moments
    (`ParamStore`); GEMM weights are stored in kernel layout (w1|w2 8-interleaved for the SwiGLU-gate epilogue);
    `import_state_dict` / `export_state_dict` convert from/to the reference's state-dict keys;
  * one fused kernel per region does AdamW + bf16 refresh + EMA teacher (vtp.py:388-401) + gradient zeroing;
  * data parallel: gradients are all-reduced (NCCL) on the flat buffer, contrastive features are all-gathered and
    their gradients reduced back — the only two collectives on the path.
Precision is the reference-under-autocast ("bf16") mode throughout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import engine as E
from . import lib
from .config import VTPConfig
from .engine import BF, F32, BlockW, Lin, TowerW, _e

_ALIGN = 64  # elements; keeps every tensor 128B-aligned in the bf16 copy (TMA needs 16B)


# ------------------------------------------------------------------------------------------------------ parameters
class ParamStore:
    def __init__(self, device):
        self.device = device
        self.specs: List[Tuple[str, Tuple[int, ...], bool, bool]] = []
        self.offset: Dict[str, int] = {}
        self.shape: Dict[str, Tuple[int, ...]] = {}
        self.regions: List[Tuple[int, int, bool, bool]] = []  # (start, end, decay, teacher)

    def add(self, name, shape, decay=True, teacher=False):
        self.specs.append((name, tuple(int(s) for s in shape), bool(decay), bool(teacher)))

    def finalize(self):
        off = 0
        for teacher in (True, False):
            for decay in (True, False):
                start = off
                for name, shape, d, t in self.specs:
                    if d == decay and t == teacher:
                        self.offset[name], self.shape[name] = off, shape
                        n = 1
                        for s in shape:
                            n *= s
                        off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
                if off > start:
                    self.regions.append((start, off, decay, teacher))
        self.n = off
        self.n_teacher = max([e for s, e, d, t in self.regions if t], default=0)
        dev = self.device
        self.p = torch.zeros(self.n, dtype=F32, device=dev)
        self.pb = torch.zeros(self.n, dtype=BF, device=dev)
        self.g = torch.zeros(self.n, dtype=F32, device=dev)
        self.m = torch.zeros(self.n, dtype=F32, device=dev)
        self.v = torch.zeros(self.n, dtype=F32, device=dev)
        self.tp = torch.zeros(self.n_teacher, dtype=F32, device=dev)
        self.tpb = torch.zeros(self.n_teacher, dtype=BF, device=dev)

    def _view(self, buf, name):
        o, shape = self.offset[name], self.shape[name]
        n = 1
        for s in shape:
            n *= s
        return buf[o:o + n].view(shape)

    def f32(self, name): return self._view(self.p, name)
    def bf16(self, name): return self._view(self.pb, name)
    def grad(self, name): return self._view(self.g, name)
    def tf32(self, name): return self._view(self.tp, name)
    def tbf16(self, name): return self._view(self.tpb, name)

    def sync_compute_copies(self, init_teacher: bool = False):
        lib.cast_f32_to_bf16(self.p, self.pb, self.n)
        if init_teacher and self.n_teacher:
            self.tp.copy_(self.p[:self.n_teacher])
            self.tpb.copy_(self.pb[:self.n_teacher])


def _vit_specs(store: ParamStore, pre: str, D: int, depth: int, hidden: int, ln: bool, teacher: bool, ffn_out: int):
    for i in range(depth):
        p = f"{pre}blocks.{i}."
        store.add(p + "n1_w", (D,), False, teacher)
        if ln: store.add(p + "n1_b", (D,), False, teacher)
        store.add(p + "qkv.w", (3 * D, D), True, teacher); store.add(p + "qkv.b", (3 * D,), False, teacher)
        store.add(p + "proj.w", (D, D), True, teacher); store.add(p + "proj.b", (D,), False, teacher)
        store.add(p + "n2_w", (D,), False, teacher)
        if ln: store.add(p + "n2_b", (D,), False, teacher)
        store.add(p + "fc1.w", (ffn_out, D), True, teacher); store.add(p + "fc1.b", (ffn_out,), False, teacher)
        store.add(p + "fc2.w", (D, hidden), True, teacher); store.add(p + "fc2.b", (D,), False, teacher)
    store.add(pre + "norm_w", (D,), False, teacher)
    if ln: store.add(pre + "norm_b", (D,), False, teacher)


def _tower_views(store: ParamStore, pre: str, tw: TowerW, depth: int, hidden: int, ln: bool, *, kind: str):
    """kind: 'param' (bf16 weights + fp32 vectors), 'teacher', or 'grad' (fp32 gradient views)."""
    if kind == "param":
        wv, fv = store.bf16, store.f32
    elif kind == "teacher":
        wv, fv = store.tbf16, store.tf32
    else:
        wv, fv = store.grad, store.grad

    def lin(name):
        w = wv(name + ".w")
        b = fv(name + ".b") if (name + ".b") in store.offset else None
        return Lin(w, b, w.shape[0], w.shape[1])

    tw.blocks = []
    for i in range(depth):
        p = f"{pre}blocks.{i}."
        tw.blocks.append(BlockW(n1_w=fv(p + "n1_w"), n1_b=fv(p + "n1_b") if ln else None, qkv=lin(p + "qkv"),
                                proj=lin(p + "proj"), n2_w=fv(p + "n2_w"), n2_b=fv(p + "n2_b") if ln else None,
                                fc1=lin(p + "fc1"), fc2=lin(p + "fc2"), hidden=hidden))
    tw.norm_w = fv(pre + "norm_w")
    tw.norm_b = fv(pre + "norm_b") if ln else None
    return lin


# ------------------------------------------------------------------------------------------------------ backward
def wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, Mtok: int):
    """dW[N_out, K_in] += dYᵀ X  (dY [Mtok, N_out] bf16, X [Mtok, K_in] bf16): TN GEMM, split-K + fp32 atomics."""
    n_out, k_in = dw.shape
    # split_k = -1: the library picks the split that fills its persistent grid most evenly for the tile shape it chose
    lib.gemm(dy, x, dw, M=n_out, N=k_in, K=Mtok, a_mn=True, b_mn=True, lda=dy.stride(0), ldb=x.stride(0), ldo=k_in,
             accumulate=True, split_k=-1, round_bf16=False)


def dgrad(dy: torch.Tensor, w: torch.Tensor, out: torch.Tensor, Mtok: int, **kw):
    """dX[Mtok, K_in] = dY[Mtok, N_out] · W[N_out, K_in]  (W consumed as an MN-major B operand, no transpose copy)."""
    n_out, k_in = w.shape
    lib.gemm(dy, w, out, M=Mtok, N=k_in, K=n_out, b_mn=True, lda=dy.stride(0), ldb=k_in, round_bf16=False, **kw)


def tower_blocks_backward(W: TowerW, G: TowerW, tape: list, g: torch.Tensor, B: int, T: int, rope, causal=False):
    """Reverse of engine.tower_blocks.  g fp32 [B*T, D]: in = dL/d(stream out), out = dL/d(stream in) (in place).
    The bf16 copy of g (dY operand of the next sub-layer to differentiate) and its column sums (that sub-layer's bias
    gradient) are by-products of the preceding norm_bwd; only the first sub-layer needs a stand-alone cast."""
    dev = g.device
    M, D, H = B * T, W.D, W.heads
    nb = len(W.blocks)
    if nb and "drop1" in tape[nb - 1]:
        return _tower_blocks_backward_drop(W, G, tape, g, B, T, rope)
    gb = _e((M, D), BF, dev)
    lib.cast_colsum(g, gb, G.blocks[nb - 1].fc2.b, M, D)
    for li in reversed(range(nb)):
        bw, gw, t = W.blocks[li], G.blocks[li], tape[li]
        Hd = bw.hidden
        # ---- FFN sub-layer (gb / fc2 bias gradient already produced)
        dhid = _e((M, Hd), BF, dev)
        dgrad(gb, bw.fc2.w, dhid, M)
        wgrad(gb, t["hid"], gw.fc2.w, M)
        dpre = torch.empty_like(t["pre"])
        if W.ffn == "swiglu":
            lib.swiglu_bwd(t["pre"], dhid, dpre, gw.fc1.b, M, Hd)
        else:
            lib.gelu_bwd(t["pre"], dhid, dpre, gw.fc1.b, M, Hd)
        dh = _e((M, D), BF, dev)
        dgrad(dpre, bw.fc1.w, dh, M)
        wgrad(dpre, t["h2"], gw.fc1.w, M)
        lib.norm_bwd(t["x_mid"], t["n2"]["rstd"], t["n2"]["mean"], bw.n2_w, dh, g, gw.n2_w, gw.n2_b, M, D,
                     gb_out=gb, g_colsum=gw.proj.b)
        # ---- attention sub-layer
        do = _e((M, D), BF, dev)
        dgrad(gb, bw.proj.w, do, M)
        wgrad(gb, t["o"], gw.proj.w, M)
        dqkv = _e((M, 3 * D), BF, dev)
        lib.attention_bwd(t["qkv"], t["o"], do, t["lse"], dqkv, B, T, H, prefix=W.prefix, causal=causal, rope=rope)
        lib.cast_colsum(dqkv, None, gw.qkv.b, M, 3 * D)
        dgrad(dqkv, bw.qkv.w, dh, M)
        wgrad(dqkv, t["h1"], gw.qkv.w, M)
        nxt = G.blocks[li - 1].fc2.b if li > 0 else None
        lib.norm_bwd(t["x_in"], t["n1"]["rstd"], t["n1"]["mean"], bw.n1_w, dh, g, gw.n1_w, gw.n1_b, M, D,
                     gb_out=gb if li > 0 else None, g_colsum=nxt)
        tape[li] = None  # free the saved activations of this block
    return g


def grad_buckets(offset: Dict[str, int], n: int) -> Dict[str, List[Tuple[int, int]]]:
    """Contiguous ranges of the flat gradient buffer per tower.  A tower's gradient is FINAL as soon as its last
    backward of the step has run: text tower + clip projection + logit scale after the contrastive objective, DINO head
    after the SSL head backward, pixel decoder after its backward inside the reconstruction objective; the trunk (all
    three objectives accumulate into it) only at the end of the step.  Each finished bucket is all-reduced right away
    (NCCL's own stream) under the remaining backward; only the trunk bucket is exposed."""
    def bucket_of(name: str) -> str:
        if name.startswith("text.") or name.startswith("visual_proj") or name == "logit_scale":
            return "text"
        if name.startswith("head."):
            return "head"
        if name.startswith("decoder."):
            return "decoder"
        return "trunk"
    spans = sorted((off, name) for name, off in offset.items())
    out: Dict[str, List[Tuple[int, int]]] = {"text": [], "head": [], "decoder": [], "trunk": []}
    for i, (off, name) in enumerate(spans):
        end = spans[i + 1][0] if i + 1 < len(spans) else n   # incl. the alignment padding behind the tensor
        r = out[bucket_of(name)]
        if r and r[-1][1] == off:
            r[-1] = (r[-1][0], end)
        else:
            r.append((off, end))
    return out


def _tower_blocks_backward_drop(W: TowerW, G: TowerW, tape: list, g: torch.Tensor, B: int, T: int, rope):
    """Reverse of engine._block_drop (batch-subset stochastic depth, layers/block.py:201-233).  The residual stream
    gradient g passes every block unchanged (identity path); each sub-layer adds, for its kept images only, the gradient
    that flows through  alpha * sublayer(norm(x[idx])) :  d(residual) = alpha * g[idx]  goes down the sub-layer, what
    comes out of its norm backward is scatter-added into g[idx]."""
    dev, D, H = g.device, W.D, W.heads
    for li in reversed(range(len(W.blocks))):
        bw, gw, t = W.blocks[li], G.blocks[li], tape[li]
        Hd = bw.hidden
        # ---- FFN sub-layer
        idx2, a2, xs2 = t["drop2"]
        M2 = idx2.numel() * T
        gs = _e((M2, D), F32, dev)
        lib.gather_images(g, gs, idx2, T, D, a2)
        gb = _e((M2, D), BF, dev)
        lib.cast_colsum(gs, gb, gw.fc2.b, M2, D)
        dhid = _e((M2, Hd), BF, dev)
        dgrad(gb, bw.fc2.w, dhid, M2)
        wgrad(gb, t["hid"], gw.fc2.w, M2)
        dpre = torch.empty_like(t["pre"])
        lib.swiglu_bwd(t["pre"], dhid, dpre, gw.fc1.b, M2, Hd)
        dh = _e((M2, D), BF, dev)
        dgrad(dpre, bw.fc1.w, dh, M2)
        wgrad(dpre, t["h2"], gw.fc1.w, M2)
        gsub = gs.zero_()
        lib.norm_bwd(xs2, t["n2"]["rstd"], t["n2"]["mean"], bw.n2_w, dh, gsub, gw.n2_w, gw.n2_b, M2, D)
        lib.scatter_add_images(gsub, g, idx2, T, D, 1.0)
        # ---- attention sub-layer
        idx1, a1, xs1 = t["drop1"]
        n1 = idx1.numel()
        M1 = n1 * T
        gs = _e((M1, D), F32, dev)
        lib.gather_images(g, gs, idx1, T, D, a1)
        gb = _e((M1, D), BF, dev)
        lib.cast_colsum(gs, gb, gw.proj.b, M1, D)
        do = _e((M1, D), BF, dev)
        dgrad(gb, bw.proj.w, do, M1)
        wgrad(gb, t["o"], gw.proj.w, M1)
        dqkv = _e((M1, 3 * D), BF, dev)
        lib.attention_bwd(t["qkv"], t["o"], do, t["lse"], dqkv, n1, T, H, prefix=W.prefix, rope=rope)
        lib.cast_colsum(dqkv, None, gw.qkv.b, M1, 3 * D)
        dh = _e((M1, D), BF, dev)
        dgrad(dqkv, bw.qkv.w, dh, M1)
        wgrad(dqkv, t["h1"], gw.qkv.w, M1)
        gsub = gs.zero_()
        lib.norm_bwd(xs1, t["n1"]["rstd"], t["n1"]["mean"], bw.n1_w, dh, gsub, gw.n1_w, gw.n1_b, M1, D)
        lib.scatter_add_images(gsub, g, idx1, T, D, 1.0)
        tape[li] = None
    return g


# ------------------------------------------------------------------------------------------------------ trainer
@dataclass
class TrainConfig:
    lr: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.95
    eps: float = 1e-8
    weight_decay: float = 0.05
    teacher_momentum: float = 0.994
    teacher_temp: float = 0.07
    student_temp: float = 0.1
    center_momentum: float = 0.9
    head_out_dim: int = 65536
    head_hidden: int = 2048
    head_bottleneck: int = 256
    n_local_crops: int = 8
    w_clip: float = 1.0
    w_ssl: float = 1.0
    w_rec: float = 1.0
    lpips_weight: float = 1.0
    # contrastive exchange (collective C2): "nccl" = all-gather + local-row logits + all-reduced cross terms;
    # "p2p" = peer-memory gather fused with the full logits / softmax-CE (csrc/clip.cu), no backward collective
    clip_exchange: str = "nccl"
    # activation-memory control (configs 3/4: VTP-Base/Large at 256 images per GPU exceed 180 GB of saved activations
    # in one piece, see vtp_b200/memory.py): source images per forward+backward pass of the SSL / reconstruction
    # objectives; gradients accumulate in the flat buffer, losses and centre statistics are those of the whole batch.
    # 0 = the whole per-GPU batch at once.
    ssl_chunk: int = 0
    rec_chunk: int = 0
    # batch-subset stochastic depth of the student trunk per objective (vtp.py:275-293 clip_drop_rate, :452-463
    # ssl_drop_rate, :487-500 rec_drop_rate; layers/block.py:20-118,201-298); 0 = the plain residual path
    clip_drop_rate: float = 0.0
    ssl_drop_rate: float = 0.0
    rec_drop_rate: float = 0.0
    # collective C3: "end" = ONE all-reduce of the flat gradient buffer after the last backward; "overlap" = every tower's
    # bucket is all-reduced as soon as it is final, under the remaining backward (env VTP_GRAD_REDUCE overrides).
    # Measured at N = 8 (profiles/r2_logs/bench_{small,large}_n8{,_end}.json): "end" 231.1 / 1317.5 ms vs "overlap"
    # 233.6 / 1325.4 ms (Small / Large) — NCCL's CTAs take SMs away from the persistent, full-machine GEMM and attention
    # kernels for longer than the ~1 ms (Small) / ~8 ms (Large, 3 GB fp32) the exposed all-reduce costs over NVLink.
    grad_reduce: str = "end"


class VTPTrainer:
    def __init__(self, cfg: VTPConfig, tc: Optional[TrainConfig] = None, device="cuda", process_group=None):
        self.cfg, self.tc = cfg, tc or TrainConfig()
        self.device = torch.device(device)
        self.pg = process_group
        import torch.distributed as dist
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        lib.check(lib.load().vtp_check_device(), "vtp_check_device")
        c = cfg
        if c.vision_norm_layer != "rmsnorm" or c.decoder_norm_layer not in ("layernorm", "layernormbf16"):
            raise NotImplementedError("trainer supports the reference defaults: rmsnorm trunk, layernorm decoder")
        from .model import _swiglu_hidden, check_head_dims
        if not (c.train_clip and c.train_reconstruction):
            raise NotImplementedError("the trainer runs all three objectives: train_clip and train_reconstruction must be on")
        check_head_dims(c)
        self.D, self.Dd, self.Dt = c.vision_embed_dim, c.decoder_embed_dim, c.text_embed_dim
        self.hs = _swiglu_hidden(self.D, c.vision_mlp_ratio, c.vision_ffn_layer)
        self.hsd = _swiglu_hidden(self.Dd, 4.0, c.decoder_ffn_layer)
        self.ht = int(self.Dt * c.text_mlp_ratio)
        self.bn = c.vision_feature_bottleneck
        st = ParamStore(self.device)
        D, Dd, Dt, K = self.D, self.Dd, self.Dt, self.tc.head_out_dim
        ps = c.vision_patch_size
        # trunk (+ clip projection + DINO head): these have an EMA teacher (vtp.py:239-262)
        st.add("trunk.patch.w", (D, 3 * ps * ps), True, True); st.add("trunk.patch.b", (D,), False, True)
        st.add("trunk.cls", (D,), False, True); st.add("trunk.mask_token", (D,), False, True)
        _vit_specs(st, "trunk.", D, c.vision_depth, self.hs, False, True, 2 * self.hs)
        st.add("trunk.bneck.w", (self.bn, D), True, True)
        st.add("visual_proj.w", (Dt, D), True, True)
        hh, hb = self.tc.head_hidden, self.tc.head_bottleneck
        st.add("head.mlp0.w", (hh, D), True, True); st.add("head.mlp0.b", (hh,), False, True)
        st.add("head.mlp2.w", (hh, hh), True, True); st.add("head.mlp2.b", (hh,), False, True)
        st.add("head.mlp4.w", (hb, hh), True, True); st.add("head.mlp4.b", (hb,), False, True)
        st.add("head.last_v", (K, hb), True, True); st.add("head.last_g", (K,), False, True)
        # decoder
        st.add("decoder.proj_in.w", (Dd, self.bn), True); st.add("decoder.proj_in.b", (Dd,), False)
        _vit_specs(st, "decoder.", Dd, c.decoder_depth, self.hsd, True, False, 2 * self.hsd)
        st.add("decoder.proj_out.w", (3 * 256, Dd), True); st.add("decoder.proj_out.b", (3 * 256,), False)
        # text
        st.add("text.tok_emb", (c.text_vocab_size, Dt), True); st.add("text.pos", (c.text_context_length, Dt), False)
        _vit_specs(st, "text.", Dt, c.text_depth, self.ht, True, False, self.ht)
        st.add("text.proj.w", (Dt, Dt), True)
        st.add("logit_scale", (1,), False)
        st.finalize()
        self.store = st
        self._build_towers()
        self.step_count = 0
        self.lpips = None  # perceptual term of the reconstruction loss, see enable_lpips()
        self.peer = None   # peer-memory exchange buffers of the contrastive objective (clip_exchange == "p2p")
        self.loss_acc = torch.zeros(8, dtype=F32, device=self.device)  # clip, dino_local, dino_global, ibot, rec
        self.center_dino = torch.zeros(K, dtype=F32, device=self.device)
        self.center_ibot = torch.zeros(K, dtype=F32, device=self.device)
        self.head_wn = _e((K, hb), BF, self.device)       # weight-normed last layer (student), rebuilt every step
        self.head_wn_t = _e((K, hb), BF, self.device)     # teacher
        self.head_vnorm = _e((K,), F32, self.device)
        # device-resident step state (csrc/backward.cu hyper_tick_kernel): {step, 1-b1^t, 1-b2^t, lr, wd, teacher momentum}
        self.hyper = torch.zeros(8, dtype=F32, device=self.device)
        self.hyper[3:6] = torch.tensor([self.tc.lr, self.tc.weight_decay, self.tc.teacher_momentum], device=self.device)
        self._sched = (None, None, None, 0)     # device tables (lr, wd, momentum) + common length
        self._buckets = self._grad_buckets()    # collective C3: gradient all-reduce per tower, overlapped with backward
        self._pending = []                      # in-flight all-reduce handles of the current step
        self._started = set()                   # buckets whose all-reduce has been started this step
        self._graph = None                      # CUDA graph of the whole step, see capture_step()
        self.reset_parameters()

    def _grad_buckets(self) -> Dict[str, List[Tuple[int, int]]]:
        return grad_buckets(self.store.offset, self.store.n)

    def _grad_reduce_mode(self) -> str:
        import os
        mode = os.environ.get("VTP_GRAD_REDUCE", self.tc.grad_reduce)
        if mode not in ("overlap", "end"):
            raise ValueError(f"grad_reduce must be 'overlap' or 'end', got {mode!r}")
        return mode

    def _reduce_bucket(self, which: str, final: bool = False):
        """Start the all-reduce of one finished gradient bucket (no-op single rank / already started this step)."""
        if self.world == 1 or which in self._started:
            return
        if not final and self._grad_reduce_mode() == "end":
            return                       # deferred: allreduce_grads() starts every bucket after the last backward
        import torch.distributed as dist
        self._started.add(which)
        for a, b in self._buckets[which]:
            self._pending.append(dist.all_reduce(self.store.g[a:b], group=self.pg, async_op=True))

    def enable_lpips(self, module=None, seed: int = 0, chunk: int = 32):
        """Attach the LPIPS term (utils/lpips.py) to the reconstruction loss: rec = L1 + lpips_weight * LPIPS.  Without
        a supplied module, frozen seeded-random VGG16/lin weights are used (the trained ones need network access)."""
        from .lpips import LPIPSLoss
        self.lpips = module if module is not None else LPIPSLoss.random_init(seed, device=self.device, chunk=chunk)
        return self.lpips

    # -------------------------------------------------------------- towers as views of the flat buffers
    def _build_towers(self):
        c, st = self.cfg, self.store

        def mk(pre, D, heads, depth, hidden, ln, norm, eps, stream_bf16, prefix, ffn, kind):
            tw = TowerW(D=D, heads=heads, norm=norm, eps=eps, stream_bf16=stream_bf16, prefix=prefix, ffn=ffn)
            lin = _tower_views(st, pre, tw, depth, hidden, ln, kind=kind)
            from .rope import rope_periods
            tw.periods = rope_periods(64)
            return tw, lin

        self.towers = {}
        for kind in ("param", "teacher", "grad"):
            tw, lin = mk("trunk.", self.D, c.vision_num_heads, c.vision_depth, self.hs, False, "rms", 1e-5, False, 1,
                         "swiglu", kind)
            fv = {"param": st.f32, "teacher": st.tf32, "grad": st.grad}[kind]
            tw.extra.update(patch=lin("trunk.patch"), patch_size=c.vision_patch_size, cls=fv("trunk.cls"),
                            mask_token=fv("trunk.mask_token"), bneck=lin("trunk.bneck"),
                            visual_proj=lin("visual_proj"), mlp0=lin("head.mlp0"), mlp2=lin("head.mlp2"),
                            mlp4=lin("head.mlp4"), last_v=fv("head.last_v"), last_g=fv("head.last_g"))
            self.towers[("trunk", kind)] = tw
        for kind in ("param", "grad"):
            eps_d = 1e-6 if c.decoder_norm_layer == "layernorm" else 1e-5
            tw, lin = mk("decoder.", self.Dd, c.decoder_num_heads, c.decoder_depth, self.hsd, True, "ln", eps_d, True, 0,
                         "swiglu", kind)
            tw.extra.update(proj_in=lin("decoder.proj_in"), proj_out=lin("decoder.proj_out"))
            self.towers[("decoder", kind)] = tw
            tw, lin = mk("text.", self.Dt, c.text_num_heads, c.text_depth, self.ht, True, "ln", 1e-5, False, 0, "gelu",
                         kind)
            fv = st.f32 if kind == "param" else st.grad
            tw.extra.update(tok_emb=fv("text.tok_emb"), pos=fv("text.pos"), proj=lin("text.proj"))
            self.towers[("text", kind)] = tw

    # -------------------------------------------------------------- init / state-dict exchange
    @torch.no_grad()
    def reset_parameters(self, seed: int = 0):
        g = torch.Generator(device="cpu").manual_seed(seed)
        st = self.store
        for name, shape, decay, _ in st.specs:
            v = st.f32(name)
            leaf = name.rsplit(".", 1)[-1]
            if name == "logit_scale":
                v.fill_(math.log(1 / 0.07))
            elif leaf in ("n1_w", "n2_w", "norm_w", "last_g"):
                v.fill_(1.0)
            elif leaf in ("b", "n1_b", "n2_b", "norm_b", "mask_token"):
                v.zero_()
            elif name == "text.pos":
                v.copy_(torch.randn(shape, generator=g) * 0.01)
            else:
                std = 0.02
                v.copy_((torch.randn(shape, generator=g) * std).clamp_(-2 * std, 2 * std))
        st.sync_compute_copies(init_teacher=True)

    @torch.no_grad()
    def import_state_dict(self, sd: Dict[str, torch.Tensor], head_sd: Optional[Dict[str, torch.Tensor]] = None):
        """Load a reference-format VTPModel state dict (+ optional DINOHead state dict with keys mlp.0.weight, ...,
        last_layer.weight_g / weight_v or the parametrizations.* spelling)."""
        st = self.store
        dev = self.device

        def put(name, t):
            st.f32(name).copy_(t.to(dev, F32).reshape(st.shape[name]))

        def vit(pre_ref, pre, depth, ln, swiglu=True):
            for i in range(depth):
                r, p = f"{pre_ref}blocks.{i}.", f"{pre}blocks.{i}."
                put(p + "n1_w", sd[r + "norm1.weight"]); put(p + "n2_w", sd[r + "norm2.weight"])
                if ln:
                    put(p + "n1_b", sd[r + "norm1.bias"]); put(p + "n2_b", sd[r + "norm2.bias"])
                put(p + "qkv.w", sd[r + "attn.qkv.weight"]); put(p + "qkv.b", sd[r + "attn.qkv.bias"])
                put(p + "proj.w", sd[r + "attn.proj.weight"]); put(p + "proj.b", sd[r + "attn.proj.bias"])
                put(p + "fc1.w", E.interleave8(sd[r + "mlp.w1.weight"], sd[r + "mlp.w2.weight"]))
                put(p + "fc1.b", E.interleave8(sd[r + "mlp.w1.bias"], sd[r + "mlp.w2.bias"]))
                put(p + "fc2.w", sd[r + "mlp.w3.weight"]); put(p + "fc2.b", sd[r + "mlp.w3.bias"])
            put(pre + "norm_w", sd[pre_ref + "norm.weight"])
            if ln: put(pre + "norm_b", sd[pre_ref + "norm.bias"])

        c = self.cfg
        put("trunk.patch.w", sd["trunk.patch_embed.proj.weight"].flatten(1)); put("trunk.patch.b", sd["trunk.patch_embed.proj.bias"])
        put("trunk.cls", sd["trunk.cls_token"]); put("trunk.mask_token", sd["trunk.mask_token"])
        vit("trunk.", "trunk.", c.vision_depth, False)
        put("trunk.bneck.w", sd["trunk.feature_bottleneck.weight"])
        put("visual_proj.w", sd["visual_proj.weight"])
        put("decoder.proj_in.w", sd["pixel_decoder.proj_in.weight"].flatten(1)); put("decoder.proj_in.b", sd["pixel_decoder.proj_in.bias"])
        vit("pixel_decoder.", "decoder.", c.decoder_depth, True)
        put("decoder.proj_out.w", sd["pixel_decoder.proj_out.weight"].flatten(1)); put("decoder.proj_out.b", sd["pixel_decoder.proj_out.bias"])
        put("text.tok_emb", sd["token_embedding.weight"]); put("text.pos", sd["positional_embedding"])
        for i in range(c.text_depth):
            r, p = f"text_transformer.resblocks.{i}.", f"text.blocks.{i}."
            put(p + "n1_w", sd[r + "ln_1.weight"]); put(p + "n1_b", sd[r + "ln_1.bias"])
            put(p + "n2_w", sd[r + "ln_2.weight"]); put(p + "n2_b", sd[r + "ln_2.bias"])
            put(p + "qkv.w", sd[r + "attn.in_proj_weight"]); put(p + "qkv.b", sd[r + "attn.in_proj_bias"])
            put(p + "proj.w", sd[r + "attn.out_proj.weight"]); put(p + "proj.b", sd[r + "attn.out_proj.bias"])
            put(p + "fc1.w", sd[r + "mlp.c_fc.weight"]); put(p + "fc1.b", sd[r + "mlp.c_fc.bias"])
            put(p + "fc2.w", sd[r + "mlp.c_proj.weight"]); put(p + "fc2.b", sd[r + "mlp.c_proj.bias"])
        put("text.norm_w", sd["ln_final.weight"]); put("text.norm_b", sd["ln_final.bias"])
        put("text.proj.w", sd["text_projection"].t())
        put("logit_scale", sd["logit_scale"].reshape(1))
        if head_sd is not None:
            for j in (0, 2, 4):
                put(f"head.mlp{j}.w", head_sd[f"mlp.{j}.weight"]); put(f"head.mlp{j}.b", head_sd[f"mlp.{j}.bias"])
            gk = "last_layer.weight_g" if "last_layer.weight_g" in head_sd else "last_layer.parametrizations.weight.original0"
            vk = "last_layer.weight_v" if "last_layer.weight_v" in head_sd else "last_layer.parametrizations.weight.original1"
            put("head.last_g", head_sd[gk].reshape(-1)); put("head.last_v", head_sd[vk])
        st.sync_compute_copies(init_teacher=True)

    @torch.no_grad()
    def export_state_dict(self) -> Dict[str, torch.Tensor]:
        """Student weights in the reference's VTPModel state-dict format."""
        st, c = self.store, self.cfg
        out = {}

        def de8(t):
            n = t.shape[0] // 2
            v = t.reshape(n // 8, 2, 8, *t.shape[1:])
            return v[:, 0].reshape(n, *t.shape[1:]).clone(), v[:, 1].reshape(n, *t.shape[1:]).clone()

        def vit(pre_ref, pre, depth, ln):
            for i in range(depth):
                r, p = f"{pre_ref}blocks.{i}.", f"{pre}blocks.{i}."
                out[r + "norm1.weight"] = st.f32(p + "n1_w").clone(); out[r + "norm2.weight"] = st.f32(p + "n2_w").clone()
                if ln:
                    out[r + "norm1.bias"] = st.f32(p + "n1_b").clone(); out[r + "norm2.bias"] = st.f32(p + "n2_b").clone()
                out[r + "attn.qkv.weight"] = st.f32(p + "qkv.w").clone(); out[r + "attn.qkv.bias"] = st.f32(p + "qkv.b").clone()
                out[r + "attn.proj.weight"] = st.f32(p + "proj.w").clone(); out[r + "attn.proj.bias"] = st.f32(p + "proj.b").clone()
                out[r + "mlp.w1.weight"], out[r + "mlp.w2.weight"] = de8(st.f32(p + "fc1.w"))
                out[r + "mlp.w1.bias"], out[r + "mlp.w2.bias"] = de8(st.f32(p + "fc1.b"))
                out[r + "mlp.w3.weight"] = st.f32(p + "fc2.w").clone(); out[r + "mlp.w3.bias"] = st.f32(p + "fc2.b").clone()
            out[pre_ref + "norm.weight"] = st.f32(pre + "norm_w").clone()
            if ln: out[pre_ref + "norm.bias"] = st.f32(pre + "norm_b").clone()

        ps = c.vision_patch_size
        out["trunk.patch_embed.proj.weight"] = st.f32("trunk.patch.w").reshape(self.D, 3, ps, ps).clone()
        out["trunk.patch_embed.proj.bias"] = st.f32("trunk.patch.b").clone()
        out["trunk.cls_token"] = st.f32("trunk.cls").reshape(1, 1, -1).clone()
        out["trunk.mask_token"] = st.f32("trunk.mask_token").reshape(1, -1).clone()
        vit("trunk.", "trunk.", c.vision_depth, False)
        out["trunk.feature_bottleneck.weight"] = st.f32("trunk.bneck.w").clone()
        out["visual_proj.weight"] = st.f32("visual_proj.w").clone()
        out["pixel_decoder.proj_in.weight"] = st.f32("decoder.proj_in.w").reshape(self.Dd, self.bn, 1, 1).clone()
        out["pixel_decoder.proj_in.bias"] = st.f32("decoder.proj_in.b").clone()
        vit("pixel_decoder.", "decoder.", c.decoder_depth, True)
        out["pixel_decoder.proj_out.weight"] = st.f32("decoder.proj_out.w").reshape(768, self.Dd, 1, 1).clone()
        out["pixel_decoder.proj_out.bias"] = st.f32("decoder.proj_out.b").clone()
        out["token_embedding.weight"] = st.f32("text.tok_emb").clone(); out["positional_embedding"] = st.f32("text.pos").clone()
        for i in range(c.text_depth):
            r, p = f"text_transformer.resblocks.{i}.", f"text.blocks.{i}."
            for a, b_ in (("ln_1.weight", "n1_w"), ("ln_1.bias", "n1_b"), ("ln_2.weight", "n2_w"), ("ln_2.bias", "n2_b"),
                          ("attn.in_proj_weight", "qkv.w"), ("attn.in_proj_bias", "qkv.b"), ("attn.out_proj.weight", "proj.w"),
                          ("attn.out_proj.bias", "proj.b"), ("mlp.c_fc.weight", "fc1.w"), ("mlp.c_fc.bias", "fc1.b"),
                          ("mlp.c_proj.weight", "fc2.w"), ("mlp.c_proj.bias", "fc2.b")):
                out[r + a] = st.f32(p + b_).clone()
        out["ln_final.weight"] = st.f32("text.norm_w").clone(); out["ln_final.bias"] = st.f32("text.norm_b").clone()
        out["text_projection"] = st.f32("text.proj.w").t().contiguous()
        out["logit_scale"] = st.f32("logit_scale").reshape(()).clone()
        from .rope import rope_periods
        out["trunk.rope_embed.periods"] = rope_periods(64).to(self.device)
        out["pixel_decoder.rope_embed.periods"] = rope_periods(64).to(self.device)
        return out

    # -------------------------------------------------------------- pieces shared by the objectives
    def _drop(self, ratio: float):
        """DropPlan of one student trunk pass (None when the objective's drop rate is 0); `drop_presets` (tests) supplies
        fixed subsets instead of random permutations."""
        if ratio <= 0.0:
            return None
        preset = self.drop_presets.pop(0) if getattr(self, "drop_presets", None) else None
        return E.DropPlan(ratio, self.world, self.rank, preset=preset)

    def _trunk_fwd(self, W: TowerW, img, tape: Optional[dict], mask_idx=None, drop=None):
        x, meta = E.trunk_forward(W, img, "bf16", mask_idx=mask_idx, tape=tape, drop=drop)
        return x, meta

    def _trunk_bwd(self, tape: dict, g: torch.Tensor, mask_idx=None):
        """g fp32 [B*T, D] = dL/d(x_prenorm) -> accumulates all trunk parameter gradients."""
        W, G = self.towers[("trunk", "param")], self.towers[("trunk", "grad")]
        B, T, gh, gw = tape["meta"]
        D, HW = W.D, gh * gw
        rope = W.rope(gh, gw, g.device)
        tower_blocks_backward(W, G, tape["blocks"], g, B, T, rope)
        gp = _e((B * HW, D), BF, g.device)
        lib.strip_prefix(g, gp, G.extra["cls"], B, T, 1, D)
        if mask_idx is not None and mask_idx.numel() > 0:
            rows = (mask_idx // HW) * T + 1 + mask_idx % HW
            tmp = _e((mask_idx.numel(), D), F32, g.device)
            lib.gather_rows(g, tmp, rows, D)
            lib.cast_colsum(tmp, None, G.extra["mask_token"], mask_idx.numel(), D)
            zeros = torch.zeros(D, dtype=F32, device=g.device)
            lib.apply_mask_tokens(gp, zeros, mask_idx, HW, HW, 0, D)
        wgrad(gp, tape["patch_a"], G.extra["patch"].w, B * HW)
        lib.cast_colsum(gp, None, G.extra["patch"].b, B * HW, D)

    # -------------------------------------------------------------- objective 1: CLIP (vtp.py:340-363 + ClipLoss)
    def clip_fwd_bwd(self, image: torch.Tensor, text: torch.Tensor, weight: float = 1.0):
        import torch.distributed as dist
        dev = self.device
        W, G = self.towers[("trunk", "param")], self.towers[("trunk", "grad")]
        Wt, Gt = self.towers[("text", "param")], self.towers[("text", "grad")]
        D, Dt = self.D, self.Dt
        # ---- image tower -> cls -> visual_proj -> normalise
        tp_i = {}
        x, (B, T, gh, gw) = self._trunk_fwd(W, image, tp_i, drop=self._drop(self.tc.clip_drop_rate))
        M = B * T
        nt = {}
        xn = E.norm(x, M, D, W.norm_w, None, W.eps, "bf16", want="f32", tape=nt)
        cls_rows = torch.arange(B, device=dev, dtype=torch.long) * T
        cls = _e((B, D), BF, dev)
        lib.gather_rows(xn, cls, cls_rows, D)
        vp: Lin = W.extra["visual_proj"]
        fi_raw = _e((B, Dt), BF, dev)
        lib.gemm(cls, vp.w, fi_raw, M=B, N=Dt, K=D)
        nrm_i = _e((B,), F32, dev)
        # ---- text tower
        tp_t = {}
        ft_raw = E.text_forward(Wt, text, "bf16", tape=tp_t)
        nrm_t = _e((B,), F32, dev)
        if self._clip_exchange() == "p2p":
            dfi, dft = self._clip_loss_p2p(fi_raw, ft_raw, nrm_i, nrm_t, B, weight)
            self._clip_backward(tp_i, tp_t, x, nt, cls, self.peer.img, self.peer.txt, nrm_i, nrm_t, dfi, dft, text, B, T)
            return
        fi = torch.empty_like(fi_raw)
        lib.l2norm_fwd(fi_raw, fi, B, Dt, 1e-12, norm_out=nrm_i)
        ft = torch.empty_like(ft_raw)
        lib.l2norm_fwd(ft_raw, ft, B, Dt, 1e-12, norm_out=nrm_t)
        # ---- contrastive loss over the global batch (features all-gathered: collective C2)
        if self.world > 1:
            fi_all = _e((self.world * B, Dt), BF, dev)
            ft_all = _e((self.world * B, Dt), BF, dev)
            dist.all_gather_into_tensor(fi_all, fi, group=self.pg)
            dist.all_gather_into_tensor(ft_all, ft, group=self.pg)
        else:
            fi_all, ft_all = fi, ft
        Bg = fi_all.shape[0]
        Bgp = (Bg + 7) // 8 * 8
        sim_i = _e((B, Bgp), F32, dev)
        sim_t = _e((B, Bgp), F32, dev)
        pad = lambda t: t if Bg == Bgp else torch.cat([t, torch.zeros(Bgp - Bg, Dt, dtype=BF, device=dev)])
        fi_allp, ft_allp = pad(fi_all), pad(ft_all)
        lib.gemm(fi, ft_allp, sim_i, M=B, N=Bgp, K=Dt, round_bf16=False)
        lib.gemm(ft, fi_allp, sim_t, M=B, N=Bgp, K=Dt, round_bf16=False)
        Gi = torch.zeros((B, Bgp), dtype=BF, device=dev)
        Gt_ = torch.zeros((B, Bgp), dtype=BF, device=dev)
        ls = self.store.f32("logit_scale")
        dls = self.store.grad("logit_scale")
        coef = weight * 0.5 / B
        lib.softmax_ce(sim_i, B, Bg, self.rank * B, Gi, coef, self.loss_acc[0:1], dls, log_scale=ls)
        lib.softmax_ce(sim_t, B, Bg, self.rank * B, Gt_, coef, self.loss_acc[0:1], dls, log_scale=ls)
        # d f_i(local) = Gi · T_all + [Gtᵀ · T_local]_(all ranks summed, local slice)   (and symmetrically for text)
        dfi = _e((B, Dt), F32, dev)
        dft = _e((B, Dt), F32, dev)
        lib.gemm(Gi, ft_allp, dfi, M=B, N=Dt, K=Bgp, b_mn=True, ldb=Dt, round_bf16=False)
        lib.gemm(Gt_, fi_allp, dft, M=B, N=Dt, K=Bgp, b_mn=True, ldb=Dt, round_bf16=False)
        if self.world > 1:
            cross_i = _e((Bgp, Dt), F32, dev)
            cross_t = _e((Bgp, Dt), F32, dev)
            lib.gemm(Gt_, ft, cross_i, M=Bgp, N=Dt, K=B, a_mn=True, b_mn=True, lda=Bgp, ldb=Dt, round_bf16=False)
            lib.gemm(Gi, fi, cross_t, M=Bgp, N=Dt, K=B, a_mn=True, b_mn=True, lda=Bgp, ldb=Dt, round_bf16=False)
            dist.all_reduce(cross_i, group=self.pg)
            dist.all_reduce(cross_t, group=self.pg)
            lib.axpby(dfi, cross_i[self.rank * B:(self.rank + 1) * B].contiguous(), 1.0, 1.0, B * Dt)
            lib.axpby(dft, cross_t[self.rank * B:(self.rank + 1) * B].contiguous(), 1.0, 1.0, B * Dt)
        else:
            lib.gemm(Gt_, ft, dfi, M=B, N=Dt, K=B, a_mn=True, b_mn=True, lda=Bgp, ldb=Dt, accumulate=True, round_bf16=False)
            lib.gemm(Gi, fi, dft, M=B, N=Dt, K=B, a_mn=True, b_mn=True, lda=Bgp, ldb=Dt, accumulate=True, round_bf16=False)
        self._clip_backward(tp_i, tp_t, x, nt, cls, fi, ft, nrm_i, nrm_t, dfi, dft, text, B, T)

    def _clip_exchange(self) -> str:
        import os
        mode = os.environ.get("VTP_CLIP_EXCHANGE", self.tc.clip_exchange)
        if mode not in ("nccl", "p2p"):
            raise ValueError(f"clip_exchange must be 'nccl' or 'p2p', got {mode!r}")
        return mode

    def _clip_loss_p2p(self, fi_raw, ft_raw, nrm_i, nrm_t, B: int, weight: float):
        """Collective C2 over NVLink peer memory (csrc/clip.cu): the normalised features are written straight into this
        rank's exchange buffer, one kernel gathers every rank's rows and forms the full Bg x Bg logits, two small kernels
        do both softmax directions; returns d(Σ_ranks L_local)/d(f_local) for images and captions (fp32 [B, E])."""
        from .comm import PeerFeatures
        dev, Dt = self.device, self.Dt
        if self.peer is None or self.peer.B != B:
            if self.peer is not None:
                self.peer.check()            # a timed-out barrier of the old buffer must not be dropped silently
                self.peer.close()
            self.peer = PeerFeatures(B, Dt, dev, self.pg, world=self.world, rank=self.rank)
        pf = self.peer
        lib.l2norm_fwd(fi_raw, pf.img, B, Dt, 1e-12, norm_out=nrm_i)
        lib.l2norm_fwd(ft_raw, pf.txt, B, Dt, 1e-12, norm_out=nrm_t)
        Bg = self.world * B
        Bgp = (Bg + 7) // 8 * 8
        S = _e((Bg, Bgp), F32, dev)
        St = _e((Bg, Bgp), F32, dev)
        fi_all = torch.zeros((Bgp, Dt), dtype=BF, device=dev)
        ft_all = torch.zeros((Bgp, Dt), dtype=BF, device=dev)
        pf.barrier(self.loss_acc[0:1])     # every rank's features are written (time-out -> NaN contrastive loss)
        lib.clip_gather_logits(pf.img_ptrs, pf.txt_ptrs, B, Dt, S, St, fi_all, ft_all)
        pf.barrier(self.loss_acc[0:1])     # every rank has read them: the buffers may be overwritten by the next step
        ls = self.store.f32("logit_scale")
        dls = self.store.grad("logit_scale")
        coef = weight * 0.5 / B
        lse = _e((2, Bg), F32, dev)
        lib.clip_lse(S, St, Bg, self.rank * B, B, ls, coef, lse, self.loss_acc[0:1], dls)
        dMi = _e((B, Bgp), BF, dev)
        dMt = _e((B, Bgp), BF, dev)
        lib.clip_grad(S, St, Bg, self.rank * B, B, ls, coef, lse, dMi, dMt)
        dfi = _e((B, Dt), F32, dev)
        dft = _e((B, Dt), F32, dev)
        lib.gemm(dMi, ft_all, dfi, M=B, N=Dt, K=Bgp, b_mn=True, ldb=Dt, round_bf16=False)
        lib.gemm(dMt, fi_all, dft, M=B, N=Dt, K=Bgp, b_mn=True, ldb=Dt, round_bf16=False)
        return dfi, dft

    def _clip_backward(self, tp_i, tp_t, x, nt, cls, fi, ft, nrm_i, nrm_t, dfi, dft, text, B: int, T: int):
        dev = self.device
        W, G = self.towers[("trunk", "param")], self.towers[("trunk", "grad")]
        D, Dt = self.D, self.Dt
        M = B * T
        vp: Lin = W.extra["visual_proj"]
        # ---- image side backward
        dfi_raw = _e((B, Dt), BF, dev)
        lib.l2norm_bwd(fi, nrm_i, dfi, dfi_raw, B, Dt)
        dxn = torch.zeros((M, D), dtype=BF, device=dev)
        dgrad(dfi_raw, vp.w, dxn, B, ldo=T * D)          # row b of d(cls) lands on token row b*T
        wgrad(dfi_raw, cls, G.extra["visual_proj"].w, B)
        g = torch.zeros((M, D), dtype=F32, device=dev)
        lib.norm_bwd(x, nt["rstd"], None, W.norm_w, dxn, g, G.norm_w, None, M, D)
        self._trunk_bwd(tp_i, g)
        # ---- text side backward
        self._text_bwd(tp_t, ft, nrm_t, dft, text)

    def _text_bwd(self, tape, ft, nrm_t, dft, ids):
        dev = self.device
        Wt, Gt = self.towers[("text", "param")], self.towers[("text", "grad")]
        B, L = tape["meta"]
        Dt, M = self.Dt, B * L
        dft_raw = _e((B, Dt), BF, dev)
        lib.l2norm_bwd(ft, nrm_t, dft, dft_raw, B, Dt)
        proj: Lin = Wt.extra["proj"]
        dpool = _e((B, Dt), F32, dev)
        dgrad(dft_raw, proj.w, dpool, B)
        wgrad(dft_raw, tape["pooled"], Gt.extra["proj"].w, B)
        dxn32 = torch.zeros((M, Dt), dtype=F32, device=dev)
        lib.scatter_add_rows(dpool, dxn32, tape["eot"], Dt)
        dxn = _e((M, Dt), BF, dev)
        lib.cast_colsum(dxn32, dxn, None, M, Dt)
        g = torch.zeros((M, Dt), dtype=F32, device=dev)
        nf = tape["nf"]
        lib.norm_bwd(tape["x_final"], nf["rstd"], nf["mean"], Wt.norm_w, dxn, g, Gt.norm_w, Gt.norm_b, M, Dt)
        tower_blocks_backward(Wt, Gt, tape["blocks"], g, B, L, None, causal=True)
        lib.scatter_add_rows(g, Gt.extra["tok_emb"], ids.reshape(-1), Dt)
        lib.cast_colsum(g, None, Gt.extra["pos"].view(-1), B, L * Dt, ldx=L * Dt)

    # -------------------------------------------------------------- objective 3: reconstruction (vtp.py:487-512)
    def rec_fwd_bwd(self, image: torch.Tensor, weight: float = 1.0, return_image: bool = False,
                    norm_B: Optional[int] = None, final_group: bool = False):
        """norm_B: batch size the loss is normalised by (the whole per-GPU batch when `image` is one chunk of it)."""
        dev = self.device
        W, G = self.towers[("trunk", "param")], self.towers[("trunk", "grad")]
        Wd, Gd = self.towers[("decoder", "param")], self.towers[("decoder", "grad")]
        D, Dd, bn = self.D, self.Dd, self.bn
        tp_e = {}
        x, (B, T, gh, gw) = self._trunk_fwd(W, image, tp_e, drop=self._drop(self.tc.rec_drop_rate))
        M, HW = B * T, gh * gw
        Md = B * HW
        nB = norm_B if norm_B is not None else B
        nt = {}
        xn = E.norm(x, M, D, W.norm_w, None, W.eps, "bf16", want="op", tape=nt)
        bneck: Lin = W.extra["bneck"]
        tok = _e((Md, bn), BF, dev)                      # latents as decoder tokens (cls rows dropped in the epilogue)
        lib.gemm(xn, bneck.w, tok, M=M, N=bn, K=D, rr_group=T, rr_skip=-1)
        # ---- decoder (decoders/pixel_decoder.py:134-162) on token-major latents
        pin: Lin = Wd.extra["proj_in"]
        xd = _e((Md, Dd), BF, dev)
        lib.gemm(tok, pin.w, xd, M=Md, N=Dd, K=bn, bias=pin.b)
        rope = Wd.rope(gh, gw, dev)
        dtape = []
        xd = E.tower_blocks(Wd, xd, B, HW, rope, "bf16", tape=dtape)
        ntd = {}
        xdn = E.norm(xd, Md, Dd, Wd.norm_w, Wd.norm_b, Wd.eps, "bf16", want="op", tape=ntd)
        pout: Lin = Wd.extra["proj_out"]
        r = 16
        rec = _e((B, 3, gh * r, gw * r), BF, dev)
        lib.gemm(xdn, pout.w, rec, M=Md, N=pout.N, K=Dd, bias=pout.b, pixel_shuffle=(r, gh, gw, 3), ldo=gw * r)
        # ---- loss: L1 (+ LPIPS gradient if a perceptual module is attached)
        dlp = None
        if getattr(self, "lpips", None) is not None and self.tc.lpips_weight > 0:
            dlp = self.lpips.loss_and_grad(rec, image, weight * self.tc.lpips_weight / nB, self.loss_acc[5:6])
        dY = _e((Md, pout.N), BF, dev)
        lib.recon_l1_grad(rec, image.contiguous(), dlp, dY, self.loss_acc[4:5], B, 3, gh, gw, r,
                          weight / (rec.numel() // B * nB))
        # ---- decoder backward
        dxdn = _e((Md, Dd), BF, dev)
        dgrad(dY, pout.w, dxdn, Md)
        wgrad(dY, xdn, Gd.extra["proj_out"].w, Md)
        lib.cast_colsum(dY, None, Gd.extra["proj_out"].b, Md, pout.N)
        g = torch.zeros((Md, Dd), dtype=F32, device=dev)
        lib.norm_bwd(xd, ntd["rstd"], ntd["mean"], Wd.norm_w, dxdn, g, Gd.norm_w, Gd.norm_b, Md, Dd)
        tower_blocks_backward(Wd, Gd, dtape, g, B, HW, rope)
        gb = _e((Md, Dd), BF, dev)
        lib.cast_colsum(g, gb, Gd.extra["proj_in"].b, Md, Dd)
        dz = torch.zeros((M, bn), dtype=BF, device=dev)  # d(latent tokens), re-expanded to [B*T] rows (cls rows = 0)
        dgrad(gb, pin.w, dz, Md, rr_group=HW, rr_skip=1)
        wgrad(gb, tok, Gd.extra["proj_in"].w, Md)
        if final_group:
            self._reduce_bucket("decoder")   # pixel-decoder gradient is final: all-reduce under the encoder backward
        # ---- encoder side
        dxn = _e((M, D), BF, dev)
        dgrad(dz, bneck.w, dxn, M)
        wgrad(dz, xn, G.extra["bneck"].w, M)
        ge = torch.zeros((M, D), dtype=F32, device=dev)
        lib.norm_bwd(x, nt["rstd"], None, W.norm_w, dxn, ge, G.norm_w, None, M, D)
        self._trunk_bwd(tp_e, ge)
        return rec if return_image else None

    # -------------------------------------------------------------- DINO head (heads/dino_head.py:65-126)
    def _head_prepare(self):
        W, Wt = self.towers[("trunk", "param")], self.towers[("trunk", "teacher")]
        K, hb = self.tc.head_out_dim, self.tc.head_bottleneck
        lib.weight_norm_fwd(W.extra["last_v"], W.extra["last_g"], self.head_wn, self.head_vnorm, K, hb)
        lib.weight_norm_fwd(Wt.extra["last_v"], Wt.extra["last_g"], self.head_wn_t, None, K, hb)

    def _head_fwd(self, W: TowerW, wn: torch.Tensor, x: torch.Tensor, tape: Optional[dict]):
        dev = self.device
        Tn, D = x.shape
        hh, hb, K = self.tc.head_hidden, self.tc.head_bottleneck, self.tc.head_out_dim
        m0, m2, m4 = W.extra["mlp0"], W.extra["mlp2"], W.extra["mlp4"]
        pre1 = _e((Tn, hh), BF, dev) if tape is not None else None
        h1 = _e((Tn, hh), BF, dev)
        lib.gemm(x, m0.w, h1, M=Tn, N=hh, K=D, bias=m0.b, act=lib.ACT_GELU, out2=pre1)
        pre2 = _e((Tn, hh), BF, dev) if tape is not None else None
        h2 = _e((Tn, hh), BF, dev)
        lib.gemm(h1, m2.w, h2, M=Tn, N=hh, K=hh, bias=m2.b, act=lib.ACT_GELU, out2=pre2)
        h3 = _e((Tn, hb), BF, dev)
        lib.gemm(h2, m4.w, h3, M=Tn, N=hb, K=hh, bias=m4.b)
        y = _e((Tn, hb), BF, dev)
        nrm = _e((Tn,), F32, dev)
        lib.l2norm_fwd(h3, y, Tn, hb, 1e-12, norm_out=nrm)
        logits = _e((Tn, K), BF, dev)
        lib.gemm(y, wn, logits, M=Tn, N=K, K=hb)
        if tape is not None:
            tape.update(x=x, pre1=pre1, h1=h1, pre2=pre2, h2=h2, y=y, nrm=nrm)
        return logits

    def _head_bwd(self, tape: dict, dlogits: torch.Tensor) -> torch.Tensor:
        dev = self.device
        W, G = self.towers[("trunk", "param")], self.towers[("trunk", "grad")]
        Tn = dlogits.shape[0]
        hh, hb, K, D = self.tc.head_hidden, self.tc.head_bottleneck, self.tc.head_out_dim, self.D
        dy = _e((Tn, hb), F32, dev)
        dgrad(dlogits, self.head_wn, dy, Tn)
        dWn = torch.zeros((K, hb), dtype=F32, device=dev)
        wgrad(dlogits, tape["y"], dWn, Tn)
        lib.weight_norm_bwd(W.extra["last_v"], W.extra["last_g"], self.head_vnorm, dWn, G.extra["last_v"],
                            G.extra["last_g"], K, hb)
        dh3 = _e((Tn, hb), BF, dev)
        lib.l2norm_bwd(tape["y"], tape["nrm"], dy, dh3, Tn, hb)
        m0, m2, m4 = W.extra["mlp0"], W.extra["mlp2"], W.extra["mlp4"]
        g0, g2, g4 = G.extra["mlp0"], G.extra["mlp2"], G.extra["mlp4"]
        lib.cast_colsum(dh3, None, g4.b, Tn, hb)
        dh2 = _e((Tn, hh), BF, dev)
        dgrad(dh3, m4.w, dh2, Tn)
        wgrad(dh3, tape["h2"], g4.w, Tn)
        dpre2 = _e((Tn, hh), BF, dev)
        lib.gelu_bwd(tape["pre2"], dh2, dpre2, g2.b, Tn, hh)
        dh1 = _e((Tn, hh), BF, dev)
        dgrad(dpre2, m2.w, dh1, Tn)
        wgrad(dpre2, tape["h1"], g2.w, Tn)
        dpre1 = _e((Tn, hh), BF, dev)
        lib.gelu_bwd(tape["pre1"], dh1, dpre1, g0.b, Tn, hh)
        dx = _e((Tn, D), BF, dev)
        dgrad(dpre1, m0.w, dx, Tn)
        wgrad(dpre1, tape["x"], g0.w, Tn)
        return dx

    # -------------------------------------------------------------- objective 2: SSL (vtp.py:365-386,410-484)
    def split_masks(self, mask_indices: torch.Tensor, masks_weight: torch.Tensor, B: int, HW: int) -> Dict[str, torch.Tensor]:
        """Per-image-group mask lists for TrainConfig.ssl_chunk (data dependent, so it synchronises): call it once per
        batch in the input pipeline and pass the result on as batch keys `mask_indices@i` / `masks_weight@i`; the step
        itself then contains no data-dependent shape and can be captured in a CUDA graph."""
        chunk = self.tc.ssl_chunk if 0 < self.tc.ssl_chunk < B else B
        out = {}
        if chunk == B:
            return out
        img = mask_indices // HW
        for i, b0 in enumerate(range(0, B, chunk)):
            b1 = min(B, b0 + chunk)
            bc = b1 - b0
            s0 = (img >= b0) & (img < b1)
            s1 = (img >= B + b0) & (img < B + b1)
            out[f"mask_indices@{i}"] = torch.cat([mask_indices[s0] - b0 * HW, mask_indices[s1] - (B + b0 - bc) * HW])
            out[f"masks_weight@{i}"] = torch.cat([masks_weight[s0], masks_weight[s1]])
        return out

    def ssl_fwd_bwd(self, global_crops, local_crops, mask_indices, masks_weight, weight: float = 1.0, mask_groups=None):
        """global_crops [2B,3,H,W] (view-major), local_crops [n_local*B,3,h,w] (crop-major), mask_indices int64 flat
        indices into [2B*HW] of masked global patches (ascending), masks_weight [n_masked] = 1/(#masked in that image).
        With TrainConfig.ssl_chunk = c the B source images are processed c at a time (all their crops together): the
        teacher targets use last step's centre either way, so chunking only changes floating-point summation order."""
        import torch.distributed as dist
        dev, tc = self.device, self.tc
        K = tc.head_out_dim
        self._head_prepare()
        B2 = global_crops.shape[0]
        B = B2 // 2
        n_m = mask_indices.numel()
        csum = torch.zeros((2, K), dtype=F32, device=dev)   # Σ raw teacher logits: [cls rows, masked-patch rows]
        chunk = tc.ssl_chunk if 0 < tc.ssl_chunk < B else B
        if chunk == B:
            self._ssl_chunk(global_crops, local_crops, mask_indices, masks_weight, weight, B, csum)
        else:
            ps = self.cfg.vision_patch_size
            HW = (global_crops.shape[-2] // ps) * (global_crops.shape[-1] // ps)
            n_loc = tc.n_local_crops
            lc = local_crops.view(n_loc, B, *local_crops.shape[1:])
            if mask_groups is None or "mask_indices@0" not in mask_groups:
                mask_groups = self.split_masks(mask_indices, masks_weight, B, HW)   # synchronises (boolean indexing)
            for i, b0 in enumerate(range(0, B, chunk)):
                b1 = min(B, b0 + chunk)
                bc = b1 - b0
                g = torch.cat([global_crops[b0:b1], global_crops[B + b0:B + b1]])
                l = lc[:, b0:b1].reshape(n_loc * bc, *local_crops.shape[1:])
                self._ssl_chunk(g, l, mask_groups[f"mask_indices@{i}"], mask_groups[f"masks_weight@{i}"], weight, B, csum,
                                final_group=(b1 == B))
        # teacher centre EMA over the whole (global) batch (DINOv2 softmax_center_teacher / update_center)
        cnt = torch.empty(2, dtype=F32, device=dev)          # fill kernels, no host copy: the step is graph-capturable
        cnt[0:1].fill_(float(B2)), cnt[1:2].fill_(float(max(n_m, 1)))
        if self.world > 1:
            dist.all_reduce(csum, group=self.pg)
            dist.all_reduce(cnt, group=self.pg)
        cm = tc.center_momentum
        mean = csum / cnt[:, None]
        lib.axpby(self.center_dino, mean[0].contiguous(), cm, 1 - cm, K)
        lib.axpby(self.center_ibot, mean[1].contiguous(), cm, 1 - cm, K)

    def _ssl_chunk(self, global_crops, local_crops, mask_indices, masks_weight, weight: float, norm_B: int, csum,
                   final_group: bool = True):
        """Teacher + student forward, DINO/iBOT losses and the full backward for one group of source images; the loss
        terms are normalised by `norm_B` (the whole per-GPU batch) and the raw teacher-logit sums are added to csum."""
        dev, tc = self.device, self.tc
        W, G, Wt = self.towers[("trunk", "param")], self.towers[("trunk", "grad")], self.towers[("trunk", "teacher")]
        D, K = self.D, tc.head_out_dim
        n_loc = tc.n_local_crops
        B2 = global_crops.shape[0]
        B = B2 // 2
        n_m = mask_indices.numel()
        # ---------------- teacher (no grad): get_teacher_forward_outputs vtp.py:410-450
        xt, (_, T, gh, gw) = self._trunk_fwd(Wt, global_crops, None)
        HW = gh * gw
        xnt = E.norm(xt, B2 * T, D, Wt.norm_w, None, Wt.eps, "bf16", want="f32")
        ar = torch.arange(B2, device=dev, dtype=torch.long)
        cls_rows = ar * T
        swapped = torch.cat([cls_rows[B:], cls_rows[:B]])                       # cat(chunk[1], chunk[0])
        m_rows = (mask_indices // HW) * T + 1 + mask_indices % HW
        Tt = B2 + n_m
        tin = _e((Tt, D), BF, dev)
        lib.gather_rows(xnt, tin, torch.cat([swapped, m_rows]), D)
        tlog = self._head_fwd(Wt, self.head_wn_t, tin, None)
        # centre statistics (sums of raw teacher logits), then centred + sharpened softmax in place
        lib.cast_colsum(tlog, None, csum[0], B2, K)
        if n_m:
            lib.cast_colsum(tlog[B2:], None, csum[1], n_m, K)
        lib.dino_teacher_probs(tlog, self.center_dino, B2, K, tc.teacher_temp)
        if n_m:
            lib.dino_teacher_probs(tlog[B2:], self.center_ibot, n_m, K, tc.teacher_temp)
        del xt, xnt
        # ---------------- student: get_student_ssl_outputs vtp.py:452-484
        tp_g, tp_l = {}, {}
        xg, _ = self._trunk_fwd(W, global_crops, tp_g, mask_idx=mask_indices, drop=self._drop(tc.ssl_drop_rate))
        xl, (Bl, Tl, ghl, gwl) = self._trunk_fwd(W, local_crops, tp_l, drop=self._drop(tc.ssl_drop_rate))
        ntg, ntl = {}, {}
        xng = E.norm(xg, B2 * T, D, W.norm_w, None, W.eps, "bf16", want="f32", tape=ntg)
        xnl = E.norm(xl, Bl * Tl, D, W.norm_w, None, W.eps, "bf16", want="f32", tape=ntl)
        l_rows = torch.arange(Bl, device=dev, dtype=torch.long) * Tl
        Ts = Bl + B2 + n_m
        sin_ = _e((Ts, D), BF, dev)
        lib.gather_rows(xnl, sin_[:Bl], l_rows, D)
        lib.gather_rows(xng, sin_[Bl:], torch.cat([cls_rows, m_rows]), D)
        htape = {}
        slog = self._head_fwd(W, self.head_wn, sin_, htape)
        # ---------------- losses (DINOv2 DINOLoss / iBOTPatchLoss, see oracle.dino_ibot_loss)
        n_terms = 2 + 2 * n_loc
        bidx = torch.arange(B, device=dev, dtype=torch.int32)
        t0 = torch.cat([bidx.repeat(n_loc), torch.arange(B2, device=dev, dtype=torch.int32),
                        B2 + torch.arange(n_m, device=dev, dtype=torch.int32)])
        t1 = torch.cat([(B + bidx).repeat(n_loc), torch.full((B2 + n_m,), -1, device=dev, dtype=torch.int32)])
        wl = weight / (norm_B * n_terms)
        wrow = torch.cat([torch.full((Bl,), wl, device=dev), torch.full((B2,), wl, device=dev),
                          masks_weight.to(F32) * (weight / norm_B)])
        lib.dino_student_ce(slog[:Bl], tlog, t0[:Bl], t1[:Bl], wrow[:Bl], Bl, K, tc.student_temp, self.loss_acc[1:2])
        lib.dino_student_ce(slog[Bl:Bl + B2], tlog, t0[Bl:Bl + B2], t1[Bl:Bl + B2], wrow[Bl:Bl + B2], B2, K,
                            tc.student_temp, self.loss_acc[2:3])
        if n_m:
            lib.dino_student_ce(slog[Bl + B2:], tlog, t0[Bl + B2:], t1[Bl + B2:], wrow[Bl + B2:], n_m, K,
                                tc.student_temp, self.loss_acc[3:4])
        del tlog
        # ---------------- backward: head -> scatter to the two trunk passes
        dsin = self._head_bwd(htape, slog)
        del slog, htape
        if final_group:
            self._reduce_bucket("head")  # DINO head gradient is final: all-reduce under the trunk backward
        dl32 = torch.zeros((Bl * Tl, D), dtype=F32, device=dev)
        lib.scatter_add_rows(dsin[:Bl], dl32, l_rows, D)
        dxl = _e((Bl * Tl, D), BF, dev)
        lib.cast_colsum(dl32, dxl, None, Bl * Tl, D)
        gl = dl32.zero_()
        lib.norm_bwd(xl, ntl["rstd"], None, W.norm_w, dxl, gl, G.norm_w, None, Bl * Tl, D)
        self._trunk_bwd(tp_l, gl)
        del gl, dl32, dxl, xl, xnl, tp_l
        dg32 = torch.zeros((B2 * T, D), dtype=F32, device=dev)
        lib.scatter_add_rows(dsin[Bl:], dg32, torch.cat([cls_rows, m_rows]), D)
        dxg = _e((B2 * T, D), BF, dev)
        lib.cast_colsum(dg32, dxg, None, B2 * T, D)
        gg = dg32.zero_()
        lib.norm_bwd(xg, ntg["rstd"], None, W.norm_w, dxg, gg, G.norm_w, None, B2 * T, D)
        self._trunk_bwd(tp_g, gg, mask_idx=mask_indices)

    # -------------------------------------------------------------- optimiser (+ EMA teacher, vtp.py:388-401)
    def allreduce_grads(self):
        """Collective C3: whatever bucket has not been started yet (always the trunk), then wait for all of them; the
        sum is averaged by grad_scale = 1/world inside the fused optimiser.  fp32 on the wire (exact accumulation)."""
        if self.world > 1:
            import torch.distributed as dist
            if not self._started:        # "end" mode: nothing in flight — one call over the whole flat buffer
                dist.all_reduce(self.store.g, group=self.pg)
                return
            for which in ("text", "head", "decoder", "trunk"):
                self._reduce_bucket(which, final=True)
            for w in self._pending:
                w.wait()                 # the compute stream waits for NCCL's stream; the host does not block
            self._pending = []

    def set_schedules(self, lr=None, weight_decay=None, teacher_momentum=None):
        """Per-step schedules for the learning rate, weight decay and EMA teacher momentum: `schedules.CosineSchedule`
        objects (the reference's CosineScheduler, text_utils.py:160-207) or plain sequences, indexed by the optimiser
        step (0-based) and held at their last value afterwards.  They live on the device; the optimiser looks them up
        with its own step counter, also inside a captured graph.  None keeps the constant from TrainConfig."""
        from .schedules import as_table, pad_tables
        given = [(i, as_table(s)) for i, s in enumerate((lr, weight_decay, teacher_momentum)) if s is not None]
        tabs = [None, None, None]
        n = 0
        if given:
            padded = pad_tables([t for _, t in given])
            n = int(padded[0].size)
            for (i, _), t in zip(given, padded):
                tabs[i] = torch.from_numpy(t).to(self.device)
        self._sched = (tabs[0], tabs[1], tabs[2], n)
        if self._graph is not None:
            raise RuntimeError("set_schedules() after capture_step(): capture again (table pointers are part of the graph)")

    def scheduled_values(self) -> Dict[str, float]:
        """{step, lr, weight_decay, teacher_momentum} the LAST optimiser step used (synchronises)."""
        h = self.hyper.cpu()
        return {"step": int(h[0]), "lr": float(h[3]), "weight_decay": float(h[4]), "teacher_momentum": float(h[5])}

    def optimizer_step(self):
        st, tc = self.store, self.tc
        self.step_count += 1
        lr_t, wd_t, mom_t, n_t = self._sched
        lib.hyper_tick(self.hyper, tc.beta1, tc.beta2, lr_t, wd_t, mom_t, n_t)
        for start, end, decay, teacher in st.regions:
            n = end - start
            lib.adamw_step(st.p[start:end], st.g[start:end], st.m[start:end], st.v[start:end], st.pb[start:end],
                           st.tp[start:end] if teacher else None, st.tpb[start:end] if teacher else None, n,
                           lr=tc.lr, beta1=tc.beta1, beta2=tc.beta2, eps=tc.eps, wd=tc.weight_decay if decay else 0.0,
                           step=self.step_count, grad_scale=1.0 / self.world, ema_momentum=tc.teacher_momentum,
                           hyper=self.hyper)

    # -------------------------------------------------------------- CUDA graph of the whole step
    def capture_step(self, batch: Dict[str, torch.Tensor], warmup: int = 2):
        """Capture ONE full training step (three objectives, collectives, optimiser + EMA) into a CUDA graph that
        `replay_step` launches with a single call: ~2 300 kernel launches per VTP-Small step cost ~200 ms of host time,
        which is exposed whenever the caller synchronises per step (reading the loss).  The step state that changes from
        step to step (Adam bias corrections, scheduled lr / wd / momentum, teacher centres) lives on the device, so the
        replay needs no host scalar.  `warmup` REAL steps run on `batch` first (they train; lazy initialisation and the
        allocator settle), then one step is captured without executing.  Batches given to `replay_step` must have the
        shapes of `batch` (incl. the number of masked patches)."""
        if self._graph is not None:
            raise RuntimeError("a step graph exists already")
        self._static = {k: v.to(self.device).clone() for k, v in batch.items()}
        B = batch["global_crops"].shape[0] // 2
        if 0 < self.tc.ssl_chunk < B and "mask_indices@0" not in self._static:
            ps = self.cfg.vision_patch_size
            HW = (batch["global_crops"].shape[-2] // ps) * (batch["global_crops"].shape[-1] // ps)
            self._static.update(self.split_masks(self._static["mask_indices"], self._static["masks_weight"], B, HW))
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.train_step(self._static)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        l0 = lib.LAUNCHES
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.train_step(self._static)
        self.step_count -= 1                      # the capture did not execute
        self.graph_launches = lib.LAUNCHES - l0   # kernels of ours inside one replay
        self._graph = graph
        return self

    def replay_step(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """One training step by graph replay.  `batch` (device or pinned-host tensors) is copied into the graph's static
        input buffers first; None re-uses their current contents.  Returns the loss vector like train_step."""
        if self._graph is None:
            raise RuntimeError("capture_step() first")
        if batch is not None and batch is not self._static:
            if any("@" in k for k in self._static) and "mask_indices@0" not in batch:
                gc, ps = self._static["global_crops"], self.cfg.vision_patch_size
                B, HW = gc.shape[0] // 2, (gc.shape[-2] // ps) * (gc.shape[-1] // ps)
                batch = dict(batch)
                batch.update(self.split_masks(batch["mask_indices"].to(self.device), batch["masks_weight"].to(self.device), B, HW))
            for k, v in self._static.items():
                if batch[k].shape != v.shape:
                    raise ValueError(f"replay_step: '{k}' has shape {tuple(batch[k].shape)}, the graph was captured for {tuple(v.shape)}")
                v.copy_(batch[k], non_blocking=True)
        self._graph.replay()
        self.step_count += 1
        lib.LAUNCHES += self.graph_launches
        return self.loss_acc

    def release_graph(self):
        """Destroy the captured step graph and its static inputs.  REQUIRED before `dist.destroy_process_group()` in a
        multi-rank job: NCCL does not tear a communicator down while a CUDA graph that captured its collectives is alive
        (the 2-GPU run of round 2 hung in destroy_process_group until the graph was released first)."""
        if self._graph is not None:
            torch.cuda.synchronize(self.device)
            self._graph.reset()
            self._graph = None
            self._static = None
            import gc
            gc.collect()
            torch.cuda.synchronize(self.device)

    @property
    def static_batch(self) -> Dict[str, torch.Tensor]:
        """The graph's input buffers (fill them directly — e.g. H2D copies on a side stream — and call replay_step())."""
        return self._static

    def check_exchange(self):
        """Raise if the peer-memory contrastive exchange reported a barrier time-out (synchronises)."""
        if self.peer is not None:
            self.peer.check()

    def train_step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """One full 3-objective step. Returns the device tensor of accumulated loss terms
        [clip, dino_local, dino_global, ibot, rec_l1, lpips, -, -] (read it with .cpu() to synchronise).  A NaN in slot 0
        with clip_exchange == "p2p" means the peer-memory barrier timed out (a rank never arrived): the step's update
        is invalid — `check_exchange()` raises in that case."""
        tc = self.tc
        self.loss_acc.zero_()
        self._started = set()
        if tc.w_clip:
            self.clip_fwd_bwd(batch["image"], batch["text"], tc.w_clip)
        self._reduce_bucket("text")      # text tower, clip projection, logit scale: final — reduce under the SSL objective
        if tc.w_ssl:
            self.ssl_fwd_bwd(batch["global_crops"], batch["local_crops"], batch["mask_indices"], batch["masks_weight"],
                             tc.w_ssl, mask_groups=batch)
        if tc.w_rec:
            img = batch["rec_image"]
            nB = img.shape[0]
            rc = tc.rec_chunk if 0 < tc.rec_chunk < nB else nB
            for b0 in range(0, nB, rc):
                self.rec_fwd_bwd(img[b0:b0 + rc], tc.w_rec, norm_B=nB, final_group=(b0 + rc >= nB))
        self.allreduce_grads()
        self.optimizer_step()
        return self.loss_acc
