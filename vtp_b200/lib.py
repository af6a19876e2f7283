"""ctypes binding of the C-ABI library (include/vtp_b200.h).  No torch types cross this boundary: only
`tensor.data_ptr()` integers, sizes and the raw CUDA stream handle.

The library is REQUIRED: there is no Python/CPU fallback for any entry point.  `load()` raises if the shared
object is missing, and every wrapper raises `VtpError` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvtp_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_SWIGLU8, ACT_ROPE = 0, 1, 2, 3


class VtpError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_int), ("a_mn_major", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_int), ("b_mn_major", C.c_int),
        ("out", C.c_void_p), ("ldo", C.c_int), ("out_dtype", C.c_int),
        ("bias", C.c_void_p),
        ("act", C.c_int), ("round_bf16", C.c_int),
        ("resid", C.c_void_p), ("ldr", C.c_int), ("resid_dtype", C.c_int),
        ("accumulate", C.c_int), ("split_k", C.c_int),
        ("rr_group", C.c_int), ("rr_skip", C.c_int),
        ("rope_sin", C.c_void_p), ("rope_cos", C.c_void_p),
        ("rope_tokens", C.c_int), ("rope_prefix", C.c_int), ("rope_cols", C.c_int),
        ("ps_r", C.c_int), ("ps_gh", C.c_int), ("ps_gw", C.c_int), ("ps_cout", C.c_int),
        ("out2", C.c_void_p), ("ldo2", C.c_int),
    ]


_lib = None

# name -> (restype, argtypes); every symbol declared in include/vtp_b200.h must be listed here
# (tests/test_abi.py checks the header against this table and against the built .so)
SIGNATURES: dict[str, tuple] = {
    "vtp_last_error": (C.c_char_p, []),
    "vtp_version": (C.c_int, []),
    "vtp_check_device": (C.c_int, []),
    "vtp_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "vtp_patchify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vtp_fill_prefix_tokens": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vtp_apply_mask_tokens": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p]),
    "vtp_norm_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                               C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vtp_split3": (C.c_int, [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]),
    "vtp_transpose_batched": (C.c_int, [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p]),
    "vtp_gather_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int,
                                  C.c_int, C.c_void_p]),
    "vtp_embed_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]),
    "vtp_l2norm_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p]),
    "vtp_attention_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p]),
    "vtp_attention_fwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VtpError(
                f"{LIB_PATH} not found — build it with `python -m vtp_b200.build` (there is no fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().vtp_last_error()
        raise VtpError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def _dt(t) -> int:
    import torch

    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise VtpError(f"unsupported dtype {t.dtype}")


def gemm(A, B, out, *, M: int, N: int, K: int, lda: int | None = None, ldb: int | None = None, ldo: int | None = None,
         a_mn: bool = False, b_mn: bool = False, bias=None, act: int = ACT_NONE, round_bf16: bool = True,
         resid=None, ldr: int | None = None, accumulate: bool = False, split_k: int = 1,
         rr_group: int = 0, rr_skip: int = 0, rope=None, pixel_shuffle=None, out2=None, ldo2: int | None = None,
         stream: int | None = None) -> None:
    """out = epi(A · Bᵀ). A/B bf16 tensors (any shape; leading dims given explicitly or inferred from stride(-2))."""
    a = GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_mn_major = _ptr(A), (lda if lda is not None else A.stride(-2)), int(a_mn)
    a.B, a.ldb, a.b_mn_major = _ptr(B), (ldb if ldb is not None else B.stride(-2)), int(b_mn)
    a.out, a.out_dtype = _ptr(out), _dt(out)
    a.ldo = ldo if ldo is not None else (out.stride(-2) if out.dim() >= 2 else N)
    a.bias = _ptr(bias)
    a.act, a.round_bf16 = act, int(round_bf16)
    if resid is not None:
        a.resid, a.resid_dtype = _ptr(resid), _dt(resid)
        a.ldr = ldr if ldr is not None else resid.stride(-2)
    a.accumulate, a.split_k = int(accumulate), split_k
    a.rr_group, a.rr_skip = rr_group, rr_skip
    if rope is not None:
        sin, cos, tokens, prefix, cols = rope
        a.rope_sin, a.rope_cos = _ptr(sin), _ptr(cos)
        a.rope_tokens, a.rope_prefix, a.rope_cols = tokens, prefix, cols
    if pixel_shuffle is not None:
        a.ps_r, a.ps_gh, a.ps_gw, a.ps_cout = pixel_shuffle
    if out2 is not None:
        a.out2, a.ldo2 = _ptr(out2), (ldo2 if ldo2 is not None else out2.stride(-2))
    check(load().vtp_gemm_bf16(C.byref(a), stream if stream is not None else current_stream()), "vtp_gemm_bf16")


def _st(stream):
    return stream if stream is not None else current_stream()


def patchify(img, out, p: int = 16, stream=None):
    B, Cc, H, W = img.shape
    check(load().vtp_patchify(_ptr(img), _ptr(out), _dt(out), B, Cc, H, W, p, _st(stream)), "vtp_patchify")


def fill_prefix_tokens(x, vec, B: int, tokens: int, nprefix: int, D: int, stream=None):
    check(load().vtp_fill_prefix_tokens(_ptr(x), _dt(x), _ptr(vec), B, tokens, nprefix, D, _st(stream)),
          "vtp_fill_prefix_tokens")


def apply_mask_tokens(x, mask_token, idx, HW: int, tokens: int, prefix: int, D: int, stream=None):
    check(load().vtp_apply_mask_tokens(_ptr(x), _dt(x), _ptr(mask_token), _ptr(idx), idx.numel(), HW, tokens, prefix, D,
                                       _st(stream)), "vtp_apply_mask_tokens")


OUT_F32, OUT_BF16, OUT_SPLIT3 = 0, 1, 2


def norm_fwd(x, y, w, b, eps: float, M: int, D: int, *, y_mode: int, ldx: int | None = None, rstd=None, mean=None,
             stream=None):
    check(load().vtp_norm_fwd(_ptr(x), _dt(x), ldx if ldx is not None else D, _ptr(y), y_mode, _ptr(w), _ptr(b), eps,
                              M, D, _ptr(rstd), _ptr(mean), _st(stream)), "vtp_norm_fwd")


def split3(x, out, M: int, K: int, *, b_side: bool, ldx: int | None = None, stream=None):
    check(load().vtp_split3(_ptr(x), ldx if ldx is not None else K, _ptr(out), M, K, int(b_side), _st(stream)),
          "vtp_split3")


def transpose_batched(inp, out, B: int, R: int, Cc: int, *, in_bstride: int | None = None,
                      out_bstride: int | None = None, in_offset: int = 0, stream=None):
    """in [B][R][Cc] (batch stride in_bstride elements, starting in_offset elements in) -> out [B][Cc][R]."""
    ip = inp.data_ptr() + in_offset * inp.element_size()
    check(load().vtp_transpose_batched(ip, _dt(inp), in_bstride if in_bstride is not None else R * Cc, _ptr(out),
                                       _dt(out), out_bstride if out_bstride is not None else R * Cc, B, R, Cc,
                                       _st(stream)), "vtp_transpose_batched")


def gather_rows(inp, out, idx, D: int, *, ld_in: int | None = None, ld_out: int | None = None, stream=None):
    check(load().vtp_gather_rows(_ptr(inp), _dt(inp), ld_in if ld_in is not None else D, _ptr(out), _dt(out),
                                 ld_out if ld_out is not None else D, _ptr(idx), idx.numel(), D, _st(stream)),
          "vtp_gather_rows")


def attention_fwd(qkv, out, B: int, T: int, H: int, *, prefix: int, causal: bool = False, lse=None, stream=None):
    check(load().vtp_attention_fwd(_ptr(qkv), _ptr(out), _ptr(lse), B, T, H, prefix, int(causal), _st(stream)),
          "vtp_attention_fwd")


def attention_fwd_f32(qkv, out, B: int, T: int, H: int, *, causal: bool = False, stream=None):
    check(load().vtp_attention_fwd_f32(_ptr(qkv), _ptr(out), B, T, H, int(causal), _st(stream)),
          "vtp_attention_fwd_f32")


def embed_tokens(ids, emb, pos, out, stream=None):
    B, L = ids.shape
    check(load().vtp_embed_tokens(_ptr(ids), _ptr(emb), _ptr(pos), _ptr(out), B * L, L, emb.shape[1], _st(stream)),
          "vtp_embed_tokens")


def l2norm_fwd(x, y, M: int, D: int, eps: float = 1e-12, norm_out=None, stream=None):
    check(load().vtp_l2norm_fwd(_ptr(x), _dt(x), _ptr(y), _dt(y), _ptr(norm_out), M, D, eps, _st(stream)),
          "vtp_l2norm_fwd")
