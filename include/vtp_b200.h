/*
 * vtp_b200.h — C ABI of the B200-native (sm_100a) VTP hot path.
 *
 * This is the drop-in boundary: the reference (MiniMax-AI/VTP) is pure Python/PyTorch and has no
 * FFI of its own, so the boundary is the set of fused stages that the reference's L1 layers dispatch to ATen for.
 * Every entry point takes raw device pointers, sizes and a CUDA stream, returns an int status (0 = ok, <0 = error,
 * message via vtp_last_error()), never throws, never allocates, never synchronises. The Python host
 * (vtp_b200/model.py, mirroring vtp/models/vtp_hf/modeling_vtp.py) binds these with ctypes.
 *
 * Each declaration cites the reference call site (relative to the reference root) that it replaces.
 */
#ifndef VTP_B200_H
#define VTP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vtp_stream_t; /* cudaStream_t */

enum {
    VTP_OK = 0,
    VTP_ERR_ARG = -1,   /* bad argument / unsupported shape */
    VTP_ERR_CUDA = -2,  /* CUDA runtime / driver error */
    VTP_ERR_ARCH = -3,  /* device is not sm_100 */
};
enum { VTP_F32 = 0, VTP_BF16 = 1 };
enum { VTP_ACT_NONE = 0, VTP_ACT_GELU = 1, VTP_ACT_SWIGLU8 = 2, VTP_ACT_ROPE = 3 };

const char* vtp_last_error(void);
int vtp_version(void);
/* 0 if the current device is compute capability 10.x, VTP_ERR_ARCH otherwise */
int vtp_check_device(void);

/* ------------------------------------------------------------------------------------------------------------
 * tcgen05 / TMA GEMM with fused epilogue:  out = epi( A · Bᵀ ),  bf16 operands, fp32 accumulation in TMEM.
 * Replaces every nn.Linear / 1x1 nn.Conv2d / 16x16-stride-16 nn.Conv2d on the path:
 *   layers/attention.py:62,64,92,94 (qkv, proj)   layers/ffn.py:73-81 (w1,w2,w3)   layers/embeddings.py:58,64
 *   encoders/vision_transformer_bottleneck.py:30,66-79   decoders/pixel_decoder.py:108,138,157,160
 *   vtp_hf/modeling_vtp.py:274,308,329   heads/dino_head.py:65-89   layers/block.py:387-412 (text tower linears)
 * and their dgrad / wgrad in the training step (operand major-ness flags select the transposes).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int M, N, K;            /* out is [M,N]; reduction length K */
    const void* A;          /* bf16. a_mn_major=0: A[M][lda] (K contiguous); =1: A[K][lda] (M contiguous) */
    int lda, a_mn_major;
    const void* B;          /* bf16. b_mn_major=0: B[N][ldb] (K contiguous, nn.Linear.weight); =1: B[K][ldb] */
    int ldb, b_mn_major;
    void* out;              /* [M][ldo] (after row remap / pixel shuffle) */
    int ldo, out_dtype;     /* VTP_F32 | VTP_BF16 */
    const float* bias;      /* [N] fp32 or NULL */
    int act;                /* VTP_ACT_* */
    int round_bf16;         /* round (acc+bias) to bf16 first (autocast nn.Linear output semantics) */
    const void* resid;      /* residual added after activation, same indexing as out; NULL = none */
    int ldr, resid_dtype;
    int accumulate;         /* 1: atomic fp32 add into out (required when split_k > 1) */
    int split_k;            /* >=1: split the reduction across CTAs */
    int rr_group, rr_skip;  /* out_row = (row/rr_group)*(rr_group+rr_skip) + rr_skip + row%rr_group; 0 = identity */
    const void* rope_sin;   /* VTP_ACT_ROPE: bf16 [rope_tokens-rope_prefix][64] tables (layers/embeddings.py:131-180) */
    const void* rope_cos;
    int rope_tokens, rope_prefix, rope_cols; /* tokens/sequence, un-rotated prefix tokens, leading columns rotated (2*D) */
    int ps_r, ps_gh, ps_gw, ps_cout;         /* ps_r>0: PixelShuffle(ps_r) NCHW store, grid gh x gw, cout channels */
    void* out2;             /* optional bf16 copy of (acc+bias) before activation (saved for backward), [M][ldo2] */
    int ldo2;
} vtp_gemm_args;

int vtp_gemm_bf16(const vtp_gemm_args* args, vtp_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * HBM-bound stages around the GEMMs (vtp_b200/csrc/elementwise.cu)
 * ------------------------------------------------------------------------------------------------------------ */
/* layers/embeddings.py:58,61-70 — input side of PatchEmbed's Conv2d(C,D,p,p): NCHW fp32 image -> bf16 im2col rows
 * (or fp32 for the accurate mode) [B*(H/p)*(W/p)][C*p*p], k = c*p*p + i*p + j (== conv weight.flatten(1)); the conv
 * itself is vtp_gemm_bf16. */
int vtp_patchify(const float* img, void* out, int out_dtype, int B, int C, int H, int W, int p, vtp_stream_t stream);
/* encoders/vision_transformer.py:198-217 — x[b, t, :] = vec[t, :] for t < nprefix (cls / storage tokens) */
int vtp_fill_prefix_tokens(void* x, int x_dtype, const float* vec, int B, int tokens, int nprefix, int D,
                           vtp_stream_t stream);
/* encoders/vision_transformer.py:195 — torch.where(masks, mask_token, x): idx = flat indices into [B*HW] */
int vtp_apply_mask_tokens(void* x, int x_dtype, const float* mask_token, const int64_t* idx, int n, int HW, int tokens,
                          int prefix, int D, vtp_stream_t stream);
/* layers/normalization.py:17-22 (RMSNorm, b == NULL) and nn.LayerNorm (encoders/vision_transformer.py:30-34,
 * layers/normalization.py:25-40).  x [M][ldx] fp32|bf16; y_mode: 0 fp32 [M][D], 1 bf16 [M][D], 2 bf16x3 split [M][3D]
 * (hi|hi|lo, operand of the fp32-accurate GEMM).  rstd_out/mean_out [M] optional (saved for backward). */
int vtp_norm_fwd(const void* x, int x_dtype, long ldx, void* y, int y_mode, const float* w, const float* b, float eps,
                 int M, int D, float* rstd_out, float* mean_out, vtp_stream_t stream);
/* fp32 [M][ldx] -> bf16 [M][3K]: A side (b_side=0) hi|hi|lo, B side (b_side=1) hi|lo|hi;  A'·B'^T = fp32-accurate */
int vtp_split3(const float* x, long ldx, void* out_bf16, long M, int K, int b_side, vtp_stream_t stream);
/* in [B][R][C] -> out [B][C][R] with dtype conversion and element batch strides (vtp_hf/modeling_vtp.py:395,
 * decoders/pixel_decoder.py:141) */
int vtp_transpose_batched(const void* in, int in_dtype, long in_bstride, void* out, int out_dtype, long out_bstride,
                          int B, int R, int C, vtp_stream_t stream);
/* out[i,:] = in[idx[i],:] (vtp.py:432-439,470-473 iBOT gather; encoders/text_transformer.py:224 argmax pool) */
int vtp_gather_rows(const void* in, int in_dtype, long ld_in, void* out, int out_dtype, long ld_out, const int64_t* idx,
                    int n, int D, vtp_stream_t stream);
/* vtp_hf/modeling_vtp.py:297-298 — out[b*L+l] = token_embedding[ids[b,l]] + positional_embedding[l] (fp32) */
int vtp_embed_tokens(const int64_t* ids, const float* emb, const float* pos, float* out, long BL, int L, int D,
                     vtp_stream_t stream);
/* F.normalize(x, dim=-1, eps) (vtp_hf/modeling_vtp.py:276,310; heads/dino_head.py:83-84); norm_out [M] optional */
int vtp_l2norm_fwd(const void* x, int x_dtype, void* y, int y_dtype, float* norm_out, int M, int D, float eps,
                   vtp_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention (vtp_b200/csrc/attention.cu)
 * ------------------------------------------------------------------------------------------------------------ */
/* layers/attention.py:110-126 after RoPE (F.scaled_dot_product_attention, scale 1/8, head_dim 64) and the causal
 * nn.MultiheadAttention of layers/block.py:387-412.  qkv bf16 [B*T][3*H*64] packed [q|k|v] x [H][64]; out bf16
 * [B*T][H*64]; lse fp32 [B][H][T] optional (saved for backward).  `prefix` leading tokens (cls) are computed on CUDA
 * cores, the other T-prefix (<=256) tokens on tcgen05. */
int vtp_attention_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int prefix, int causal,
                      vtp_stream_t stream);
/* same op on fp32 tensors (CUDA cores) for the fp32-accurate inference mode */
int vtp_attention_fwd_f32(const float* qkv, float* out, int B, int T, int H, int causal, vtp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VTP_B200_H */
