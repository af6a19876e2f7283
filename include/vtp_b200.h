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

#ifdef __cplusplus
}
#endif
#endif /* VTP_B200_H */
