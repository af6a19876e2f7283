// vtp_b200 — the data formats either side of the encode/decode path: the uint8 image that
// generation/tokenizer/vtp_tokenizer.py:106-119 (decode_to_images) and tools/test_reconstruction_hf.py:371-372,401-402
// produce from the decoder output, and the per-channel latent statistics behind latents_stats.pt
// (generation/tools/extract_features_vtp.py:128-131).  Single-pass, HBM-bound, integer outputs bit-exact.
#include "host.h"
#include "ptx.cuh"

namespace vtp {

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
// 4 consecutive elements with one vector load (pointers are 16 / 8 byte aligned: W % 4 == 0, torch allocations)
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = bf16_lo(q.x), v[1] = bf16_hi(q.x), v[2] = bf16_lo(q.y), v[3] = bf16_hi(q.y);
}

// torchvision Normalize(inv_mean, inv_std) = (x − sub) / div in fp32 (IEEE sub, div — no contraction), then ·255,
// clamp to [0,255], truncate: exactly `torch.clamp(t * 255, 0, 255).to(torch.uint8)`.
__device__ __forceinline__ unsigned to_u8(float x, float sub, float div) {
    float t = __fmul_rn(__fdiv_rn(__fsub_rn(x, sub), div), 255.f);
    t = fminf(fmaxf(t, 0.f), 255.f);  // NaN -> 0 (fmaxf returns the non-NaN operand)
    return (unsigned)t;                // truncation toward zero
}

// img NCHW [B][3][H][W] (fp32 | bf16) -> out NHWC uint8 [B][H][W][3].  One thread = 4 consecutive pixels of a row:
// three 4-element reads (one per channel plane) and one 12-byte write.
template <typename T>
__global__ void image_to_u8_kernel(const T* __restrict__ img, const float* __restrict__ sub3,
                                   const float* __restrict__ div3, unsigned* __restrict__ out, long HW, long total4) {
    const float s0 = sub3[0], s1 = sub3[1], s2 = sub3[2], d0 = div3[0], d1 = div3[1], d2 = div3[2];
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total4; t += (long)gridDim.x * blockDim.x) {
        const long p = t * 4;              // pixel index within [B*H*W]
        const long b = p / HW, q = p - b * HW;
        const T* base = img + b * 3 * HW + q;
        float r[4], g[4], bl[4];
        ld4(base, r), ld4(base + HW, g), ld4(base + 2 * HW, bl);
        unsigned u[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            u[3 * i + 0] = to_u8(r[i], s0, d0);
            u[3 * i + 1] = to_u8(g[i], s1, d1);
            u[3 * i + 2] = to_u8(bl[i], s2, d2);
        }
        unsigned* o = out + t * 3;
#pragma unroll
        for (int w = 0; w < 3; ++w)
            o[w] = u[4 * w] | (u[4 * w + 1] << 8) | (u[4 * w + 2] << 16) | (u[4 * w + 3] << 24);
    }
}

// lat [B][C][HW] (fp32 | bf16): sum[c] += Σ x, sumsq[c] += Σ x² in fp64.  grid (C, slices): a block owns channel
// blockIdx.x and the (b, hw) pairs  blockIdx.y, blockIdx.y + gridDim.y, ...  in units of blockDim.x elements.
template <typename T>
__global__ void latent_stats_kernel(const T* __restrict__ lat, int B, int C, int HW, double* __restrict__ sum,
                                    double* __restrict__ sumsq) {
    __shared__ double sh[2][32];
    const int c = blockIdx.x;
    const long n = (long)B * HW;
    double s = 0.0, ss = 0.0;
    for (long i = (long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.y * blockDim.x) {
        const long b = i / HW, q = i - b * HW;
        const double v = (double)ldf(lat + (b * C + c) * HW + q);
        s += v;
        ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[0][warp] = s, sh[1][warp] = ss;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        s = lane < nw ? sh[0][lane] : 0.0;
        ss = lane < nw ? sh[1][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) {
            atomicAdd(sum + c, s);
            atomicAdd(sumsq + c, ss);
        }
    }
}

}  // namespace vtp

using namespace vtp;

extern "C" int vtp_image_to_u8(const void* img, int img_dtype, const float* sub3, const float* div3, uint8_t* out_nhwc,
                               int B, int H, int W, vtp_stream_t st) {
    VTP_CHECK_ARG(img && sub3 && div3 && out_nhwc && B > 0 && H > 0 && W > 0 && W % 4 == 0 && ((uintptr_t)img & 15) == 0 &&
                      ((uintptr_t)out_nhwc & 3) == 0,
                  "image_to_u8: bad args (W %% 4 == 0, 16-byte aligned input)");
    const long HW = (long)H * W, total4 = (long)B * HW / 4;
    long g = (total4 + 255) / 256;
    const int grid = (int)(g < (long)num_sms() * 16 ? g : (long)num_sms() * 16);
    if (img_dtype == VTP_F32)
        image_to_u8_kernel<float><<<grid, 256, 0, (cudaStream_t)st>>>((const float*)img, sub3, div3, (unsigned*)out_nhwc, HW,
                                                                     total4);
    else
        image_to_u8_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)st>>>((const __nv_bfloat16*)img, sub3, div3,
                                                                             (unsigned*)out_nhwc, HW, total4);
    VTP_LAUNCH_CHECK();
    return VTP_OK;
}

extern "C" int vtp_latent_stats(const void* lat, int dtype, int B, int C, int HW, double* sum, double* sumsq,
                                vtp_stream_t st) {
    VTP_CHECK_ARG(lat && sum && sumsq && B > 0 && C > 0 && HW > 0, "latent_stats: bad args");
    const long n = (long)B * HW;
    int slices = (int)((n + 1023) / 1024);
    const int cap = ceil_div(num_sms() * 8, C);
    if (slices > cap) slices = cap;
    if (slices < 1) slices = 1;
    const dim3 grid(C, slices);
    if (dtype == VTP_F32)
        latent_stats_kernel<float><<<grid, 256, 0, (cudaStream_t)st>>>((const float*)lat, B, C, HW, sum, sumsq);
    else
        latent_stats_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)st>>>((const __nv_bfloat16*)lat, B, C, HW, sum, sumsq);
    VTP_LAUNCH_CHECK();
    return VTP_OK;
}
