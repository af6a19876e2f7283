// vtp_b200 — input side of the training step (absent from the reference, which leaves multi-crop /
// mask generation / tokenisation to an un-released DINOv2-style CPU data loader): ONE kernel turns the decoded uint8 source
// images (NHWC, as a JPEG decoder / DataLoader delivers them) into every normalised fp32 NCHW crop of the step.
//
//   crop n :  source image src_idx[n], box (x0, y0, w, h) in source pixels (fractional allowed), optional horizontal flip,
//             bilinear resample to S x S with half-pixel centres (== F.interpolate(..., mode="bilinear",
//             align_corners=False, antialias=False) of the cropped region), then (v / 255 - mean[c]) / std[c].
// One thread = one output pixel (3 channels): the 4 taps x 3 bytes come from L2 (a 256 x 256 x 3 source is 192 KB), the
// store is three coalesced fp32 writes.  HBM-bound on the output: 12 B per output pixel.
#include "host.h"
#include "ptx.cuh"

namespace vtp {

__global__ void crop_resize_norm_kernel(const uint8_t* __restrict__ src, int H, int W, const int* __restrict__ src_idx,
                                        const float* __restrict__ boxes, const uint8_t* __restrict__ flips,
                                        float* __restrict__ out, int N, int S, float m0, float m1, float m2, float is0,
                                        float is1, float is2) {
    const long per = (long)S * S, total = (long)N * per;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int n = (int)(t / per), r = (int)(t - (long)n * per);
        const int oy = r / S, ox = r - oy * S;
        const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + n);   // x0, y0, w, h
        const int xo = flips && flips[n] ? S - 1 - ox : ox;
        // half-pixel centres: source coordinate of the output pixel centre, clamped like torch's bilinear kernel
        float sx = bx.x + (xo + 0.5f) * (bx.z / S) - 0.5f;
        float sy = bx.y + (oy + 0.5f) * (bx.w / S) - 0.5f;
        // torch clamps the coordinate inside the CROPPED tensor; in source coordinates that is [x0, x0 + w - 1]
        sx = fminf(fmaxf(sx, bx.x), bx.x + bx.z - 1.f);
        sy = fminf(fmaxf(sy, bx.y), bx.y + bx.w - 1.f);
        const int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
        const float fx = sx - x0, fy = sy - y0;
        const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
        const int ya = min(max(y0, 0), H - 1), yb = min(max(y0 + 1, 0), H - 1);
        const uint8_t* img = src + (long)src_idx[n] * H * W * 3;
        const uint8_t* p00 = img + ((long)ya * W + xa) * 3;
        const uint8_t* p01 = img + ((long)ya * W + xb) * 3;
        const uint8_t* p10 = img + ((long)yb * W + xa) * 3;
        const uint8_t* p11 = img + ((long)yb * W + xb) * 3;
        const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = w00 * p00[c] + w01 * p01[c] + w10 * p10[c] + w11 * p11[c];
        float* o = out + (long)n * 3 * per + r;
        o[0] = (v[0] * (1.f / 255.f) - m0) * is0;
        o[per] = (v[1] * (1.f / 255.f) - m1) * is1;
        o[2 * per] = (v[2] * (1.f / 255.f) - m2) * is2;
    }
}

}  // namespace vtp

extern "C" int vtp_crop_resize_norm(const uint8_t* src_nhwc, int B, int H, int W, const int* src_idx, const float* boxes_xywh,
                                    const uint8_t* flips, float* out_nchw, int N, int S, const float* mean3,
                                    const float* std3, vtp_stream_t st) {
    VTP_CHECK_ARG(src_nhwc && src_idx && boxes_xywh && out_nchw && mean3 && std3 && B > 0 && H > 0 && W > 0 && N > 0 && S > 0,
                  "crop_resize_norm: bad args");
    VTP_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes_xywh) & 15) == 0, "crop_resize_norm: boxes must be 16B aligned");
    const long total = (long)N * S * S;
    long g = (total + 255) / 256;
    const long cap = (long)vtp::num_sms() * 16;
    vtp::crop_resize_norm_kernel<<<(int)(g < cap ? g : cap), 256, 0, (cudaStream_t)st>>>(
        src_nhwc, H, W, src_idx, boxes_xywh, flips, out_nchw, N, S, mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1],
        1.f / std3[2]);
    VTP_LAUNCH_CHECK();
    return VTP_OK;
}
