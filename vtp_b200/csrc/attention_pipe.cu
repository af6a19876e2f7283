// vtp_b200 — persistent, software-pipelined self-attention forward for 128 < HW <= 256 (the ViT trunk, T = 1 + 256, and the
// pixel decoder, T = 256).  First hardware run in round 2: bit-identical to attn_fwd_kernel and 1.11x faster; the default for
// these shapes since (VTP_ATTN_FWD_PIPE=0 selects the one-tile-per-CTA kernel).
//
// FULL (HW == 256, every key valid): the row threads run a predicate-free softmax — ncu of the generic loop showed ~25
// executed instructions per score element, three quarters of them mask predicates, address arithmetic and selects
// (profiles/ncu_attn_r2a.md).  The FULL path: max pass with four independent FMNMX chains over two TMEM loads in flight;
// exp pass fully unrolled with the TMEM load of the next 32 columns in flight, hoisted swizzle offsets, two row-sum chains;
// the second 128-key half is exponentiated into registers BEFORE waiting for the first half's P.V (which frees the P buffer).
//
// Why: the one-tile-per-CTA kernel (attention.cu) spends ~12.9 us per 128x256 tile for ~3 us of issue slots
// (profiles/ncu_attn_r1b_before_cls_fix.md: 16 % warps active, 5 % tensor pipe) and a variant with twice the row threads
// is not faster (profiles/hbm_kernels_r1.md) — the time goes into the per-CTA latency chain
//     TMEM alloc -> TMA(Q,K,V) -> S MMA -> softmax -> P.V MMA -> store -> teardown
// with only two chains in flight per SM.  This kernel keeps ONE persistent CTA per SM that walks a list of (head, image)
// jobs.  Each job has two 128-row query tiles that share K and V (loaded once instead of twice); two softmax warpgroups
// ping-pong on them, so the S / P.V MMAs and TMA latency of one tile hide under the other tile's softmax, and the K/V
// stage of the NEXT job is prefetched while the current one computes.
//
//   warp 0        TMA producer        K|V double-buffered (2 x 64 KB), Q0|Q1 single-buffered (32 KB)
//   warps 1, 3    MMA issuers (tile 0 / tile 1)   S_w = Q_w K^T into TMEM columns [256 w, 256 w + 256);  O_w = P_w V into [256 w, +64)
//   warp 2        prefix (cls) query row of the job on CUDA cores (as in attention.cu), from the K/V stage in smem
//   warps 4-7     softmax warpgroup 0 (query tile 0): one thread per row, exact single-pass softmax out of TMEM
//   warps 8-11    softmax warpgroup 1 (query tile 1)
// smem: 128 K (K/V stages) + 32 K (Q) + 64 K (P: one 128-key half per warpgroup at a time) + 1 K = 225 KB, 1 CTA/SM, TMEM 512 cols.
#include <stdlib.h>

#include "attention.h"
#include "host.h"
#include "ptx.cuh"

namespace vtp {

#ifndef VTP_ATTN_TSO_DEFAULT
#define VTP_ATTN_TSO_DEFAULT 1  // TMA-store output epilogue of the FULL path (measured: 199.1 -> 171.8 us, same bits)
#endif
static constexpr int PIPE_THREADS = 384;
static constexpr int P_KV = 0;                    // 2 stages x (K 32768 | V 32768)
static constexpr int P_Q = 131072;                // Q0 16384 | Q1 16384
static constexpr int P_P = P_Q + 32768;           // P of warpgroup 0 (32768) | warpgroup 1 (32768)
static constexpr int P_PCLS = P_P + 65536;        // bf16 [256] numerators of the cls query row
static constexpr int P_BAR = P_PCLS + 512;        // mbarriers
static constexpr int PIPE_SMEM = P_BAR + 256;     // 230144 B

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t swz(int row, int col /* bf16 element 0..63 */) {
    return row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1));
}

// TSO (FULL only; round 2, from the warp-state samples of profiles/ncu_attn_r2b: 29 % of the row threads' time sat in the
// output epilogue — eight 16-byte global stores per thread, each warp instruction touching 32 different 128-byte lines —
// and 13 % in the cls-key score waiting for its global loads): the normalised O row goes through the warp's own 4 KB of the
// (by then idle) P buffer and leaves as ONE TMA store per warp; the cls key / value rows are prefetched into L1 at the top of the job.
template <bool FULL, bool TSO>
__global__ void __launch_bounds__(PIPE_THREADS, 1)
attn_fwd_pipe_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmo, const AttnDev p) {
    static_assert(!TSO || FULL, "the TMA-store epilogue writes whole 128-row tiles");
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_BAR);
    uint64_t* kv_full = bars + 0;    // [2]  K,V of stage s landed (TMA tx)
    uint64_t* kv_empty = bars + 2;   // [2]  stage s free again: last P.V of the job committed (+ cls warp done)
    uint64_t* q_full = bars + 4;     //      Q0,Q1 landed
    uint64_t* q_empty = bars + 5;    //      both S chains committed + all 256 row threads have read their q row
    uint64_t* s_full = bars + 6;     // [2]  S_w complete in TMEM
    uint64_t* s_empty = bars + 8;    // [2]  warpgroup w has drained O_w: the TMEM region may take the next job's S_w
    uint64_t* p_full = bars + 10;    // [2][2] P half written by warpgroup w (128 arrivals)
    uint64_t* pv_done = bars + 14;   // [2][2] P.V half committed (half 0: P buffer reusable; half 1: O_w complete)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.D, T = p.T, prefix = p.prefix, HW = p.HW, H = p.H;
    const int njobs = p.B * H;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm);
        if (TSO) tma_prefetch_desc(&tmo);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], prefix > 0 ? 3 : 2);   // two MMA chains (+ the cls warp)
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], 128);
            for (int h2 = 0; h2 < 2; ++h2) mbar_init(&p_full[s * 2 + h2], 128), mbar_init(&pv_done[s * 2 + h2], 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 2 + 256);                     // two S chains + every row thread
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0;
            for (int job = blockIdx.x; job < njobs; job += gridDim.x, ++it) {
                const int b = job / H, h = job - b * H;
                const int row_k = b * T + prefix;
                const int s = it & 1, u = it >> 1;
                if (u >= 1) mbar_wait(&kv_empty[s], (u - 1) & 1);
                uint8_t* kv = smem + P_KV + s * 65536;
                mbar_expect_tx(&kv_full[s], 65536);
                tma_load_2d(kv, &tm, &kv_full[s], D + h * 64, row_k);
                tma_load_2d(kv + 16384, &tm, &kv_full[s], D + h * 64, row_k + 128);
                tma_load_2d(kv + 32768, &tm, &kv_full[s], 2 * D + h * 64, row_k);
                tma_load_2d(kv + 49152, &tm, &kv_full[s], 2 * D + h * 64, row_k + 128);
                if (it >= 1) mbar_wait(q_empty, (it - 1) & 1);
                mbar_expect_tx(q_full, 32768);
                tma_load_2d(smem + P_Q, &tm, q_full, h * 64, row_k);
                tma_load_2d(smem + P_Q + 16384, &tm, q_full, h * 64, row_k + 128);
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ------------------------------------------------------------------------------------------ MMA issuers
        // one issuing thread per query tile: the two chains  S_w -> P_w.V(half 0) -> P_w.V(half 1)  are independent, so a
        // warpgroup that is ahead never waits for the other one's softmax (the tiles drift into a ping-pong by themselves)
        if (lane == 0) {
            const int w = warp >> 1;  // warp 1 -> tile 0, warp 3 -> tile 1
            const uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
            const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
            const uint32_t acc = tmem + w * 256;
            const uint32_t qa = smem_u32(smem + P_Q + w * 16384);
            const uint32_t pa = smem_u32(smem + P_P + w * 32768);
            int it = 0;
            for (int job = blockIdx.x; job < njobs; job += gridDim.x, ++it) {
                const int s = it & 1, u = it >> 1;
                const uint32_t ka = smem_u32(smem + P_KV + s * 65536), va = ka + 32768;
                mbar_wait(q_full, it & 1);
                mbar_wait(&kv_full[s], u & 1);
                if (it >= 1) mbar_wait(&s_empty[w], (it - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_bf16_ss(acc, umma_desc_sw128(qa + j * 32, 0, 1024), umma_desc_sw128(ka + j * 32, 0, 1024), idesc_s, j > 0);
                umma_commit(&s_full[w]);
                umma_commit(q_empty);  // this chain has consumed Q_w
                for (int half = 0; half < 2; ++half) {
                    mbar_wait(&p_full[w * 2 + half], it & 1);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {  // 8 k-steps of 16 keys
                        const uint64_t ad = umma_desc_sw128(pa + (j >> 2) * 16384 + (j & 3) * 32, 0, 1024);
                        const uint64_t bd = umma_desc_sw128(va + half * 16384 + j * 2048, 8192, 1024);
                        umma_bf16_ss(acc, ad, bd, idesc_o, (half > 0 || j > 0) ? 1u : 0u);
                    }
                    umma_commit(&pv_done[w * 2 + half]);
                }
                umma_commit(&kv_empty[s]);  // this chain's last read of the K/V stage has been issued; arrives when it completes
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------------------------------ prefix (cls) query rows
        if (prefix > 0) {
            int it = 0;
            for (int job = blockIdx.x; job < njobs; job += gridDim.x, ++it) {
                const int b = job / H, h = job - b * H;
                const long seq_row0 = (long)b * T;
                const int s = it & 1, u = it >> 1;
                const uint8_t* ks = smem + P_KV + s * 65536;
                const uint8_t* vs = ks + 32768;
                mbar_wait(&kv_full[s], u & 1);
                for (int j = 0; j < prefix; ++j) {
                    float qf[64];
                    {
                        const uint4* qp = reinterpret_cast<const uint4*>(p.qkv + (seq_row0 + j) * 3 * D + h * 64);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 w = __ldg(qp + c);
                            qf[c * 8 + 0] = bf16_lo(w.x), qf[c * 8 + 1] = bf16_hi(w.x), qf[c * 8 + 2] = bf16_lo(w.y);
                            qf[c * 8 + 3] = bf16_hi(w.y), qf[c * 8 + 4] = bf16_lo(w.z), qf[c * 8 + 5] = bf16_hi(w.z);
                            qf[c * 8 + 6] = bf16_lo(w.w), qf[c * 8 + 7] = bf16_hi(w.w);
                        }
                    }
                    auto dot_row = [&](const uint4* kp, bool from_smem, int row) {
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 w = from_smem ? *reinterpret_cast<const uint4*>(ks + swz(row, c * 8)) : __ldg(kp + c);
                            acc += qf[c * 8 + 0] * bf16_lo(w.x) + qf[c * 8 + 1] * bf16_hi(w.x) + qf[c * 8 + 2] * bf16_lo(w.y) +
                                   qf[c * 8 + 3] * bf16_hi(w.y) + qf[c * 8 + 4] * bf16_lo(w.z) + qf[c * 8 + 5] * bf16_hi(w.z) +
                                   qf[c * 8 + 6] * bf16_lo(w.w) + qf[c * 8 + 7] * bf16_hi(w.w);
                        }
                        return acc;
                    };
                    float sp[ATT_MAX_PREFIX], sc[8];
                    float m = -INFINITY;
#pragma unroll
                    for (int t = 0; t < ATT_MAX_PREFIX; ++t) {
                        sp[t] = -INFINITY;
                        if (t < prefix)
                            sp[t] = dot_row(reinterpret_cast<const uint4*>(p.qkv + (seq_row0 + t) * 3 * D + D + h * 64), false, 0);
                        m = fmaxf(m, sp[t]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int kk = lane + 32 * i;
                        sc[i] = -INFINITY;
                        if (kk < HW) sc[i] = dot_row(nullptr, true, kk);
                        m = fmaxf(m, sc[i]);
                    }
                    m = warp_max(m);
                    const float msc = m * p.scale_log2;
                    float l = 0.f;
                    __nv_bfloat16* pcls = reinterpret_cast<__nv_bfloat16*>(smem + P_PCLS);
                    __syncwarp();  // the previous row / job has finished reading pcls
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float e = (sc[i] == -INFINITY) ? 0.f : ex2a(sc[i] * p.scale_log2 - msc);
                        l += e;
                        pcls[lane + 32 * i] = __float2bfloat16_rn(e);
                    }
                    l = warp_sum(l);
                    __syncwarp();
                    float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
                    for (int t = 0; t < ATT_MAX_PREFIX; ++t) {
                        if (t < prefix && sp[t] != -INFINITY) {
                            const float pe = ex2a(sp[t] * p.scale_log2 - msc);
                            l += pe;
                            const uint32_t w =
                                __ldg(reinterpret_cast<const uint32_t*>(p.qkv + (seq_row0 + t) * 3 * D + 2 * D + h * 64) + lane);
                            a0 += bf16_round(pe) * bf16_lo(w), a1 += bf16_round(pe) * bf16_hi(w);
                        }
                    }
                    for (int k8 = 0; k8 < HW; k8 += 8) {  // HW % 8 == 0 is checked by the launcher
                        const uint4 pw = *reinterpret_cast<const uint4*>(pcls + k8);
                        const uint32_t pr[4] = {pw.x, pw.y, pw.z, pw.w};
#pragma unroll
                        for (int v = 0; v < 8; ++v) {
                            const float pk = (v & 1) ? bf16_hi(pr[v >> 1]) : bf16_lo(pr[v >> 1]);
                            const uint32_t w = *reinterpret_cast<const uint32_t*>(vs + swz(k8 + v, 2 * lane));
                            if (v & 1) c0 += pk * bf16_lo(w), c1 += pk * bf16_hi(w);
                            else a0 += pk * bf16_lo(w), a1 += pk * bf16_hi(w);
                        }
                    }
                    a0 += c0, a1 += c1;
                    const float inv = 1.f / l;
                    *reinterpret_cast<uint32_t*>(p.out + (seq_row0 + j) * D + h * 64 + 2 * lane) = pack_bf16x2(a0 * inv, a1 * inv);
                    if (p.lse && lane == 0) p.lse[((long)b * H + h) * T + j] = m * p.scale + logf(l);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&kv_empty[s]);  // this warp no longer reads the K/V stage
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------------------------------ softmax warpgroups
        const int w = (warp - 4) >> 2;   // warpgroup = query tile of the job
        const int q4 = warp & 3;         // TMEM lane quarter (== warp id % 4)
        const int r = q4 * 32 + lane;    // row within the tile == TMEM lane
        const uint32_t trow = tmem + (uint32_t(q4 * 32) << 16) + w * 256;
        uint8_t* pbuf = smem + P_P + w * 32768;
        const int qpos = 128 * w + r;
        const bool row_valid = qpos < HW;
        int it = 0;
        for (int job = blockIdx.x; job < njobs; job += gridDim.x, ++it) {
            const int b = job / H, h = job - b * H;
            const long seq_row0 = (long)b * T;
            const int qtok = prefix + qpos;
            const uint32_t ph = it & 1;
            // scores against the prefix keys (CUDA cores): q row from smem, k rows from global
            float s_pre[ATT_MAX_PREFIX];
            if constexpr (TSO) {  // the job's prefix key / value rows (one 128-byte line each) into L1 before anybody needs them
                if (lane < 2 * prefix)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p.qkv + (seq_row0 + (lane >> 1)) * 3 * D + (1 + (lane & 1)) * D + h * 64));
            }
            mbar_wait(q_full, ph);
            if (prefix > 0) {
                float qf[64];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 x = *reinterpret_cast<const uint4*>(smem + P_Q + w * 16384 + swz(r, c * 8));
                    qf[c * 8 + 0] = bf16_lo(x.x), qf[c * 8 + 1] = bf16_hi(x.x), qf[c * 8 + 2] = bf16_lo(x.y);
                    qf[c * 8 + 3] = bf16_hi(x.y), qf[c * 8 + 4] = bf16_lo(x.z), qf[c * 8 + 5] = bf16_hi(x.z);
                    qf[c * 8 + 6] = bf16_lo(x.w), qf[c * 8 + 7] = bf16_hi(x.w);
                }
#pragma unroll
                for (int j = 0; j < ATT_MAX_PREFIX; ++j) {
                    s_pre[j] = -INFINITY;
                    if (j < prefix) {
                        const uint4* kp = reinterpret_cast<const uint4*>(p.qkv + (seq_row0 + j) * 3 * D + D + h * 64);
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 x = __ldg(kp + c);
                            acc += qf[c * 8 + 0] * bf16_lo(x.x) + qf[c * 8 + 1] * bf16_hi(x.x) + qf[c * 8 + 2] * bf16_lo(x.y) +
                                   qf[c * 8 + 3] * bf16_hi(x.y) + qf[c * 8 + 4] * bf16_lo(x.z) + qf[c * 8 + 5] * bf16_hi(x.z) +
                                   qf[c * 8 + 6] * bf16_lo(x.w) + qf[c * 8 + 7] * bf16_hi(x.w);
                        }
                        s_pre[j] = acc;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < ATT_MAX_PREFIX; ++j) s_pre[j] = -INFINITY;
            }
            mbar_arrive(q_empty);  // my q row is in registers: the Q buffers may take the next job once the S MMAs are done

            mbar_wait(&s_full[w], ph);
            tc_fence_after();
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < ATT_MAX_PREFIX; ++j) m = fmaxf(m, s_pre[j]);
            float l = 0.f;
            float p_pre[ATT_MAX_PREFIX];
            if constexpr (FULL) {
                // ---- pass 1: row max, 64 columns per round (two TMEM loads in flight), four independent chains
                float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                for (int c = 0; c < 256; c += 64) {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32(trow + c, ra);
                    tmem_ld_32x32(trow + c + 32, rb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(ra[i + 1]), __uint_as_float(rb[i + 1])));
                        m2 = fmaxf(m2, fmaxf(__uint_as_float(ra[i + 2]), __uint_as_float(rb[i + 2])));
                        m3 = fmaxf(m3, fmaxf(__uint_as_float(ra[i + 3]), __uint_as_float(rb[i + 3])));
                    }
                }
                m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                const float sl2 = p.scale_log2, msc = m * sl2;   // a full row always has a finite maximum
#pragma unroll
                for (int j = 0; j < ATT_MAX_PREFIX; ++j) {
                    p_pre[j] = (s_pre[j] == -INFINITY) ? 0.f : ex2a(s_pre[j] * sl2 - msc);
                    l += p_pre[j];
                    p_pre[j] = bf16_round(p_pre[j]);
                }
                // ---- pass 2: p = exp2(s * scale*log2e - m * scale*log2e) as bf16 into the swizzled P tile.
                // this row's eight 16-byte chunk slots of a 128-byte swizzled line (chunk ^ (row & 7)), computed once
                uint32_t poff[8];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) poff[ch] = (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
                float l0 = 0.f, l1 = 0.f;
                uint32_t cur[32], nxt[32];
                tmem_ld_32x32(trow, cur);
                if constexpr (TSO) {  // the previous job's O tile (TMA store out of this warp's rows of the P buffer) has been read
                    if (lane == 0) bulk_wait_read0();
                    __syncwarp();
                }
                tmem_ld_wait();
                // half 0: straight into the P buffer (free: the previous job's P.V of this tile completed before s_full)
#pragma unroll
                for (int c32 = 0; c32 < 4; ++c32) {
                    tmem_ld_32x32(trow + 32 * (c32 + 1), nxt);           // next 32 columns in flight (c32 = 3: first of half 1)
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float e0 = ex2a(fmaf(__uint_as_float(cur[i]), sl2, -msc));
                        const float e1 = ex2a(fmaf(__uint_as_float(cur[i + 1]), sl2, -msc));
                        l0 += e0, l1 += e1;
                        pk[i >> 1] = pack_bf16x2(e0, e1);
                    }
                    uint8_t* pb = pbuf + (c32 >> 1) * 16384;
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4)
                        *reinterpret_cast<uint4*>(pb + poff[(c32 & 1) * 4 + v4]) = make_uint4(pk[v4 * 4], pk[v4 * 4 + 1], pk[v4 * 4 + 2], pk[v4 * 4 + 3]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&p_full[w * 2 + 0]);
                // half 1: exponentiate into registers while the tensor core consumes half 0, store once the buffer is free
                uint32_t pk1[64];
#pragma unroll
                for (int c32 = 0; c32 < 4; ++c32) {   // (no load-ahead here: 64 packed registers are live already)
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float e0 = ex2a(fmaf(__uint_as_float(cur[i]), sl2, -msc));
                        const float e1 = ex2a(fmaf(__uint_as_float(cur[i + 1]), sl2, -msc));
                        l0 += e0, l1 += e1;
                        pk1[c32 * 16 + (i >> 1)] = pack_bf16x2(e0, e1);
                    }
                    if (c32 < 3) {
                        tmem_ld_32x32(trow + 128 + 32 * (c32 + 1), cur);
                        tmem_ld_wait();
                    }
                }
                l += l0 + l1;
                mbar_wait(&pv_done[w * 2 + 0], ph);  // the P buffer is free again
#pragma unroll
                for (int c32 = 0; c32 < 4; ++c32) {
                    uint8_t* pb = pbuf + (c32 >> 1) * 16384;
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4)
                        *reinterpret_cast<uint4*>(pb + poff[(c32 & 1) * 4 + v4]) =
                            make_uint4(pk1[c32 * 16 + v4 * 4], pk1[c32 * 16 + v4 * 4 + 1], pk1[c32 * 16 + v4 * 4 + 2], pk1[c32 * 16 + v4 * 4 + 3]);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&p_full[w * 2 + 1]);
            } else {
            const int kmax = HW;
            for (int c = 0; c < 256; c += 32) {
                if (c >= kmax) continue;  // uniform: tcgen05.ld is warp-collective
                uint32_t rr[32];
                tmem_ld_32x32(trow + c, rr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c + i < kmax) m = fmaxf(m, __uint_as_float(rr[i]));
            }
            const float msc = (m == -INFINITY) ? 0.f : m * p.scale_log2;
#pragma unroll
            for (int j = 0; j < ATT_MAX_PREFIX; ++j) {
                p_pre[j] = (s_pre[j] == -INFINITY) ? 0.f : ex2a(s_pre[j] * p.scale_log2 - msc);
                l += p_pre[j];
                p_pre[j] = bf16_round(p_pre[j]);
            }
            for (int half = 0; half < 2; ++half) {
                if (half == 1) mbar_wait(&pv_done[w * 2 + 0], ph);  // the P buffer is free again
#pragma unroll 1
                for (int c32 = 0; c32 < 4; ++c32) {
                    const int c = half * 128 + c32 * 32;
                    uint32_t pk[16];
                    if (c >= kmax) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    } else {
                        uint32_t rr[32];
                        tmem_ld_32x32(trow + c, rr);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float e0 = (c + i < kmax) ? ex2a(__uint_as_float(rr[i]) * p.scale_log2 - msc) : 0.f;
                            const float e1 = (c + i + 1 < kmax) ? ex2a(__uint_as_float(rr[i + 1]) * p.scale_log2 - msc) : 0.f;
                            l += e0 + e1;
                            pk[i >> 1] = pack_bf16x2(e0, e1);
                        }
                    }
                    uint8_t* pb = pbuf + (c32 >> 1) * 16384;
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        const int col = (c32 & 1) * 32 + v4 * 8;
                        *reinterpret_cast<uint4*>(pb + swz(r, col)) = make_uint4(pk[v4 * 4], pk[v4 * 4 + 1], pk[v4 * 4 + 2], pk[v4 * 4 + 3]);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&p_full[w * 2 + half]);
            }
            }
            // epilogue
            mbar_wait(&pv_done[w * 2 + 1], ph);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld_32x32(trow, o0);
            tmem_ld_32x32(trow + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_empty[w]);  // O_w is in registers: the TMEM region may take the next job's S_w
            float o[64];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(o0[i]), o[32 + i] = __uint_as_float(o1[i]);
#pragma unroll
            for (int j = 0; j < ATT_MAX_PREFIX; ++j) {
                if (j < prefix && p_pre[j] != 0.f) {
                    const uint4* vp = reinterpret_cast<const uint4*>(p.qkv + (seq_row0 + j) * 3 * D + 2 * D + h * 64);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 x = __ldg(vp + c);
                        o[c * 8 + 0] += p_pre[j] * bf16_lo(x.x), o[c * 8 + 1] += p_pre[j] * bf16_hi(x.x);
                        o[c * 8 + 2] += p_pre[j] * bf16_lo(x.y), o[c * 8 + 3] += p_pre[j] * bf16_hi(x.y);
                        o[c * 8 + 4] += p_pre[j] * bf16_lo(x.z), o[c * 8 + 5] += p_pre[j] * bf16_hi(x.z);
                        o[c * 8 + 6] += p_pre[j] * bf16_lo(x.w), o[c * 8 + 7] += p_pre[j] * bf16_hi(x.w);
                    }
                }
            }
            if constexpr (TSO) {
                // every row of a FULL tile is valid.  The P buffer is idle (its last P.V has completed): my row goes into the
                // same swizzled 128-byte line layout the tensor map expects; lane 0 stores the warp's 32 rows
                const float inv = 1.f / l;
                uint8_t* ob = pbuf + (uint32_t)r * 128u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 x;
                    x.x = pack_bf16x2(o[c * 8] * inv, o[c * 8 + 1] * inv), x.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
                    x.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv), x.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
                    *reinterpret_cast<uint4*>(ob + ((c ^ (r & 7)) << 4)) = x;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmo, pbuf + q4 * 4096, h * 64, (int)(seq_row0 + prefix) + 128 * w + q4 * 32);
                    bulk_commit();
                }
                if (p.lse) p.lse[((long)b * H + h) * T + qtok] = m * p.scale + logf(l);
            } else if (row_valid) {
                const float inv = 1.f / l;
                __nv_bfloat16* op = p.out + (seq_row0 + qtok) * D + h * 64;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 x;
                    x.x = pack_bf16x2(o[c * 8] * inv, o[c * 8 + 1] * inv), x.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
                    x.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv), x.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
                    *reinterpret_cast<uint4*>(op + c * 8) = x;
                }
                if (p.lse) p.lse[((long)b * H + h) * T + qtok] = m * p.scale + logf(l);
            }
        }
    }

    if (TSO && warp >= 4 && lane == 0) bulk_wait0();  // this warp's last O store has left shared memory and is complete
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// launcher used by vtp_attention_fwd (attention.cu) when VTP_ATTN_FWD_PIPE=1 and the shape qualifies
int attn_fwd_pipe_launch(const CUtensorMap& tm, const AttnDev& p, cudaStream_t st) {
    VTP_CHECK_ARG(!p.pack && !p.causal && p.nkt == 2 && p.HW > 128 && p.HW <= 256 && p.HW % 8 == 0 && p.prefix <= ATT_MAX_PREFIX,
                  "attention_fwd(pipe): needs 128 < HW <= 256, HW %% 8 == 0, no causal mask");
    static bool configured = false;
    if (!configured) {
        VTP_CUDA(cudaFuncSetAttribute(attn_fwd_pipe_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM));
        VTP_CUDA(cudaFuncSetAttribute(attn_fwd_pipe_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM));
        VTP_CUDA(cudaFuncSetAttribute(attn_fwd_pipe_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM));
        configured = true;
    }
    const int njobs = p.B * p.H;
    const int grid = njobs < num_sms() ? njobs : num_sms();
    if (p.HW == 256 && getenv("VTP_ATTN_PIPE_GENERIC") == nullptr) {
        const char* ts = getenv("VTP_ATTN_PIPE_TSO");
        if (ts ? ts[0] != '0' : VTP_ATTN_TSO_DEFAULT) {
            CUtensorMap tmo;  // out [B*T][D] bf16: 32-row x 64-column boxes (one head, one warp's rows)
            uint64_t dims[2] = {(uint64_t)p.D, (uint64_t)p.B * p.T}, strides[1] = {(uint64_t)p.D * 2};
            uint32_t box[2] = {64, 32};
            int rc = make_tmap_bf16(&tmo, p.out, 2, dims, strides, box);
            if (rc) return rc;
            attn_fwd_pipe_kernel<true, true><<<grid, PIPE_THREADS, PIPE_SMEM, st>>>(tm, tmo, p);
        } else {
            attn_fwd_pipe_kernel<true, false><<<grid, PIPE_THREADS, PIPE_SMEM, st>>>(tm, tm, p);
        }
    } else {
        attn_fwd_pipe_kernel<false, false><<<grid, PIPE_THREADS, PIPE_SMEM, st>>>(tm, tm, p);
    }
    VTP_LAUNCH_CHECK();
    return VTP_OK;
}

}  // namespace vtp
