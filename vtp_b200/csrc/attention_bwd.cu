// vtp_b200 — fused self-attention BACKWARD for short sequences (T = prefix + HW, prefix <= 1, HW <= 256) on tcgen05.
//
// Gradient of layers/attention.py:110-126 (RoPE + SDPA) w.r.t. the packed pre-RoPE qkv projection output:
//   P = exp(s·QKᵀ − lse)      dV = Pᵀ dO      dP = dO Vᵀ      dS = s · P ∘ (dP − δ),  δ_i = Σ_d dO_id O_id
//   dQ = dS K                 dK = dSᵀ Q      then RoPEᵀ on dQ, dK (rotation is linear: layers/attention.py:12-23)
//
// One CTA per (head, image).  All five GEMMs run on tcgen05 with the operands exactly as TMA lands them (128-row x
// 64-col bf16 tiles of Q, K, V out of the packed qkv buffer and of dO): the transposes are expressed through the UMMA
// major-ness bits —  Pᵀ / dSᵀ are MN-major A operands read from the same swizzled smem tile that serves dS as a K-major
// A operand for dQ; dO, Q, K are MN-major B operands.  Work is split in (key half kh, query tile t) steps:
//   MMA1: S = Q_t K_khᵀ, dP = dO_t V_khᵀ   (TMEM cols 256..511)      -> 128 row-threads build P, dS (bf16, smem)
//   MMA2: dV_kh += Pᵀ dO_t, dK_kh += dSᵀ Q_t (TMEM 0..127), dQ_t += dS K_kh (TMEM 128..255)
// TMEM: dK|dV (128) + dQ_0|dQ_1 (128) + S|dP (256) = 512 columns.
// The cls token (prefix) is handled on CUDA cores as in the forward kernel: as an extra key column by the row threads
// (rank-1 updates + a warp-reduced column for dK_0/dV_0) and as an extra query row by a spare warp.
#include <stdlib.h>

#include "host.h"
#include "ptx.cuh"

namespace vtp {

#ifndef VTP_ATTN_BWD_TSO_DEFAULT
#define VTP_ATTN_BWD_TSO_DEFAULT 1  // TMA-store row epilogues of the FULL path (measured: 502.5 -> 450.0 us, same bits)
#endif
static constexpr int AB_THREADS = 320;  // 8 row warps (2 per scheduler) + TMA/MMA warp + cls warp
static constexpr int BQ = 0, BK_ = 32768, BV = 65536, BDO = 98304, BP = 131072, BDS = 163840, BX = 196608;
// extras after BX: p0[264] | ds0[264] | dk0[64] | dv0[64] | pcol[2][128] | dscol[2][128] | barriers
static constexpr int X_P0 = 0, X_DS0 = 1056, X_DK0 = 2112, X_DV0 = 2368, X_PCOL = 2624, X_DSCOL = 3648, X_BAR = 4672;
static constexpr int AB_SMEM = BX + X_BAR + 128;

struct AttnBwdDev {
    const __nv_bfloat16* qkv;   // [B*T][3D] post-RoPE q,k ; v
    const __nv_bfloat16* o;     // [B*T][D]
    const __nv_bfloat16* dout;  // [B*T][D]
    const float* lse;           // [B][H][T]
    __nv_bfloat16* dqkv;        // [B*T][3D] gradient w.r.t. the PRE-RoPE qkv
    const __nv_bfloat16* rope_sin;  // [HW][64] or null (no RoPE: text tower)
    const __nv_bfloat16* rope_cos;
    int B, T, H, D, prefix, HW, causal, nkt;
    float scale, scale_log2;
    // packed mode (T <= 64): `pack` whole sequences share the 128-row tile, their prefix tokens (`rprefix` per sequence)
    // are ordinary rows / key columns (prefix == 0 above) and P, dS are masked block-diagonally
    int pack, rprefix;
    int tso;       // FULL only: last-key-half dK/dV rows and the dQ rows leave through shared memory + one TMA store per warp
    int prefetch;  // row threads pull their O rows (delta = dO.O) and lse towards L2/L1 before the tile loads are waited for
};

__device__ __forceinline__ float ex2f(float x) {  // ex2.approx.ftz: no denormal slow path (exp2f() costs 4 extra instr)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t sw_off(int row, int col) {
    return row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1));
}
__device__ __forceinline__ void load_row64(const uint8_t* tile, int row, float (&f)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 w = *reinterpret_cast<const uint4*>(tile + sw_off(row, c * 8));
        f[c * 8 + 0] = bf16_lo(w.x), f[c * 8 + 1] = bf16_hi(w.x), f[c * 8 + 2] = bf16_lo(w.y), f[c * 8 + 3] = bf16_hi(w.y);
        f[c * 8 + 4] = bf16_lo(w.z), f[c * 8 + 5] = bf16_hi(w.z), f[c * 8 + 6] = bf16_lo(w.w), f[c * 8 + 7] = bf16_hi(w.w);
    }
}
__device__ __forceinline__ void load_grow64(const __nv_bfloat16* g, float (&f)[64]) {
    const uint4* p = reinterpret_cast<const uint4*>(g);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 w = __ldg(p + c);
        f[c * 8 + 0] = bf16_lo(w.x), f[c * 8 + 1] = bf16_hi(w.x), f[c * 8 + 2] = bf16_lo(w.y), f[c * 8 + 3] = bf16_hi(w.y);
        f[c * 8 + 4] = bf16_lo(w.z), f[c * 8 + 5] = bf16_hi(w.z), f[c * 8 + 6] = bf16_lo(w.w), f[c * 8 + 7] = bf16_hi(w.w);
    }
}
__device__ __forceinline__ void store_row64(__nv_bfloat16* g, const float (&f)[64]) {
    uint4* p = reinterpret_cast<uint4*>(g);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 w;
        w.x = pack_bf16x2(f[c * 8], f[c * 8 + 1]), w.y = pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]);
        w.z = pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]), w.w = pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]);
        p[c] = w;
    }
}
// dx = RoPEᵀ dy :  dx[i] = dy[i] cos[i] + dy[i+32] sin[i+32] ;  dx[i+32] = dy[i+32] cos[i+32] − dy[i] sin[i]
__device__ __forceinline__ void rope_bwd64(float (&g)[64], const __nv_bfloat16* sin_row, const __nv_bfloat16* cos_row) {
    float sn[64], cs[64];
    load_grow64(sin_row, sn);
    load_grow64(cos_row, cs);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float a = g[i], b = g[i + 32];
        g[i] = a * cs[i] + b * sn[i + 32];
        g[i + 32] = b * cs[i + 32] - a * sn[i];
    }
}

// bf16 row (64 values) of a warp's 32-row slab -> the warp's 4 KB staging tile (128-byte swizzle, as the tensor map expects),
// then ONE TMA store of the [32 rows x 64 columns] box: replaces eight scattered 16-byte global stores per thread whose
// drain stalled the row threads at the end of every CTA (profiles/ncu_attn_r2b warp-state samples)
__device__ __forceinline__ void stage_store_row64(const CUtensorMap* tmap, uint8_t* warp_stage, int lane, const float (&f)[64], int x,
                                                  int y) {
    uint8_t* ob = warp_stage + lane * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 w;
        w.x = pack_bf16x2(f[c * 8], f[c * 8 + 1]), w.y = pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]);
        w.z = pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]), w.w = pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]);
        *reinterpret_cast<uint4*>(ob + ((c ^ (lane & 7)) << 4)) = w;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(tmap, warp_stage, x, y);
        bulk_commit();
    }
}

// FULL: HW == 256, no packing, no causal mask — every (query, key) pair of every step is valid, so the P / dS loop runs
// without per-element predicates (ncu, profiles/ncu_attn_r2a.md: ~30 executed instructions per score element in the
// generic loop, mostly mask predicates, selects and address arithmetic) and with hoisted swizzle offsets.
template <bool FULL>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const AttnBwdDev p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023) __trap();
    float* p0 = reinterpret_cast<float*>(smem + BX + X_P0);
    float* ds0 = reinterpret_cast<float*>(smem + BX + X_DS0);
    float* dk0 = reinterpret_cast<float*>(smem + BX + X_DK0);
    float* dv0 = reinterpret_cast<float*>(smem + BX + X_DV0);
    float* pcol = reinterpret_cast<float*>(smem + BX + X_PCOL);
    float* dscol = reinterpret_cast<float*>(smem + BX + X_DSCOL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BX + X_BAR);
    uint64_t* bar_ld = bars + 0;       // [2] tile loads
    uint64_t* bar_sdp = bars + 2;      // S,dP in TMEM
    uint64_t* bar_pds = bars + 3;      // P,dS in smem (128 arrivals), S/dP TMEM consumed
    uint64_t* bar_mma2 = bars + 4;     // dV/dK/dQ MMAs of a step complete
    uint64_t* bar_accfree = bars + 5;  // dK/dV accumulators drained by the epilogue (128 arrivals)
    uint64_t* bar_cls = bars + 6;      // p0/ds0 rows written by the cls warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = p.pack ? blockIdx.y * p.pack : blockIdx.y;
    const int D = p.D, T = p.T, prefix = p.prefix, HW = p.HW, nkt = p.nkt;
    const long row0 = (long)b * T;
    const int nsteps = nkt * nkt;
    const float lse_l2 = 1.4426950408889634f;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        if (FULL && p.tso) tma_prefetch_desc(&tm_dq);
        mbar_init(&bar_ld[0], 1), mbar_init(&bar_ld[1], 1), mbar_init(bar_sdp, 1), mbar_init(bar_pds, 256);
        mbar_init(bar_mma2, 1), mbar_init(bar_accfree, 256), mbar_init(bar_cls, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 128) dk0[threadIdx.x & 63] = 0.f, dv0[threadIdx.x & 63] = 0.f;
    if (p.prefetch && !p.pack && warp < 8) {
        // warp-state samples (profiles/ncu_attn_r2b): ~10 % of the row threads' time was the first-touch latency of their
        // O rows (one 128-byte line per row and head, straight from HBM) at the start of every CTA
        const int rr = (warp & 3) * 32 + lane, tt = warp >> 2;   // group g = warp >> 2 prefetches query tile g
        if (tt < nkt && 128 * tt + rr < HW)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.o + (row0 + prefix + 128 * tt + rr) * D + h * 64));
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns
    const uint32_t C_DV = 0, C_DK = 64, C_DQ = 128 /* + 64*t */, C_S = 256, C_DP = 384;

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------ TMA: tile i = rows [128i, 128i+128) of Q,K,V,dO
            for (int i = 0; i < nkt; ++i) {
                const int r = (int)row0 + prefix + 128 * i;
                mbar_expect_tx(&bar_ld[i], 4 * 16384);
                tma_load_2d(smem + BQ + i * 16384, &tm_qkv, &bar_ld[i], h * 64, r);
                tma_load_2d(smem + BK_ + i * 16384, &tm_qkv, &bar_ld[i], D + h * 64, r);
                tma_load_2d(smem + BV + i * 16384, &tm_qkv, &bar_ld[i], 2 * D + h * 64, r);
                tma_load_2d(smem + BDO + i * 16384, &tm_do, &bar_ld[i], h * 64, r);
            }
            const uint32_t id_nt = umma_idesc_bf16(128, 128, 0, 0);  // S, dP: A K-major, B K-major
            const uint32_t id_tn = umma_idesc_bf16(128, 64, 1, 1);   // dV, dK: A MN-major (Pᵀ), B MN-major
            const uint32_t id_nn = umma_idesc_bf16(128, 64, 0, 1);   // dQ: A K-major (dS), B MN-major (K)
            const uint32_t aQ = smem_u32(smem + BQ), aK = smem_u32(smem + BK_), aV = smem_u32(smem + BV);
            const uint32_t aDO = smem_u32(smem + BDO), aP = smem_u32(smem + BP), aDS = smem_u32(smem + BDS);
            auto mma1 = [&](int kh, int t) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_bf16_ss(tmem + C_S, umma_desc_sw128(aQ + t * 16384 + j * 32, 0, 1024),
                                 umma_desc_sw128(aK + kh * 16384 + j * 32, 0, 1024), id_nt, j > 0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_bf16_ss(tmem + C_DP, umma_desc_sw128(aDO + t * 16384 + j * 32, 0, 1024),
                                 umma_desc_sw128(aV + kh * 16384 + j * 32, 0, 1024), id_nt, j > 0);
                umma_commit(bar_sdp);
            };
            mbar_wait(&bar_ld[0], 0);
            tc_fence_after();
            mma1(0, 0);
            bool ld1_waited = false;
            for (int n = 0; n < nsteps; ++n) {
                const int kh = n / nkt, t = n % nkt;
                mbar_wait(bar_pds, n & 1);
                tc_fence_after();
                if (n + 1 < nsteps) {
                    if (!ld1_waited) mbar_wait(&bar_ld[1], 0), ld1_waited = true;
                    mma1((n + 1) / nkt, (n + 1) % nkt);
                }
                if (kh > 0 && t == 0) {  // dK/dV accumulators of the previous key half must be drained
                    mbar_wait(bar_accfree, (kh - 1) & 1);
                    tc_fence_after();
                }
                // P / dS tile: [128 q][128 keys] as 2 chunks of 64 keys.  As MN-major A (M = keys): LBO = chunk
                // stride 16384, SBO = 1024 (8 queries), K-step of 16 queries = 2048 B.  As K-major A (M = queries):
                // k-step 32 B inside a chunk, next chunk +16384.
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // reduction over 128 queries
                    const uint64_t bdo = umma_desc_sw128(aDO + t * 16384 + j * 2048, 8192, 1024);
                    const uint64_t bq = umma_desc_sw128(aQ + t * 16384 + j * 2048, 8192, 1024);
                    umma_bf16_ss(tmem + C_DV, umma_desc_sw128(aP + j * 2048, 16384, 1024), bdo, id_tn, (t > 0 || j > 0));
                    umma_bf16_ss(tmem + C_DK, umma_desc_sw128(aDS + j * 2048, 16384, 1024), bq, id_tn, (t > 0 || j > 0));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // reduction over 128 keys
                    const uint64_t ads = umma_desc_sw128(aDS + (j >> 2) * 16384 + (j & 3) * 32, 0, 1024);
                    const uint64_t bk = umma_desc_sw128(aK + kh * 16384 + j * 2048, 8192, 1024);
                    umma_bf16_ss(tmem + C_DQ + 64 * t, ads, bk, id_nn, (kh > 0 || j > 0));
                }
                umma_commit(bar_mma2);
            }
        }
    } else if (warp < 8) {
        // ---------------------------------------------------- row threads: 2 groups x 128 (one TMEM lane each).
        // group g builds the P/dS columns of 64-key chunk g in every step, owns query tile g (cls-key column terms,
        // dQ epilogue) and one of the two accumulators in the key-half epilogue (g=0: dV, g=1: dK).
        const int g = warp >> 2, q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t trow = tmem + (uint32_t(q4 * 32) << 16);
        // packed mode: row r = token (r % T) of sequence b + r / T; it sees the key columns of its own sequence only
        const int pseq = p.pack ? r / T : 0;
        const bool pvalid = p.pack && pseq < p.pack && b + pseq < p.B;
        const int ptok = r - pseq * T;
        const int my_tile = (nkt == 2) ? g : 0;
        const bool owns_tile = (nkt == 2) || (g == 0);
        float lse_i[2], delta_i[2];
        float ds_own = 0.f;
        const __nv_bfloat16* kcls = p.qkv + row0 * 3 * D + D + h * 64;
        const __nv_bfloat16* vcls = p.qkv + row0 * 3 * D + 2 * D + h * 64;

        for (int n = 0; n < nsteps; ++n) {
            const int kh = n / nkt, t = n % nkt;
            const int qi = 128 * t + r;  // patch index of my query row in this step
            const bool qvalid = p.pack ? pvalid : qi < HW;
            if (kh == 0) {
                // per-query-tile scalars (first visit of tile t): lse, delta = dO·O; the owner group also does the
                // cls-key column
                mbar_wait(&bar_ld[t], 0);
                lse_i[t] = 0.f, delta_i[t] = 0.f;
                const bool mine = owns_tile && t == my_tile && prefix > 0;
                float p0v = 0.f, ds0v = 0.f;
                if (qvalid) {
                    const long grow = row0 + prefix + qi;
                    lse_i[t] = p.pack ? p.lse[((long)(b + pseq) * p.H + h) * T + ptok] : p.lse[((long)b * p.H + h) * T + prefix + qi];
                    float dof[64], tmpf[64];
                    load_row64(smem + BDO + t * 16384, r, dof);
                    load_grow64(p.o + grow * D + h * 64, tmpf);
                    float dl = 0.f;
#pragma unroll
                    for (int d = 0; d < 64; ++d) dl += dof[d] * tmpf[d];
                    delta_i[t] = dl;
                    if (mine) {
                        load_grow64(vcls, tmpf);
                        float dp0 = 0.f;
#pragma unroll
                        for (int d = 0; d < 64; ++d) dp0 += dof[d] * tmpf[d];
                        load_row64(smem + BQ + t * 16384, r, dof);  // q row
                        load_grow64(kcls, tmpf);
                        float s0 = 0.f;
#pragma unroll
                        for (int d = 0; d < 64; ++d) s0 += dof[d] * tmpf[d];
                        p0v = ex2f(s0 * p.scale_log2 - lse_i[t] * lse_l2);
                        ds0v = p.scale * p0v * (dp0 - dl);
                    }
                }
                if (mine) {
                    // column reductions for the cls key: dV_0 += Σ_i p_i0 dO_i ; dK_0 += Σ_i ds_i0 q_i.  The per-row
                    // scalars go through smem, then thread (which, d) walks the 128 rows of the dO / Q tile.
                    ds_own = ds0v;
                    pcol[g * 128 + r] = p0v, dscol[g * 128 + r] = ds0v;
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
                    const int d = r & 63;
                    const bool isk = r >= 64;
                    const uint8_t* tile = smem + (isk ? BQ : BDO) + t * 16384;
                    const float* colv = (isk ? dscol : pcol) + g * 128;
                    // element (row i, dim d) of a swizzled 128-byte-row tile: row i = 8 a + k sits at a*1024 + k*128 +
                    // (((d >> 3) ^ k) << 4) + 2 (d & 7): the eight k-offsets are loop invariant, a*1024 an immediate
                    uint32_t ok[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) ok[k] = (uint32_t)(k * 128 + ((((d >> 3) ^ k) << 4) | ((d & 7) << 1)));
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int a = 0; a < 16; ++a) {
                        const float4 c0 = *reinterpret_cast<const float4*>(colv + 8 * a);
                        const float4 c1 = *reinterpret_cast<const float4*>(colv + 8 * a + 4);
                        const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t e = *reinterpret_cast<const uint16_t*>(tile + a * 1024 + ok[k]);
                            if (k & 1) acc1 = fmaf(cv[k], __uint_as_float(e << 16), acc1);
                            else acc0 = fmaf(cv[k], __uint_as_float(e << 16), acc0);
                        }
                    }
                    atomicAdd(isk ? &dk0[d] : &dv0[d], acc0 + acc1);
                }
            }
            mbar_wait(bar_sdp, n & 1);
            tc_fence_after();
            if (n > 0) mbar_wait(bar_mma2, (n - 1) & 1);  // P/dS smem tiles free again
            const float lsc = lse_i[t] * lse_l2, dl = delta_i[t];
            if constexpr (FULL) {
                const float sl2 = p.scale_log2, sc = p.scale, dls = dl * p.scale;
                uint8_t* pb_ = smem + BP + g * 16384 + r * 128;   // my row of 64-key chunk g of the P / dS tiles
                uint8_t* db_ = smem + BDS + g * 16384 + r * 128;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c32 = 2 * g + cc;
                    uint32_t rs[32], rd[32];
                    tmem_ld_32x32(trow + C_S + c32 * 32, rs);
                    tmem_ld_32x32(trow + C_DP + c32 * 32, rd);
                    tmem_ld_wait();
                    uint32_t pk[16], dk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float pa = ex2f(fmaf(__uint_as_float(rs[i]), sl2, -lsc));
                        const float pb = ex2f(fmaf(__uint_as_float(rs[i + 1]), sl2, -lsc));
                        const float da = pa * fmaf(__uint_as_float(rd[i]), sc, -dls);       // s P (dP - delta)
                        const float db = pb * fmaf(__uint_as_float(rd[i + 1]), sc, -dls);
                        pk[i >> 1] = pack_bf16x2(pa, pb);
                        dk[i >> 1] = pack_bf16x2(da, db);
                    }
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        const uint32_t off = (uint32_t)(((cc * 4 + v4) ^ (r & 7)) << 4);
                        *reinterpret_cast<uint4*>(pb_ + off) = make_uint4(pk[v4 * 4], pk[v4 * 4 + 1], pk[v4 * 4 + 2], pk[v4 * 4 + 3]);
                        *reinterpret_cast<uint4*>(db_ + off) = make_uint4(dk[v4 * 4], dk[v4 * 4 + 1], dk[v4 * 4 + 2], dk[v4 * 4 + 3]);
                    }
                }
            } else {
            const int kmin = p.pack ? pseq * T : 0;
            const int kmax = p.pack ? kmin + T : (p.causal ? min(HW, qi + 1) : HW);
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c32 = 2 * g + cc;
                uint32_t rs[32], rd[32];
                tmem_ld_32x32(trow + C_S + c32 * 32, rs);
                tmem_ld_32x32(trow + C_DP + c32 * 32, rd);
                tmem_ld_wait();
                uint32_t pk[16], dk[16];
                const int kbase = 128 * kh + c32 * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float pa = 0.f, pb = 0.f, da = 0.f, db = 0.f;
                    if (qvalid && kbase + i >= kmin && kbase + i < kmax) {
                        pa = ex2f(__uint_as_float(rs[i]) * p.scale_log2 - lsc);
                        da = p.scale * pa * (__uint_as_float(rd[i]) - dl);
                    }
                    if (qvalid && kbase + i + 1 >= kmin && kbase + i + 1 < kmax) {
                        pb = ex2f(__uint_as_float(rs[i + 1]) * p.scale_log2 - lsc);
                        db = p.scale * pb * (__uint_as_float(rd[i + 1]) - dl);
                    }
                    pk[i >> 1] = pack_bf16x2(pa, pb);
                    dk[i >> 1] = pack_bf16x2(da, db);
                }
                uint8_t* pb_ = smem + BP + g * 16384;   // 64-key chunk g of the P / dS tiles
                uint8_t* db_ = smem + BDS + g * 16384;
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const uint32_t off = sw_off(r, cc * 32 + v4 * 8);
                    *reinterpret_cast<uint4*>(pb_ + off) = make_uint4(pk[v4 * 4], pk[v4 * 4 + 1], pk[v4 * 4 + 2], pk[v4 * 4 + 3]);
                    *reinterpret_cast<uint4*>(db_ + off) = make_uint4(dk[v4 * 4], dk[v4 * 4 + 1], dk[v4 * 4 + 2], dk[v4 * 4 + 3]);
                }
            }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar_pds);

            if (t == nkt - 1) {
                // ------------ epilogue of key half kh: I own key row kj = 128*kh + r; group 0 -> dV, group 1 -> dK
                mbar_wait(bar_mma2, n & 1);
                tc_fence_after();
                const int kj = 128 * kh + r;
                // last key half: the P / dS tiles are idle (their last MMAs have completed) and serve as store staging
                const bool tso = FULL && p.tso && kh == nkt - 1;
                uint32_t a0[32], a1[32];
                float gq[64];
                if (prefix > 0) mbar_wait(bar_cls, 0);
                tmem_ld_32x32(trow + (g == 0 ? C_DV : C_DK), a0);
                tmem_ld_32x32(trow + (g == 0 ? C_DV : C_DK) + 32, a1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) gq[i] = __uint_as_float(a0[i]), gq[32 + i] = __uint_as_float(a1[i]);
                tc_fence_before();
                mbar_arrive(bar_accfree);
                if (p.pack ? pvalid : kj < HW) {
                    if (g == 0) {
                        if (prefix > 0) {
                            float f[64];
                            load_grow64(p.dout + row0 * D + h * 64, f);  // dO of the cls query
                            const float pc = p0[kj];
#pragma unroll
                            for (int d = 0; d < 64; ++d) gq[d] += pc * f[d];
                        }
                        if (tso)
                            stage_store_row64(&tm_dq, smem + BP + q4 * 4096, lane, gq, 2 * D + h * 64,
                                              (int)(row0 + prefix) + 128 * kh + q4 * 32);
                        else store_row64(p.dqkv + (row0 + prefix + kj) * 3 * D + 2 * D + h * 64, gq);
                    } else {
                        if (prefix > 0) {
                            float f[64];
                            load_grow64(p.qkv + row0 * 3 * D + h * 64, f);  // q of the cls query
                            const float dc = ds0[kj];
#pragma unroll
                            for (int d = 0; d < 64; ++d) gq[d] += dc * f[d];
                        }
                        const int pos = p.pack ? ptok - p.rprefix : kj;  // patch position (prefix tokens are not rotated)
                        if (p.rope_sin && pos >= 0) rope_bwd64(gq, p.rope_sin + (long)pos * 64, p.rope_cos + (long)pos * 64);
                        if (tso)
                            stage_store_row64(&tm_dq, smem + BP + 16384 + q4 * 4096, lane, gq, D + h * 64,
                                              (int)(row0 + prefix) + 128 * kh + q4 * 32);
                        else store_row64(p.dqkv + (row0 + prefix + kj) * 3 * D + D + h * 64, gq);
                    }
                }
            }
        }
        // ------------ dQ epilogue of my query tile (all steps done; last bar_mma2 phase already observed above)
        if (owns_tile) {
            const int t = my_tile;
            const int qi = 128 * t + r;
            uint32_t a0[32], a1[32];
            tmem_ld_32x32(trow + C_DQ + 64 * t, a0);
            tmem_ld_32x32(trow + C_DQ + 64 * t + 32, a1);
            tmem_ld_wait();
            if (p.pack ? pvalid : qi < HW) {
                float gq[64];
#pragma unroll
                for (int i = 0; i < 32; ++i) gq[i] = __uint_as_float(a0[i]), gq[32 + i] = __uint_as_float(a1[i]);
                if (prefix > 0) {
                    float f[64];
                    load_grow64(kcls, f);
#pragma unroll
                    for (int d = 0; d < 64; ++d) gq[d] += ds_own * f[d];
                }
                const int pos = p.pack ? ptok - p.rprefix : qi;
                if (p.rope_sin && pos >= 0) rope_bwd64(gq, p.rope_sin + (long)pos * 64, p.rope_cos + (long)pos * 64);
                if (FULL && p.tso)
                    stage_store_row64(&tm_dq, smem + BDS + g * 16384 + q4 * 4096, lane, gq, h * 64,
                                      (int)(row0 + prefix) + 128 * t + q4 * 32);
                else store_row64(p.dqkv + (row0 + prefix + qi) * 3 * D + h * 64, gq);
            }
        }
        if (FULL && p.tso && lane == 0) bulk_wait0();  // my warp's row stores have left shared memory and are complete
        tc_fence_before();
    } else {
        // ---------------------------------------------------- warp 9: the cls query row (prefix == 1)
        if (prefix > 0) {
            for (int i = 0; i < nkt; ++i) mbar_wait(&bar_ld[i], 0);
            float q0[64], do0[64];
            load_grow64(p.qkv + row0 * 3 * D + h * 64, q0);
            load_grow64(p.dout + row0 * D + h * 64, do0);
            float delta0 = 0.f;
            {
                float o0[64];
                load_grow64(p.o + row0 * D + h * 64, o0);
#pragma unroll
                for (int d = 0; d < 64; ++d) delta0 += do0[d] * o0[d];
            }
            const float lse0 = p.lse[((long)b * p.H + h) * T] * lse_l2;
            for (int i = 0; i < 8; ++i) {
                const int kj = lane + 32 * i;
                if (kj < 128 * nkt) {
                    float pv = 0.f, dsv = 0.f;
                    if (kj < HW) {
                        float f[64];
                        load_row64(smem + BK_, kj, f);  // K tiles are contiguous: row kj of [256][128B]
                        float s = 0.f;
#pragma unroll
                        for (int d = 0; d < 64; ++d) s += q0[d] * f[d];
                        load_row64(smem + BV, kj, f);
                        float dp = 0.f;
#pragma unroll
                        for (int d = 0; d < 64; ++d) dp += do0[d] * f[d];
                        pv = ex2f(s * p.scale_log2 - lse0);
                        dsv = p.scale * pv * (dp - delta0);
                    }
                    p0[kj] = pv, ds0[kj] = dsv;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_cls);
            // cls-cls term
            float kc[64], vc[64];
            load_grow64(p.qkv + row0 * 3 * D + D + h * 64, kc);
            load_grow64(p.qkv + row0 * 3 * D + 2 * D + h * 64, vc);
            float s00 = 0.f, dp00 = 0.f;
#pragma unroll
            for (int d = 0; d < 64; ++d) s00 += q0[d] * kc[d], dp00 += do0[d] * vc[d];
            const float p00 = ex2f(s00 * p.scale_log2 - lse0);
            const float ds00 = p.scale * p00 * (dp00 - delta0);
            // dQ_0[d] = Σ_j ds_0j k_j[d] + ds_00 k_0[d]; lane owns dims 2*lane, 2*lane+1
            float a0 = 0.f, a1 = 0.f;
            for (int kj = 0; kj < HW; ++kj) {
                const float dsv = ds0[kj];
                const uint32_t w = *reinterpret_cast<const uint32_t*>(smem + BK_ + sw_off(kj, 2 * lane));
                a0 += dsv * bf16_lo(w), a1 += dsv * bf16_hi(w);
            }
            {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p.qkv + row0 * 3 * D + D + h * 64) + lane);
                a0 += ds00 * bf16_lo(w), a1 += ds00 * bf16_hi(w);
            }
            *reinterpret_cast<uint32_t*>(p.dqkv + row0 * 3 * D + h * 64 + 2 * lane) = pack_bf16x2(a0, a1);
            // stash p00/ds00 for the final dK_0/dV_0 write
            if (lane == 0) p0[260] = p00, ds0[260] = ds00;
        }
    }

    __syncthreads();
    if (warp == 9 && prefix > 0) {
        // dV_0 = Σ_i p_i0 dO_i (+ p_00 dO_0) ;  dK_0 = Σ_i ds_i0 q_i (+ ds_00 q_0)   — cls key row, no RoPE
        const float p00 = p0[260], ds00 = ds0[260];
        const uint32_t wdo = __ldg(reinterpret_cast<const uint32_t*>(p.dout + row0 * D + h * 64) + lane);
        const uint32_t wq = __ldg(reinterpret_cast<const uint32_t*>(p.qkv + row0 * 3 * D + h * 64) + lane);
        const float v0 = dv0[2 * lane] + p00 * bf16_lo(wdo), v1 = dv0[2 * lane + 1] + p00 * bf16_hi(wdo);
        const float k0 = dk0[2 * lane] + ds00 * bf16_lo(wq), k1 = dk0[2 * lane + 1] + ds00 * bf16_hi(wq);
        *reinterpret_cast<uint32_t*>(p.dqkv + row0 * 3 * D + 2 * D + h * 64 + 2 * lane) = pack_bf16x2(v0, v1);
        *reinterpret_cast<uint32_t*>(p.dqkv + row0 * 3 * D + D + h * 64 + 2 * lane) = pack_bf16x2(k0, k1);
    }
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace vtp

using namespace vtp;

extern "C" int vtp_attention_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv,
                                 const void* rope_sin, const void* rope_cos, int B, int T, int H, int prefix, int causal,
                                 vtp_stream_t st) {
    VTP_CHECK_ARG(qkv && o && dout && lse && dqkv && B > 0 && T > 0 && H > 0, "attention_bwd: bad args");
    VTP_CHECK_ARG(prefix == 0 || prefix == 1, "attention_bwd: prefix must be 0 or 1");
    VTP_CHECK_ARG(!(prefix && causal), "attention_bwd: causal + prefix is not supported");
    VTP_CHECK_ARG((rope_sin == nullptr) == (rope_cos == nullptr), "attention_bwd: rope tables");
    const int HW = T - prefix;
    VTP_CHECK_ARG(HW >= 1 && HW <= 256, "attention_bwd: %d non-prefix tokens not in [1,256]", HW);
    VTP_CHECK_ARG(B <= 65535, "attention_bwd: grid too large");
    const int D = H * 64;
    AttnBwdDev p;
    p.qkv = (const __nv_bfloat16*)qkv, p.o = (const __nv_bfloat16*)o, p.dout = (const __nv_bfloat16*)dout;
    p.lse = lse, p.dqkv = (__nv_bfloat16*)dqkv;
    p.rope_sin = (const __nv_bfloat16*)rope_sin, p.rope_cos = (const __nv_bfloat16*)rope_cos;
    p.B = B, p.T = T, p.H = H, p.D = D, p.prefix = prefix, p.HW = HW, p.causal = causal;
    p.nkt = HW > 128 ? 2 : 1;
    p.scale = 0.125f, p.scale_log2 = 0.125f * 1.4426950408889634f;
    p.pack = 0, p.rprefix = prefix;
    p.prefetch = getenv("VTP_ATTN_BWD_NO_PREFETCH") == nullptr;
    {
        const char* ts = getenv("VTP_ATTN_BWD_TSO");
        p.tso = ts ? ts[0] != '0' : VTP_ATTN_BWD_TSO_DEFAULT;
    }
    if (!causal && T <= 64 && B > 1 && getenv("VTP_ATTN_NO_PACK") == nullptr) {
        p.pack = 128 / T;
        p.prefix = 0, p.HW = T, p.nkt = 1;
    }
    CUtensorMap tq, td, tdq;
    {   // dqkv [B*T][3D]: 32-row x 64-column boxes (one warp's rows of one head's dq / dk / dv)
        uint64_t dims[2] = {(uint64_t)3 * D, (uint64_t)B * T}, strides[1] = {(uint64_t)3 * D * 2};
        uint32_t box[2] = {64, 32};
        int rc = make_tmap_bf16(&tdq, dqkv, 2, dims, strides, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)3 * D, (uint64_t)B * T}, strides[1] = {(uint64_t)3 * D * 2};
        uint32_t box[2] = {64, 128};
        int rc = make_tmap_bf16(&tq, qkv, 2, dims, strides, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)D, (uint64_t)B * T}, strides[1] = {(uint64_t)D * 2};
        uint32_t box[2] = {64, 128};
        int rc = make_tmap_bf16(&td, dout, 2, dims, strides, box);
        if (rc) return rc;
    }
    static bool configured = false;
    if (!configured) {
        VTP_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
        VTP_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
        configured = true;
    }
    const dim3 grid(H, p.pack ? ceil_div(B, p.pack) : B);
    if (p.HW == 256 && !p.pack && !p.causal && getenv("VTP_ATTN_BWD_GENERIC") == nullptr)
        attn_bwd_kernel<true><<<grid, AB_THREADS, AB_SMEM, (cudaStream_t)st>>>(tq, td, tdq, p);
    else
        attn_bwd_kernel<false><<<grid, AB_THREADS, AB_SMEM, (cudaStream_t)st>>>(tq, td, tdq, p);
    VTP_LAUNCH_CHECK();
    return VTP_OK;
}
