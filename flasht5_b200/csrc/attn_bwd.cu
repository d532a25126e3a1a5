// FlashAttention-2 backward with additive (T5) bias for sm_100a: dQ, dK, dV and the dS tiles that
// become dBias, in ONE fused kernel (5 tensor-core contractions per tile instead of the
// reference's 7), plus three small helper kernels (delta, dQ convert, dBias reduce).
//
// Replaces /root/reference/src/model/ops/flash_attention_v2_bias.py:
//   _bwd_preprocess :516-556, _bwd_kv_kernel :559-745, _bwd_q_kernel :748-905, ds.sum(0) :214-215.
//
// One CTA = one (batch, head, 128-key block).  K and V stay resident in shared memory; the CTA
// walks the query sequence in 128-row blocks:
//     S  = Q K^T            (SS MMA -> TMEM)          dP = dO V^T          (SS MMA -> TMEM)
//     P  = exp(S*s + bias - L),  dS = P * (dP - delta)      (two warpgroups, 64 columns each)
//     P, dS -> shared memory as 16-bit, 128B-swizzled [m][n] tiles
//     dV += P^T dO,  dK += dS^T Q   (A = MN-major smem)     dQ_blk = dS K  (A = K-major smem)
//     dS tile  -> global with a TMA store (per-batch dS workspace or dBias directly)
//     dQ_blk   -> fp32 accumulator in global with a TMA reduce-add (cp.reduce.async.bulk.tensor)
//
//   warp 4 : TMA producer (K, V once; Q / dO ring)     warp 6 : TMA producer for bias halves
//   warp 5 : tcgen05.mma issuer                        warps 0-3 and 8-11 : compute warpgroups
//
// TMEM columns: S [0,128) | dP [128,256) | dV [256,256+D) | dK [256+D,256+2D) | dQ [256+2D, 256+3D)
// (D = 128: dQ aliases the S columns).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kBM = 128;
constexpr int kBN = 128;
constexpr int kHalfBytes = 128 * 64 * 2;   // one [128][64] 16-bit swizzled half tile
constexpr float kLog2e = 1.4426950408889634f;

template <int kD>
struct BwdCfg {
    static constexpr int kRowBytes = (kD >= 64 ? 64 : kD) * 2;
    static constexpr int kBoxes = kD >= 64 ? kD / 64 : 1;
    static constexpr int kBoxBytes = kBM * kRowBytes;
    static constexpr int kTileBytes = kBM * kD * 2;
    static constexpr uint32_t kSwizzle = kRowBytes == 128 ? kSwz128 : (kRowBytes == 64 ? kSwz64 : kSwz32);
    static constexpr int kQStages = kD <= 64 ? 2 : 1;
    static constexpr bool kLookahead = kQStages == 2;      // issue S/dP of block i+1 before the dV/dK/dQ of block i
    // dQ staging (16-bit, [128][min(D,64)] boxes, 128B swizzle when D >= 64) aliases the P tile
    static constexpr int kDqBoxCols = kD >= 64 ? 64 : kD;
    static constexpr int kDqBoxes = kD / kDqBoxCols;
    static constexpr int kDqBoxBytes = kBM * kDqBoxCols * 2;
    static constexpr bool kDqAliasesDs = kDqBoxes * kDqBoxBytes > 2 * kHalfBytes;
    static constexpr int kK = 0;
    static constexpr int kV = kK + kTileBytes;
    static constexpr int kQ = kV + kTileBytes;
    static constexpr int kDO = kQ + kQStages * kTileBytes;
    static constexpr int kBias = kDO + kQStages * kTileBytes;
    static constexpr int kP = kBias + 2 * kHalfBytes;
    static constexpr int kDS = kP + 2 * kHalfBytes;
    static constexpr int kBars = kDS + 2 * kHalfBytes;
    static constexpr int kNumBars = 1 + 2 * kQStages + 4 + 6;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static constexpr int kColS = 0;
    static constexpr int kColDP = 128;
    static constexpr int kColDV = 256;
    static constexpr int kColDK = 256 + kD;
    static constexpr bool kDqAliasS = (256 + 3 * kD) > 512;
    static constexpr int kColDQ = kDqAliasS ? 0 : 256 + 2 * kD;
    // columns of dQ each compute warpgroup drains (D = 16: warpgroup 0 takes all of them)
    static constexpr int kDqColsPerWg = kD >= 32 ? kD / 2 : kD;
};

template <int kN>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
    if constexpr (kN == 32) tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(r));
    else tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(r));
}

}  // namespace

template <int kD, bool kBf16, int kBiasMode, bool kCausal>
__global__ void __launch_bounds__(384, 1)
attn_bwd_kernel(const __grid_constant__ AttnBwdKernelParams p) {
    using C = BwdCfg<kD>;
    extern __shared__ __align__(1024) uint8_t smem[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- work decode: batch fastest (bias tiles shared in L2), late key blocks first when causal ----
    const int nnb = p.num_n_blocks;
    int bid = blockIdx.x;
    const int b = bid % p.B;
    bid /= p.B;
    const int nb = kCausal ? (bid % nnb) : (nnb - 1 - bid % nnb);   // causal: early key blocks see the most rows
    const int h = bid / nnb;
    const int col0 = nb * kBN;
    const int pseq = p.N - p.M;

    int i_start = 0;
    if (kCausal) {
        const int first_row = col0 - pseq;                      // first query row that sees key col0
        i_start = first_row <= 0 ? 0 : first_row / kBM;
    }
    const int n_iter = p.num_m_blocks > i_start ? p.num_m_blocks - i_start : 0;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
    uint64_t* kv_full = bars;
    uint64_t* qdo_full = bars + 1;
    uint64_t* qdo_empty = qdo_full + C::kQStages;
    uint64_t* b_full = qdo_empty + C::kQStages;    // [2] one per 64-column half
    uint64_t* b_empty = b_full + 2;
    uint64_t* sdp_full = b_empty + 2;
    uint64_t* sdp_empty = sdp_full + 1;
    uint64_t* pds_full = sdp_empty + 1;
    uint64_t* dq_full = pds_full + 1;
    uint64_t* dq_empty = dq_full + 1;
    uint64_t* acc_full = dq_empty + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kTmemSlot);

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("b200t5: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        mbar_init(kv_full, 1);
        for (int i = 0; i < C::kQStages; ++i) {
            mbar_init(qdo_full + i, 1);
            mbar_init(qdo_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 4);
        }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_empty, 8);
        mbar_init(pds_full, 1);
        mbar_init(dq_full, 1);
        mbar_init(dq_empty, 1);
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc<512>(tmem_slot);
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&p.map_q);
        tma_prefetch_desc(&p.map_k);
        tma_prefetch_desc(&p.map_v);
        tma_prefetch_desc(&p.map_do);
        tma_prefetch_desc(&p.map_dq);
        if (kBiasMode == 1) tma_prefetch_desc(&p.map_bias);
        if (kBiasMode != 0) tma_prefetch_desc(&p.map_ds);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const bool is_compute = (warp < 4) || (warp >= 8);

    if (!is_compute) {
        // =============================== control warps (4..7) ===============================
        setmaxnreg_dec<64>();
        if (warp == 4 && lane == 0 && n_iter > 0) {
            // ---- K, V once; then the Q / dO ring ----
            mbar_arrive_expect_tx(kv_full, 2 * C::kTileBytes);
#pragma unroll
            for (int bx = 0; bx < C::kBoxes; ++bx) {
                tma_load_4d(smem + C::kK + bx * C::kBoxBytes, &p.map_k, kv_full, bx * 64, col0, h, b);
                tma_load_4d(smem + C::kV + bx * C::kBoxBytes, &p.map_v, kv_full, bx * 64, col0, h, b);
            }
            for (int k = 0; k < n_iter; ++k) {
                const int s = k % C::kQStages;
                const int mrow0 = (i_start + k) * kBM;
                mbar_wait_producer(qdo_empty + s, ((k / C::kQStages) & 1) ^ 1);
                mbar_arrive_expect_tx(qdo_full + s, 2 * C::kTileBytes);
#pragma unroll
                for (int bx = 0; bx < C::kBoxes; ++bx) {
                    tma_load_4d(smem + C::kQ + s * C::kTileBytes + bx * C::kBoxBytes, &p.map_q, qdo_full + s, bx * 64,
                                mrow0, h, b);
                    tma_load_4d(smem + C::kDO + s * C::kTileBytes + bx * C::kBoxBytes, &p.map_do, qdo_full + s,
                                bx * 64, mrow0, h, b);
                }
            }
        } else if (warp == 6 && lane == 0 && kBiasMode == 1) {
            const int hb = p.bias_h_bcast ? 0 : h;
            const int bb = p.bias_b_bcast ? 0 : b;
            for (int k = 0; k < n_iter; ++k) {
                const int mrow0 = (i_start + k) * kBM;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    mbar_wait_producer(b_empty + hh, (k & 1) ^ 1);
                    mbar_arrive_expect_tx(b_full + hh, kHalfBytes);
                    tma_load_4d(smem + C::kBias + hh * kHalfBytes, &p.map_bias, b_full + hh, col0 + hh * 64, mrow0, hb,
                                bb);
                }
            }
        } else if (warp == 5 && n_iter > 0) {
            // ---- MMA issuer: whole warp runs the loop (warp-uniform descriptor math), one elected lane issues ----
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, 128, 128, false, false);    // S, dP
            constexpr uint32_t idesc_dkv = make_idesc(kBf16, 128, kD, true, true);     // dV, dK
            constexpr uint32_t idesc_dq = make_idesc(kBf16, 128, kD, false, true);     // dQ
            constexpr uint32_t sbo = 8 * C::kRowBytes;
            constexpr uint32_t hi_op = sdesc_hi(sbo, C::kSwizzle);       // Q, dO, K, V tiles (either major)
            constexpr uint32_t hi_pds = sdesc_hi(1024, kSwz128);         // P, dS tiles (either major)
            const uint32_t k_lo = sdesc_lo(smem_u32(smem + C::kK), 16);
            const uint32_t v_lo = sdesc_lo(smem_u32(smem + C::kV), 16);
            const uint32_t k_mn_lo = sdesc_lo(smem_u32(smem + C::kK), C::kBoxBytes);
            const uint32_t q_lo0 = sdesc_lo(smem_u32(smem + C::kQ), 16);
            const uint32_t do_lo0 = sdesc_lo(smem_u32(smem + C::kDO), 16);
            const uint32_t q_mn_lo0 = sdesc_lo(smem_u32(smem + C::kQ), C::kBoxBytes);
            const uint32_t do_mn_lo0 = sdesc_lo(smem_u32(smem + C::kDO), C::kBoxBytes);
            const uint32_t p_mn_lo = sdesc_lo(smem_u32(smem + C::kP), kHalfBytes);
            const uint32_t ds_mn_lo = sdesc_lo(smem_u32(smem + C::kDS), kHalfBytes);
            const uint32_t ds_k_lo = sdesc_lo(smem_u32(smem + C::kDS), 16);
            const uint32_t tm_s = tmem_base + C::kColS;
            const uint32_t tm_dp = tmem_base + C::kColDP;
            const uint32_t tm_dv = tmem_base + C::kColDV;
            const uint32_t tm_dk = tmem_base + C::kColDK;
            const uint32_t tm_dq = tmem_base + C::kColDQ;

            auto issue_s_dp = [&](int k) {
                const uint32_t so = (k % C::kQStages) * (C::kTileBytes >> 4);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk) {
                        const uint32_t off = ((kk / 4) * C::kBoxBytes + (kk % 4) * 32) >> 4;
                        umma_ss2(tm_s, q_lo0 + so + off, hi_op, k_lo + off, hi_op, idesc_s, kk > 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk) {
                        const uint32_t off = ((kk / 4) * C::kBoxBytes + (kk % 4) * 32) >> 4;
                        umma_ss2(tm_dp, do_lo0 + so + off, hi_op, v_lo + off, hi_op, idesc_s, kk > 0 ? 1u : 0u);
                    }
                    umma_commit(sdp_full);
                }
                __syncwarp();
            };
            auto issue_grads = [&](int k) {
                const int s = k % C::kQStages;
                const uint32_t so = s * (C::kTileBytes >> 4);
                const uint32_t acc = k > 0 ? 1u : 0u;
                if (leader) {
                    // dV += P^T dO ; dK += dS^T Q      (K dimension = the 128 query rows of this block)
#pragma unroll
                    for (int kk = 0; kk < kBM / 16; ++kk)
                        umma_ss2(tm_dv, p_mn_lo + kk * (2048 >> 4), hi_pds,
                                 do_mn_lo0 + so + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dkv,
                                 (acc | (kk > 0)) ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < kBM / 16; ++kk)
                        umma_ss2(tm_dk, ds_mn_lo + kk * (2048 >> 4), hi_pds,
                                 q_mn_lo0 + so + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dkv,
                                 (acc | (kk > 0)) ? 1u : 0u);
                    umma_commit(qdo_empty + s);
                }
                __syncwarp();
                // dQ_blk = dS K                    (K dimension = the 128 keys of this CTA)
                if (k > 0) {
                    mbar_wait(dq_empty, (k - 1) & 1);
                    tc_fence_after();
                }
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk)
                        umma_ss2(tm_dq, ds_k_lo + (((kk / 4) * kHalfBytes + (kk % 4) * 32) >> 4), hi_pds,
                                 k_mn_lo + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dq, kk > 0 ? 1u : 0u);
                    umma_commit(dq_full);
                }
                __syncwarp();
            };

            mbar_wait(kv_full, 0);
            if (C::kLookahead) {
                mbar_wait(qdo_full + 0, 0);
                tc_fence_after();
                issue_s_dp(0);
                for (int k = 0; k < n_iter; ++k) {
                    if (k + 1 < n_iter) {
                        const int kn = k + 1;
                        mbar_wait(qdo_full + (kn % C::kQStages), (kn / C::kQStages) & 1);
                        mbar_wait(sdp_empty, k & 1);
                        tc_fence_after();
                        issue_s_dp(kn);
                    }
                    mbar_wait(pds_full, k & 1);
                    tc_fence_after();
                    issue_grads(k);
                }
            } else {
                for (int k = 0; k < n_iter; ++k) {
                    mbar_wait(qdo_full + 0, k & 1);
                    if (k > 0) {
                        mbar_wait(sdp_empty, (k - 1) & 1);
                        if (C::kDqAliasS) mbar_wait(dq_empty, (k - 1) & 1);
                    }
                    tc_fence_after();
                    issue_s_dp(k);
                    mbar_wait(pds_full, k & 1);
                    tc_fence_after();
                    issue_grads(k);
                }
            }
            if (leader) umma_commit(acc_full);
            __syncwarp();
        }
    } else {
        // =============================== compute warpgroups ===============================
        setmaxnreg_inc<208>();
        const int wg = warp >= 8 ? 1 : 0;                    // which 64-column half of the tile
        const int r = (warp & 3) * 32 + lane;                // row in the block == TMEM lane
        const int ctid = wg * 128 + r;                       // 0..255 among compute threads
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tm_s = tmem_base + lane_off + C::kColS + wg * 64;
        const uint32_t tm_dp = tmem_base + lane_off + C::kColDP + wg * 64;
        const uint32_t tm_dq = tmem_base + lane_off + C::kColDQ + wg * C::kDqColsPerWg;
        uint8_t* sP = smem + C::kP + wg * kHalfBytes + r * 128;
        uint8_t* sDS = smem + C::kDS + wg * kHalfBytes + r * 128;
        const uint8_t* sB = smem + C::kBias + wg * kHalfBytes + r * 128;
        const int hb = p.bias_h_bcast ? 0 : h;
        const int bb = p.bias_b_bcast ? 0 : b;
        const float scale_log2 = p.sm_scale * kLog2e;

        // bias mode 3 (T5 relative-position bias computed in the kernel, see attn_fwd.cu): the per-head band of
        // bias values replaces the dense bias halves in shared memory
        const float* band = reinterpret_cast<const float*>(smem + C::kBias);
        if (kBiasMode == 3) {
            float* dst = reinterpret_cast<float*>(smem + C::kBias);
            const float* src = p.rpe.band + (int64_t)h * p.rpe.band_len;
            for (int i = ctid; i < p.rpe.band_len; i += 256) dst[i] = __ldg(src + i);
            named_bar_sync(4, 256);
        }

        for (int k = 0; k < n_iter; ++k) {
            const int mrow0 = (i_start + k) * kBM;
            const int grow = mrow0 + r;
            const bool row_ok = grow < p.M;
            float L_log2 = INFINITY, dlt = 0.f;              // out-of-range rows: P = 0
            if (row_ok) {
                const int64_t ri = ((int64_t)b * p.H + h) * p.M + grow;
                const float Lv = __ldg(p.lse + ri);
                dlt = __ldg(p.delta + ri);
                L_log2 = (Lv < -1e37f) ? INFINITY : Lv * kLog2e;   // no visible key, or every key carries the finfo.min mask: P = 0
            }
            int lim = p.N - col0 - wg * 64;                  // first masked column, relative to my half
            if (kCausal) {
                const int cl = grow + pseq + 1 - col0 - wg * 64;
                lim = cl < lim ? cl : lim;
            }
            const bool need_mask = (col0 + kBN > p.N) || (kCausal && (col0 + kBN - 1 > mrow0 + pseq));
            // bias mode 3: relative positions n - m of this tile are [col0 - mrow0 - 127, col0 - mrow0 + 127]
            bool rpe_const = false;
            float rpe_cval = 0.f;
            if (kBiasMode == 3) {
                const int rel_min = col0 - mrow0 - (kBM - 1);
                const int rel_max = col0 - mrow0 + (kBN - 1);
                rpe_const = rel_max <= p.rpe.const_lo || rel_min >= p.rpe.const_hi;
                if (rpe_const)
                    rpe_cval = band[(rel_max <= p.rpe.const_lo ? p.rpe.const_lo : p.rpe.const_hi) - p.rpe.band_lo];
            }

            mbar_wait(sdp_full, k & 1);
            tc_fence_after();
            if (kBiasMode == 1) mbar_wait(b_full + wg, k & 1);

#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t sr[32], dr[32];
                tmem_ld32(tm_s + ch * 32, sr);
                tmem_ld32(tm_dp + ch * 32, dr);
                tmem_ld_wait();
                if (ch == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sdp_empty);
                }
                uint32_t pp[16], dd[16];
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float bv[8];
                    if (kBiasMode == 1) {
                        const uint4 u = *reinterpret_cast<const uint4*>(sB + (((ch * 4 + c8) ^ (r & 7)) << 4));
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = unpack2<kBf16>(w[e]);
                            bv[2 * e] = f.x;
                            bv[2 * e + 1] = f.y;
                        }
                    } else if (kBiasMode == 2) {
                        const uint16_t* bp = reinterpret_cast<const uint16_t*>(p.bias) + (int64_t)bb * p.bias_sb +
                                             (int64_t)hb * p.bias_sh + (int64_t)grow * p.bias_sm;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = col0 + wg * 64 + ch * 32 + c8 * 8 + e;
                            bv[e] = (row_ok && c < p.N) ? to_float16bit<kBf16>(__ldg(bp + (int64_t)c * p.bias_sn)) : 0.f;
                        }
                    } else if (kBiasMode == 3) {
                        if (rpe_const) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) bv[e] = rpe_cval;
                        } else {
                            const float* bp = band + (col0 + wg * 64 + ch * 32 + c8 * 8 - grow - p.rpe.band_lo);
#pragma unroll
                            for (int e = 0; e < 8; ++e) bv[e] = bp[e];
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) bv[e] = 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        float pv[2], dv[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int ci = c8 * 8 + e + q;                      // column inside this 32-chunk
                            const float xs = fmaf(__uint_as_float(sr[ci]), scale_log2, bv[e + q] * kLog2e);
                            float pe = ex2_approx(xs - L_log2);
                            if (need_mask && (ch * 32 + ci >= lim)) pe = 0.f;
                            pv[q] = pe;
                            dv[q] = pe * (__uint_as_float(dr[ci]) - dlt);
                        }
                        pp[c8 * 4 + e / 2] = pack2<kBf16>(pv[0], pv[1]);
                        dd[c8 * 4 + e / 2] = pack2<kBf16>(dv[0], dv[1]);
                    }
                }
                if (ch == 0 && k > 0) {
                    // previous block's TMA reads of the P / dS / staging buffers must be finished
                    named_bar_sync(2, 256);
                }
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const int off = ((ch * 4 + c8) ^ (r & 7)) << 4;
                    *reinterpret_cast<uint4*>(sP + off) = make_uint4(pp[c8 * 4], pp[c8 * 4 + 1], pp[c8 * 4 + 2], pp[c8 * 4 + 3]);
                    *reinterpret_cast<uint4*>(sDS + off) = make_uint4(dd[c8 * 4], dd[c8 * 4 + 1], dd[c8 * 4 + 2], dd[c8 * 4 + 3]);
                }
            }
            // one proxy fence covers both hazards: the generic reads of the bias tile must complete before
            // TMA overwrites it (WAR), and the generic writes of P / dS must be visible to UMMA / TMA (RAW)
            fence_proxy_async_smem();
            if (kBiasMode == 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(b_empty + wg);
            }
            named_bar_sync(1, 256);
            if (ctid == 0) {
                mbar_arrive(pds_full);
                if (kBiasMode != 0) {
                    const int g = b % p.ds_groups;
                    if (p.ds_use_reduce) {
                        tma_reduce_add_4d(&p.map_ds, smem + C::kDS, col0, mrow0, h, g);
                        tma_reduce_add_4d(&p.map_ds, smem + C::kDS + kHalfBytes, col0 + 64, mrow0, h, g);
                    } else {
                        tma_store_4d(&p.map_ds, smem + C::kDS, col0, mrow0, h, g);
                        tma_store_4d(&p.map_ds, smem + C::kDS + kHalfBytes, col0 + 64, mrow0, h, g);
                    }
                    bulk_commit_group();
                    if (C::kDqAliasesDs) bulk_wait_group_read<0>();
                }
            }

            // ---- dQ block: TMEM -> swizzled fp32 staging (aliases P) -> TMA reduce-add ----
            mbar_wait(dq_full, k & 1);           // dV, dK, dQ MMAs of this block are complete
            tc_fence_after();
            if (C::kDqAliasesDs) named_bar_sync(3, 256);      // dS store has drained the dS tile
            if (wg == 0 || kD >= 32) {
                constexpr int kCols = C::kDqColsPerWg;
                constexpr int kChunk = kCols >= 32 ? 32 : kCols;
#pragma unroll
                for (int c0 = 0; c0 < kCols; c0 += kChunk) {
                    uint32_t q[kChunk];
                    tmem_ld_n<kChunk>(tm_dq + c0, q);
                    tmem_ld_wait();
                    const int gcol = wg * C::kDqColsPerWg + c0;            // first dQ column of this chunk
                    uint8_t* box = smem + C::kP + (gcol / C::kDqBoxCols) * C::kDqBoxBytes + r * (C::kDqBoxCols * 2);
#pragma unroll
                    for (int i = 0; i < kChunk; i += 8) {
                        const int c16 = ((gcol % C::kDqBoxCols) + i) / 8;  // 16-byte chunk inside the box row
                        const int off = (C::kDqBoxCols == 64) ? ((c16 ^ (r & 7)) << 4) : (c16 << 4);
                        uint4 v;
                        v.x = pack2<kBf16>(__uint_as_float(q[i + 0]), __uint_as_float(q[i + 1]));
                        v.y = pack2<kBf16>(__uint_as_float(q[i + 2]), __uint_as_float(q[i + 3]));
                        v.z = pack2<kBf16>(__uint_as_float(q[i + 4]), __uint_as_float(q[i + 5]));
                        v.w = pack2<kBf16>(__uint_as_float(q[i + 6]), __uint_as_float(q[i + 7]));
                        *reinterpret_cast<uint4*>(box + off) = v;
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (ctid == 0) {
                mbar_arrive(dq_empty);
#pragma unroll
                for (int bx = 0; bx < C::kDqBoxes; ++bx)
                    tma_reduce_add_4d(&p.map_dq, smem + C::kP + bx * C::kDqBoxBytes, bx * C::kDqBoxCols, mrow0, h,
                                      (nb % p.dq_groups) * p.B + b);
                bulk_commit_group();
                bulk_wait_group_read<0>();
            }
            // (the matching named_bar_sync(2) sits in front of the next block's first smem write)
        }

        // ---- epilogue: dV (warpgroup 0) and dK * sm_scale (warpgroup 1) ----
        {
            const int gn = col0 + r;
            const bool row_ok = gn < p.N;
            uint8_t* out_row = wg == 0
                ? reinterpret_cast<uint8_t*>(p.dv) + 2 * ((int64_t)b * p.dv_sb + (int64_t)h * p.dv_sh + (int64_t)gn * p.dv_sn)
                : reinterpret_cast<uint8_t*>(p.dk) + 2 * ((int64_t)b * p.dk_sb + (int64_t)h * p.dk_sh + (int64_t)gn * p.dk_sn);
            const float sc = wg == 0 ? 1.f : p.sm_scale;
            if (n_iter > 0) {
                mbar_wait(acc_full, 0);
                tc_fence_after();
                const uint32_t tm_acc = tmem_base + lane_off + (wg == 0 ? C::kColDV : C::kColDK);
                constexpr int kChunk = kD >= 32 ? 32 : 16;
#pragma unroll
                for (int c0 = 0; c0 < kD; c0 += kChunk) {
                    uint32_t a[kChunk];
                    tmem_ld_n<kChunk>(tm_acc + c0, a);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < kChunk; i += 8) {
                            uint4 out;
                            out.x = pack2<kBf16>(__uint_as_float(a[i + 0]) * sc, __uint_as_float(a[i + 1]) * sc);
                            out.y = pack2<kBf16>(__uint_as_float(a[i + 2]) * sc, __uint_as_float(a[i + 3]) * sc);
                            out.z = pack2<kBf16>(__uint_as_float(a[i + 4]) * sc, __uint_as_float(a[i + 5]) * sc);
                            out.w = pack2<kBf16>(__uint_as_float(a[i + 6]) * sc, __uint_as_float(a[i + 7]) * sc);
                            *reinterpret_cast<uint4*>(out_row + 2 * (c0 + i)) = out;
                        }
                    }
                }
                tc_fence_before();
            } else if (row_ok) {
#pragma unroll
                for (int c = 0; c < kD; c += 8) *reinterpret_cast<uint4*>(out_row + 2 * c) = make_uint4(0, 0, 0, 0);
            }
        }
        if (ctid == 0) bulk_wait_group<0>();     // all TMA stores / reductions of this CTA have landed
    }

    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// helper kernels
// ------------------------------------------------------------------------------------------
// delta = rowsum(O * dO) in fp32 (reference: _bwd_preprocess :516-556) + zero the 16-bit dQ group surface.
template <int kD, bool kBf16>
__device__ __forceinline__ void attn_bwd_preprocess_body(const uint8_t* __restrict__ o, int64_t o_sb, int64_t o_sh, int64_t o_sm,
                                                         const uint8_t* __restrict__ dout, int64_t do_sb, int64_t do_sh,
                                                         int64_t do_sm, float* __restrict__ delta, uint4* __restrict__ dq_ws,
                                                         int dq_groups, int B, int H, int M, uint4* __restrict__ zero_ptr,
                                                         int64_t zero_chunks, int vblock, int vgrid,
                                                         const float* __restrict__ lse = nullptr, float* __restrict__ nl_out = nullptr,
                                                         int m_pad = 0) {
    // nl_out != NULL (v3 kernel): the row statistics are written in the form the kernel consumes, in rows padded to m_pad
    // (a multiple of 128) entries:  nl = -L * log2e (-inf for a row without visible key, for a row whose every key carries
    // the finfo.min mask, and for the padding), delta -> -delta (0 in the padding).
    constexpr int kTpr = kD / 8;                              // threads per row, 8 elements (16 B) each
    const int64_t gid = (int64_t)vblock * blockDim.x + threadIdx.x;
    const int64_t row = gid / kTpr;
    const int part = static_cast<int>(gid % kTpr);
    const int64_t rows = (int64_t)B * H * M;
    float acc = 0.f;
    if (row < rows) {
        const int m = static_cast<int>(row % M);
        const int64_t bh = row / M;
        const int hh = static_cast<int>(bh % H);
        const int64_t bb = bh / H;
        const uint4 ov = *reinterpret_cast<const uint4*>(o + 2 * (bb * o_sb + hh * o_sh + m * o_sm + part * 8));
        const uint4 dv = *reinterpret_cast<const uint4*>(dout + 2 * (bb * do_sb + hh * do_sh + m * do_sm + part * 8));
        const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w};
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = unpack2<kBf16>(ow[e]);
            const float2 c = unpack2<kBf16>(dw[e]);
            acc = fmaf(a.x, c.x, acc);
            acc = fmaf(a.y, c.y, acc);
        }
        for (int g = 0; g < dq_groups; ++g) dq_ws[(g * rows + row) * kTpr + part] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int off = kTpr / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (row < rows && part == 0) {
        if (nl_out == nullptr) {
            delta[row] = acc;
        } else {
            // one 512-byte record per 64 padded rows: [64 x nl | 64 x -delta] (the kernel fetches a record with one bulk copy)
            const int mm = static_cast<int>(row % M);
            const int64_t prow = (row / M) * (2 * (int64_t)m_pad) + (mm >> 6) * 128 + (mm & 63);
            const float Lv = __ldg(lse + row);
            nl_out[prow] = Lv < -1e37f ? -INFINITY : -Lv * 1.4426950408889634f;
            nl_out[prow + 64] = -acc;
        }
    }
    if (nl_out != nullptr && m_pad > M) {
        const int64_t pad = m_pad - M, total = (int64_t)B * H * pad;
        for (int64_t i = gid; i < total; i += (int64_t)vgrid * blockDim.x) {
            const int mm = M + static_cast<int>(i % pad);
            const int64_t prow = (i / pad) * (2 * (int64_t)m_pad) + (mm >> 6) * 128 + (mm & 63);
            nl_out[prow] = -INFINITY;
            nl_out[prow + 64] = 0.f;
        }
    }
    // zero-fill of the dS batch-group surface (replaces a separate memset node)
    for (int64_t i = gid; i < zero_chunks; i += (int64_t)vgrid * blockDim.x) zero_ptr[i] = make_uint4(0, 0, 0, 0);
}

template <int kD, bool kBf16>
__global__ void attn_bwd_preprocess_kernel(const uint8_t* __restrict__ o, int64_t o_sb, int64_t o_sh, int64_t o_sm,
                                           const uint8_t* __restrict__ dout, int64_t do_sb, int64_t do_sh,
                                           int64_t do_sm, float* __restrict__ delta, uint4* __restrict__ dq_ws,
                                           int dq_groups, int B, int H, int M, uint4* __restrict__ zero_ptr,
                                           int64_t zero_chunks) {
    attn_bwd_preprocess_body<kD, kBf16>(o, o_sb, o_sh, o_sm, dout, do_sb, do_sh, do_sm, delta, dq_ws, dq_groups, B, H, M, zero_ptr,
                                        zero_chunks, blockIdx.x, gridDim.x);
}

template <int kD, bool kBf16>
__device__ __forceinline__ void attn_bwd_dq_convert_body(const uint4* __restrict__ dq_ws, int dq_groups, uint8_t* __restrict__ dq,
                                                         int64_t sb, int64_t sh, int64_t sm, int B, int H, int M, float scale,
                                                         int vblock) {
    constexpr int kTpr = kD / 8;
    const int64_t gid = (int64_t)vblock * blockDim.x + threadIdx.x;
    const int64_t row = gid / kTpr;
    const int part = static_cast<int>(gid % kTpr);
    const int64_t rows = (int64_t)B * H * M;
    if (row >= rows) return;
    const int m = static_cast<int>(row % M);
    const int64_t bh = row / M;
    const int hh = static_cast<int>(bh % H);
    const int64_t bb = bh / H;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int g = 0; g < dq_groups; ++g) {
        const uint4 u = __ldg(dq_ws + (g * rows + row) * kTpr + part);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<kBf16>(w[e]);
            acc[2 * e] += f.x;
            acc[2 * e + 1] += f.y;
        }
    }
    uint4 out;
    out.x = pack2<kBf16>(acc[0] * scale, acc[1] * scale);
    out.y = pack2<kBf16>(acc[2] * scale, acc[3] * scale);
    out.z = pack2<kBf16>(acc[4] * scale, acc[5] * scale);
    out.w = pack2<kBf16>(acc[6] * scale, acc[7] * scale);
    *reinterpret_cast<uint4*>(dq + 2 * (bb * sb + hh * sh + m * sm + part * 8)) = out;
}

template <int kD, bool kBf16>
__global__ void attn_bwd_dq_convert_kernel(const uint4* __restrict__ dq_ws, int dq_groups, uint8_t* __restrict__ dq,
                                           int64_t sb, int64_t sh, int64_t sm, int B, int H, int M, float scale) {
    attn_bwd_dq_convert_body<kD, kBf16>(dq_ws, dq_groups, dq, sb, sh, sm, B, H, M, scale, blockIdx.x);
}

// dBias = sum of the per-(batch, head) dS tiles over every broadcast dimension, fp32 accumulation,
// one rounding (reference: ds.sum(0) :214-215; the head sum is the fix described in SURVEY.md section 4).
template <bool kBf16>
__global__ void dbias_reduce_kernel(const uint16_t* __restrict__ ws, int pitch, uint16_t* __restrict__ out,
                                    int64_t o_sb, int64_t o_sh, int64_t o_sm, int64_t o_sn, int B, int H, int M, int N,
                                    int reduce_b, int reduce_h, int causal) {
    const int64_t mn = (int64_t)M * N;
    const int64_t mp = (int64_t)M * pitch;
    const int ob_n = reduce_b ? 1 : B;
    const int oh_n = reduce_h ? 1 : H;
    const int64_t total = (int64_t)ob_n * oh_n * mn;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int n = static_cast<int>(idx % N);
        int64_t t = idx / N;
        const int m = static_cast<int>(t % M);
        t /= M;
        const int oh = static_cast<int>(t % oh_n);
        const int ob = static_cast<int>(t / oh_n);
        float acc = 0.f;
        if (!(causal && n > m + (N - M))) {
            const int b0 = reduce_b ? 0 : ob, b1 = reduce_b ? B : ob + 1;
            const int h0 = reduce_h ? 0 : oh, h1 = reduce_h ? H : oh + 1;
            for (int bb = b0; bb < b1; ++bb)
                for (int hh = h0; hh < h1; ++hh)
                    acc += to_float16bit<kBf16>(ws[((int64_t)bb * H + hh) * mp + (int64_t)m * pitch + n]);
        }
        const uint32_t packed = pack2<kBf16>(acc, 0.f);
        out[ob * o_sb + oh * o_sh + m * o_sm + n * o_sn] = static_cast<uint16_t>(packed & 0xFFFFu);
    }
}

// 8 columns per thread (16-byte loads), N % 8 == 0 and contiguous output rows
template <bool kBf16>
__global__ void dbias_reduce_vec8_kernel(const uint4* __restrict__ ws, int pitch8, uint4* __restrict__ out, int B,
                                         int H, int M, int N, int reduce_b, int reduce_h, int causal) {
    const int n8 = N / 8;
    const int64_t mn8 = (int64_t)M * n8;
    const int64_t mp8 = (int64_t)M * pitch8;
    const int ob_n = reduce_b ? 1 : B;
    const int oh_n = reduce_h ? 1 : H;
    const int64_t total = (int64_t)ob_n * oh_n * mn8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = static_cast<int>(idx % n8);
        int64_t t = idx / n8;
        const int m = static_cast<int>(t % M);
        t /= M;
        const int oh = static_cast<int>(t % oh_n);
        const int ob = static_cast<int>(t / oh_n);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        const int vis = causal ? (m + (N - M) + 1 - c8 * 8) : 8;      // visible columns of this group
        if (vis > 0) {
            const int b0 = reduce_b ? 0 : ob, b1 = reduce_b ? B : ob + 1;
            const int h0 = reduce_h ? 0 : oh, h1 = reduce_h ? H : oh + 1;
            for (int bb = b0; bb < b1; ++bb)
                for (int hh = h0; hh < h1; ++hh) {
                    const uint4 u = __ldg(ws + ((int64_t)bb * H + hh) * mp8 + (int64_t)m * pitch8 + c8);
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = unpack2<kBf16>(w[e]);
                        acc[2 * e] += f.x;
                        acc[2 * e + 1] += f.y;
                    }
                }
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e >= vis) acc[e] = 0.f;
        }
        uint4 o;
        o.x = pack2<kBf16>(acc[0], acc[1]);
        o.y = pack2<kBf16>(acc[2], acc[3]);
        o.z = pack2<kBf16>(acc[4], acc[5]);
        o.w = pack2<kBf16>(acc[6], acc[7]);
        out[idx] = o;
    }
}

// dq convert + dBias reduce in ONE launch (they are independent; one grid, block-index split): blocks
// [0, cvt_blocks) run the dQ conversion, the rest run the vectorised dBias reduction.
template <int kD, bool kBf16>
__global__ void attn_bwd_finalize_kernel(const uint4* __restrict__ dq_ws, int dq_groups, uint8_t* __restrict__ dq,
                                         int64_t sb, int64_t sh, int64_t sm, int B, int H, int M, float scale,
                                         int cvt_blocks, const uint4* __restrict__ ds_ws, int pitch8,
                                         uint4* __restrict__ dbias, int G, int N, int reduce_b, int reduce_h,
                                         int causal) {
    if (static_cast<int>(blockIdx.x) < cvt_blocks) {
        constexpr int kTpr = kD / 8;
        const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t row = gid / kTpr;
        const int part = static_cast<int>(gid % kTpr);
        const int64_t rows = (int64_t)B * H * M;
        if (row >= rows) return;
        const int m = static_cast<int>(row % M);
        const int64_t bh = row / M;
        const int hh = static_cast<int>(bh % H);
        const int64_t bb = bh / H;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int g = 0; g < dq_groups; ++g) {
            const uint4 u = __ldg(dq_ws + (g * rows + row) * kTpr + part);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = unpack2<kBf16>(w[e]);
                acc[2 * e] += f.x;
                acc[2 * e + 1] += f.y;
            }
        }
        uint4 out;
        out.x = pack2<kBf16>(acc[0] * scale, acc[1] * scale);
        out.y = pack2<kBf16>(acc[2] * scale, acc[3] * scale);
        out.z = pack2<kBf16>(acc[4] * scale, acc[5] * scale);
        out.w = pack2<kBf16>(acc[6] * scale, acc[7] * scale);
        *reinterpret_cast<uint4*>(dq + 2 * (bb * sb + hh * sh + m * sm + part * 8)) = out;
        return;
    }
    // ---- dBias: sum of the G group slices (and of a broadcast head dim), fp32, one rounding ----
    const int n8 = N / 8;
    const int64_t mn8 = (int64_t)M * n8;
    const int64_t mp8 = (int64_t)M * pitch8;
    const int ob_n = reduce_b ? 1 : G;
    const int oh_n = reduce_h ? 1 : H;
    const int64_t total = (int64_t)ob_n * oh_n * mn8;
    const int64_t nthreads = (int64_t)(gridDim.x - cvt_blocks) * blockDim.x;
    for (int64_t idx = (int64_t)(blockIdx.x - cvt_blocks) * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
        const int c8 = static_cast<int>(idx % n8);
        int64_t t = idx / n8;
        const int m = static_cast<int>(t % M);
        t /= M;
        const int oh = static_cast<int>(t % oh_n);
        const int ob = static_cast<int>(t / oh_n);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        const int vis = causal ? (m + (N - M) + 1 - c8 * 8) : 8;
        if (vis > 0) {
            const int b0 = reduce_b ? 0 : ob, b1 = reduce_b ? G : ob + 1;
            const int h0 = reduce_h ? 0 : oh, h1 = reduce_h ? H : oh + 1;
            for (int bb = b0; bb < b1; ++bb)
                for (int hh = h0; hh < h1; ++hh) {
                    const uint4 u = __ldg(ds_ws + ((int64_t)bb * H + hh) * mp8 + (int64_t)m * pitch8 + c8);
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = unpack2<kBf16>(w[e]);
                        acc[2 * e] += f.x;
                        acc[2 * e + 1] += f.y;
                    }
                }
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e >= vis) acc[e] = 0.f;
        }
        uint4 o;
        o.x = pack2<kBf16>(acc[0], acc[1]);
        o.y = pack2<kBf16>(acc[2], acc[3]);
        o.z = pack2<kBf16>(acc[4], acc[5]);
        o.w = pack2<kBf16>(acc[6], acc[7]);
        dbias[idx] = o;
    }
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kBiasMode, bool kCausal>
static cudaError_t launch_bwd_inst(const AttnBwdKernelParams& kp, cudaStream_t stream) {
    using C = BwdCfg<kD>;
    auto kern = attn_bwd_kernel<kD, kBf16, kBiasMode, kCausal>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
    if (e != cudaSuccess) return e;
    const int grid = kp.B * kp.H * kp.num_n_blocks;
    kern<<<grid, 384, C::kTotal, stream>>>(kp);
    count_launch();
    return cudaGetLastError();
}

template <int kD, bool kBf16>
static cudaError_t launch_bwd_d(const AttnBwdKernelParams& kp, int bias_mode, bool causal, cudaStream_t stream) {
    switch (bias_mode * 2 + (causal ? 1 : 0)) {
        case 0: return launch_bwd_inst<kD, kBf16, 0, false>(kp, stream);
        case 1: return launch_bwd_inst<kD, kBf16, 0, true>(kp, stream);
        case 2: return launch_bwd_inst<kD, kBf16, 1, false>(kp, stream);
        case 3: return launch_bwd_inst<kD, kBf16, 1, true>(kp, stream);
        case 4: return launch_bwd_inst<kD, kBf16, 2, false>(kp, stream);
        case 5: return launch_bwd_inst<kD, kBf16, 2, true>(kp, stream);
        case 6: return launch_bwd_inst<kD, kBf16, 3, false>(kp, stream);
        case 7: return launch_bwd_inst<kD, kBf16, 3, true>(kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

// This kernel is the D = 128 path only (its TMEM budget is what forces the simpler schedule); head dims 16 / 32 / 64 take
// the transposed-formulation kernel of attn_bwd_v3.cu.
cudaError_t launch_attn_bwd(const AttnBwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                            cudaStream_t stream) {
    if (D != 128) return cudaErrorInvalidValue;
    return bf16 ? launch_bwd_d<128, true>(kp, bias_mode, causal, stream) : launch_bwd_d<128, false>(kp, bias_mode, causal, stream);
}

// ------------------------------------------------------------------------------------------
// helpers of the transposed-formulation kernel (attn_bwd_v3.cu)
// ------------------------------------------------------------------------------------------
// Repacked dense bias for the v3 kernel (see kernels.h): one block per (bh, key block, query block of 32): reads 32 rows of
// 128 bias values (coalesced along n), writes 4 x 128 16-byte words (thread = key row order the attention kernel reads).
__device__ __forceinline__ void bias_repack_body(const uint16_t* __restrict__ in, int64_t s_b, int64_t s_h, int64_t s_m,
                                                 int64_t s_n, uint4* __restrict__ out, int Hb, int M, int N, int qb, int kb, int bh,
                                                 int n_qb, int n_kb) {
    __shared__ uint16_t tile[32][130];
    const int bb = bh / Hb, hb = bh % Hb;
    const int m0 = qb * 32, n0 = kb * 128;
    const uint16_t* src = in + bb * s_b + hb * s_h;
    const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int ml = ty + 2 * i, m = m0 + ml, n = n0 + tx;
        tile[ml][tx] = (m < M && n < N) ? __ldg(src + (int64_t)m * s_m + (int64_t)n * s_n) : (uint16_t)0;
    }
    __syncthreads();
    uint4* dst = out + (((int64_t)bh * n_kb + kb) * n_qb + qb) * (4 * 128);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = threadIdx.x + 256 * i;          // c * 128 + n
        const int c = idx >> 7, n = idx & 127;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            w[e] = (uint32_t)tile[c * 8 + 2 * e][n] | ((uint32_t)tile[c * 8 + 2 * e + 1][n] << 16);
        dst[idx] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void __launch_bounds__(256) bias_repack_kernel(const uint16_t* __restrict__ in, int64_t s_b, int64_t s_h, int64_t s_m,
                                                          int64_t s_n, uint4* __restrict__ out, int Hb, int M, int N) {
    bias_repack_body(in, s_b, s_h, s_m, s_n, out, Hb, M, N, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, gridDim.y);
}

// delta / zero-fill and the bias repack in ONE launch (independent work; block-index split)
template <int kD, bool kBf16>
__global__ void __launch_bounds__(256) attn_bwd_pre_fused_kernel(const uint8_t* __restrict__ o, int64_t o_sb, int64_t o_sh, int64_t o_sm,
                                                                 const uint8_t* __restrict__ dout, int64_t do_sb, int64_t do_sh, int64_t do_sm,
                                                                 float* __restrict__ delta, uint4* __restrict__ dq_ws, int dq_groups, int B,
                                                                 int H, int M, uint4* __restrict__ zero_ptr, int64_t zero_chunks,
                                                                 int pre_blocks, const uint16_t* __restrict__ bias, int64_t s_b, int64_t s_h,
                                                                 int64_t s_m, int64_t s_n, uint4* __restrict__ bias_p, int Hb, int N,
                                                                 int n_qb, int n_kb, const float* __restrict__ lse,
                                                                 float* __restrict__ nl_out, int m_pad) {
    if (static_cast<int>(blockIdx.x) < pre_blocks) {
        attn_bwd_preprocess_body<kD, kBf16>(o, o_sb, o_sh, o_sm, dout, do_sb, do_sh, do_sm, delta, dq_ws, dq_groups, B, H, M, zero_ptr,
                                            zero_chunks, blockIdx.x, pre_blocks, lse, nl_out, m_pad);
        return;
    }
    const int id = static_cast<int>(blockIdx.x) - pre_blocks;
    bias_repack_body(bias, s_b, s_h, s_m, s_n, bias_p, Hb, M, N, id % n_qb, (id / n_qb) % n_kb, id / (n_qb * n_kb), n_qb, n_kb);
}

cudaError_t launch_bias_repack(const void* bias, const int64_t* s, void* bias_p, int Bb, int Hb, int M, int N, cudaStream_t stream) {
    const dim3 grid(4 * ((M + 127) / 128), (N + 127) / 128, Bb * Hb);   // whole 128-query tiles (zero beyond M)
    if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
    bias_repack_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(bias), s[0], s[1], s[2], s[3], static_cast<uint4*>(bias_p),
                                                 Hb, M, N);
    count_launch();
    return cudaGetLastError();
}

// dbias[ob, oh, m, n] = sum over the reduced group / head slices of ds_t[g, h, n, m]; fp32 accumulation, one rounding.
// 64 x 64 tiles: read rows of the transposed surface (m contiguous), transpose through shared memory, write rows of dbias.
template <bool kBf16, bool kOutF32>
__device__ __forceinline__ void dbias_reduce_t_body(const uint16_t* __restrict__ ws, int m_pitch, void* __restrict__ out_v,
                                                    int64_t o_sb, int64_t o_sh, int64_t o_sm, int64_t o_sn, int G, int H,
                                                    int M, int N, int reduce_b, int reduce_h, int causal, int bx, int by, int bz,
                                                    bool accumulate = false) {
    // thread (ty, tx) = (tid / 8, tid % 8): loads 16 bytes = 8 consecutive m of rows n0 + ty and n0 + ty + 32 (a warp reads
    // four 128-byte row segments per instruction), writes 8 consecutive n of rows m0 + ty and m0 + ty + 32 the same way.
    __shared__ float tile[64][65];
    const int m0 = bx * 64, n0 = by * 64;
    const int oh_n = reduce_h ? 1 : H;
    const int ob = bz / oh_n, oh = bz % oh_n;
    const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
    const int pseq = N - M;
    const bool tile_masked = causal && (n0 > m0 + 63 + pseq);              // every (m, n) of the tile is above the diagonal
    // accumulate (fp32 output only): out += sum -- a tile that is entirely masked adds nothing
    if (kOutF32 && accumulate && tile_masked) return;
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
    if (!tile_masked) {
        const int g0 = reduce_b ? 0 : ob, g1 = reduce_b ? G : ob + 1;
        const int h0 = reduce_h ? 0 : oh, h1 = reduce_h ? H : oh + 1;
        const int m = m0 + 8 * tx;                                         // (the row pitch is a multiple of 8: m < M => in the row)
        for (int g = g0; g < g1; ++g)
            for (int hh = h0; hh < h1; ++hh) {
                const uint16_t* src = ws + ((int64_t)g * H + hh) * (int64_t)N * m_pitch;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int n = n0 + ty + 32 * j;
                    if (n < N && m < M) {
                        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)n * m_pitch + m));
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // (unpack + FADD: the mixed-precision add.rn.f32.bf16 runs at a quarter of the FADD rate -- measured
                            //  +13 us on the dQ conversion when it was tried here)
                            const float2 f = unpack2<kBf16>(w[e]);
                            acc[j][2 * e] += f.x;
                            acc[j][2 * e + 1] += f.y;
                        }
                    }
                }
            }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) tile[ty + 32 * j][8 * tx + e] = acc[j][e];          // tile[n][m]
    __syncthreads();
    const int64_t obase = ob * o_sb + oh * o_sh;
    const bool vec_ok = o_sn == 1 && (o_sm % 8) == 0 && (obase % 8) == 0 &&
                        (reinterpret_cast<uintptr_t>(out_v) % (kOutF32 ? 32 : 16)) == 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int ml = ty + 32 * j, m = m0 + ml;
        if (m >= M) continue;
        const int nb = n0 + 8 * tx;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            v[e] = tile[8 * tx + e][ml];
            if (causal && nb + e > m + pseq) v[e] = 0.f;                    // select, not multiply: unwritten tiles may hold anything
        }
        const int64_t oi = obase + (int64_t)m * o_sm + (int64_t)nb * o_sn;
        if (vec_ok && nb + 8 <= N) {
            if constexpr (kOutF32) {
                float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out_v) + oi);
                if (accumulate) {
                    const float4 a0 = dst[0], a1 = dst[1];
                    v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
                    v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
                }
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                *reinterpret_cast<uint4*>(static_cast<uint16_t*>(out_v) + oi) =
                    make_uint4(pack2<kBf16>(v[0], v[1]), pack2<kBf16>(v[2], v[3]), pack2<kBf16>(v[4], v[5]), pack2<kBf16>(v[6], v[7]));
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (nb + e >= N) continue;
                if constexpr (kOutF32) static_cast<float*>(out_v)[oi + e * o_sn] = accumulate ? static_cast<float*>(out_v)[oi + e * o_sn] + v[e] : v[e];
                else static_cast<uint16_t*>(out_v)[oi + e * o_sn] = static_cast<uint16_t>(pack2<kBf16>(v[e], 0.f) & 0xFFFFu);
            }
        }
    }
}

template <bool kBf16, bool kOutF32>
__global__ void __launch_bounds__(256) dbias_reduce_t_kernel(const uint16_t* __restrict__ ws, int m_pitch, void* __restrict__ out_v,
                                                             int64_t o_sb, int64_t o_sh, int64_t o_sm, int64_t o_sn, int G, int H,
                                                             int M, int N, int reduce_b, int reduce_h, int causal, int accumulate) {
    dbias_reduce_t_body<kBf16, kOutF32>(ws, m_pitch, out_v, o_sb, o_sh, o_sm, o_sn, G, H, M, N, reduce_b, reduce_h, causal, blockIdx.x,
                                        blockIdx.y, blockIdx.z, accumulate != 0);
}

// dQ conversion and the transposing dBias reduction in ONE launch (independent work; block-index split)
template <int kD, bool kBf16, bool kOutF32>
__global__ void __launch_bounds__(256, 6) attn_bwd_post_fused_kernel(const uint4* __restrict__ dq_ws, int dq_groups, uint8_t* __restrict__ dq,
                                                                  int64_t sb, int64_t sh, int64_t sm, int B, int H, int M, float scale,
                                                                  int cvt_blocks, const uint16_t* __restrict__ ws, int m_pitch,
                                                                  void* __restrict__ dbias, int64_t o_sb, int64_t o_sh, int64_t o_sm,
                                                                  int64_t o_sn, int G, int N, int reduce_b, int reduce_h, int causal, int gx,
                                                                  int gy, int accumulate) {
    // the (heavier) reduction blocks come first: the light conversion blocks fill the tail of the grid
    const int red_blocks = static_cast<int>(gridDim.x) - cvt_blocks;
    if (static_cast<int>(blockIdx.x) >= red_blocks) {
        attn_bwd_dq_convert_body<kD, kBf16>(dq_ws, dq_groups, dq, sb, sh, sm, B, H, M, scale, static_cast<int>(blockIdx.x) - red_blocks);
        return;
    }
    const int id = static_cast<int>(blockIdx.x);
    dbias_reduce_t_body<kBf16, kOutF32>(ws, m_pitch, dbias, o_sb, o_sh, o_sm, o_sn, G, H, M, N, reduce_b, reduce_h, causal, id % gx,
                                        (id / gx) % gy, id / (gx * gy), accumulate != 0);
}

cudaError_t launch_dbias_reduce_t(const void* ds_t, int m_pitch, void* dbias, const int64_t* s, int G, int H, int M, int N,
                                  int reduce_b, int reduce_h, bool causal, bool bf16, bool out_f32, cudaStream_t stream, bool accumulate) {
    const int ob = reduce_b ? 1 : G, oh = reduce_h ? 1 : H;
    const dim3 grid((M + 63) / 64, (N + 63) / 64, ob * oh);
    if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
#define B200T5_DBR(BF, F32)                                                                                                   \
    dbias_reduce_t_kernel<BF, F32><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(ds_t), m_pitch, dbias, s[0], s[1], s[2], \
                                                             s[3], G, H, M, N, reduce_b, reduce_h, causal ? 1 : 0, accumulate ? 1 : 0)
    if (bf16) { if (out_f32) B200T5_DBR(true, true); else B200T5_DBR(true, false); }
    else { if (out_f32) B200T5_DBR(false, true); else B200T5_DBR(false, false); }
#undef B200T5_DBR
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_attn_bwd_preprocess(const void* o, const int64_t* os, const void* dout, const int64_t* ds,
                                       float* delta, void* dq_ws, int dq_groups, int B, int H, int M, int D,
                                       bool bf16, void* zero_ptr, size_t zero_bytes, cudaStream_t stream) {
    const int64_t threads = (int64_t)B * H * M * (D / 8);
    const int block = 256;
    const int grid = static_cast<int>((threads + block - 1) / block);
#define B200T5_PRE(DD, BF)                                                                                         \
    attn_bwd_preprocess_kernel<DD, BF><<<grid, block, 0, stream>>>(                                                \
        static_cast<const uint8_t*>(o), os[0], os[1], os[2], static_cast<const uint8_t*>(dout), ds[0], ds[1],     \
        ds[2], delta, static_cast<uint4*>(dq_ws), dq_groups, B, H, M, static_cast<uint4*>(zero_ptr),            \
        static_cast<int64_t>(zero_bytes / 16))
    switch (D) {
        case 16: if (bf16) B200T5_PRE(16, true); else B200T5_PRE(16, false); break;
        case 32: if (bf16) B200T5_PRE(32, true); else B200T5_PRE(32, false); break;
        case 64: if (bf16) B200T5_PRE(64, true); else B200T5_PRE(64, false); break;
        case 128: if (bf16) B200T5_PRE(128, true); else B200T5_PRE(128, false); break;
        default: return cudaErrorInvalidValue;
    }
#undef B200T5_PRE
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_attn_bwd_dq_convert(const void* dq_ws, int dq_groups, void* dq, const int64_t* s, int B, int H,
                                       int M, int D, float sm_scale, bool bf16, cudaStream_t stream) {
    const int64_t threads = (int64_t)B * H * M * (D / 8);
    const int block = 256;
    const int grid = static_cast<int>((threads + block - 1) / block);
#define B200T5_CVT(DD, BF)                                                                                  \
    attn_bwd_dq_convert_kernel<DD, BF><<<grid, block, 0, stream>>>(static_cast<const uint4*>(dq_ws), dq_groups, \
                                                                    static_cast<uint8_t*>(dq), s[0], s[1], s[2], B, \
                                                                    H, M, sm_scale)
    switch (D) {
        case 16: if (bf16) B200T5_CVT(16, true); else B200T5_CVT(16, false); break;
        case 32: if (bf16) B200T5_CVT(32, true); else B200T5_CVT(32, false); break;
        case 64: if (bf16) B200T5_CVT(64, true); else B200T5_CVT(64, false); break;
        case 128: if (bf16) B200T5_CVT(128, true); else B200T5_CVT(128, false); break;
        default: return cudaErrorInvalidValue;
    }
#undef B200T5_CVT
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_dbias_reduce(const void* ds_ws, int ws_pitch, void* dbias, const int64_t* s, int B, int H, int M,
                                int N, int reduce_b, int reduce_h, bool causal, bool bf16, cudaStream_t stream) {
    const int ob = reduce_b ? 1 : B, oh = reduce_h ? 1 : H;
    const bool contiguous = s[3] == 1 && s[2] == N && (oh == 1 || s[1] == (int64_t)M * N) &&
                            (ob == 1 || s[0] == (int64_t)oh * M * N);
    const bool vec = contiguous && (N % 8 == 0) && ((reinterpret_cast<uintptr_t>(dbias) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(ds_ws) & 15) == 0);
    const int block = 256;
    if (vec) {
        const int64_t total = (int64_t)ob * oh * M * (N / 8);
        const int grid = static_cast<int>(std::min<int64_t>((total + block - 1) / block, 148 * 16));
        if (bf16)
            dbias_reduce_vec8_kernel<true><<<grid, block, 0, stream>>>(static_cast<const uint4*>(ds_ws), ws_pitch / 8,
                                                                       static_cast<uint4*>(dbias), B, H, M, N,
                                                                       reduce_b, reduce_h, causal ? 1 : 0);
        else
            dbias_reduce_vec8_kernel<false><<<grid, block, 0, stream>>>(static_cast<const uint4*>(ds_ws), ws_pitch / 8,
                                                                        static_cast<uint4*>(dbias), B, H, M, N,
                                                                        reduce_b, reduce_h, causal ? 1 : 0);
    } else {
        const int64_t total = (int64_t)ob * oh * M * N;
        const int grid = static_cast<int>(std::min<int64_t>((total + block - 1) / block, 148 * 16));
        if (bf16)
            dbias_reduce_kernel<true><<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(ds_ws), ws_pitch,
                                                                  static_cast<uint16_t*>(dbias), s[0], s[1], s[2],
                                                                  s[3], B, H, M, N, reduce_b, reduce_h, causal ? 1 : 0);
        else
            dbias_reduce_kernel<false><<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(ds_ws), ws_pitch,
                                                                   static_cast<uint16_t*>(dbias), s[0], s[1], s[2],
                                                                   s[3], B, H, M, N, reduce_b, reduce_h, causal ? 1 : 0);
    }
    count_launch();
    return cudaGetLastError();
}

// dQ conversion and (when there is a bias) the dBias reduction.  One fused launch when dBias is contiguous and
// 16-byte addressable, two launches otherwise.
cudaError_t launch_attn_bwd_finalize(const void* dq_ws, int dq_groups, void* dq, const int64_t* dqs, int B, int H, int M,
                                     int N, int D, float sm_scale, bool bf16, const void* ds_ws, int ws_pitch, void* dbias,
                                     const int64_t* dbs, int G, int reduce_b, int reduce_h, bool causal,
                                     cudaStream_t stream) {
    bool fused = dbias != nullptr;
    if (fused) {
        const int ob = reduce_b ? 1 : G, oh = reduce_h ? 1 : H;
        const bool contiguous = dbs[3] == 1 && dbs[2] == N && (oh == 1 || dbs[1] == (int64_t)M * N) &&
                                (ob == 1 || dbs[0] == (int64_t)oh * M * N);
        fused = contiguous && (N % 8 == 0) && ((reinterpret_cast<uintptr_t>(dbias) & 15) == 0) &&
                ((reinterpret_cast<uintptr_t>(ds_ws) & 15) == 0);
    }
    if (!fused) {
        cudaError_t e = launch_attn_bwd_dq_convert(dq_ws, dq_groups, dq, dqs, B, H, M, D, sm_scale, bf16, stream);
        if (e != cudaSuccess || dbias == nullptr) return e;
        return launch_dbias_reduce(ds_ws, ws_pitch, dbias, dbs, G, H, M, N, reduce_b, reduce_h, causal, bf16, stream);
    }
    const int block = 256;
    const int64_t cvt_threads = (int64_t)B * H * M * (D / 8);
    const int cvt_blocks = static_cast<int>((cvt_threads + block - 1) / block);
    const int ob = reduce_b ? 1 : G, oh = reduce_h ? 1 : H;
    const int64_t red_total = (int64_t)ob * oh * M * (N / 8);
    const int red_blocks = static_cast<int>(std::min<int64_t>((red_total + block - 1) / block, 148 * 16));
    const int grid = cvt_blocks + red_blocks;
#define B200T5_FIN(DD, BF)                                                                                          \
    attn_bwd_finalize_kernel<DD, BF><<<grid, block, 0, stream>>>(                                                   \
        static_cast<const uint4*>(dq_ws), dq_groups, static_cast<uint8_t*>(dq), dqs[0], dqs[1], dqs[2], B, H, M,    \
        sm_scale, cvt_blocks, static_cast<const uint4*>(ds_ws), ws_pitch / 8, static_cast<uint4*>(dbias), G, N,     \
        reduce_b, reduce_h, causal ? 1 : 0)
    switch (D) {
        case 16: if (bf16) B200T5_FIN(16, true); else B200T5_FIN(16, false); break;
        case 32: if (bf16) B200T5_FIN(32, true); else B200T5_FIN(32, false); break;
        case 64: if (bf16) B200T5_FIN(64, true); else B200T5_FIN(64, false); break;
        case 128: if (bf16) B200T5_FIN(128, true); else B200T5_FIN(128, false); break;
        default: return cudaErrorInvalidValue;
    }
#undef B200T5_FIN
    count_launch();
    return cudaGetLastError();
}

// v3 kernel: row statistics (-L * log2e and -delta, padded rows) + zero-fills + (bias != NULL) the bias repack in one launch
cudaError_t launch_attn_bwd_pre_fused(const void* o, const int64_t* os, const void* dout, const int64_t* ds, const float* lse,
                                      float* nl_out, float* ndelta_out, int m_pad, void* dq_ws, int dq_groups, int B, int H, int M,
                                      int N, int D, bool bf16, void* zero_ptr, size_t zero_bytes, const void* bias, const int64_t* bs,
                                      void* bias_p, int Bb, int Hb, cudaStream_t stream) {
    const int64_t threads = (int64_t)B * H * M * (D / 8);
    const int pre_blocks = static_cast<int>((threads + 255) / 256);
    const int n_qb = 4 * ((M + 127) / 128), n_kb = (N + 127) / 128;
    const int64_t rep_blocks = bias ? (int64_t)n_qb * n_kb * Bb * Hb : 0;
    if (pre_blocks + rep_blocks > 0x7FFFFFFFLL) return cudaErrorInvalidValue;
    const int grid = pre_blocks + static_cast<int>(rep_blocks);
    const int64_t zero[4] = {0, 0, 0, 0};
    if (!bias) bs = zero;
#define B200T5_PREF(DD, BF)                                                                                                  \
    attn_bwd_pre_fused_kernel<DD, BF><<<grid, 256, 0, stream>>>(                                                              \
        static_cast<const uint8_t*>(o), os[0], os[1], os[2], static_cast<const uint8_t*>(dout), ds[0], ds[1], ds[2], ndelta_out, \
        static_cast<uint4*>(dq_ws), dq_groups, B, H, M, static_cast<uint4*>(zero_ptr), static_cast<int64_t>(zero_bytes / 16), \
        pre_blocks, static_cast<const uint16_t*>(bias), bs[0], bs[1], bs[2], bs[3], static_cast<uint4*>(bias_p), Hb, N, n_qb, n_kb, \
        lse, nl_out, m_pad)
    switch (D) {
        case 16: if (bf16) B200T5_PREF(16, true); else B200T5_PREF(16, false); break;
        case 32: if (bf16) B200T5_PREF(32, true); else B200T5_PREF(32, false); break;
        case 64: if (bf16) B200T5_PREF(64, true); else B200T5_PREF(64, false); break;
        default: return cudaErrorInvalidValue;
    }
#undef B200T5_PREF
    count_launch();
    return cudaGetLastError();
}

// dQ conversion + transposing dBias reduction in one launch.  dbias == nullptr: plain dQ conversion.
cudaError_t launch_attn_bwd_post_fused(const void* dq_ws, int dq_groups, void* dq, const int64_t* dqs, int B, int H, int M, int N,
                                       int D, float sm_scale, bool bf16, const void* ds_t, int m_pitch, void* dbias,
                                       const int64_t* dbs, int G, int reduce_b, int reduce_h, bool causal, bool out_f32,
                                       cudaStream_t stream, bool accumulate) {
    if (dbias == nullptr) return launch_attn_bwd_dq_convert(dq_ws, dq_groups, dq, dqs, B, H, M, D, sm_scale, bf16, stream);
    const int64_t cvt_threads = (int64_t)B * H * M * (D / 8);
    const int cvt_blocks = static_cast<int>((cvt_threads + 255) / 256);
    const int ob = reduce_b ? 1 : G, oh = reduce_h ? 1 : H;
    const int gx = (M + 63) / 64, gy = (N + 63) / 64;
    const int64_t red_blocks = (int64_t)gx * gy * ob * oh;
    if (cvt_blocks + red_blocks > 0x7FFFFFFFLL) return cudaErrorInvalidValue;
    const int grid = cvt_blocks + static_cast<int>(red_blocks);
#define B200T5_POSTF(DD, BF, F32)                                                                                              \
    attn_bwd_post_fused_kernel<DD, BF, F32><<<grid, 256, 0, stream>>>(                                                          \
        static_cast<const uint4*>(dq_ws), dq_groups, static_cast<uint8_t*>(dq), dqs[0], dqs[1], dqs[2], B, H, M, sm_scale,      \
        cvt_blocks, static_cast<const uint16_t*>(ds_t), m_pitch, dbias, dbs[0], dbs[1], dbs[2], dbs[3], G, N, reduce_b, reduce_h, \
        causal ? 1 : 0, gx, gy, accumulate ? 1 : 0)
#define B200T5_POSTF2(DD)                                                                            \
    if (bf16) { if (out_f32) B200T5_POSTF(DD, true, true); else B200T5_POSTF(DD, true, false); }     \
    else { if (out_f32) B200T5_POSTF(DD, false, true); else B200T5_POSTF(DD, false, false); }
    switch (D) {
        case 16: B200T5_POSTF2(16) break;
        case 32: B200T5_POSTF2(32) break;
        case 64: B200T5_POSTF2(64) break;
        default: return cudaErrorInvalidValue;
    }
#undef B200T5_POSTF2
#undef B200T5_POSTF
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200t5
