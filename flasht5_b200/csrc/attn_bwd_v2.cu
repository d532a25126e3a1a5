// FlashAttention-2 backward with additive (T5) bias for sm_100a, head dims 16 / 32 / 64: the
// software-pipelined version of the fused kernel in attn_bwd.cu (which stays as the D = 128 path).
//
// Replaces /root/reference/src/model/ops/flash_attention_v2_bias.py:
//   _bwd_kv_kernel :559-745 and _bwd_q_kernel :748-905 (one kernel, 5 tensor-core contractions per tile).
//
// One CTA = one (batch, head, 128-key block); K and V stay resident in shared memory; the CTA walks
// the query sequence in 128-row blocks k = 0, 1, ...:
//
//   tensor pipe (1 thread):  S,dP(k+1) | dV,dK(k) | dQ(k) | S,dP(k+2) | ...
//   compute warpgroups:      [B] P,dS(k+1) in registers   (overlaps dV,dK,dQ(k) and S,dP(k+2))
//                            [C] wait until dV,dK,dQ(k) have finished reading the P / dS tiles
//                            [E] drain dQ(k) TMEM -> 16-bit staging tile (double buffered) -> TMA reduce-add
//                                into the dQ group surface
//                            [D] P,dS(k+1) -> shared memory (16-bit, 128B swizzle) -> signal the MMA thread,
//                                dS tile -> global through TMA (reduce-add into the batch-group surface)
//
// The difference from attn_bwd.cu is the order of [B]..[E]: the dQ drain of block k is deferred until
// after the math of block k+1, and it has its own staging tile, so the elementwise work no longer sits
// between two dependent groups of MMAs.
//
//   warp 4 : TMA producer (K, V once; Q / dO 2-stage ring)      warp 6 : TMA producer for the bias halves
//   warp 5 : tcgen05.mma issuer                                 warps 0-3, 8-11 : compute (64 columns each)
//
// TMEM columns: S [0,128) | dP [128,256) | dV [256,256+D) | dK [256+D,256+2D) | dQ [256+2D,256+3D).
//
// Bias mode 3 (T5 relative-position bias computed in the kernel, see attn_fwd.cu): the per-head band of bias values
// lives in the shared memory of the dense bias halves; dS takes the same route as for a dense (1, H, M, N) bias.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kBM = 128;
constexpr int kBN = 128;
constexpr int kHalfBytes = 128 * 64 * 2;   // one [128][64] 16-bit swizzled half tile
constexpr float kLog2e = 1.4426950408889634f;
#ifndef B200T5_EXP2_POLY
#define B200T5_EXP2_POLY 0      // developer switch, see attn_fwd.cu
#endif
#ifndef B200T5_BIAS_FHADD
#define B200T5_BIAS_FHADD 0     // developer switch, see attn_fwd.cu
#endif
// Developer switch: the two compute warpgroups of a CTA work on the two column halves of the same tile in lock-step, so the
// two warps that share a scheduler run their MUFU-bound P / dS chunks at the same time (each at half rate) and load S / dP /
// bias at the same time (MUFU idle) -- the pattern the forward timeline exposed (attn_fwd_pingpong.cu).  With
// B200T5_BWD_PINGPONG=1 the warpgroups pass a token through two named barriers around every 32-column chunk of math, so one
// computes while the other waits for its tcgen05.ld and reads its bias.  Not yet run on hardware.
#ifndef B200T5_BWD_PINGPONG
#define B200T5_BWD_PINGPONG 0
#endif
static_assert(!(B200T5_BWD_PINGPONG && B200T5_BIAS_FHADD), "the two developer switches are not combined yet");
constexpr int kBwdToken0 = 4;       // named barriers 4, 5 (1-3 are taken)

template <int kD>
struct Bwd2Cfg {
    static_assert(kD == 16 || kD == 32 || kD == 64, "v2 backward covers head dims 16, 32, 64");
    static constexpr int kRowBytes = kD * 2;
    static constexpr int kTileBytes = kBM * kD * 2;
    static constexpr uint32_t kSwizzle = kRowBytes == 128 ? kSwz128 : (kRowBytes == 64 ? kSwz64 : kSwz32);
    static constexpr int kQStages = 2;
    // dQ staging: one [128][D] tile in the io dtype (128B swizzle at D = 64, linear below), double buffered
    static constexpr int kDqStageBytes = kBM * kD * 2;
    static constexpr int kDqColsPerWg = kD >= 32 ? kD / 2 : kD;     // D = 16: warpgroup 0 drains everything
    static constexpr int kK = 0;
    static constexpr int kV = kK + kTileBytes;
    static constexpr int kQ = kV + kTileBytes;
    static constexpr int kDO = kQ + kQStages * kTileBytes;
    static constexpr int kBias = kDO + kQStages * kTileBytes;
    static constexpr int kP = kBias + 2 * kHalfBytes;
    static constexpr int kDS = kP + 2 * kHalfBytes;
    static constexpr int kDQ = kDS + 2 * kHalfBytes;
    static constexpr int kBars = kDQ + 2 * kDqStageBytes;
    static constexpr int kNumBars = 1 + 2 * kQStages + 4 + 5;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static_assert(kTotal <= 232448, "shared memory budget");
    static constexpr int kColS = 0;
    static constexpr int kColDP = 128;
    static constexpr int kColDV = 256;
    static constexpr int kColDK = 256 + kD;
    static constexpr int kColDQ = 256 + 2 * kD;
};

template <int kN>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
    if constexpr (kN == 32) tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(r));
    else tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(r));
}

// P and dS for 32 columns of one row.  sr / dr: S and dP accumulators (fp32 bits); bv: bias (already fp32);
// outputs packed 16-bit pairs.  kMask: this tile crosses the causal diagonal or the key tail.  kSum: also add the
// (unrounded) dS values of the chunk to ds_sum (bias mode 3, tiles whose bias is one constant: see the kernel).
template <bool kBf16, bool kMask, bool kSum = false>
__device__ __forceinline__ void p_ds_chunk(const uint32_t (&sr)[32], const uint32_t (&dr)[32], const float (&bv)[32],
                                           float scale_log2, float neg_L_log2, float dlt, int lim, uint32_t (&pp)[16],
                                           uint32_t (&dd)[16], float* ds_sum = nullptr) {
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        float pe[2], de[2];
#if B200T5_EXP2_POLY > 0
        // developer build: of every 8 column pairs, B200T5_EXP2_POLY take the FMA-pipe exp2 (common.cuh).  P <= 1, so the
        // argument never exceeds 0 by more than rounding; rows with L = -inf (no visible key) have a = -inf, which the
        // polynomial clamps to 2^-125: they are zeroed explicitly, like masked columns.
        const float a0 = fmaf(__uint_as_float(sr[c]), scale_log2, fmaf(bv[c], kLog2e, neg_L_log2));
        const float a1 = fmaf(__uint_as_float(sr[c + 1]), scale_log2, fmaf(bv[c + 1], kLog2e, neg_L_log2));
        if (((c / 2) % 8) < B200T5_EXP2_POLY) {
            ex2_poly_pair(a0, a1, pe[0], pe[1]);
            if (neg_L_log2 == -INFINITY) pe[0] = pe[1] = 0.f;
        } else {
            pe[0] = ex2_approx(a0);
            pe[1] = ex2_approx(a1);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float e = pe[q];
            if (kMask && (c + q >= lim)) e = 0.f;
            pe[q] = e;
            de[q] = e * (__uint_as_float(dr[c + q]) - dlt);
        }
#else
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            // exp2(S*scale*log2e + bias*log2e - L*log2e)
            const float t = fmaf(bv[c + q], kLog2e, neg_L_log2);
            float e = ex2_approx(fmaf(__uint_as_float(sr[c + q]), scale_log2, t));
            if (kMask && (c + q >= lim)) e = 0.f;
            pe[q] = e;
            de[q] = e * (__uint_as_float(dr[c + q]) - dlt);
        }
#endif
        pp[c / 2] = pack2<kBf16>(pe[0], pe[1]);
        dd[c / 2] = pack2<kBf16>(de[0], de[1]);
        if constexpr (kSum) {
            sum0 += de[0];
            sum1 += de[1];
        }
    }
    if constexpr (kSum) *ds_sum += sum0 + sum1;
}

#if B200T5_BIAS_FHADD
// Developer build, dense bias with sm_scale == 1: the score is rebuilt exactly as the forward builds it -- a = S + bias with
// one mixed-precision add straight from the packed 16-bit pair, then exp2(a * log2e - L * log2e) -- which saves the unpack
// and one FMA per element (5.5 instead of 6.5 instructions).  bw: the 16 packed bias words of this 32-column chunk.
template <bool kBf16, bool kMask>
__device__ __forceinline__ void p_ds_chunk_fhadd(const uint32_t (&sr)[32], const uint32_t (&dr)[32], const uint32_t (&bw)[16],
                                                 float neg_L_log2, float dlt, int lim, uint32_t (&pp)[16], uint32_t (&dd)[16]) {
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        float a0, a1, pe[2], de[2];
        add_f32_16x2<kBf16>(bw[c / 2], __uint_as_float(sr[c]), __uint_as_float(sr[c + 1]), a0, a1);
        pe[0] = ex2_approx(fmaf(a0, kLog2e, neg_L_log2));
        pe[1] = ex2_approx(fmaf(a1, kLog2e, neg_L_log2));
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float e = pe[q];
            if (kMask && (c + q >= lim)) e = 0.f;
            pe[q] = e;
            de[q] = e * (__uint_as_float(dr[c + q]) - dlt);
        }
        pp[c / 2] = pack2<kBf16>(pe[0], pe[1]);
        dd[c / 2] = pack2<kBf16>(de[0], de[1]);
    }
}
#endif

}  // namespace

#ifdef B200T5_BWD_TIMING
__device__ long long g_bwd_ts[2][16][8];      // [role][iteration][slot]
#define BWD_TS(role, k, slot)                                                          \
    do {                                                                               \
        if (blockIdx.x == 777 && (k) < 16) g_bwd_ts[role][k][slot] = clock64();        \
    } while (0)
#else
#define BWD_TS(role, k, slot) do { } while (0)
#endif

template <int kD, bool kBf16, int kBiasMode, bool kCausal>
__global__ void __launch_bounds__(384, 1)
attn_bwd_kernel_v2(const __grid_constant__ AttnBwdKernelParams p) {
    using C = Bwd2Cfg<kD>;
    extern __shared__ __align__(1024) uint8_t smem[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- work decode: batch fastest (bias tiles shared in L2), long key blocks first when causal ----
    const int nnb = p.num_n_blocks;
    int bid = blockIdx.x;
    const int b = bid % p.B;
    bid /= p.B;
    const int nb = kCausal ? (bid % nnb) : (nnb - 1 - bid % nnb);
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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kTmemSlot);

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
#if B200T5_BWD_PINGPONG
        *reinterpret_cast<float*>(smem + C::kTmemSlot + 4) = 0.f;
#endif
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
            tma_load_4d(smem + C::kK, &p.map_k, kv_full, 0, col0, h, b);
            tma_load_4d(smem + C::kV, &p.map_v, kv_full, 0, col0, h, b);
            for (int k = 0; k < n_iter; ++k) {
                const int s = k % C::kQStages;
                const int mrow0 = (i_start + k) * kBM;
                mbar_wait_producer(qdo_empty + s, ((k / C::kQStages) & 1) ^ 1);
                mbar_arrive_expect_tx(qdo_full + s, 2 * C::kTileBytes);
                tma_load_4d(smem + C::kQ + s * C::kTileBytes, &p.map_q, qdo_full + s, 0, mrow0, h, b);
                tma_load_4d(smem + C::kDO + s * C::kTileBytes, &p.map_do, qdo_full + s, 0, mrow0, h, b);
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
            // ---- MMA issuer: the whole warp runs the loop (descriptor arithmetic stays warp-uniform), one
            //      elected lane issues the tcgen05 instructions ----
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, 128, 128, false, false);    // S, dP
            constexpr uint32_t idesc_dkv = make_idesc(kBf16, 128, kD, true, true);     // dV, dK
            constexpr uint32_t idesc_dq = make_idesc(kBf16, 128, kD, false, true);     // dQ
            constexpr uint32_t sbo = 8 * C::kRowBytes;
            constexpr uint32_t hi_op = sdesc_hi(sbo, C::kSwizzle);       // Q, dO, K, V tiles (either major)
            constexpr uint32_t hi_pds = sdesc_hi(1024, kSwz128);         // P, dS tiles (either major)
            const uint32_t k_lo = sdesc_lo(smem_u32(smem + C::kK), 16);           // K-major (S)
            const uint32_t v_lo = sdesc_lo(smem_u32(smem + C::kV), 16);           // K-major (dP)
            const uint32_t k_mn_lo = sdesc_lo(smem_u32(smem + C::kK), C::kTileBytes);   // MN-major (dQ)
            const uint32_t q_lo0 = sdesc_lo(smem_u32(smem + C::kQ), 16);
            const uint32_t do_lo0 = sdesc_lo(smem_u32(smem + C::kDO), 16);
            const uint32_t q_mn_lo0 = sdesc_lo(smem_u32(smem + C::kQ), C::kTileBytes);
            const uint32_t do_mn_lo0 = sdesc_lo(smem_u32(smem + C::kDO), C::kTileBytes);
            const uint32_t p_mn_lo = sdesc_lo(smem_u32(smem + C::kP), kHalfBytes);      // A = P^T  (MN-major)
            const uint32_t ds_mn_lo = sdesc_lo(smem_u32(smem + C::kDS), kHalfBytes);    // A = dS^T (MN-major)
            const uint32_t ds_k_lo = sdesc_lo(smem_u32(smem + C::kDS), 16);             // A = dS   (K-major)
            const uint32_t tm_s = tmem_base + C::kColS;
            const uint32_t tm_dp = tmem_base + C::kColDP;
            const uint32_t tm_dv = tmem_base + C::kColDV;
            const uint32_t tm_dk = tmem_base + C::kColDK;
            const uint32_t tm_dq = tmem_base + C::kColDQ;

            auto issue_s_dp = [&](int k) {
                const uint32_t so = (k % C::kQStages) * (C::kTileBytes >> 4);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_ss2(tm_s, q_lo0 + so + kk * 2, hi_op, k_lo + kk * 2, hi_op, idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_ss2(tm_dp, do_lo0 + so + kk * 2, hi_op, v_lo + kk * 2, hi_op, idesc_s, kk > 0 ? 1u : 0u);
                    umma_commit(sdp_full);
                }
                __syncwarp();
            };
            auto issue_dv_dk = [&](int k) {
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
            };
            auto issue_dq = [&]() {
                if (leader) {
                    // dQ_blk = dS K                    (K dimension = the 128 keys of this CTA)
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk)
                        umma_ss2(tm_dq, ds_k_lo + (((kk / 4) * kHalfBytes + (kk % 4) * 32) >> 4), hi_pds,
                                 k_mn_lo + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dq, kk > 0 ? 1u : 0u);
                    umma_commit(dq_full);
                }
                __syncwarp();
            };

            mbar_wait(kv_full, 0);
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
                if (lane == 0) BWD_TS(1, k, 0);
                mbar_wait(pds_full, k & 1);
                tc_fence_after();
                if (lane == 0) BWD_TS(1, k, 1);
                issue_dv_dk(k);
                if (k > 0) {
                    mbar_wait(dq_empty, (k - 1) & 1);    // dQ(k-1) has been drained out of TMEM
                    tc_fence_after();
                }
                if (lane == 0) BWD_TS(1, k, 2);
                issue_dq();
                if (lane == 0) BWD_TS(1, k, 3);
            }
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
        const int g = b % p.ds_groups;
        const float scale_log2 = p.sm_scale * kLog2e;
        const int64_t stat_base = ((int64_t)b * p.H + h) * p.M;

        // dQ : TMEM -> 16-bit staging tile `buf` (my 32 / 16 columns of my row)
        auto drain_dq = [&](int buf) {
            if (wg == 0 || kD >= 32) {
                constexpr int kCols = C::kDqColsPerWg;
                constexpr int kChunk = kCols >= 32 ? 32 : kCols;
                uint8_t* row_base = smem + C::kDQ + buf * C::kDqStageBytes + r * (kD * 2);
#pragma unroll
                for (int c0 = 0; c0 < kCols; c0 += kChunk) {
                    uint32_t q[kChunk];
                    tmem_ld_cols<kChunk>(tm_dq + c0, q);
                    tmem_ld_wait();
                    const int gcol = wg * C::kDqColsPerWg + c0;            // first dQ column of this chunk
#pragma unroll
                    for (int i = 0; i < kChunk; i += 8) {
                        const int c16 = (gcol + i) / 8;                    // 16-byte chunk inside the row
                        const int off = (kD == 64) ? ((c16 ^ (r & 7)) << 4) : (c16 << 4);
                        uint4 v;
                        v.x = pack2<kBf16>(__uint_as_float(q[i + 0]), __uint_as_float(q[i + 1]));
                        v.y = pack2<kBf16>(__uint_as_float(q[i + 2]), __uint_as_float(q[i + 3]));
                        v.z = pack2<kBf16>(__uint_as_float(q[i + 4]), __uint_as_float(q[i + 5]));
                        v.w = pack2<kBf16>(__uint_as_float(q[i + 6]), __uint_as_float(q[i + 7]));
                        *reinterpret_cast<uint4*>(row_base + off) = v;
                    }
                }
            }
        };
        const int dq_c3 = (nb % p.dq_groups) * p.B + b;      // (group, batch) slice of the dQ surface
        auto issue_dq_reduce = [&](int buf, int mrow0) {
            tma_reduce_add_4d(&p.map_dq, smem + C::kDQ + buf * C::kDqStageBytes, 0, mrow0, h, dq_c3);
            bulk_commit_group();
        };

        const float* band = reinterpret_cast<const float*>(smem + C::kBias);   // [bias mode 3]
        if (kBiasMode == 3) {
            float* dst = reinterpret_cast<float*>(smem + C::kBias);
            const float* src = p.rpe.band + (int64_t)h * p.rpe.band_len;
            for (int i = ctid; i < p.rpe.band_len; i += 256) dst[i] = __ldg(src + i);
            named_bar_sync(3, 256);
        }
        // bias mode 3 with p.rpe.dconst: tiles entirely beyond a constant end of the bucket table send no dS tile to the
        // surface; their dS is summed here (fp32) and leaves the CTA as two numbers (one per side)
        const bool rpe_skip = kBiasMode == 3 && p.rpe.dconst != nullptr;
        float ds_const_lo = 0.f, ds_const_hi = 0.f;
#if B200T5_BWD_PINGPONG
        const uint32_t token_zero = smem_u32(smem + C::kTmemSlot + 4);            // holds 0.0f
        const uint32_t token_dump = smem_u32(smem + C::kTmemSlot + 8 + 4 * wg);   // write-only scratch word
        if (wg == 1) named_bar_arrive(kBwdToken0 + 0, 256);                       // the token starts with warpgroup 0
#endif

        // row statistics are prefetched one block ahead (global latency off the critical path)
        float L_next = 0.f, dlt_next = 0.f;
        if (n_iter > 0 && i_start * kBM + r < p.M) {
            L_next = __ldg(p.lse + stat_base + i_start * kBM + r);
            dlt_next = __ldg(p.delta + stat_base + i_start * kBM + r);
        }

        for (int k = 0; k < n_iter; ++k) {
            const int mrow0 = (i_start + k) * kBM;
            const int grow = mrow0 + r;
            const bool row_ok = grow < p.M;
            const float Lv = L_next, dlt = row_ok ? dlt_next : 0.f;
            if (k + 1 < n_iter && grow + kBM < p.M) {
                L_next = __ldg(p.lse + stat_base + grow + kBM);
                dlt_next = __ldg(p.delta + stat_base + grow + kBM);
            }
            // out-of-range rows and rows with no visible key (L = -inf): P = 0
            const float neg_L_log2 = (row_ok && Lv != -INFINITY) ? -Lv * kLog2e : -INFINITY;
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

            // ---------------- [B] P and dS of this block, in registers ----------------
            if (ctid == 0) BWD_TS(0, k, 0);
            mbar_wait(sdp_full, k & 1);
            tc_fence_after();
            if (ctid == 0) BWD_TS(0, k, 1);
            if (kBiasMode == 1) mbar_wait(b_full + wg, k & 1);
            uint32_t pp[2][16], dd[2][16];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t sr[32], dr[32];
                tmem_ld32(tm_s + ch * 32, sr);
                tmem_ld32(tm_dp + ch * 32, dr);
                tmem_ld_wait();
                if (ch == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sdp_empty);   // S / dP columns may be overwritten by block k+1
                }
#if B200T5_BIAS_FHADD
                if (kBiasMode == 1 && p.sm_scale == 1.f) {
                    uint32_t bwp[16];
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const uint4 u = *reinterpret_cast<const uint4*>(sB + (((ch * 4 + c8) ^ (r & 7)) << 4));
                        bwp[c8 * 4 + 0] = u.x;
                        bwp[c8 * 4 + 1] = u.y;
                        bwp[c8 * 4 + 2] = u.z;
                        bwp[c8 * 4 + 3] = u.w;
                    }
                    if (need_mask)
                        p_ds_chunk_fhadd<kBf16, true>(sr, dr, bwp, neg_L_log2, dlt, lim - ch * 32, pp[ch], dd[ch]);
                    else
                        p_ds_chunk_fhadd<kBf16, false>(sr, dr, bwp, neg_L_log2, dlt, 0, pp[ch], dd[ch]);
                    continue;
                }
#endif
                float bv[32];
                if (kBiasMode == 1) {
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const uint4 u = *reinterpret_cast<const uint4*>(sB + (((ch * 4 + c8) ^ (r & 7)) << 4));
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = unpack2<kBf16>(w[e]);
                            bv[c8 * 8 + 2 * e] = f.x;
                            bv[c8 * 8 + 2 * e + 1] = f.y;
                        }
                    }
                } else if (kBiasMode == 2) {
                    const uint16_t* bp = reinterpret_cast<const uint16_t*>(p.bias) + (int64_t)bb * p.bias_sb +
                                         (int64_t)hb * p.bias_sh + (int64_t)grow * p.bias_sm;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int c = col0 + wg * 64 + ch * 32 + e;
                        bv[e] = (row_ok && c < p.N) ? to_float16bit<kBf16>(__ldg(bp + (int64_t)c * p.bias_sn)) : 0.f;
                    }
                } else if (kBiasMode == 3) {
                    if (rpe_const) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) bv[e] = rpe_cval;
                    } else {
                        const float* bp = band + (col0 + wg * 64 + ch * 32 - grow - p.rpe.band_lo);
#pragma unroll
                        for (int e = 0; e < 32; ++e) bv[e] = bp[e];
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) bv[e] = 0.f;
                }
#if B200T5_BWD_PINGPONG
                // my turn for the MUFU-bound chunk.  ptxas moves register arithmetic across barriers (attn_fwd_pingpong.cu): the
                // operand every exp2 of the chunk shares takes a zero read from shared memory after the barrier ...
                named_bar_sync(kBwdToken0 + wg, 256);
                const float nl = neg_L_log2 + ld_shared_volatile_f32(token_zero);
#else
                const float nl = neg_L_log2;
#endif
                if (kBiasMode == 3 && rpe_skip && rpe_const) {
                    float* acc = (col0 - mrow0 + (kBN - 1) <= p.rpe.const_lo) ? &ds_const_lo : &ds_const_hi;
                    if (need_mask)
                        p_ds_chunk<kBf16, true, true>(sr, dr, bv, scale_log2, nl, dlt, lim - ch * 32, pp[ch], dd[ch], acc);
                    else
                        p_ds_chunk<kBf16, false, true>(sr, dr, bv, scale_log2, nl, dlt, 0, pp[ch], dd[ch], acc);
                } else if (need_mask)
                    p_ds_chunk<kBf16, true>(sr, dr, bv, scale_log2, nl, dlt, lim - ch * 32, pp[ch], dd[ch]);
                else
                    p_ds_chunk<kBf16, false>(sr, dr, bv, scale_log2, nl, dlt, 0, pp[ch], dd[ch]);
#if B200T5_BWD_PINGPONG
                // ... and a word that depends on the first and the last results of the chunk is stored before the token moves on
                st_shared_volatile_f32(token_dump, __uint_as_float(pp[ch][0] ^ pp[ch][15] ^ dd[ch][0] ^ dd[ch][15]));
                named_bar_arrive(kBwdToken0 + (wg ^ 1), 256);
#endif
            }
            if (kBiasMode == 1) {
                fence_proxy_async_smem();                    // bias reads complete before TMA refills the half
                __syncwarp();
                if (lane == 0) mbar_arrive(b_empty + wg);
            }

            // ---------------- [C] the previous block's dV / dK / dQ MMAs are done with the P / dS tiles ------
            if (ctid == 0) BWD_TS(0, k, 2);
            if (k > 0) {
                mbar_wait(dq_full, (k - 1) & 1);
                tc_fence_after();
            }
            if (ctid == 0) BWD_TS(0, k, 3);

            // ---------------- [E] drain dQ of the previous block (its MMAs completed at [C]) ----------------
            // Staging buffer (k-1)&1 was last read by the TMA reduce of dQ(k-3), certified at barrier 2 of the
            // previous iteration; the dS tile was last read by the TMA reduce of dS(k-1), issued one whole
            // [B] ago -- thread 0 waits for it here so that [D] below may overwrite the tile.
            if (k > 0) drain_dq((k - 1) & 1);
            tc_fence_before();
            fence_proxy_async_smem();
            if (ctid == 0) BWD_TS(0, k, 4);
            if (ctid == 0) bulk_wait_group_read<0>();
            if (ctid == 0) BWD_TS(0, k, 5);
            named_bar_sync(2, 256);
            if (ctid == 0) BWD_TS(0, k, 6);
            if (ctid == 0 && k > 0) {
                mbar_arrive(dq_empty);                       // TMEM dQ columns are free for block k
                issue_dq_reduce((k - 1) & 1, mrow0 - kBM);
            }

            // ---------------- [D] P, dS -> shared memory; hand over to the MMA thread; dS tile -> global -----
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const int off = ((ch * 4 + c8) ^ (r & 7)) << 4;
                    *reinterpret_cast<uint4*>(sP + off) =
                        make_uint4(pp[ch][c8 * 4], pp[ch][c8 * 4 + 1], pp[ch][c8 * 4 + 2], pp[ch][c8 * 4 + 3]);
                    *reinterpret_cast<uint4*>(sDS + off) =
                        make_uint4(dd[ch][c8 * 4], dd[ch][c8 * 4 + 1], dd[ch][c8 * 4 + 2], dd[ch][c8 * 4 + 3]);
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (ctid == 0) BWD_TS(0, k, 7);
            if (ctid == 0) {
                mbar_arrive(pds_full);
                if (kBiasMode != 0 && !(kBiasMode == 3 && rpe_skip && rpe_const)) {
                    if (p.ds_use_reduce) {
                        tma_reduce_add_4d(&p.map_ds, smem + C::kDS, col0, mrow0, h, g);
                        tma_reduce_add_4d(&p.map_ds, smem + C::kDS + kHalfBytes, col0 + 64, mrow0, h, g);
                    } else {
                        tma_store_4d(&p.map_ds, smem + C::kDS, col0, mrow0, h, g);
                        tma_store_4d(&p.map_ds, smem + C::kDS + kHalfBytes, col0 + 64, mrow0, h, g);
                    }
                    bulk_commit_group();
                }
            }
        }

#if B200T5_BWD_PINGPONG
        if (wg == 0) named_bar_sync(kBwdToken0 + 0, 256);        // consume the arrival warpgroup 1 posted after its last chunk
#endif
        // ---- tail: dQ of the last block, then dV (warpgroup 0) and dK * sm_scale (warpgroup 1) ----
        if (n_iter > 0) {
            mbar_wait(dq_full, (n_iter - 1) & 1);            // every MMA of this CTA has completed
            tc_fence_after();
            drain_dq((n_iter - 1) & 1);                      // that buffer's last reader was certified at barrier 2
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar_sync(2, 256);
            if (ctid == 0) issue_dq_reduce((n_iter - 1) & 1, (i_start + n_iter - 1) * kBM);
        }
        if (kBiasMode == 3 && rpe_skip) {
            // 256 threads x 2 sums -> 8 warps x 2 -> 2 atomics per CTA.  The band's shared memory is free again: every
            // thread has left the loop when it reaches the barrier (barrier 2 above when n_iter > 0, the first one below).
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ds_const_lo += __shfl_xor_sync(0xffffffffu, ds_const_lo, off);
                ds_const_hi += __shfl_xor_sync(0xffffffffu, ds_const_hi, off);
            }
            float* red = reinterpret_cast<float*>(smem + C::kBias);
            named_bar_sync(3, 256);
            if (lane == 0) {
                red[(ctid >> 5) * 2 + 0] = ds_const_lo;
                red[(ctid >> 5) * 2 + 1] = ds_const_hi;
            }
            named_bar_sync(3, 256);
            if (ctid < 2) {
                float t = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) t += red[w8 * 2 + ctid];
                if (t != 0.f) atomicAdd(p.rpe.dconst + h * 2 + ctid, t);
            }
        }
        {
            const int gn = col0 + r;
            const bool row_ok = gn < p.N;
            uint8_t* out_row = wg == 0
                ? reinterpret_cast<uint8_t*>(p.dv) + 2 * ((int64_t)b * p.dv_sb + (int64_t)h * p.dv_sh + (int64_t)gn * p.dv_sn)
                : reinterpret_cast<uint8_t*>(p.dk) + 2 * ((int64_t)b * p.dk_sb + (int64_t)h * p.dk_sh + (int64_t)gn * p.dk_sn);
            const float sc = wg == 0 ? 1.f : p.sm_scale;
            if (n_iter > 0) {
                const uint32_t tm_acc = tmem_base + lane_off + (wg == 0 ? C::kColDV : C::kColDK);
                constexpr int kChunk = kD >= 32 ? 32 : 16;
#pragma unroll
                for (int c0 = 0; c0 < kD; c0 += kChunk) {
                    uint32_t a[kChunk];
                    tmem_ld_cols<kChunk>(tm_acc + c0, a);
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
// host-side launcher
// ------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kBiasMode, bool kCausal>
static cudaError_t launch_bwd2_inst(const AttnBwdKernelParams& kp, cudaStream_t stream) {
    using C = Bwd2Cfg<kD>;
    auto kern = attn_bwd_kernel_v2<kD, kBf16, kBiasMode, kCausal>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
    if (e != cudaSuccess) return e;
    const int grid = kp.B * kp.H * kp.num_n_blocks;
    kern<<<grid, 384, C::kTotal, stream>>>(kp);
    count_launch();
#ifdef B200T5_BWD_TIMING
    {
        cudaDeviceSynchronize();
        long long ts[2][16][8];
        cudaMemcpyFromSymbol(ts, g_bwd_ts, sizeof(ts));
        const long long t0 = ts[0][0][0];
        printf("compute thread 0 (cycles rel. to first stamp): k: [B wait-start, B start, B end, C end, E drained, tma-wait end, bar2, bar1]\n");
        for (int k = 0; k < 8; ++k) {
            printf(" k=%d:", k);
            for (int j = 0; j < 8; ++j) printf(" %7lld", ts[0][k][j] - t0);
            printf("\n");
        }
        printf("mma thread: k: [wait pds_full start, pds_full, after dq_empty, issued dq]\n");
        for (int k = 0; k < 8; ++k) {
            printf(" k=%d:", k);
            for (int j = 0; j < 4; ++j) printf(" %7lld", ts[1][k][j] - t0);
            printf("\n");
        }
    }
#endif
    return cudaGetLastError();
}

template <int kD, bool kBf16>
static cudaError_t launch_bwd2_d(const AttnBwdKernelParams& kp, int bias_mode, bool causal, cudaStream_t stream) {
    switch (bias_mode * 2 + (causal ? 1 : 0)) {
        case 0: return launch_bwd2_inst<kD, kBf16, 0, false>(kp, stream);
        case 1: return launch_bwd2_inst<kD, kBf16, 0, true>(kp, stream);
        case 2: return launch_bwd2_inst<kD, kBf16, 1, false>(kp, stream);
        case 3: return launch_bwd2_inst<kD, kBf16, 1, true>(kp, stream);
        case 4: return launch_bwd2_inst<kD, kBf16, 2, false>(kp, stream);
        case 5: return launch_bwd2_inst<kD, kBf16, 2, true>(kp, stream);
        case 6: return launch_bwd2_inst<kD, kBf16, 3, false>(kp, stream);
        case 7: return launch_bwd2_inst<kD, kBf16, 3, true>(kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_attn_bwd_v2(const AttnBwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                               cudaStream_t stream) {
#ifdef B200T5_HEADLINE_ONLY
    if (D == 64 && bf16) return launch_bwd2_d<64, true>(kp, bias_mode, causal, stream);
    return cudaErrorInvalidValue;
#else
    switch (D) {
        case 16: return bf16 ? launch_bwd2_d<16, true>(kp, bias_mode, causal, stream) : launch_bwd2_d<16, false>(kp, bias_mode, causal, stream);
        case 32: return bf16 ? launch_bwd2_d<32, true>(kp, bias_mode, causal, stream) : launch_bwd2_d<32, false>(kp, bias_mode, causal, stream);
        case 64: return bf16 ? launch_bwd2_d<64, true>(kp, bias_mode, causal, stream) : launch_bwd2_d<64, false>(kp, bias_mode, causal, stream);
        default: return cudaErrorInvalidValue;
    }
#endif
}

}  // namespace b200t5
