// FlashAttention-2 backward with additive (T5) bias for sm_100a, head dims 16 / 32 / 64 -- transposed formulation.
//
// Replaces /root/reference/src/model/ops/flash_attention_v2_bias.py:
//   _bwd_kv_kernel :559-745 and _bwd_q_kernel :748-905 (one kernel, 5 tensor-core contractions per tile).
//
// Why this shape (measured, round 2): the previous kernel (lanes = query rows, P / dS handed to the dV / dK / dQ MMAs
// through shared memory) moved ~430 KB per 128x128 tile through the 128 B/clk shared-memory port (3 400 of its 3 900
// cycles per tile), and its stages ran back to back.  Here the CTA computes the TRANSPOSED score tile,
//     S^T = K Q^T   and   dP^T = V dO^T          (TMEM lanes = keys, columns = queries),
// with K and V held in TMEM for the whole CTA as the A operands (written once with tcgen05.st), so that
//     dV += P^T dO   and   dK += dS^T Q
// take P^T / dS^T straight FROM TMEM (the compute warps write them back over S^T / dP^T as packed 16-bit pairs), and
// only dS goes to shared memory -- once, in the layout the dQ MMA reads as an MN-major A operand and the TMA unit
// reduces into the (transposed) dBias surface.  Shared-memory traffic per tile drops from ~430 KB to ~210 KB (+96 KB
// with a dense bias).
//
// One CTA = one (batch, head, 128-key block); it walks the query sequence in 128-row tiles, each processed as two
// 64-query half tiles t = 2k, 2k+1 that alternate between two S^T / dP^T buffers in TMEM and two compute warpgroups:
//
//   tensor pipe (one thread)   S,dP(t+1) | dV,dK(t) | [dQ(k) after the second half] | S,dP(t+2) | ...
//   compute warpgroup t & 1    wait S,dP(t) -> P, dS in registers -> P^T, dS^T -> TMEM, dS -> smem -> signal
//   drain warpgroup            dQ(k): TMEM -> 16-bit staging tile -> TMA reduce-add into the dQ group surface
//
//   warps 0-3 / 4-7 : compute warpgroups (thread = key row)     warps 8-11 : K, V -> TMEM (prologue), dQ drain
//   warp 12 : TMA producer K, Q / dO ring      warp 13 : tcgen05.mma issuer      warp 14 : TMA producer of the bias tiles
//
// TMEM columns: S^T [0,64) [64,128) | dP^T [128,192) [192,256) | dV [256,256+D) | dK [256+D,256+2D) |
//               dQ [256+2D,256+3D) | K [448,448+D/2) | V [480,480+D/2)       (P^T / dS^T alias the first 32 columns of
//               their S^T / dP^T buffer).
//
// Bias modes: 0 none | 1 dense bias, read through a TRANSPOSED copy (B|1, H|1, N, M) that api.cu makes in the workspace
// (so that a thread = key row reads its bias values with 16-byte loads) | 3 T5 relative-position bias from the band.
// dS leaves the CTA transposed as well: the surface is (G, H, N, M) and the finalize kernel transposes while it reduces.
#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kBM = 128;            // queries per tile
constexpr int kBN = 128;            // keys per CTA
constexpr int kSub = 64;            // queries per half tile
constexpr int kBoxBytes = 128 * 64 * 2;   // [128 keys][64 queries] 16-bit, 128B-swizzled rows
constexpr float kLog2e = 1.4426950408889634f;

template <int kD>
struct Bwd3Cfg {
    static_assert(kD == 16 || kD == 32 || kD == 64, "v3 backward covers head dims 16, 32, 64");
    static constexpr int kRowBytes = kD * 2;
    static constexpr int kTileBytes = 128 * kD * 2;
    static constexpr int kHalfTileBytes = 64 * kD * 2;
    static constexpr uint32_t kSwizzle = kRowBytes == 128 ? kSwz128 : (kRowBytes == 64 ? kSwz64 : kSwz32);
    static constexpr int kQStages = 2;
    static constexpr int kBiasStages = 4;
    static constexpr int kK = 0;
    static constexpr int kQ = kK + kTileBytes;
    static constexpr int kDO = kQ + kQStages * kTileBytes;
    static constexpr int kBias = kDO + kQStages * kTileBytes;          // bias ring (mode 1) or the band (mode 3)
    static constexpr int kDS = kBias + kBiasStages * kBoxBytes;        // [2 boxes] dS^T of the current tile
    static constexpr int kDQ = kDS + 2 * kBoxBytes;                    // dQ staging tile [128][D] io dtype
    static constexpr int kStats = kDQ + kTileBytes;                    // [2 wg][2 buf][2][64] fp32: -L*log2e, -delta
    static constexpr int kBars = kStats + 2 * 2 * 2 * 64 * 4;
    static constexpr int kNumBars = 2 + 2 * kQStages + 2 * kBiasStages + 2 + 2 + 2;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static_assert(kTotal <= 232448, "shared memory budget");
    static_assert(kTileBytes % 1024 == 0, "swizzled tiles need 1024-byte alignment");
    static constexpr int kColS = 0;
    static constexpr int kColDP = 128;
    static constexpr int kColDV = 256;
    static constexpr int kColDK = 256 + kD;
    static constexpr int kColDQ = 256 + 2 * kD;
    static constexpr int kColKt = 448;
    static constexpr int kColVt = 480;
};

template <int kN>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
    if constexpr (kN == 32) tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(r));
    else tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(r));
}
template <int kN>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r) {
    if constexpr (kN == 32) tmem_st32(taddr, *reinterpret_cast<const uint32_t(*)[32]>(r));
    else tmem_st16(taddr, *reinterpret_cast<const uint32_t(*)[16]>(r));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// P^T and dS^T for 32 query columns of one key row.
//   sr / dr : S^T, dP^T accumulators (fp32 bits), column c <-> query m0 + c
//   nl / nd : shared-memory rows of -L * log2e and -delta for those 32 queries (broadcast 16-byte loads)
//   bias    : kBiasMode 1: the 64 bytes of this row's transposed-bias chunk are at brow + ((chunk16 ^ (r & 7)) << 4)
//             kBiasMode 3: band pointer such that bias(c) = bp[-c]  (or the constant bconst when kConst)
//   vis_lo  : [kMask] columns c < vis_lo are masked (causal: query before the key; whole row when the key is out of range)
template <bool kBf16, int kBiasMode, bool kMask, bool kConst, bool kSum>
__device__ __forceinline__ void v3_chunk(const uint32_t (&sr)[32], const uint32_t (&dr)[32], const float* nl, const float* nd,
                                         const uint8_t* brow, int chunk16_0, int rx, const float* bp, float bconst,
                                         float scale_log2, int vis_lo, uint32_t (&pp)[16], uint32_t (&dd)[16], float& ds_sum) {
    const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
    const f32x2 l2e2 = f2_pack(kLog2e, kLog2e);
    f32x2 sum2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {                       // 8 columns per group
        const float4 nla = *reinterpret_cast<const float4*>(nl + g * 8), nlb = *reinterpret_cast<const float4*>(nl + g * 8 + 4);
        const float4 nda = *reinterpret_cast<const float4*>(nd + g * 8), ndb = *reinterpret_cast<const float4*>(nd + g * 8 + 4);
        const float nlv[8] = {nla.x, nla.y, nla.z, nla.w, nlb.x, nlb.y, nlb.z, nlb.w};
        const float ndv[8] = {nda.x, nda.y, nda.z, nda.w, ndb.x, ndb.y, ndb.z, ndb.w};
        uint32_t bw[4] = {0, 0, 0, 0};
        if (kBiasMode == 1) {
            const uint4 u = *reinterpret_cast<const uint4*>(brow + (((chunk16_0 + g) ^ rx) << 4));
            bw[0] = u.x; bw[1] = u.y; bw[2] = u.z; bw[3] = u.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = g * 8 + 2 * e;
            f32x2 t = f2_pack(nlv[2 * e], nlv[2 * e + 1]);
            if (kBiasMode == 1) {
                const float2 bf = unpack2<kBf16>(bw[e]);
                t = f2_fma(f2_pack(bf.x, bf.y), l2e2, t);
            } else if (kBiasMode == 3) {
                if (kConst) t = f2_add(t, f2_pack(bconst, bconst));          // bconst already times log2e
                else t = f2_fma(f2_pack(bp[-c], bp[-c - 1]), l2e2, t);
            }
            float a0, a1;
            f2_unpack(f2_fma(f2_pack_bits(sr[c], sr[c + 1]), sc2, t), a0, a1);
            float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
            if (kMask) {
                if (c < vis_lo) p0 = 0.f;
                if (c + 1 < vis_lo) p1 = 0.f;
            }
            const f32x2 p2 = f2_pack(p0, p1);
            const f32x2 ds2 = f2_mul(p2, f2_add(f2_pack_bits(dr[c], dr[c + 1]), f2_pack(ndv[2 * e], ndv[2 * e + 1])));
            float g0, g1;
            f2_unpack(ds2, g0, g1);
            pp[c / 2] = pack2<kBf16>(p0, p1);
            dd[c / 2] = pack2<kBf16>(g0, g1);
            if (kSum) sum2 = f2_add(sum2, ds2);
        }
    }
    if (kSum) {
        float s0, s1;
        f2_unpack(sum2, s0, s1);
        ds_sum += s0 + s1;
    }
}

}  // namespace

#ifdef B200T5_BWD_TIMING
__device__ long long g_bwd3_ts[4][24][8];      // [role: wg0, wg1, mma, drain][iteration][slot]
#define BWD3_TS(role, k, slot)                                                          \
    do {                                                                                \
        if (blockIdx.x == 777 && (k) < 24) g_bwd3_ts[role][k][slot] = clock64();        \
    } while (0)
#else
#define BWD3_TS(role, k, slot) do { } while (0)
#endif

template <int kD, bool kBf16, int kBiasMode, bool kCausal>
__global__ void __launch_bounds__(512, 1)
attn_bwd_kernel_v3(const __grid_constant__ AttnBwdKernelParams p) {
    using C = Bwd3Cfg<kD>;
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
    const int T = 2 * n_iter;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
    uint64_t* k_full = bars;
    uint64_t* kt_ready = bars + 1;
    uint64_t* qdo_full = bars + 2;
    uint64_t* qdo_empty = qdo_full + C::kQStages;
    uint64_t* b_full = qdo_empty + C::kQStages;
    uint64_t* b_empty = b_full + C::kBiasStages;
    uint64_t* sdp_full = b_empty + C::kBiasStages;      // [2] one per S^T / dP^T buffer
    uint64_t* pds_full = sdp_full + 2;                  // [2]
    uint64_t* dq_full = pds_full + 2;
    uint64_t* dq_empty = dq_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kTmemSlot);

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        mbar_init(k_full, 1);
        mbar_init(kt_ready, 4);
        for (int i = 0; i < C::kQStages; ++i) {
            mbar_init(qdo_full + i, 1);
            mbar_init(qdo_empty + i, 1);
        }
        for (int i = 0; i < C::kBiasStages; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 4);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(sdp_full + i, 1);
            mbar_init(pds_full + i, 1);
        }
        mbar_init(dq_full, 1);
        mbar_init(dq_empty, 4);
        fence_mbar_init();
    }
    if (warp == 13) tmem_alloc<512>(tmem_slot);
    if (warp == 12 && lane == 0) {
        tma_prefetch_desc(&p.map_q);
        tma_prefetch_desc(&p.map_k);
        tma_prefetch_desc(&p.map_do);
        tma_prefetch_desc(&p.map_dq);
        if (kBiasMode == 1) tma_prefetch_desc(&p.map_bias);
        if (kBiasMode != 0) tma_prefetch_desc(&p.map_ds);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 12) {
        // =============================== control warps (12..15) ===============================
        setmaxnreg_dec<40>();
        if (warp == 12 && lane == 0 && n_iter > 0) {
            // ---- K once; then the Q / dO ring (one stage = one 128-query tile) ----
            mbar_arrive_expect_tx(k_full, C::kTileBytes);
            tma_load_4d(smem + C::kK, &p.map_k, k_full, 0, col0, h, b);
            for (int k = 0; k < n_iter; ++k) {
                const int s = k % C::kQStages;
                const int mrow0 = (i_start + k) * kBM;
                mbar_wait_producer(qdo_empty + s, ((k / C::kQStages) & 1) ^ 1);
                mbar_arrive_expect_tx(qdo_full + s, 2 * C::kTileBytes);
                tma_load_4d(smem + C::kQ + s * C::kTileBytes, &p.map_q, qdo_full + s, 0, mrow0, h, b);
                tma_load_4d(smem + C::kDO + s * C::kTileBytes, &p.map_do, qdo_full + s, 0, mrow0, h, b);
            }
        } else if (warp == 14 && lane == 0 && kBiasMode == 1) {
            // ---- transposed-bias tiles: box [64 queries][128 keys] per half tile ----
            const int hb = p.bias_h_bcast ? 0 : h;
            const int bb = p.bias_b_bcast ? 0 : b;
            for (int t = 0; t < T; ++t) {
                const int s = t % C::kBiasStages;
                const int m0 = (i_start + (t >> 1)) * kBM + (t & 1) * kSub;
                mbar_wait_producer(b_empty + s, ((t / C::kBiasStages) & 1) ^ 1);
                mbar_arrive_expect_tx(b_full + s, kBoxBytes);
                tma_load_4d(smem + C::kBias + s * kBoxBytes, &p.map_bias, b_full + s, m0, col0, hb, bb);
            }
        } else if (warp == 13 && n_iter > 0) {
            // ---- MMA issuer: the whole warp runs the loop (descriptor arithmetic stays warp-uniform), one elected
            //      lane issues the tcgen05 instructions ----
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, 128, kSub, false, false);   // S^T, dP^T : A tmem, B K-major
            constexpr uint32_t idesc_dkv = make_idesc(kBf16, 128, kD, false, true);    // dV, dK     : A tmem, B MN-major
            constexpr uint32_t idesc_dq = make_idesc(kBf16, 128, kD, true, true);      // dQ         : A, B MN-major
            constexpr uint32_t sbo = 8 * C::kRowBytes;
            constexpr uint32_t hi_op = sdesc_hi(sbo, C::kSwizzle);       // Q, dO, K tiles (either major)
            constexpr uint32_t hi_ds = sdesc_hi(1024, kSwz128);          // dS^T boxes
            const uint32_t q_lo0 = sdesc_lo(smem_u32(smem + C::kQ), 16);
            const uint32_t do_lo0 = sdesc_lo(smem_u32(smem + C::kDO), 16);
            const uint32_t q_mn_lo0 = sdesc_lo(smem_u32(smem + C::kQ), C::kTileBytes);
            const uint32_t do_mn_lo0 = sdesc_lo(smem_u32(smem + C::kDO), C::kTileBytes);
            const uint32_t k_mn_lo = sdesc_lo(smem_u32(smem + C::kK), C::kTileBytes);
            const uint32_t ds_mn_lo = sdesc_lo(smem_u32(smem + C::kDS), kBoxBytes);     // A = dS (M = queries, 2 boxes)
            const uint32_t tm_kt = tmem_base + C::kColKt;
            const uint32_t tm_vt = tmem_base + C::kColVt;
            const uint32_t tm_dv = tmem_base + C::kColDV;
            const uint32_t tm_dk = tmem_base + C::kColDK;
            const uint32_t tm_dq = tmem_base + C::kColDQ;

            auto issue_s_dp = [&](int t) {
                const int k = t >> 1, hh = t & 1;
                const uint32_t so = (k % C::kQStages) * (C::kTileBytes >> 4) + hh * (C::kHalfTileBytes >> 4);
                const uint32_t tm_s = tmem_base + C::kColS + hh * kSub;
                const uint32_t tm_dp = tmem_base + C::kColDP + hh * kSub;
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_ts2(tm_s, tm_kt + kk * 8, q_lo0 + so + kk * 2, hi_op, idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_ts2(tm_dp, tm_vt + kk * 8, do_lo0 + so + kk * 2, hi_op, idesc_s, kk > 0 ? 1u : 0u);
                    umma_commit(sdp_full + hh);
                }
                __syncwarp();
            };
            auto issue_dv_dk = [&](int t) {
                const int k = t >> 1, hh = t & 1;
                const uint32_t so = (k % C::kQStages) * (C::kTileBytes >> 4) + hh * (C::kHalfTileBytes >> 4);
                const uint32_t tm_p = tmem_base + C::kColS + hh * kSub;      // P^T  (packed 16-bit) over S^T
                const uint32_t tm_ds = tmem_base + C::kColDP + hh * kSub;    // dS^T (packed 16-bit) over dP^T
                if (leader) {
                    // dV += P^T dO ; dK += dS^T Q      (K dimension = the 64 queries of this half tile)
#pragma unroll
                    for (int kk = 0; kk < kSub / 16; ++kk)
                        umma_ts2(tm_dv, tm_p + kk * 8, do_mn_lo0 + so + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dkv,
                                 (t > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < kSub / 16; ++kk)
                        umma_ts2(tm_dk, tm_ds + kk * 8, q_mn_lo0 + so + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dkv,
                                 (t > 0 || kk > 0) ? 1u : 0u);
                }
                __syncwarp();
            };
            auto issue_dq = [&](int k) {
                if (leader) {
                    // dQ_tile = dS K      (K dimension = the 128 keys of this CTA; A = dS^T boxes read MN-major)
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk)
                        umma_ss2(tm_dq, ds_mn_lo + kk * (2048 >> 4), hi_ds, k_mn_lo + ((kk * 16 * C::kRowBytes) >> 4), hi_op,
                                 idesc_dq, kk > 0 ? 1u : 0u);
                    umma_commit(dq_full);
                    umma_commit(qdo_empty + (k % C::kQStages));
                }
                __syncwarp();
            };

            mbar_wait(kt_ready, 0);
            mbar_wait(qdo_full + 0, 0);
            tc_fence_after();
            issue_s_dp(0);
            for (int t = 0; t < T; ++t) {
                const int k = t >> 1, hh = t & 1;
                if (t + 1 < T) {
                    const int tn = t + 1, kn = tn >> 1;
                    if ((tn & 1) == 0) mbar_wait(qdo_full + (kn % C::kQStages), (kn / C::kQStages) & 1);
                    // buffer (tn & 1) was last used by half tile t - 1: its P^T / dS^T were waited for (pds_full) and its
                    // dV / dK MMAs issued in the previous iteration; MMAs execute in issue order
                    tc_fence_after();
                    issue_s_dp(tn);
                }
                if (lane == 0) BWD3_TS(2, t, 0);
                mbar_wait(pds_full + hh, k & 1);
                tc_fence_after();
                if (lane == 0) BWD3_TS(2, t, 1);
                issue_dv_dk(t);
                if (hh == 1) {
                    if (k == 0) mbar_wait(k_full, 0);
                    else mbar_wait(dq_empty, (k - 1) & 1);    // dQ(k-1) has been drained out of TMEM
                    tc_fence_after();
                    if (lane == 0) BWD3_TS(2, t, 2);
                    issue_dq(k);
                }
                if (lane == 0) BWD3_TS(2, t, 3);
            }
        }
    } else if (warp >= 8) {
        // =============================== drain warpgroup (8..11) ===============================
        setmaxnreg_dec<72>();
        const int r = (warp & 3) * 32 + lane;                 // TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        if (n_iter > 0) {
            // ---- prologue: K and V rows of this key block -> TMEM (the A operands of S^T and dP^T) ----
            const int gn = col0 + r;
            constexpr int kWords = kD / 2;
            uint32_t kr[kWords], vr[kWords];
            if (gn < p.N) {
                const uint4* kp4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.k) +
                                                                  2 * ((int64_t)b * p.k_sb + (int64_t)h * p.k_sh + (int64_t)gn * p.k_sn));
                const uint4* vp4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.v) +
                                                                  2 * ((int64_t)b * p.v_sb + (int64_t)h * p.v_sh + (int64_t)gn * p.v_sn));
#pragma unroll
                for (int i = 0; i < kWords / 4; ++i) {
                    const uint4 a = __ldg(kp4 + i), c = __ldg(vp4 + i);
                    kr[4 * i] = a.x; kr[4 * i + 1] = a.y; kr[4 * i + 2] = a.z; kr[4 * i + 3] = a.w;
                    vr[4 * i] = c.x; vr[4 * i + 1] = c.y; vr[4 * i + 2] = c.z; vr[4 * i + 3] = c.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < kWords; ++i) kr[i] = vr[i] = 0u;
            }
            if constexpr (kWords == 8) {
                tmem_st8(tmem_base + lane_off + C::kColKt, kr);
                tmem_st8(tmem_base + lane_off + C::kColVt, vr);
            } else {
                tmem_st_n<kWords>(tmem_base + lane_off + C::kColKt, kr);
                tmem_st_n<kWords>(tmem_base + lane_off + C::kColVt, vr);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(kt_ready);
        }
        const uint32_t tm_dq = tmem_base + lane_off + C::kColDQ;
        const int dq_c3 = (nb % p.dq_groups) * p.B + b;       // (group, batch) slice of the dQ surface
        uint8_t* stage_row = smem + C::kDQ + r * (kD * 2);
        for (int k = 0; k < n_iter; ++k) {
            if (r == 0) BWD3_TS(3, k, 0);
            mbar_wait(dq_full, k & 1);
            tc_fence_after();
            if (r == 0) BWD3_TS(3, k, 1);
            // (72 registers per thread here: convert chunk by chunk, keep only the packed words)
            uint32_t qp[kD / 2];
            constexpr int kChunk = kD >= 32 ? 32 : 16;
#pragma unroll
            for (int c0 = 0; c0 < kD; c0 += kChunk) {
                uint32_t q[kChunk];
                tmem_ld_n<kChunk>(tm_dq + c0, q);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < kChunk; i += 2) qp[(c0 + i) / 2] = pack2<kBf16>(__uint_as_float(q[i]), __uint_as_float(q[i + 1]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);             // the dQ columns may be overwritten by tile k + 1
            if (r == 0) bulk_wait_group_read<0>();            // staging tile: the previous reduce has read it
            named_bar_sync(3, 128);
#pragma unroll
            for (int i = 0; i < kD; i += 8) {
                const int c16 = i / 8;
                const int off = (kD == 64) ? ((c16 ^ (r & 7)) << 4) : (c16 << 4);
                *reinterpret_cast<uint4*>(stage_row + off) = make_uint4(qp[i / 2], qp[i / 2 + 1], qp[i / 2 + 2], qp[i / 2 + 3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            if (r == 0) {
                tma_reduce_add_4d(&p.map_dq, smem + C::kDQ, 0, (i_start + k) * kBM, h, dq_c3);
                bulk_commit_group();
            }
            if (r == 0) BWD3_TS(3, k, 2);
        }
        if (r == 0) bulk_wait_group<0>();
    } else {
        // =============================== compute warpgroups (0..3, 4..7) ===============================
        setmaxnreg_inc<200>();
        const int wg = warp >> 2;                             // half tile parity this warpgroup serves
        const int r = (warp & 3) * 32 + lane;                 // key row in the block == TMEM lane
        const int gn = col0 + r;                              // global key index
        const bool key_ok = gn < p.N;
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tm_s = tmem_base + lane_off + C::kColS + wg * kSub;
        const uint32_t tm_dp = tmem_base + lane_off + C::kColDP + wg * kSub;
        uint8_t* const sDS = smem + C::kDS + wg * kBoxBytes + r * 128;
        float* const stats = reinterpret_cast<float*>(smem + C::kStats) + wg * 256;     // [buf][2][64]
        const int bar_id = 1 + wg;
        const int rx = r & 7;
        const float scale_log2 = p.sm_scale * kLog2e;
        const int64_t stat_base = ((int64_t)b * p.H + h) * p.M;
        const int g_ds = b % p.ds_groups;

        const float* band = reinterpret_cast<const float*>(smem + C::kBias);   // [bias mode 3]
        if (kBiasMode == 3) {
            float* dst = reinterpret_cast<float*>(smem + C::kBias);
            const float* src = p.rpe.band + (int64_t)h * p.rpe.band_len;
            for (int i = threadIdx.x; i < p.rpe.band_len; i += 256) dst[i] = __ldg(src + i);
            named_bar_sync(4, 256);
        }
        const bool rpe_skip = kBiasMode == 3 && p.rpe.dconst != nullptr;
        float ds_const_lo = 0.f, ds_const_hi = 0.f;

        // row statistics of a half tile: thread i < 64 loads L, thread 64 + i loads delta of query m0 + i
        auto load_stat = [&](int m0) -> float {
            const int m = m0 + (r & 63);
            float v = 0.f;
            if (r < 64) {
                v = -INFINITY;                                               // out-of-range query or L = -inf: P = 0
                if (m < p.M) {
                    const float Lv = __ldg(p.lse + stat_base + m);
                    if (Lv != -INFINITY) v = -Lv * kLog2e;
                }
            } else if (m < p.M) {
                v = -__ldg(p.delta + stat_base + m);
            }
            return v;
        };
        if (n_iter > 0) {
            stats[r] = load_stat(i_start * kBM + wg * kSub);                 // buffer 0: [0,64) -L*log2e, [64,128) -delta
            named_bar_sync(bar_id, 128);
        }

        for (int k = 0; k < n_iter; ++k) {
            const int mrow0 = (i_start + k) * kBM;
            const int m0 = mrow0 + wg * kSub;
            const int t = 2 * k + wg;
            const float* st = stats + (k & 1) * 128;
            float stat_next = 0.f;
            if (k + 1 < n_iter) stat_next = load_stat(m0 + kBM);

            // masks: key tail (whole row) and causal (query m sees key n iff n <= m + pseq, i.e. c >= gn - pseq - m0)
            const bool need_mask = (col0 + kBN > p.N) || (kCausal && (col0 + kBN - 1 - pseq > m0));
            int vis_lo = 0;
            if (kCausal) vis_lo = gn - pseq - m0;
            if (!key_ok) vis_lo = kSub;
            // bias mode 3: relative positions n - m of this half tile
            bool rpe_const = false;
            float rpe_cval = 0.f;
            bool const_is_lo = false;
            if (kBiasMode == 3) {
                const int rel_min = col0 - (m0 + kSub - 1);
                const int rel_max = col0 + (kBN - 1) - m0;
                const_is_lo = rel_max <= p.rpe.const_lo;
                rpe_const = const_is_lo || rel_min >= p.rpe.const_hi;
                if (rpe_const) rpe_cval = band[(const_is_lo ? p.rpe.const_lo : p.rpe.const_hi) - p.rpe.band_lo] * kLog2e;
            }

            // ---------------- P^T and dS^T of this half tile, in registers ----------------
            if (r == 0) BWD3_TS(wg, k, 0);
            mbar_wait(sdp_full + wg, k & 1);
            tc_fence_after();
            if (r == 0) BWD3_TS(wg, k, 1);
            const int bstage = t % C::kBiasStages;
            if (kBiasMode == 1) mbar_wait(b_full + bstage, (t / C::kBiasStages) & 1);
            const uint8_t* brow = smem + C::kBias + bstage * kBoxBytes + r * 128;
            uint32_t pp[2][16], dd[2][16];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t sr[32], dr[32];
                tmem_ld32(tm_s + ch * 32, sr);
                tmem_ld32(tm_dp + ch * 32, dr);
                tmem_ld_wait();
                const float* nl = st + ch * 32;
                const float* nd = st + 64 + ch * 32;
                // bias mode 3, element c of the chunk: band[(gn - (m0 + ch*32 + c)) - band_lo]
                const float* bp = band + (gn - m0 - ch * 32 - p.rpe.band_lo);
                const int vl = vis_lo - ch * 32;
                if (kBiasMode == 3 && rpe_const) {
                    float* acc = const_is_lo ? &ds_const_lo : &ds_const_hi;
                    if (rpe_skip) {
                        if (need_mask) v3_chunk<kBf16, 3, true, true, true>(sr, dr, nl, nd, nullptr, 0, rx, bp, rpe_cval, scale_log2, vl, pp[ch], dd[ch], *acc);
                        else v3_chunk<kBf16, 3, false, true, true>(sr, dr, nl, nd, nullptr, 0, rx, bp, rpe_cval, scale_log2, 0, pp[ch], dd[ch], *acc);
                    } else {
                        if (need_mask) v3_chunk<kBf16, 3, true, true, false>(sr, dr, nl, nd, nullptr, 0, rx, bp, rpe_cval, scale_log2, vl, pp[ch], dd[ch], *acc);
                        else v3_chunk<kBf16, 3, false, true, false>(sr, dr, nl, nd, nullptr, 0, rx, bp, rpe_cval, scale_log2, 0, pp[ch], dd[ch], *acc);
                    }
                } else {
                    float dummy = 0.f;
                    if (need_mask) v3_chunk<kBf16, kBiasMode, true, false, false>(sr, dr, nl, nd, brow, ch * 4, rx, bp, 0.f, scale_log2, vl, pp[ch], dd[ch], dummy);
                    else v3_chunk<kBf16, kBiasMode, false, false, false>(sr, dr, nl, nd, brow, ch * 4, rx, bp, 0.f, scale_log2, 0, pp[ch], dd[ch], dummy);
                }
            }
            if (kBiasMode == 1) {
                fence_proxy_async_smem();                    // bias reads complete before TMA refills the stage
                __syncwarp();
                if (lane == 0) mbar_arrive(b_empty + bstage);
            }
            if (r == 0) BWD3_TS(wg, k, 2);

            // ---------------- the dS^T box of this warpgroup is free again ----------------
            // (last readers: the dQ MMAs of tile k - 1 and the TMA reduce of this warpgroup's previous box)
            if (k > 0) mbar_wait(dq_full, (k - 1) & 1);
            if (r == 0) bulk_wait_group_read<0>();
            named_bar_sync(bar_id, 128);
            if (r == 0) BWD3_TS(wg, k, 3);

            // ---------------- P^T, dS^T -> TMEM (A operands of dV, dK); dS^T -> shared memory (dQ, dBias) ----------------
            tmem_st16(tm_s + 0, pp[0]);
            tmem_st16(tm_s + 16, pp[1]);
            tmem_st16(tm_dp + 0, dd[0]);
            tmem_st16(tm_dp + 16, dd[1]);
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const int off = ((ch * 4 + c8) ^ rx) << 4;
                    *reinterpret_cast<uint4*>(sDS + off) =
                        make_uint4(dd[ch][c8 * 4], dd[ch][c8 * 4 + 1], dd[ch][c8 * 4 + 2], dd[ch][c8 * 4 + 3]);
                }
            }
            if (k + 1 < n_iter) stats[((k + 1) & 1) * 128 + r] = stat_next;
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar_sync(bar_id, 128);
            if (r == 0) BWD3_TS(wg, k, 4);
            if (r == 0) {
                mbar_arrive(pds_full + wg);
                if (kBiasMode != 0 && !(kBiasMode == 3 && rpe_skip && rpe_const)) {
                    if (p.ds_use_reduce) tma_reduce_add_4d(&p.map_ds, smem + C::kDS + wg * kBoxBytes, m0, col0, h, g_ds);
                    else tma_store_4d(&p.map_ds, smem + C::kDS + wg * kBoxBytes, m0, col0, h, g_ds);
                    bulk_commit_group();
                }
            }
        }

        // ---- tail: dV (warpgroup 0) and dK * sm_scale (warpgroup 1) once every MMA of the CTA has completed ----
        if (kBiasMode == 3 && rpe_skip) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ds_const_lo += __shfl_xor_sync(0xffffffffu, ds_const_lo, off);
                ds_const_hi += __shfl_xor_sync(0xffffffffu, ds_const_hi, off);
            }
            // (the stats buffers are free: every thread of the warpgroup passed the last barrier of the loop)
            float* red = stats;
            named_bar_sync(bar_id, 128);
            if (lane == 0) {
                red[(warp & 3) * 2 + 0] = ds_const_lo;
                red[(warp & 3) * 2 + 1] = ds_const_hi;
            }
            named_bar_sync(bar_id, 128);
            if (r < 2) {
                const float tsum = red[r] + red[2 + r] + red[4 + r] + red[6 + r];
                if (tsum != 0.f) atomicAdd(p.rpe.dconst + h * 2 + r, tsum);
            }
        }
        {
            uint8_t* out_row = wg == 0
                ? reinterpret_cast<uint8_t*>(p.dv) + 2 * ((int64_t)b * p.dv_sb + (int64_t)h * p.dv_sh + (int64_t)gn * p.dv_sn)
                : reinterpret_cast<uint8_t*>(p.dk) + 2 * ((int64_t)b * p.dk_sb + (int64_t)h * p.dk_sh + (int64_t)gn * p.dk_sn);
            const float sc = wg == 0 ? 1.f : p.sm_scale;
            if (n_iter > 0) {
                mbar_wait(dq_full, (n_iter - 1) & 1);            // every MMA of this CTA has completed
                tc_fence_after();
                const uint32_t tm_acc = tmem_base + lane_off + (wg == 0 ? C::kColDV : C::kColDK);
                constexpr int kChunk = kD >= 32 ? 32 : 16;
#pragma unroll
                for (int c0 = 0; c0 < kD; c0 += kChunk) {
                    uint32_t a[kChunk];
                    tmem_ld_n<kChunk>(tm_acc + c0, a);
                    tmem_ld_wait();
                    if (key_ok) {
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
            } else if (key_ok) {
#pragma unroll
                for (int c = 0; c < kD; c += 8) *reinterpret_cast<uint4*>(out_row + 2 * c) = make_uint4(0, 0, 0, 0);
            }
        }
        if (r == 0) bulk_wait_group<0>();     // all TMA stores / reductions of this warpgroup have landed
    }

    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kBiasMode, bool kCausal>
static cudaError_t launch_bwd3_inst(const AttnBwdKernelParams& kp, cudaStream_t stream) {
    using C = Bwd3Cfg<kD>;
    auto kern = attn_bwd_kernel_v3<kD, kBf16, kBiasMode, kCausal>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
    if (e != cudaSuccess) return e;
    const int grid = kp.B * kp.H * kp.num_n_blocks;
    kern<<<grid, 512, C::kTotal, stream>>>(kp);
    count_launch();
#ifdef B200T5_BWD_TIMING
    {
        cudaDeviceSynchronize();
        static long long ts[4][24][8];
        cudaMemcpyFromSymbol(ts, g_bwd3_ts, sizeof(ts));
        const long long t0 = ts[0][0][0];
        const char* names[4] = {"wg0 [wait S, S ready, math done, box free, stored]", "wg1", "mma [wait P/dS(t), P/dS ready, dq gate, issued]  (per half tile)",
                                "drain [wait dQ, dQ ready, reduce issued]"};
        for (int role = 0; role < 4; ++role) {
            printf("BWD3_TIMING %s\n", names[role]);
            for (int k = 0; k < (role == 2 ? 16 : 8); ++k) {
                printf("  %2d:", k);
                for (int j = 0; j < 5; ++j) printf(" %7lld", ts[role][k][j] ? ts[role][k][j] - t0 : 0);
                printf("\n");
            }
        }
        fflush(stdout);
    }
#endif
    return cudaGetLastError();
}

template <int kD, bool kBf16>
static cudaError_t launch_bwd3_d(const AttnBwdKernelParams& kp, int bias_mode, bool causal, cudaStream_t stream) {
    switch (bias_mode * 2 + (causal ? 1 : 0)) {
        case 0: return launch_bwd3_inst<kD, kBf16, 0, false>(kp, stream);
        case 1: return launch_bwd3_inst<kD, kBf16, 0, true>(kp, stream);
        case 2: return launch_bwd3_inst<kD, kBf16, 1, false>(kp, stream);
        case 3: return launch_bwd3_inst<kD, kBf16, 1, true>(kp, stream);
        case 6: return launch_bwd3_inst<kD, kBf16, 3, false>(kp, stream);
        case 7: return launch_bwd3_inst<kD, kBf16, 3, true>(kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_attn_bwd_v3(const AttnBwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                               cudaStream_t stream) {
#ifdef B200T5_HEADLINE_ONLY
    if (D == 64 && bf16) return launch_bwd3_d<64, true>(kp, bias_mode, causal, stream);
    return cudaErrorInvalidValue;
#else
    switch (D) {
        case 16: return bf16 ? launch_bwd3_d<16, true>(kp, bias_mode, causal, stream) : launch_bwd3_d<16, false>(kp, bias_mode, causal, stream);
        case 32: return bf16 ? launch_bwd3_d<32, true>(kp, bias_mode, causal, stream) : launch_bwd3_d<32, false>(kp, bias_mode, causal, stream);
        case 64: return bf16 ? launch_bwd3_d<64, true>(kp, bias_mode, causal, stream) : launch_bwd3_d<64, false>(kp, bias_mode, causal, stream);
        default: return cudaErrorInvalidValue;
    }
#endif
}

}  // namespace b200t5
