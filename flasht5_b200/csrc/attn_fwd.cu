// FlashAttention-2 forward with additive (T5) bias for sm_100a.
//
// Replaces /root/reference/src/model/ops/flash_attention_v2_bias.py:327-483 (`_fwd_kernel`).
//
// One CTA = one (batch, head, 128-row query block); it walks the key/value sequence in 128-wide
// tiles.  Two CTAs are resident per SM (D <= 64) so one CTA's softmax overlaps the other's MMAs.
//
//   warp 4 (1 lane)  : TMA producer for Q (once) and the K / V ring (2 stages each)
//   warp 6 (1 lane)  : TMA producer for the bias ring (2 stages of 128 rows x 64 columns)
//   warp 5 (1 lane)  : tcgen05.mma issuer   S = Q K^T (SS)  and  O += P V (A = P from TMEM)
//   warps 0-3        : one thread per query row: tcgen05.ld S, bias add, online softmax (lazy
//                      rescale of O in TMEM), P -> TMEM as packed 16-bit, epilogue O / l and LSE
//
// TMEM columns: S [0,128) fp32 | O [128,128+D) fp32 | P [128+D, 128+D+64) packed 16-bit pairs.
//
// Bias modes: 0 none | 1 dense bias through TMA | 2 dense bias through pointers (rows not 16-byte aligned) |
// 3 T5 relative-position bias computed in the kernel (the reference's `fa2_rpe` surface,
//   /root/reference/src/model/modeling_flash_t5.py:275-279): the per-head band of bias values over relative
//   positions sits in shared memory (where modes 1/2 keep the bias ring); tiles whose relative positions are all
//   beyond the last distinct bucket add one scalar, the others read band[n - m] -- no (H, M, N) tensor exists.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kBM = 128;   // query rows per CTA
constexpr int kBN = 128;   // keys per tile
constexpr int kKVStages = 2;
constexpr int kBiasStages = 2;
constexpr int kBiasHalfBytes = kBM * 64 * 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// keep a stale reference max while exp(x - m) stays small enough for the io dtype of P: <= 2^8 for fp16 (max 65 504, and the
// P V accumulation adds 128 of them), <= 2^16 for bf16 (fp32 exponent range; the fp32 accumulators are nowhere near overflow).
// A larger bound means fewer O rescales in TMEM (~300 cycles per tile whenever one row of a warp needs it).
template <bool kBf16> constexpr float kRescaleThreshold = (kBf16 ? 16.0f : 8.0f) * kLn2;

template <int kD>
struct FwdSmem {
    static constexpr int kRowBytes = (kD >= 64 ? 64 : kD) * 2;   // bytes per smem row inside one swizzle box
    static constexpr int kBoxes = kD >= 64 ? kD / 64 : 1;        // 64-column TMA boxes per operand row
    static constexpr int kTileBytes = kBM * kD * 2;              // Q, K or V tile
    static constexpr int kBoxBytes = kBM * kRowBytes;
    static constexpr int kQ = 0;
    static constexpr int kK = kQ + kTileBytes;
    static constexpr int kV = kK + kKVStages * kTileBytes;
    static constexpr int kBias = kV + kKVStages * kTileBytes;
    static constexpr int kBars = kBias + kBiasStages * kBiasHalfBytes;
    static constexpr int kNumBars = 1 + 4 * kKVStages + 2 * kBiasStages + 4;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static constexpr uint32_t kSwizzle = kRowBytes == 128 ? kSwz128 : (kRowBytes == 64 ? kSwz64 : kSwz32);
    static constexpr int kTmemCols = (128 + kD + 64) <= 256 ? 256 : 512;
    static constexpr int kOCol = 128;
    static constexpr int kPCol = 128 + kD;
};

struct FwdBars {
    uint64_t* q_full;
    uint64_t* k_full;    // [kKVStages]
    uint64_t* k_empty;
    uint64_t* v_full;
    uint64_t* v_empty;
    uint64_t* b_full;    // [kBiasStages]
    uint64_t* b_empty;
    uint64_t* s_full;
    uint64_t* s_empty;
    uint64_t* p_full;
    uint64_t* pv_done;
};

}  // namespace

// Developer build (-DB200T5_FWD_TIMING): every CTA that lands on SM `kTimedSm` stamps clock64 at the phase
// boundaries of softmax thread 0 and of the MMA warp; the launcher prints the timeline.  Off in the product build.
#ifdef B200T5_FWD_TIMING
constexpr int kTimedSm = 17;
__device__ int g_fwd_ts_slots;
__device__ long long g_fwd_ts[32][2][12][8];   // [slot][role][tile][stamp]
__device__ int g_fwd_ts_bid[32];
__device__ __forceinline__ long long clk64() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#define FWD_TS(role, j, slot_)                                                              \
    do {                                                                                    \
        if (ts_slot >= 0 && (j) < 12) g_fwd_ts[ts_slot][role][j][slot_] = clk64();          \
    } while (0)
#else
#define FWD_TS(role, j, slot_) do { } while (0)
#endif

template <int kD, bool kBf16, int kBiasMode, bool kCausal>
__global__ void __launch_bounds__(256, 2)   // 128 regs at launch; setmaxnreg re-splits them per role
attn_fwd_kernel(const __grid_constant__ AttnFwdKernelParams p) {
    using L = FwdSmem<kD>;
    extern __shared__ __align__(1024) uint8_t smem[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- work decode: batch fastest so the CTAs sharing a bias tile run together (L2 reuse) ----
    const int nmb = p.num_m_blocks;
    int bid = blockIdx.x;
    const int b = bid % p.B;
    bid /= p.B;
    const int mb = nmb - 1 - (bid % nmb);     // heavy (late) causal blocks first
    const int h = bid / nmb;
    const int row0 = mb * kBM;
    const int pseq = p.N - p.M;

    int num_tiles = (p.N + kBN - 1) / kBN;
    if (kCausal) {
        const int last_col = row0 + kBM - 1 + pseq;          // last visible key of the last row
        const int t = last_col < 0 ? 0 : last_col / kBN + 1;
        num_tiles = t < num_tiles ? t : num_tiles;
    }

    FwdBars bars;
    {
        uint64_t* bb = reinterpret_cast<uint64_t*>(smem + L::kBars);
        bars.q_full = bb;
        bars.k_full = bb + 1;
        bars.k_empty = bars.k_full + kKVStages;
        bars.v_full = bars.k_empty + kKVStages;
        bars.v_empty = bars.v_full + kKVStages;
        bars.b_full = bars.v_empty + kKVStages;
        bars.b_empty = bars.b_full + kBiasStages;
        bars.s_full = bars.b_empty + kBiasStages;
        bars.s_empty = bars.s_full + 1;
        bars.p_full = bars.s_empty + 1;
        bars.pv_done = bars.p_full + 1;
    }
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    if (warp == 4 && lane == 0) {                 // the producer lane initialises the barriers: its first loads leave at once
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("b200t5: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        mbar_init(bars.q_full, 1);
        for (int i = 0; i < kKVStages; ++i) {
            mbar_init(bars.k_full + i, 1);
            mbar_init(bars.k_empty + i, 1);
            mbar_init(bars.v_full + i, 1);
            mbar_init(bars.v_empty + i, 1);
        }
        for (int i = 0; i < kBiasStages; ++i) {
            mbar_init(bars.b_full + i, 1);
            mbar_init(bars.b_empty + i, 4);
        }
        mbar_init(bars.s_full, 1);
        mbar_init(bars.s_empty, 4);
        mbar_init(bars.p_full, 4);
        mbar_init(bars.pv_done, 1);
        fence_mbar_init();
        fence_proxy_async_smem();                 // barrier words: written through the generic proxy, completed through the async one
        tma_prefetch_desc(&p.map_q);
        tma_prefetch_desc(&p.map_k);
        tma_prefetch_desc(&p.map_v);
        if (kBiasMode == 1) tma_prefetch_desc(&p.map_bias);
        if (num_tiles > 0) {
            // Q and the first K tile are on the critical path of the prologue (first S ready ~2 900 cycles after launch): they
            // leave before the CTA-wide setup barrier, while warp 5 allocates TMEM
            mbar_arrive_expect_tx(bars.q_full, L::kTileBytes);
#pragma unroll
            for (int bx = 0; bx < L::kBoxes; ++bx)
                tma_load_4d(smem + L::kQ + bx * L::kBoxBytes, &p.map_q, bars.q_full, bx * 64, row0, h, b);
            mbar_arrive_expect_tx(bars.k_full + 0, L::kTileBytes);
#pragma unroll
            for (int bx = 0; bx < L::kBoxes; ++bx)
                tma_load_4d(smem + L::kK + bx * L::kBoxBytes, &p.map_k, bars.k_full + 0, bx * 64, 0, h, b);
        }
    }
    if (warp == 5) tmem_alloc<L::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
#ifdef B200T5_FWD_TIMING
    int ts_slot = -1;
    {
        // (no static shared memory: 4 static bytes pad to 1 KB in front of the 1024-aligned dynamic block and cost
        //  the second resident CTA; the TMEM slot has spare words)
        volatile int& s_ts_slot = *reinterpret_cast<volatile int*>(smem + L::kTmemSlot + 8);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (threadIdx.x == 0) {
            s_ts_slot = -1;
            if (smid == kTimedSm) {
                const int sl = atomicAdd(&g_fwd_ts_slots, 1);
                if (sl < 32) {
                    s_ts_slot = sl;
                    g_fwd_ts_bid[sl] = blockIdx.x;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 || (warp == 5 && lane == 0)) ts_slot = s_ts_slot;
    }
#endif

    if (warp >= 4) {
        // =============================== control warps ===============================
        setmaxnreg_dec<48>();
        if (warp == 4 && lane == 0 && num_tiles > 0) {
            // ---- Q / K / V producer (Q and K tile 0 were requested before the setup barrier) ----
            for (int j = 0; j < num_tiles; ++j) {
                const int s = j % kKVStages;
                const uint32_t par = ((j / kKVStages) & 1) ^ 1;
                if (j > 0) {
                    mbar_wait_producer(bars.k_empty + s, par);
                    mbar_arrive_expect_tx(bars.k_full + s, L::kTileBytes);
#pragma unroll
                    for (int bx = 0; bx < L::kBoxes; ++bx)
                        tma_load_4d(smem + L::kK + s * L::kTileBytes + bx * L::kBoxBytes, &p.map_k, bars.k_full + s,
                                    bx * 64, j * kBN, h, b);
                }
                mbar_wait_producer(bars.v_empty + s, par);
                mbar_arrive_expect_tx(bars.v_full + s, L::kTileBytes);
#pragma unroll
                for (int bx = 0; bx < L::kBoxes; ++bx)
                    tma_load_4d(smem + L::kV + s * L::kTileBytes + bx * L::kBoxBytes, &p.map_v, bars.v_full + s,
                                bx * 64, j * kBN, h, b);
            }
        } else if (warp == 6 && lane == 0 && kBiasMode == 1) {
            // ---- bias producer: two 64-column halves per tile ----
            const int hb = p.bias_h_bcast ? 0 : h;
            const int bb = p.bias_b_bcast ? 0 : b;
            for (int it = 0; it < 2 * num_tiles; ++it) {
                const int s = it % kBiasStages;
                const uint32_t par = ((it / kBiasStages) & 1) ^ 1;
                mbar_wait_producer(bars.b_empty + s, par);
                mbar_arrive_expect_tx(bars.b_full + s, kBiasHalfBytes);
                tma_load_4d(smem + L::kBias + s * kBiasHalfBytes, &p.map_bias, bars.b_full + s,
                            (it >> 1) * kBN + (it & 1) * 64, row0, hb, bb);
            }
        } else if (warp == 5 && num_tiles > 0) {
            // ---- MMA issuer: the whole warp runs the loop (descriptor arithmetic stays warp-uniform), one
            //      elected lane issues the tcgen05 instructions ----
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, kBM, kBN, false, false);
            constexpr uint32_t idesc_pv = make_idesc(kBf16, kBM, kD, false, true);
            constexpr uint32_t sbo = 8 * L::kRowBytes;
            constexpr uint32_t hi_k = sdesc_hi(sbo, L::kSwizzle);          // Q, K (K-major) and V (MN-major) share it
            const uint32_t q_lo = sdesc_lo(smem_u32(smem + L::kQ), 16);
            const uint32_t k_lo0 = sdesc_lo(smem_u32(smem + L::kK), 16);
            const uint32_t v_lo0 = sdesc_lo(smem_u32(smem + L::kV), L::kBoxBytes);
            const uint32_t tm_s = tmem_base;
            const uint32_t tm_o = tmem_base + L::kOCol;
            const uint32_t tm_p = tmem_base + L::kPCol;

            auto issue_s = [&](int j) {
                const int s = j % kKVStages;
                const uint32_t k_lo = k_lo0 + s * (L::kTileBytes >> 4);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk) {
                        // K-major operands: 16 elements of K per MMA = 32 bytes inside a swizzled row
                        const uint32_t off = ((kk / 4) * L::kBoxBytes + (kk % 4) * 32) >> 4;
                        umma_ss2(tm_s, q_lo + off, hi_k, k_lo + off, hi_k, idesc_s, kk > 0 ? 1u : 0u);
                    }
                    umma_commit(bars.s_full);
                    umma_commit(bars.k_empty + s);
                }
                __syncwarp();
            };

            mbar_wait(bars.q_full, 0);
            mbar_wait(bars.k_full + 0, 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < num_tiles; ++j) {
                if (j + 1 < num_tiles) {
                    const int jn = j + 1;
                    mbar_wait(bars.k_full + (jn % kKVStages), (jn / kKVStages) & 1);
                    mbar_wait(bars.s_empty, j & 1);
                    tc_fence_after();
                    FWD_TS(1, jn, 0);
                    issue_s(jn);
                    FWD_TS(1, jn, 1);
                }
                const int s = j % kKVStages;
                mbar_wait(bars.v_full + s, (j / kKVStages) & 1);
                mbar_wait(bars.p_full, j & 1);
                tc_fence_after();
                FWD_TS(1, j, 2);
                const uint32_t v_lo = v_lo0 + s * (L::kTileBytes >> 4);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk) {
                        // B = V, MN-major: 16 key rows per MMA; LBO = stride between 64-wide d chunks
                        umma_ts2(tm_o, tm_p + kk * 8, v_lo + ((kk * 16 * L::kRowBytes) >> 4), hi_k, idesc_pv,
                                 (j > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(bars.pv_done);
                    umma_commit(bars.v_empty + s);
                }
                __syncwarp();
                FWD_TS(1, j, 3);
            }
        }
    } else {
        // =============================== softmax warpgroup ===============================
        setmaxnreg_inc<208>();
        const int r = threadIdx.x;                 // row inside the tile == TMEM lane
        const int grow = row0 + r;                 // global query row
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tm_s = tmem_base + lane_off;
        const uint32_t tm_o = tmem_base + lane_off + L::kOCol;
        const uint32_t tm_p = tmem_base + lane_off + L::kPCol;

        float m_ref = -INFINITY;   // running reference max (natural units, scaled + biased scores)
        float l_sum = 0.f;         // sum of exp(x - m_ref), unrounded: L = m_ref + ln(l_sum)
        float l_hat = 0.f;         // sum of the same values after rounding to the io dtype: normalises O

        const float* band = reinterpret_cast<const float*>(smem + L::kBias);   // [bias mode 3]
        if (kBiasMode == 3) {
            float* dst = reinterpret_cast<float*>(smem + L::kBias);
            const float* src = p.rpe.band + (int64_t)h * p.rpe.band_len;
            for (int i = threadIdx.x; i < p.rpe.band_len; i += 128) dst[i] = __ldg(src + i);
            named_bar_sync(1, 128);
        }

        const uint16_t* bias_row = nullptr;
        if (kBiasMode == 2) {
            bias_row = reinterpret_cast<const uint16_t*>(p.bias) + (p.bias_b_bcast ? 0 : (int64_t)b * p.bias_sb) +
                       (p.bias_h_bcast ? 0 : (int64_t)h * p.bias_sh) + (int64_t)grow * p.bias_sm;
        }

        const f32x2 scale2 = f2_pack(p.sm_scale, p.sm_scale);
        const f32x2 l2e2 = f2_pack(kLog2e, kLog2e);
        const bool unit_scale = p.sm_scale == 1.0f;
        for (int j = 0; j < num_tiles; ++j) {
            const int col0 = j * kBN;
            float x[kBN];

            FWD_TS(0, j, 0);
            mbar_wait(bars.s_full, j & 1);
            tc_fence_after();
            FWD_TS(0, j, 1);
            {
                uint32_t(&xr)[kBN] = reinterpret_cast<uint32_t(&)[kBN]>(x);
                tmem_ld32(tm_s + 0, reinterpret_cast<uint32_t(&)[32]>(xr[0]));
                tmem_ld32(tm_s + 32, reinterpret_cast<uint32_t(&)[32]>(xr[32]));
                tmem_ld32(tm_s + 64, reinterpret_cast<uint32_t(&)[32]>(xr[64]));
                tmem_ld32(tm_s + 96, reinterpret_cast<uint32_t(&)[32]>(xr[96]));
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars.s_empty);
            FWD_TS(0, j, 2);

            // ---- scores = S * sm_scale + bias ----
            if (kBiasMode == 1) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int it = 2 * j + hh;
                    const int s = it % kBiasStages;
                    mbar_wait(bars.b_full + s, (it / kBiasStages) & 1);
                    const uint8_t* brow = smem + L::kBias + s * kBiasHalfBytes + r * 128;
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        const uint4 u = *reinterpret_cast<const uint4*>(brow + ((c8 ^ (r & 7)) << 4));
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                        if (unit_scale) {
                            // sm_scale == 1 (T5): S + bias straight from the packed 16-bit pair, one mixed-precision add per element
                            // (same single rounding as fma(S, 1, bias))
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int c = hh * 64 + c8 * 8 + e * 2;
                                add_f32_16x2<kBf16>(w[e], x[c], x[c + 1], x[c], x[c + 1]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = unpack2<kBf16>(w[e]);
                                const int c = hh * 64 + c8 * 8 + e * 2;
                                f2_unpack(f2_fma(f2_pack(x[c], x[c + 1]), scale2, f2_pack(f.x, f.y)), x[c], x[c + 1]);   // one FFMA2
                            }
                        }
                    }
                }
                // WAR across proxies: these generic-proxy reads must have completed before the TMA (async proxy) may
                // overwrite the stages.  Without the fence ~0.1% of rows picked up bytes of the NEXT tile when two CTAs
                // shared an SM (measured on B200).  One fence for both halves of the tile: it costs ~100 cycles of latency.
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bars.b_empty + (2 * j) % kBiasStages);
                    mbar_arrive(bars.b_empty + (2 * j + 1) % kBiasStages);
                }
            } else if (kBiasMode == 2) {
#pragma unroll
                for (int c = 0; c < kBN; ++c) {
                    float bv = 0.f;
                    if (grow < p.M && col0 + c < p.N)
                        bv = to_float16bit<kBf16>(__ldg(bias_row + (int64_t)(col0 + c) * p.bias_sn));
                    x[c] = fmaf(x[c], p.sm_scale, bv);
                }
            } else if (kBiasMode == 3) {
                // relative positions n - m of this tile: [col0 - row0 - 127, col0 - row0 + 127]
                const int rel_min = col0 - row0 - (kBM - 1);
                const int rel_max = col0 - row0 + (kBN - 1);
                if (rel_max <= p.rpe.const_lo || rel_min >= p.rpe.const_hi) {
                    const float bv = band[(rel_max <= p.rpe.const_lo ? p.rpe.const_lo : p.rpe.const_hi) - p.rpe.band_lo];
#pragma unroll
                    for (int c = 0; c < kBN; c += 2) f2_unpack(f2_fma(f2_pack(x[c], x[c + 1]), scale2, f2_pack(bv, bv)), x[c], x[c + 1]);
                } else {
                    const float* bp = band + (col0 - grow - p.rpe.band_lo);
#pragma unroll
                    for (int c = 0; c < kBN; c += 2) f2_unpack(f2_fma(f2_pack(x[c], x[c + 1]), scale2, f2_pack(bp[c], bp[c + 1])), x[c], x[c + 1]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < kBN; c += 2) f2_unpack(f2_mul(f2_pack(x[c], x[c + 1]), scale2), x[c], x[c + 1]);
            }

            FWD_TS(0, j, 3);
            // ---- masks: key tail and (bottom-right aligned) causal ----
            {
                int lim = p.N - col0;
                if (kCausal) {
                    const int cl = grow + pseq + 1 - col0;
                    lim = cl < lim ? cl : lim;
                }
                const bool need_mask = (col0 + kBN > p.N) || (kCausal && (col0 + kBN - 1 > row0 + pseq));
                if (need_mask) {
#pragma unroll
                    for (int c = 0; c < kBN; ++c)
                        if (c >= lim) x[c] = -INFINITY;
                }
            }

            // ---- online softmax with lazy rescale ----
            float t0 = x[0], t1 = x[1], t2 = x[2], t3 = x[3];
#pragma unroll
            for (int c = 4; c < kBN; c += 4) {
                t0 = fmaxf(t0, x[c]);
                t1 = fmaxf(t1, x[c + 1]);
                t2 = fmaxf(t2, x[c + 2]);
                t3 = fmaxf(t3, x[c + 3]);
            }
            const float tmax = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
            float alpha = 1.f;
            if (tmax > m_ref + kRescaleThreshold<kBf16>) {       // also true for the first finite tile (m_ref = -inf)
                alpha = __expf(m_ref - tmax);              // exp(-inf) = 0 on the first tile
                m_ref = tmax;
            }
            // Rows whose running max is the dtype's most negative value (the reference model folds padding masks into the bias
            // as finfo.min, modeling_flash_t5.py:266-270): m * log2e would overflow, so such rows subtract first, as the
            // reference does (exp2((s - m) * log2e), flash_attention_v2_bias.py:453-454).  Never taken for ordinary rows.
            const bool huge_m = m_ref < -1e37f && m_ref != -INFINITY;
            if (__any_sync(0xffffffffu, huge_m)) {
                const float sub = huge_m ? m_ref : 0.f;
#pragma unroll
                for (int c = 0; c < kBN; ++c) x[c] -= sub;
            }
            const float m_safe = (m_ref == -INFINITY || huge_m) ? 0.f : m_ref;
            const float neg_m_log2 = -m_safe * kLog2e;
            FWD_TS(0, j, 4);

            // P = exp2(.) rounded to the io dtype feeds the P V contraction; O is normalised by the sum of the ROUNDED values
            // (l_hat), so the rounding of a dominant P cancels in the ratio, while L = m + ln(sum of the unrounded values)
            // stays exact for the backward.  (Measured against the reference Triton kernel, profiles/r2a_triton_parity_*:
            // with the stale reference max of the lazy rescale a dominant P is no longer exactly 1, and normalising by the
            // exact sum left O at 1.96e-3 relative error where Triton has 1.66e-3.)
            uint32_t pk[kBN / 2];
            float s0, s1, r0 = 0.f, r1 = 0.f;
            const f32x2 negm2 = f2_pack(neg_m_log2, neg_m_log2);
            f32x2 sum2 = f2_pack(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < kBN; c += 2) {
                float a0, a1;
                f2_unpack(f2_fma(f2_pack(x[c], x[c + 1]), l2e2, negm2), a0, a1);      // FFMA2 / FADD2: two columns per issue slot
                const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
                sum2 = f2_add(sum2, f2_pack(e0, e1));
                pk[c / 2] = pack2<kBf16>(e0, e1);
                add_f32_16x2<kBf16>(pk[c / 2], r0, r1, r0, r1);
            }
            f2_unpack(sum2, s0, s1);
            l_sum = l_sum * alpha + (s0 + s1);
            l_hat = l_hat * alpha + (r0 + r1);
            FWD_TS(0, j, 5);

            if (j > 0) {
                mbar_wait(bars.pv_done, (j - 1) & 1);     // O and the P buffer are free again
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
                    for (int c0 = 0; c0 < kD; c0 += 32) {
                        if constexpr (kD >= 32) {
                            uint32_t o[32];
                            tmem_ld32(tm_o + c0, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st32(tm_o + c0, o);
                        } else {
                            uint32_t o[16];
                            tmem_ld16(tm_o + c0, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st16(tm_o + c0, o);
                        }
                    }
                }
            }
            FWD_TS(0, j, 6);
            tmem_st32(tm_p + 0, reinterpret_cast<const uint32_t(&)[32]>(pk[0]));
            tmem_st32(tm_p + 32, reinterpret_cast<const uint32_t(&)[32]>(pk[32]));
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars.p_full);
            FWD_TS(0, j, 7);
        }

        // ---- epilogue: O / l -> global (row-contiguous 16-byte stores), LSE ----
        const bool row_ok = grow < p.M;
        uint8_t* o_row = reinterpret_cast<uint8_t*>(p.o) +
                         2 * ((int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)grow * p.o_sm);
        if (num_tiles > 0) {
            mbar_wait(bars.pv_done, (num_tiles - 1) & 1);
            tc_fence_after();
            const float inv_l = l_hat > 0.f ? 1.f / l_hat : 0.f;
            constexpr int kChunk = kD >= 32 ? 32 : 16;
#pragma unroll
            for (int c0 = 0; c0 < kD; c0 += kChunk) {
                uint32_t o[kChunk];
                if constexpr (kChunk == 32) tmem_ld32(tm_o + c0, o);
                else tmem_ld16(tm_o + c0, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                tmem_ld_wait();
                if constexpr (kD == 64) {
                    // D = 64: the tile leaves through shared memory (Q's tile: every S MMA is long done) and ONE TMA store.  A
                    // thread = row store to global memory costs an L1 request per 16 bytes: 1 024 requests per CTA.  Rows beyond
                    // M are clipped by the tensor map.
#pragma unroll
                    for (int i = 0; i < kChunk; i += 8) {
                        uint4 out;
                        out.x = pack2<kBf16>(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
                        out.y = pack2<kBf16>(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
                        out.z = pack2<kBf16>(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
                        out.w = pack2<kBf16>(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(smem + L::kQ + r * 128 + ((((c0 + i) / 8) ^ (r & 7)) << 4)) = out;
                    }
                } else if (row_ok) {
#pragma unroll
                    for (int i = 0; i < kChunk; i += 8) {
                        uint4 out;
                        out.x = pack2<kBf16>(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
                        out.y = pack2<kBf16>(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
                        out.z = pack2<kBf16>(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
                        out.w = pack2<kBf16>(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(o_row + 2 * (c0 + i)) = out;
                    }
                }
            }
            tc_fence_before();
            if constexpr (kD == 64) {
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (r == 0) {
                    tma_store_4d(&p.map_o, smem + L::kQ, 0, row0, h, b);
                    bulk_commit_group();
                    bulk_wait_group_read<0>();                     // shared memory must outlive the read; the write completes by itself
                }
            }
        } else if (row_ok) {
            // every key is masked for this whole block (causal, M > N): O = 0, L = -inf
#pragma unroll
            for (int c = 0; c < kD; c += 8) *reinterpret_cast<uint4*>(o_row + 2 * c) = make_uint4(0, 0, 0, 0);
        }
        if (row_ok) {
            const float lse = (l_sum > 0.f) ? (m_ref + __logf(l_sum)) : -INFINITY;
            p.lse[((int64_t)b * p.H + h) * p.M + grow] = lse;
        }
    }

    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc<L::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// Bias rows that a tensor map cannot address (N % 8 != 0, odd strides, unaligned base -- e.g. the reference test's
// N = 1045): copied once into a workspace with rows padded to a multiple of 8 elements, so that the kernels take the
// TMA path.  The per-element pointer path they would use otherwise reads 128 strided 2-byte values per thread and tile
// (measured on B200: forward 143 us vs 76 us for the reference's Triton kernel at (2, 4, 1024, 1045, 64)).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bias_align_copy_kernel(const uint16_t* __restrict__ src, int64_t s_b, int64_t s_h, int64_t s_m,
                                                              int64_t s_n, uint16_t* __restrict__ dst, int Hb, int M, int N, int pitch,
                                                              int chunks) {
    const int64_t row = blockIdx.x / chunks;               // (bb * Hb + hh) * M + m
    const int n = static_cast<int>(blockIdx.x % chunks) * 256 + threadIdx.x;
    if (n >= pitch) return;
    const int m = static_cast<int>(row % M);
    const int64_t bh = row / M;
    const int hh = static_cast<int>(bh % Hb);
    const int64_t bb = bh / Hb;
    dst[row * pitch + n] = n < N ? __ldg(src + bb * s_b + hh * s_h + (int64_t)m * s_m + (int64_t)n * s_n) : uint16_t(0);
}

cudaError_t launch_bias_align_copy(const void* bias, const int64_t* s, void* dst, int Bb, int Hb, int M, int N, int pitch,
                                   cudaStream_t stream) {
    const int chunks = (pitch + 255) / 256;
    const int64_t blocks = (int64_t)Bb * Hb * M * chunks;
    if (blocks > 0x7FFFFFFFLL) return cudaErrorInvalidValue;
    bias_align_copy_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const uint16_t*>(bias), s[0], s[1], s[2], s[3],
                                                                              static_cast<uint16_t*>(dst), Hb, M, N, pitch, chunks);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kBiasMode, bool kCausal>
static cudaError_t launch_fwd_inst(const AttnFwdKernelParams& kp, cudaStream_t stream) {
    using L = FwdSmem<kD>;
    auto kern = attn_fwd_kernel<kD, kBf16, kBiasMode, kCausal>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    const int grid = kp.B * kp.H * kp.num_m_blocks;
#ifdef B200T5_FWD_TIMING
    {
        int zero = 0;
        cudaMemcpyToSymbol(g_fwd_ts_slots, &zero, sizeof(int));
    }
#endif
    kern<<<grid, 256, L::kTotal, stream>>>(kp);
    count_launch();
#ifdef B200T5_FWD_TIMING
    {
        cudaDeviceSynchronize();
        static long long ts[32][2][12][8];
        int bids[32], n = 0;
        cudaMemcpyFromSymbol(ts, g_fwd_ts, sizeof(ts));
        cudaMemcpyFromSymbol(bids, g_fwd_ts_bid, sizeof(bids));
        cudaMemcpyFromSymbol(&n, g_fwd_ts_slots, sizeof(int));
        if (n > 32) n = 32;
        const int tiles = (kp.N + kBN - 1) / kBN < 12 ? (kp.N + kBN - 1) / kBN : 12;
        long long t0 = n > 0 ? ts[0][0][0][0] : 0;
        for (int s = 0; s < n; ++s)
            if (ts[s][0][0][0] < t0) t0 = ts[s][0][0][0];
        printf("FWD_TIMING sm %d: %d CTAs; softmax thread 0: [wait S, S ready, S in regs, bias added, max done, exp done, PV(j-1) done, P stored]; "
               "mma: [S(j) issue start, S(j) issued, P(j) ready, PV(j) issued]\n", kTimedSm, n);
        for (int s = 0; s < n; ++s) {
            printf(" CTA slot %d (block %d)\n", s, bids[s]);
            for (int j = 0; j < tiles; ++j) {
                printf("  j=%d sm:", j);
                for (int q = 0; q < 8; ++q) printf(" %7lld", ts[s][0][j][q] - t0);
                printf("   mma:");
                for (int q = 0; q < 4; ++q) printf(" %7lld", ts[s][1][j][q] - t0);
                printf("\n");
            }
        }
        fflush(stdout);
    }
#endif
    return cudaGetLastError();
}

template <int kD, bool kBf16>
static cudaError_t launch_fwd_d(const AttnFwdKernelParams& kp, int bias_mode, bool causal, cudaStream_t stream) {
    switch (bias_mode * 2 + (causal ? 1 : 0)) {
        case 0: return launch_fwd_inst<kD, kBf16, 0, false>(kp, stream);
        case 1: return launch_fwd_inst<kD, kBf16, 0, true>(kp, stream);
        case 2: return launch_fwd_inst<kD, kBf16, 1, false>(kp, stream);
        case 3: return launch_fwd_inst<kD, kBf16, 1, true>(kp, stream);
        case 4: return launch_fwd_inst<kD, kBf16, 2, false>(kp, stream);
        case 5: return launch_fwd_inst<kD, kBf16, 2, true>(kp, stream);
        case 6: return launch_fwd_inst<kD, kBf16, 3, false>(kp, stream);
        case 7: return launch_fwd_inst<kD, kBf16, 3, true>(kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_attn_fwd(const AttnFwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                            cudaStream_t stream) {
#ifdef B200T5_FWD_TIMING
    // the timing build only carries the headline instantiations (D = 64, bf16, non-causal)
    if (D == 64 && bf16 && !causal) {
        if (bias_mode == 0) return launch_fwd_inst<64, true, 0, false>(kp, stream);
        if (bias_mode == 1) return launch_fwd_inst<64, true, 1, false>(kp, stream);
        if (bias_mode == 3) return launch_fwd_inst<64, true, 3, false>(kp, stream);
    }
    return cudaErrorInvalidValue;
#elif defined(B200T5_HEADLINE_ONLY)
    // developer variant libraries (tools/build_variant.sh --headline): D = 64, bf16 only -- a fifth of the compile time
    if (D == 64 && bf16) return launch_fwd_d<64, true>(kp, bias_mode, causal, stream);
    return cudaErrorInvalidValue;
#else
#define B200T5_FWD_CASE(DD)                                                              \
    case DD:                                                                             \
        return bf16 ? launch_fwd_d<DD, true>(kp, bias_mode, causal, stream)              \
                    : launch_fwd_d<DD, false>(kp, bias_mode, causal, stream);
    switch (D) {
        B200T5_FWD_CASE(16)
        B200T5_FWD_CASE(32)
        B200T5_FWD_CASE(64)
        B200T5_FWD_CASE(128)
        default: return cudaErrorInvalidValue;
    }
#undef B200T5_FWD_CASE
#endif
}

}  // namespace b200t5
