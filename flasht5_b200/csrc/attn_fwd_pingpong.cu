// FlashAttention-2 forward with additive (T5) bias for sm_100a -- two query tiles per CTA, exp phases in anti-phase.
//
// Replaces /root/reference/src/model/ops/flash_attention_v2_bias.py:327-483 (`_fwd_kernel`).  Head dims 16 / 32 / 64.
// DEVELOPER KERNEL (B200T5_FWD_PINGPONG=1): written from the measurements of this round, not yet run on hardware.
//
// Why.  The cycle-stamped timeline of attn_fwd.cu (profiles/r1e_fwd_timeline_*.txt) shows that one softmax warp already
// saturates its scheduler's MUFU while it computes the 128 exp2 of its row (1 024 of the 2 900 cycles a tile takes),
// and that with two CTAs per SM the two warps that share a scheduler spend part of the time in that phase together (each
// at half rate) and part of the time both outside it (MUFU idle): 3 541 cycles per tile and CTA where max(2 E, E + R)
// = 2 800 would do (E = exp phase, R = everything else).  Two independent CTAs cannot be kept out of step; a persistent
// schedule starts them in lock-step, which is the worst case (attn_fwd_persist.cu gained nothing).  So here ONE CTA per SM
// owns TWO 128-row query tiles of the same (batch, head) -- two softmax warpgroups, two MMA-issuing warps, one K/V ring
// shared by both (half the K/V traffic per query row) -- and the two warpgroups hand an "exp token" back and forth
// through named barriers: the exp phases alternate strictly, E0 E1 E0 E1 ..., each overlapping the other tile's R.
//
//   warps 0-3   : softmax warpgroup of query tile 0        warps 4-7 : softmax warpgroup of query tile 1
//   warp 8      : TMA producer: Q0, Q1 (per work item), the K / V ring (2 stages each)
//   warp 9 / 10 : tcgen05.mma issuer of tile 0 / tile 1    (S_i = Q_i K^T, O_i += P_i V; disjoint TMEM columns)
//   warp 11 / 12: TMA producer of the bias halves of tile 0 / tile 1          warps 13-15: idle (setmaxnreg is per warpgroup)
//
// TMEM (512 columns): S0 [0,128) S1 [128,256) | O0 [256,320) O1 [320,384) | P0 [384,448) P1 [448,512).
// Shared memory (D = 64): Q0 Q1 32 KB | K ring 32 KB | V ring 32 KB | bias halves 2 x 32 KB (or the relative-position
// band) = 160 KB.  Persistent: grid = number of SMs, work item = (batch, head, pair of query blocks), batch fastest.
//
// Barrier phases come from running counters: Tk = K/V tiles so far (shared sequence: producer and both MMA warps), Ti =
// score tiles of query tile i so far, W_i = work items in which tile i had any work.  A K/V stage is released by two
// arrivals (one per MMA warp; a warp whose tile does not need that key tile -- causal -- arrives without reading it).
// The arithmetic is that of attn_fwd.cu: outputs must be bit-identical (tools/fwd_persist_check.py compares them).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kBM = 128;   // rows per query tile
constexpr int kBN = 128;   // keys per tile
constexpr int kKVStages = 2;
constexpr int kBiasHalfBytes = kBM * 64 * 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f * kLn2;
constexpr int kThreads = 512;
constexpr int kTokenBar0 = 2;      // named barriers 2, 3: the exp token of warpgroup 0 / 1

template <int kD>
struct PPSmem {
    static_assert(kD == 16 || kD == 32 || kD == 64, "two query tiles per CTA need 2 x (128 + D + 64) <= 512 TMEM columns");
    static constexpr int kRowBytes = kD * 2;
    static constexpr int kTileBytes = kBM * kD * 2;
    static constexpr int kQ = 0;                                   // [2] tiles
    static constexpr int kK = kQ + 2 * kTileBytes;
    static constexpr int kV = kK + kKVStages * kTileBytes;
    static constexpr int kBias = kV + kKVStages * kTileBytes;      // [2 tiles][2 halves], or the band (mode 3)
    static constexpr int kBars = kBias + 4 * kBiasHalfBytes;
    static constexpr int kNumBars = 4 + 4 * kKVStages + 2 * (4 + 4);
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static constexpr uint32_t kSwizzle = kRowBytes == 128 ? kSwz128 : (kRowBytes == 64 ? kSwz64 : kSwz32);
    static constexpr int kSCol = 0, kOCol = 256, kPCol = 384;     // + i * 128 / 64 / 64
};

struct PPBars {
    uint64_t* q_full;    // [2]
    uint64_t* q_empty;   // [2]
    uint64_t* k_full;    // [kKVStages]
    uint64_t* k_empty;   // 2 arrivals
    uint64_t* v_full;
    uint64_t* v_empty;   // 2 arrivals
    uint64_t* s_full;    // [2] per query tile from here on
    uint64_t* s_empty;
    uint64_t* p_full;
    uint64_t* pv_done;
    uint64_t* b_full;    // [2 tiles][2 halves]
    uint64_t* b_empty;
};

// scalars only (no arrays: a runtime index would send the struct to local memory); tile i is picked with row0(i) / nt(i)
struct PPWork {
    int b, h, row0_0, row0_1, nt_0, nt_1, nt_max;
    __device__ __forceinline__ int row0(int i) const { return i ? row0_1 : row0_0; }
    __device__ __forceinline__ int nt(int i) const { return i ? nt_1 : nt_0; }
};
template <bool kCausal>
__device__ __forceinline__ int pp_tiles_of(int row0, const AttnFwdKernelParams& p) {
    int nt = row0 < p.M ? (p.N + kBN - 1) / kBN : 0;       // a pair may hang over the end of the query sequence
    if (kCausal && nt > 0) {
        const int last_col = row0 + kBM - 1 + (p.N - p.M);   // last visible key of the last row
        const int t = last_col < 0 ? 0 : last_col / kBN + 1;
        nt = t < nt ? t : nt;
    }
    return nt;
}
template <bool kCausal>
__device__ __forceinline__ PPWork pp_decode(int w, const AttnFwdKernelParams& p, int npairs) {
    PPWork it;
    it.b = w % p.B;
    w /= p.B;
    const int pb = npairs - 1 - (w % npairs);        // late (long, when causal) query blocks first
    it.h = w / npairs;
    it.row0_0 = (2 * pb) * kBM;
    it.row0_1 = (2 * pb + 1) * kBM;
    it.nt_0 = pp_tiles_of<kCausal>(it.row0_0, p);
    it.nt_1 = pp_tiles_of<kCausal>(it.row0_1, p);
    it.nt_max = it.nt_0 > it.nt_1 ? it.nt_0 : it.nt_1;
    return it;
}

}  // namespace

// Developer build (-DB200T5_FWD_TIMING): the CTA resident on SM `kPPTimedSm` stamps clock64 at the phase boundaries of
// thread 0 of each softmax warpgroup for its first kPPTimedTiles score tiles; the launcher prints the two timelines.
#ifdef B200T5_FWD_TIMING
constexpr int kPPTimedSm = 17;
constexpr int kPPTimedTiles = 28;
__device__ long long g_pp_ts[2][kPPTimedTiles][8];   // [warpgroup][tile][stamp]
__device__ int g_pp_ts_bid;
__device__ __forceinline__ long long ppclk64() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#define PP_TS(t_, slot_)                                                                                   \
    do {                                                                                                   \
        if (ts_on && (t_) < (uint32_t)kPPTimedTiles) g_pp_ts[qi][t_][slot_] = ppclk64();                   \
    } while (0)
#else
#define PP_TS(t_, slot_) do { } while (0)
#endif

template <int kD, bool kBf16, int kBiasMode, bool kCausal>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_pingpong_kernel(const __grid_constant__ AttnFwdKernelParams p, const int total_work, const int npairs) {
    using L = PPSmem<kD>;
    extern __shared__ __align__(1024) uint8_t smem[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int pseq = p.N - p.M;

    PPBars bars;
    {
        uint64_t* bb = reinterpret_cast<uint64_t*>(smem + L::kBars);
        bars.q_full = bb;
        bars.q_empty = bb + 2;
        bars.k_full = bb + 4;
        bars.k_empty = bars.k_full + kKVStages;
        bars.v_full = bars.k_empty + kKVStages;
        bars.v_empty = bars.v_full + kKVStages;
        bars.s_full = bars.v_empty + kKVStages;
        bars.s_empty = bars.s_full + 2;
        bars.p_full = bars.s_empty + 2;
        bars.pv_done = bars.p_full + 2;
        bars.b_full = bars.pv_done + 2;
        bars.b_empty = bars.b_full + 4;
    }
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("b200t5: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        *reinterpret_cast<float*>(smem + L::kTmemSlot + 4) = 0.f;      // the zero word of the exp token (see the softmax loop)
        for (int i = 0; i < 2; ++i) {
            mbar_init(bars.q_full + i, 1);
            mbar_init(bars.q_empty + i, 1);
            mbar_init(bars.s_full + i, 1);
            mbar_init(bars.s_empty + i, 4);
            mbar_init(bars.p_full + i, 4);
            mbar_init(bars.pv_done + i, 1);
        }
        for (int i = 0; i < kKVStages; ++i) {
            mbar_init(bars.k_full + i, 1);
            mbar_init(bars.k_empty + i, 2);
            mbar_init(bars.v_full + i, 1);
            mbar_init(bars.v_empty + i, 2);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(bars.b_full + i, 1);
            mbar_init(bars.b_empty + i, 4);
        }
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_slot);
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&p.map_q);
        tma_prefetch_desc(&p.map_k);
        tma_prefetch_desc(&p.map_v);
        if (kBiasMode == 1) tma_prefetch_desc(&p.map_bias);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
        // =============================== control warps ===============================
        setmaxnreg_dec<48>();
        if (warp == 8 && lane == 0) {
            // ---- Q0 / Q1 / K / V producer ----
            uint32_t Tk = 0, W0 = 0, W1 = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const PPWork it = pp_decode<kCausal>(w, p, npairs);
                if (it.nt_max == 0) continue;
                if (it.nt_0 > 0) {
                    if (W0 > 0) mbar_wait_producer(bars.q_empty + 0, (W0 - 1) & 1);
                    mbar_arrive_expect_tx(bars.q_full + 0, L::kTileBytes);
                    tma_load_4d(smem + L::kQ, &p.map_q, bars.q_full + 0, 0, it.row0_0, it.h, it.b);
                    ++W0;
                }
                if (it.nt_1 > 0) {
                    if (W1 > 0) mbar_wait_producer(bars.q_empty + 1, (W1 - 1) & 1);
                    mbar_arrive_expect_tx(bars.q_full + 1, L::kTileBytes);
                    tma_load_4d(smem + L::kQ + L::kTileBytes, &p.map_q, bars.q_full + 1, 0, it.row0_1, it.h, it.b);
                    ++W1;
                }
                for (int j = 0; j < it.nt_max; ++j, ++Tk) {
                    const int s = Tk % kKVStages;
                    const uint32_t par = ((Tk / kKVStages) & 1) ^ 1;
                    mbar_wait_producer(bars.k_empty + s, par);
                    mbar_arrive_expect_tx(bars.k_full + s, L::kTileBytes);
                    tma_load_4d(smem + L::kK + s * L::kTileBytes, &p.map_k, bars.k_full + s, 0, j * kBN, it.h, it.b);
                    mbar_wait_producer(bars.v_empty + s, par);
                    mbar_arrive_expect_tx(bars.v_full + s, L::kTileBytes);
                    tma_load_4d(smem + L::kV + s * L::kTileBytes, &p.map_v, bars.v_full + s, 0, j * kBN, it.h, it.b);
                }
            }
        } else if ((warp == 11 || warp == 12) && lane == 0 && kBiasMode == 1) {
            // ---- bias producer of query tile i: two 64-column halves per score tile ----
            const int i = warp - 11;
            uint32_t I = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const PPWork it = pp_decode<kCausal>(w, p, npairs);
                const int hb = p.bias_h_bcast ? 0 : it.h;
                const int bb = p.bias_b_bcast ? 0 : it.b;
                for (int i2 = 0; i2 < 2 * it.nt(i); ++i2, ++I) {
                    const int s = i * 2 + (I & 1);
                    mbar_wait_producer(bars.b_empty + s, ((I >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(bars.b_full + s, kBiasHalfBytes);
                    tma_load_4d(smem + L::kBias + s * kBiasHalfBytes, &p.map_bias, bars.b_full + s,
                                (i2 >> 1) * kBN + (i2 & 1) * 64, it.row0(i), hb, bb);
                }
            }
        } else if (warp == 9 || warp == 10) {
            // ---- MMA issuer of query tile i (whole warp runs the loop, one elected lane issues) ----
            const int i = warp - 9;
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, kBM, kBN, false, false);
            constexpr uint32_t idesc_pv = make_idesc(kBf16, kBM, kD, false, true);
            constexpr uint32_t sbo = 8 * L::kRowBytes;
            constexpr uint32_t hi_k = sdesc_hi(sbo, L::kSwizzle);
            const uint32_t q_lo = sdesc_lo(smem_u32(smem + L::kQ + i * L::kTileBytes), 16);
            const uint32_t k_lo0 = sdesc_lo(smem_u32(smem + L::kK), 16);
            const uint32_t v_lo0 = sdesc_lo(smem_u32(smem + L::kV), L::kTileBytes);
            const uint32_t tm_s = tmem_base + L::kSCol + i * 128;
            const uint32_t tm_o = tmem_base + L::kOCol + i * 64;
            const uint32_t tm_p = tmem_base + L::kPCol + i * 64;

            auto issue_s = [&](uint32_t tk, bool last) {
                const int s = tk % kKVStages;
                const uint32_t k_lo = k_lo0 + s * (L::kTileBytes >> 4);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_ss2(tm_s, q_lo + kk * 2, hi_k, k_lo + kk * 2, hi_k, idesc_s, kk > 0 ? 1u : 0u);
                    umma_commit(bars.s_full + i);
                    umma_commit(bars.k_empty + s);
                    if (last) umma_commit(bars.q_empty + i);
                }
                __syncwarp();
            };

            uint32_t Tk = 0, Ti = 0, Wi = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const PPWork it = pp_decode<kCausal>(w, p, npairs);
                const int nt = it.nt(i);
                if (nt > 0) {
                    mbar_wait(bars.q_full + i, Wi & 1);
                    mbar_wait(bars.k_full + (Tk % kKVStages), (Tk / kKVStages) & 1);
                    if (Ti > 0) mbar_wait(bars.s_empty + i, (Ti - 1) & 1);
                    tc_fence_after();
                    issue_s(Tk, nt == 1);
                    ++Wi;
                }
                for (int j = 0; j < it.nt_max; ++j, ++Tk) {
                    const int s = Tk % kKVStages;
                    if (j < nt) {
                        if (j + 1 < nt) {
                            const uint32_t tn = Tk + 1;
                            mbar_wait(bars.k_full + (tn % kKVStages), (tn / kKVStages) & 1);
                            mbar_wait(bars.s_empty + i, Ti & 1);
                            tc_fence_after();
                            issue_s(tn, j + 2 == nt);
                        }
                        mbar_wait(bars.v_full + s, (Tk / kKVStages) & 1);
                        mbar_wait(bars.p_full + i, Ti & 1);
                        tc_fence_after();
                        const uint32_t v_lo = v_lo0 + s * (L::kTileBytes >> 4);
                        if (leader) {
#pragma unroll
                            for (int kk = 0; kk < kBN / 16; ++kk)
                                umma_ts2(tm_o, tm_p + kk * 8, v_lo + ((kk * 16 * L::kRowBytes) >> 4), hi_k, idesc_pv,
                                         (j > 0 || kk > 0) ? 1u : 0u);
                            umma_commit(bars.pv_done + i);
                            umma_commit(bars.v_empty + s);
                        }
                        __syncwarp();
                        ++Ti;
                    } else {
                        // this key tile is beyond what query tile i sees (causal): release the stage without reading it.
                        // Waiting for the data first keeps the arrival inside the right phase of the empty barriers.
                        mbar_wait(bars.k_full + s, (Tk / kKVStages) & 1);
                        mbar_wait(bars.v_full + s, (Tk / kKVStages) & 1);
                        if (lane == 0) {
                            mbar_arrive(bars.k_empty + s);
                            mbar_arrive(bars.v_empty + s);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // =============================== softmax warpgroups ===============================
        setmaxnreg_inc<208>();
        const int qi = warp >> 2;                   // query tile of this warpgroup
        const int r = threadIdx.x & 127;            // row inside the tile == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tm_s = tmem_base + lane_off + L::kSCol + qi * 128;
        const uint32_t tm_o = tmem_base + lane_off + L::kOCol + qi * 64;
        const uint32_t tm_p = tmem_base + lane_off + L::kPCol + qi * 64;
        uint64_t* const b_full = bars.b_full + qi * 2;
        uint64_t* const b_empty = bars.b_empty + qi * 2;
        const uint8_t* const bias_smem = smem + L::kBias + qi * 2 * kBiasHalfBytes;
        const float* band = reinterpret_cast<const float*>(smem + L::kBias);   // [bias mode 3] shared by both warpgroups
        int band_h = -1;
        const uint32_t token_zero = smem_u32(smem + L::kTmemSlot + 4);          // holds 0.0f (written before the first barrier)
        const uint32_t token_dump = smem_u32(smem + L::kTmemSlot + 8 + 4 * qi); // write-only scratch word of this warpgroup

#ifdef B200T5_FWD_TIMING
        bool ts_on = false;
        {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            ts_on = smid == kPPTimedSm && r == 0;
            if (ts_on && qi == 0) g_pp_ts_bid = blockIdx.x;
        }
#endif
        // the exp token starts with warpgroup 0: warpgroup 1 pre-arrives on warpgroup 0's barrier
        if (qi == 1) named_bar_arrive(kTokenBar0 + 0, 256);

        uint32_t T = 0;       // score tiles of this query tile so far
        uint32_t I = 0;       // bias halves so far
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const PPWork it = pp_decode<kCausal>(w, p, npairs);
            const int b = it.b, h = it.h, row0 = it.row0(qi), num_tiles = it.nt(qi);
            const int grow = row0 + r;

            if (kBiasMode == 3 && h != band_h) {
                // (re)load the band of this head: all 256 softmax threads have left the previous item's tiles
                named_bar_sync(1, 256);
                float* dst = reinterpret_cast<float*>(smem + L::kBias);
                const float* src = p.rpe.band + (int64_t)h * p.rpe.band_len;
                for (int i = threadIdx.x; i < p.rpe.band_len; i += 256) dst[i] = __ldg(src + i);
                named_bar_sync(1, 256);
                band_h = h;
            }

            float m_ref = -INFINITY;
            float l_sum = 0.f;

            for (int j = 0; j < it.nt_max; ++j) {
                if (j >= num_tiles) {
                    // the other query tile still has key tiles: keep the token moving
                    named_bar_sync(kTokenBar0 + qi, 256);
                    named_bar_arrive(kTokenBar0 + (qi ^ 1), 256);
                    continue;
                }
                const int col0 = j * kBN;
                float x[kBN];
                PP_TS(T, 0);

                // ---- dense bias tile -> registers before the scores are needed ----
                uint32_t bw[kBN / 2];
                if (kBiasMode == 1) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int s = (I + hh) & 1;
                        mbar_wait(b_full + s, ((I + hh) >> 1) & 1);
                        const uint8_t* brow = bias_smem + s * kBiasHalfBytes + r * 128;
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) {
                            const uint4 u = *reinterpret_cast<const uint4*>(brow + ((c8 ^ (r & 7)) << 4));
                            bw[hh * 32 + c8 * 4 + 0] = u.x;
                            bw[hh * 32 + c8 * 4 + 1] = u.y;
                            bw[hh * 32 + c8 * 4 + 2] = u.z;
                            bw[hh * 32 + c8 * 4 + 3] = u.w;
                        }
                    }
                    fence_proxy_async_smem();         // generic reads done before the TMA refills the halves (see attn_fwd.cu)
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(b_empty + (I & 1));
                        mbar_arrive(b_empty + ((I + 1) & 1));
                    }
                    I += 2;
                }

                mbar_wait(bars.s_full + qi, T & 1);
                tc_fence_after();
                PP_TS(T, 1);
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
                if (lane == 0) mbar_arrive(bars.s_empty + qi);
                PP_TS(T, 2);

                // ---- scores = S * sm_scale + bias and their row max, one straight-line block ----
                auto row_max = [&]() -> float {
                    float t0 = x[0], t1 = x[1], t2 = x[2], t3 = x[3], t4 = x[4], t5 = x[5], t6 = x[6], t7 = x[7];
#pragma unroll
                    for (int c = 8; c < kBN; c += 8) {
                        t0 = fmaxf(t0, x[c]);
                        t1 = fmaxf(t1, x[c + 1]);
                        t2 = fmaxf(t2, x[c + 2]);
                        t3 = fmaxf(t3, x[c + 3]);
                        t4 = fmaxf(t4, x[c + 4]);
                        t5 = fmaxf(t5, x[c + 5]);
                        t6 = fmaxf(t6, x[c + 6]);
                        t7 = fmaxf(t7, x[c + 7]);
                    }
                    return fmaxf(fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)), fmaxf(fmaxf(t4, t5), fmaxf(t6, t7)));
                };
                float tmax;
                if (kBiasMode == 1) {
#pragma unroll
                    for (int c = 0; c < kBN; c += 2) {
                        const float2 f = unpack2<kBf16>(bw[c / 2]);
                        x[c] = fmaf(x[c], p.sm_scale, f.x);
                        x[c + 1] = fmaf(x[c + 1], p.sm_scale, f.y);
                    }
                    tmax = row_max();
                } else if (kBiasMode == 3) {
                    const int rel_min = col0 - row0 - (kBM - 1);
                    const int rel_max = col0 - row0 + (kBN - 1);
                    if (rel_max <= p.rpe.const_lo || rel_min >= p.rpe.const_hi) {
                        const float bv = band[(rel_max <= p.rpe.const_lo ? p.rpe.const_lo : p.rpe.const_hi) - p.rpe.band_lo];
#pragma unroll
                        for (int c = 0; c < kBN; ++c) x[c] = fmaf(x[c], p.sm_scale, bv);
                        tmax = row_max();
                    } else {
                        const float* bp = band + (col0 - grow - p.rpe.band_lo);
#pragma unroll
                        for (int c = 0; c < kBN; ++c) x[c] = fmaf(x[c], p.sm_scale, bp[c]);
                        tmax = row_max();
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < kBN; ++c) x[c] *= p.sm_scale;
                    tmax = row_max();
                }

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
                        tmax = row_max();
                    }
                }

                // ---- online softmax with lazy rescale ----
                float alpha = 1.f;
                if (tmax > m_ref + kRescaleThreshold) {
                    alpha = __expf(m_ref - tmax);
                    m_ref = tmax;
                }
                const float m_safe = (m_ref == -INFINITY) ? 0.f : m_ref;
                float neg_m_log2 = -m_safe * kLog2e;

                // ---- exp phase: MUFU-bound; the two warpgroups take turns ----
                PP_TS(T, 3);
                named_bar_sync(kTokenBar0 + qi, 256);
                PP_TS(T, 4);
                // exp2 is pure register arithmetic, which ptxas schedules freely around a barrier (first build: 124 of the
                // 128 MUFU.EX2 sat ABOVE the BAR.SYNC).  Memory operations are ordered against barriers, so the common
                // operand of every exp2 takes a (zero) word read from shared memory after the barrier ...
                neg_m_log2 += ld_shared_volatile_f32(token_zero);
                uint32_t pk[kBN / 2];
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int c = 0; c < kBN; c += 2) {
                    const float e0 = ex2_approx(fmaf(x[c], kLog2e, neg_m_log2));
                    const float e1 = ex2_approx(fmaf(x[c + 1], kLog2e, neg_m_log2));
                    s0 += e0;
                    s1 += e1;
                    pk[c / 2] = pack2<kBf16>(e0, e1);
                }
                // ... and the sum every exp2 feeds is written to shared memory before the token is handed over
                st_shared_volatile_f32(token_dump, s0 + s1);
                named_bar_arrive(kTokenBar0 + (qi ^ 1), 256);
                PP_TS(T, 5);
                l_sum = l_sum * alpha + (s0 + s1);

                if (j > 0) {
                    mbar_wait(bars.pv_done + qi, (T - 1) & 1);     // O and the P buffer are free again
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
                PP_TS(T, 6);
                tmem_st32(tm_p + 0, reinterpret_cast<const uint32_t(&)[32]>(pk[0]));
                tmem_st32(tm_p + 32, reinterpret_cast<const uint32_t(&)[32]>(pk[32]));
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars.p_full + qi);
                PP_TS(T, 7);
                ++T;
            }

            // ---- epilogue: O / l -> global (row-contiguous 16-byte stores), LSE ----
            const bool row_ok = grow < p.M;
            uint8_t* o_row = reinterpret_cast<uint8_t*>(p.o) +
                             2 * ((int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)grow * p.o_sm);
            if (num_tiles > 0) {
                mbar_wait(bars.pv_done + qi, (T - 1) & 1);
                tc_fence_after();
                const float inv_l = l_sum > 0.f ? 1.f / l_sum : 0.f;
                constexpr int kChunk = kD >= 32 ? 32 : 16;
#pragma unroll
                for (int c0 = 0; c0 < kD; c0 += kChunk) {
                    uint32_t o[kChunk];
                    if constexpr (kChunk == 32) tmem_ld32(tm_o + c0, o);
                    else tmem_ld16(tm_o + c0, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                    tmem_ld_wait();
                    if (row_ok) {
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
        // balance the token: warpgroup 0 consumes the arrival warpgroup 1 posted after its last exp phase
        if (qi == 0) named_bar_sync(kTokenBar0 + 0, 256);
    }

    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kBiasMode, bool kCausal>
static cudaError_t launch_pp_inst(const AttnFwdKernelParams& kp, cudaStream_t stream) {
    using L = PPSmem<kD>;
    auto kern = attn_fwd_pingpong_kernel<kD, kBf16, kBiasMode, kCausal>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    int dev = 0, num_sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const int npairs = (kp.num_m_blocks + 1) / 2;
    const long long total = (long long)kp.B * kp.H * npairs;
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    kern<<<grid, kThreads, L::kTotal, stream>>>(kp, static_cast<int>(total), npairs);
    count_launch();
#ifdef B200T5_FWD_TIMING
    {
        cudaDeviceSynchronize();
        static long long ts[2][kPPTimedTiles][8];
        int bid = -1;
        cudaMemcpyFromSymbol(ts, g_pp_ts, sizeof(ts));
        cudaMemcpyFromSymbol(&bid, g_pp_ts_bid, sizeof(int));
        const long long t0 = ts[0][0][0] < ts[1][0][0] ? ts[0][0][0] : ts[1][0][0];
        printf("PP_TIMING sm %d (block %d, grid %d); thread 0 of each softmax warpgroup per score tile: [tile start, S ready, S in regs, "
               "bias+max done, token acquired, exp done (token passed), P V(t-1) done, P stored]\n", kPPTimedSm, bid, grid);
        for (int j = 0; j < kPPTimedTiles; ++j) {
            for (int q = 0; q < 2; ++q) {
                printf("  t=%d wg%d:", j, q);
                for (int k = 0; k < 8; ++k) printf(" %7lld", ts[q][j][k] - t0);
                printf(q ? "\n" : "   |");
            }
        }
        fflush(stdout);
    }
#endif
    return cudaGetLastError();
}

template <int kD, bool kBf16>
static cudaError_t launch_pp_d(const AttnFwdKernelParams& kp, int bias_mode, bool causal, cudaStream_t stream) {
    switch (bias_mode * 2 + (causal ? 1 : 0)) {
        case 0: return launch_pp_inst<kD, kBf16, 0, false>(kp, stream);
        case 1: return launch_pp_inst<kD, kBf16, 0, true>(kp, stream);
        case 2: return launch_pp_inst<kD, kBf16, 1, false>(kp, stream);
        case 3: return launch_pp_inst<kD, kBf16, 1, true>(kp, stream);
        case 6: return launch_pp_inst<kD, kBf16, 3, false>(kp, stream);
        case 7: return launch_pp_inst<kD, kBf16, 3, true>(kp, stream);
        default: return cudaErrorInvalidValue;       // the pointer path (mode 2) stays with attn_fwd.cu
    }
}

cudaError_t launch_attn_fwd_pingpong(const AttnFwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                                     cudaStream_t stream) {
#ifdef B200T5_HEADLINE_ONLY
    if (D == 64 && bf16) return launch_pp_d<64, true>(kp, bias_mode, causal, stream);
    return cudaErrorInvalidValue;
#else
    switch (D) {
        case 16: return bf16 ? launch_pp_d<16, true>(kp, bias_mode, causal, stream) : launch_pp_d<16, false>(kp, bias_mode, causal, stream);
        case 32: return bf16 ? launch_pp_d<32, true>(kp, bias_mode, causal, stream) : launch_pp_d<32, false>(kp, bias_mode, causal, stream);
        case 64: return bf16 ? launch_pp_d<64, true>(kp, bias_mode, causal, stream) : launch_pp_d<64, false>(kp, bias_mode, causal, stream);
        default: return cudaErrorInvalidValue;
    }
#endif
}

}  // namespace b200t5
