// FlashAttention-2 forward with additive (T5) bias for sm_100a -- persistent version of attn_fwd.cu.
//
// Replaces /root/reference/src/model/ops/flash_attention_v2_bias.py:327-483 (`_fwd_kernel`).
//
// Same tile pipeline, warp roles, TMEM layout and arithmetic as attn_fwd.cu (the outputs are bit-identical: checked on
// B200 over 17 shapes / modes, profiles/r1e_fwd_persist_check.jsonl); what changes is the scheduling.  attn_fwd.cu
// launches one CTA per (batch, head, 128-row query block): a cycle-stamped build (tools/fwd_timeline.py,
// profiles/r1e_fwd_timeline_bias_2cta.txt) showed that of the ~37 k cycles such a CTA occupies its half of an SM,
// ~3.3 k pass before the first score tile is ready (barrier init, TMEM allocation, the first Q/K loads) and ~5 k between
// its last tile and the first stamp of the CTA that replaces it (epilogue, exit, launch).  Here the grid is 2 CTAs
// per SM (1 at D = 128) and every CTA walks work items w = blockIdx.x, blockIdx.x + gridDim.x, ...: barriers and TMEM
// are set up once, the K/V ring and the bias ring run ahead across work items, Q of the next item is loaded as soon
// as the last QK^T of the current one has retired, and the first QK^T of the next item is issued while the softmax
// warps are still writing the current output.  Two further changes in the softmax loop: the dense bias tile is pulled
// into registers before the wait for S (its shared-memory reads, proxy fence and ring release leave the critical
// path), and the row max is taken in the same straight-line block as the bias adds.
//
// MEASURED (B200, headline shape): 142.4 us against 142.8 us for attn_fwd.cu -- no gain, +2..7 % at S = 512 / D = 128,
// -6 % at S = 4096 (static round-robin tail).  The schedule is therefore opt-in (B200T5_FWD_PERSIST=1) and the file is
// kept as the starting point for the next round (first suspect: the two resident CTAs now start in lock-step, so
// their MUFU-bound exp phases coincide instead of interleaving).
//
//   warp 4 (1 lane)  : TMA producer for Q (per work item) and the K / V ring (2 stages each)
//   warp 6 (1 lane)  : TMA producer for the bias ring (2 stages of 128 rows x 64 columns)
//   warp 5           : tcgen05.mma issuer   S = Q K^T (SS)  and  O += P V (A = P from TMEM)
//   warps 0-3        : one thread per query row: tcgen05.ld S, bias add, online softmax (lazy rescale of O in
//                      TMEM), P -> TMEM as packed 16-bit, epilogue O / l and LSE
//
// Barrier phases are derived from running counters (T = score tiles so far, W = non-empty work items so far):
//   k/v_full[s], k/v_empty[s] : ring position T % 2, phase (T / 2) & 1        s_full, p_full, pv_done : phase T & 1
//   s_empty : S(T+1) may overwrite S(T) -- waited with phase T & 1            q_full, q_empty : phase W & 1
// q_empty is a tcgen05.commit issued after the last QK^T of a work item; O needs no barrier of its own: the first
// P V of the next item waits for p_full, which the softmax warps signal after they have read O in their epilogue.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kBM = 128;   // query rows per work item
constexpr int kBN = 128;   // keys per tile
constexpr int kKVStages = 2;
constexpr int kBiasStages = 2;
constexpr int kBiasHalfBytes = kBM * 64 * 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f * kLn2;   // keep a stale max while exp(x - m) <= 256
#ifndef B200T5_EXP2_POLY
#define B200T5_EXP2_POLY 0      // developer switch, see attn_fwd.cu
#endif
#ifndef B200T5_BIAS_FHADD
#define B200T5_BIAS_FHADD 0     // developer switch, see attn_fwd.cu
#endif

template <int kD>
struct PFwdSmem {
    static constexpr int kRowBytes = (kD >= 64 ? 64 : kD) * 2;
    static constexpr int kBoxes = kD >= 64 ? kD / 64 : 1;
    static constexpr int kTileBytes = kBM * kD * 2;
    static constexpr int kBoxBytes = kBM * kRowBytes;
    static constexpr int kQ = 0;
    static constexpr int kK = kQ + kTileBytes;
    static constexpr int kV = kK + kKVStages * kTileBytes;
    static constexpr int kBias = kV + kKVStages * kTileBytes;   // dense bias ring, or the relative-position band
    static constexpr int kBars = kBias + kBiasStages * kBiasHalfBytes;
    static constexpr int kNumBars = 2 + 4 * kKVStages + 2 * kBiasStages + 4;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static constexpr uint32_t kSwizzle = kRowBytes == 128 ? kSwz128 : (kRowBytes == 64 ? kSwz64 : kSwz32);
    static constexpr int kTmemCols = (128 + kD + 64) <= 256 ? 256 : 512;
    static constexpr int kOCol = 128;
    static constexpr int kPCol = 128 + kD;
    static constexpr int kCtasPerSm = kTmemCols <= 256 && 2 * (kTotal + 1024) <= 232448 ? 2 : 1;
};

struct PFwdBars {
    uint64_t* q_full;
    uint64_t* q_empty;
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

// one work item = one (batch, head, query block); batch fastest so that the CTAs running together share bias tiles
// in L2, late (long, when causal) query blocks first
struct WorkItem {
    int b, h, row0, num_tiles;
};
template <bool kCausal>
__device__ __forceinline__ WorkItem decode_work(int w, const AttnFwdKernelParams& p) {
    WorkItem it;
    const int nmb = p.num_m_blocks;
    it.b = w % p.B;
    w /= p.B;
    const int mb = nmb - 1 - (w % nmb);
    it.h = w / nmb;
    it.row0 = mb * kBM;
    int nt = (p.N + kBN - 1) / kBN;
    if (kCausal) {
        const int last_col = it.row0 + kBM - 1 + (p.N - p.M);     // last visible key of the last row
        const int t = last_col < 0 ? 0 : last_col / kBN + 1;
        nt = t < nt ? t : nt;
    }
    it.num_tiles = nt;
    return it;
}

}  // namespace

// Developer build (-DB200T5_FWD_TIMING, as in attn_fwd.cu): the CTAs resident on SM `kPTimedSm` stamp clock64 at the phase
// boundaries of softmax thread 0 and of the MMA warp for their first kPTimedTiles score tiles (work-item boundaries show
// up as the gap between "P stored" of one tile and "wait S" of the next); the launcher prints the timeline.
#ifdef B200T5_FWD_TIMING
constexpr int kPTimedSm = 17;
constexpr int kPTimedTiles = 28;
__device__ int g_pfwd_ts_slots;
__device__ long long g_pfwd_ts[4][2][kPTimedTiles][8];   // [slot][role][tile][stamp]
__device__ int g_pfwd_ts_bid[4];
__device__ __forceinline__ long long pclk64() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#define PFWD_TS(role, t_, slot_)                                                                       \
    do {                                                                                               \
        if (ts_slot >= 0 && (t_) < (uint32_t)kPTimedTiles) g_pfwd_ts[ts_slot][role][t_][slot_] = pclk64(); \
    } while (0)
#else
#define PFWD_TS(role, t_, slot_) do { } while (0)
#endif

// Developer switch: the CTAs of the second half of the grid (the second CTA of every SM) start their softmax this many
// nanoseconds late, so that the two resident CTAs' MUFU-bound exp phases do not begin in lock-step.  0 = off.
#ifndef B200T5_PERSIST_STAGGER_NS
#define B200T5_PERSIST_STAGGER_NS 0
#endif

template <int kD, bool kBf16, int kBiasMode, bool kCausal>
__global__ void __launch_bounds__(256, 2)   // 128 regs at launch; setmaxnreg re-splits them per role
attn_fwd_persist_kernel(const __grid_constant__ AttnFwdKernelParams p, const int total_work) {
    using L = PFwdSmem<kD>;
    extern __shared__ __align__(1024) uint8_t smem[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int pseq = p.N - p.M;

    PFwdBars bars;
    {
        uint64_t* bb = reinterpret_cast<uint64_t*>(smem + L::kBars);
        bars.q_full = bb;
        bars.q_empty = bb + 1;
        bars.k_full = bb + 2;
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

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("b200t5: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        mbar_init(bars.q_full, 1);
        mbar_init(bars.q_empty, 1);
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
    }
    if (warp == 5) tmem_alloc<L::kTmemCols>(tmem_slot);
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&p.map_q);
        tma_prefetch_desc(&p.map_k);
        tma_prefetch_desc(&p.map_v);
        if (kBiasMode == 1) tma_prefetch_desc(&p.map_bias);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
#ifdef B200T5_FWD_TIMING
    int ts_slot = -1;
    {
        volatile int& s_ts_slot = *reinterpret_cast<volatile int*>(smem + L::kTmemSlot + 8);   // spare word, no static smem
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (threadIdx.x == 0) {
            s_ts_slot = -1;
            if (smid == kPTimedSm) {
                const int sl = atomicAdd(&g_pfwd_ts_slots, 1);
                if (sl < 4) {
                    s_ts_slot = sl;
                    g_pfwd_ts_bid[sl] = blockIdx.x;
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
        if (warp == 4 && lane == 0) {
            // ---- Q / K / V producer ----
            uint32_t T = 0, W = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const WorkItem it = decode_work<kCausal>(w, p);
                if (it.num_tiles == 0) continue;
                if (W > 0) mbar_wait_producer(bars.q_empty, (W - 1) & 1);      // every QK^T of the previous item has read Q
                mbar_arrive_expect_tx(bars.q_full, L::kTileBytes);
#pragma unroll
                for (int bx = 0; bx < L::kBoxes; ++bx)
                    tma_load_4d(smem + L::kQ + bx * L::kBoxBytes, &p.map_q, bars.q_full, bx * 64, it.row0, it.h, it.b);
                for (int j = 0; j < it.num_tiles; ++j, ++T) {
                    const int s = T % kKVStages;
                    const uint32_t par = ((T / kKVStages) & 1) ^ 1;
                    mbar_wait_producer(bars.k_empty + s, par);
                    mbar_arrive_expect_tx(bars.k_full + s, L::kTileBytes);
#pragma unroll
                    for (int bx = 0; bx < L::kBoxes; ++bx)
                        tma_load_4d(smem + L::kK + s * L::kTileBytes + bx * L::kBoxBytes, &p.map_k, bars.k_full + s,
                                    bx * 64, j * kBN, it.h, it.b);
                    mbar_wait_producer(bars.v_empty + s, par);
                    mbar_arrive_expect_tx(bars.v_full + s, L::kTileBytes);
#pragma unroll
                    for (int bx = 0; bx < L::kBoxes; ++bx)
                        tma_load_4d(smem + L::kV + s * L::kTileBytes + bx * L::kBoxBytes, &p.map_v, bars.v_full + s,
                                    bx * 64, j * kBN, it.h, it.b);
                }
                ++W;
            }
        } else if (warp == 6 && lane == 0 && kBiasMode == 1) {
            // ---- bias producer: two 64-column halves per tile ----
            uint32_t I = 0;                                       // halves issued so far
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const WorkItem it = decode_work<kCausal>(w, p);
                const int hb = p.bias_h_bcast ? 0 : it.h;
                const int bb = p.bias_b_bcast ? 0 : it.b;
                for (int i2 = 0; i2 < 2 * it.num_tiles; ++i2, ++I) {
                    const int s = I % kBiasStages;
                    const uint32_t par = ((I / kBiasStages) & 1) ^ 1;
                    mbar_wait_producer(bars.b_empty + s, par);
                    mbar_arrive_expect_tx(bars.b_full + s, kBiasHalfBytes);
                    tma_load_4d(smem + L::kBias + s * kBiasHalfBytes, &p.map_bias, bars.b_full + s,
                                (i2 >> 1) * kBN + (i2 & 1) * 64, it.row0, hb, bb);
                }
            }
        } else if (warp == 5) {
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

            // S for score tile number t (ring position t % 2); `last` = last tile of its work item
            auto issue_s = [&](uint32_t t, bool last) {
                const int s = t % kKVStages;
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
                    if (last) umma_commit(bars.q_empty);
                }
                __syncwarp();
            };

            uint32_t T = 0, W = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const WorkItem it = decode_work<kCausal>(w, p);
                if (it.num_tiles == 0) continue;
                // first score tile of this item
                mbar_wait(bars.q_full, W & 1);
                mbar_wait(bars.k_full + (T % kKVStages), (T / kKVStages) & 1);
                if (T > 0) mbar_wait(bars.s_empty, (T - 1) & 1);
                tc_fence_after();
                PFWD_TS(1, T, 0);
                issue_s(T, it.num_tiles == 1);
                PFWD_TS(1, T, 1);
                for (int j = 0; j < it.num_tiles; ++j, ++T) {
                    if (j + 1 < it.num_tiles) {
                        const uint32_t tn = T + 1;
                        mbar_wait(bars.k_full + (tn % kKVStages), (tn / kKVStages) & 1);
                        mbar_wait(bars.s_empty, T & 1);
                        tc_fence_after();
                        PFWD_TS(1, tn, 0);
                        issue_s(tn, j + 2 == it.num_tiles);
                        PFWD_TS(1, tn, 1);
                    }
                    const int s = T % kKVStages;
                    mbar_wait(bars.v_full + s, (T / kKVStages) & 1);
                    mbar_wait(bars.p_full, T & 1);      // (j == 0: also means the previous item's O has been read)
                    tc_fence_after();
                    PFWD_TS(1, T, 2);
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
                    PFWD_TS(1, T, 3);
                }
                ++W;
            }
        }
    } else {
        // =============================== softmax warpgroup ===============================
        setmaxnreg_inc<208>();
        const int r = threadIdx.x;                 // row inside the tile == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tm_s = tmem_base + lane_off;
        const uint32_t tm_o = tmem_base + lane_off + L::kOCol;
        const uint32_t tm_p = tmem_base + lane_off + L::kPCol;
        const float* band = reinterpret_cast<const float*>(smem + L::kBias);   // [bias mode 3]
        int band_h = -1;

        uint32_t T = 0;       // score tiles so far
        uint32_t I = 0;       // bias halves so far
        if (B200T5_PERSIST_STAGGER_NS > 0 && blockIdx.x >= gridDim.x / 2) __nanosleep(B200T5_PERSIST_STAGGER_NS);
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const WorkItem it = decode_work<kCausal>(w, p);
            const int b = it.b, h = it.h, row0 = it.row0, num_tiles = it.num_tiles;
            const int grow = row0 + r;                 // global query row

            if (kBiasMode == 3 && h != band_h) {
                // (re)load the band of this head; every softmax thread has left the previous item's tiles
                if (band_h >= 0) named_bar_sync(1, 128);
                float* dst = reinterpret_cast<float*>(smem + L::kBias);
                const float* src = p.rpe.band + (int64_t)h * p.rpe.band_len;
                for (int i = threadIdx.x; i < p.rpe.band_len; i += 128) dst[i] = __ldg(src + i);
                named_bar_sync(1, 128);
                band_h = h;
            }

            float m_ref = -INFINITY;   // running reference max (natural units, scaled + biased scores)
            float l_sum = 0.f;

            const uint16_t* bias_row = nullptr;
            if (kBiasMode == 2) {
                bias_row = reinterpret_cast<const uint16_t*>(p.bias) + (p.bias_b_bcast ? 0 : (int64_t)b * p.bias_sb) +
                           (p.bias_h_bcast ? 0 : (int64_t)h * p.bias_sh) + (int64_t)grow * p.bias_sm;
            }

            for (int j = 0; j < num_tiles; ++j, ++T) {
                const int col0 = j * kBN;
                float x[kBN];

                // ---- dense bias tile -> registers (packed 16-bit pairs) before the scores are needed: the shared-memory
                //      reads, the proxy fence and the release of the ring overlap the wait for S ----
                PFWD_TS(0, T, 0);
                uint32_t bw[kBN / 2];
                if (kBiasMode == 1) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int s = (I + hh) % kBiasStages;
                        mbar_wait(bars.b_full + s, ((I + hh) / kBiasStages) & 1);
                        const uint8_t* brow = smem + L::kBias + s * kBiasHalfBytes + r * 128;
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) {
                            const uint4 u = *reinterpret_cast<const uint4*>(brow + ((c8 ^ (r & 7)) << 4));
                            bw[hh * 32 + c8 * 4 + 0] = u.x;
                            bw[hh * 32 + c8 * 4 + 1] = u.y;
                            bw[hh * 32 + c8 * 4 + 2] = u.z;
                            bw[hh * 32 + c8 * 4 + 3] = u.w;
                        }
                    }
                    // WAR across proxies: these generic-proxy reads must have completed before the TMA (async proxy)
                    // may overwrite the stages (see attn_fwd.cu); one fence covers both halves.
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bars.b_empty + (I % kBiasStages));
                        mbar_arrive(bars.b_empty + ((I + 1) % kBiasStages));
                    }
                    I += 2;
                }

                mbar_wait(bars.s_full, T & 1);
                tc_fence_after();
                PFWD_TS(0, T, 1);
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
                PFWD_TS(0, T, 2);

                // ---- scores = S * sm_scale + bias, and their row max.  The max is taken in the same straight-line
                //      block as the adds (its dependent chain hides inside them); a tile that needs masking (the
                //      causal diagonal, the key tail) masks afterwards and takes the max again. ----
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
#if B200T5_BIAS_FHADD
                    if (p.sm_scale == 1.f) {
                        // developer build: bias add straight from the packed 16-bit pair (one FHADD per element)
#pragma unroll
                        for (int c = 0; c < kBN; c += 2) add_f32_16x2<kBf16>(bw[c / 2], x[c], x[c + 1], x[c], x[c + 1]);
                    } else
#endif
#pragma unroll
                    for (int c = 0; c < kBN; c += 2) {
                        const float2 f = unpack2<kBf16>(bw[c / 2]);
                        x[c] = fmaf(x[c], p.sm_scale, f.x);
                        x[c + 1] = fmaf(x[c + 1], p.sm_scale, f.y);
                    }
                    tmax = row_max();
                } else if (kBiasMode == 2) {
#pragma unroll
                    for (int c = 0; c < kBN; ++c) {
                        float bv = 0.f;
                        if (grow < p.M && col0 + c < p.N)
                            bv = to_float16bit<kBf16>(__ldg(bias_row + (int64_t)(col0 + c) * p.bias_sn));
                        x[c] = fmaf(x[c], p.sm_scale, bv);
                    }
                    tmax = row_max();
                } else if (kBiasMode == 3) {
                    // relative positions n - m of this tile: [col0 - row0 - 127, col0 - row0 + 127]
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

                PFWD_TS(0, T, 3);
                PFWD_TS(0, T, 3);                                 // scores + bias + row max done
                // ---- online softmax with lazy rescale ----
                float alpha = 1.f;
                if (tmax > m_ref + kRescaleThreshold) {       // also true for the first finite tile (m_ref = -inf)
                    alpha = __expf(m_ref - tmax);              // exp(-inf) = 0 on the first tile
                    m_ref = tmax;
                }
                const float m_safe = (m_ref == -INFINITY) ? 0.f : m_ref;
                const float neg_m_log2 = -m_safe * kLog2e;

                uint32_t pk[kBN / 2];
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int c = 0; c < kBN; c += 2) {
                    float e0, e1;
                    if (B200T5_EXP2_POLY > 0 && ((c / 2) % 8) < B200T5_EXP2_POLY) {
                        ex2_poly_pair(fmaf(x[c], kLog2e, neg_m_log2), fmaf(x[c + 1], kLog2e, neg_m_log2), e0, e1);
                    } else {
                        e0 = ex2_approx(fmaf(x[c], kLog2e, neg_m_log2));
                        e1 = ex2_approx(fmaf(x[c + 1], kLog2e, neg_m_log2));
                    }
                    s0 += e0;
                    s1 += e1;
                    pk[c / 2] = pack2<kBf16>(e0, e1);
                }
                l_sum = l_sum * alpha + (s0 + s1);
                PFWD_TS(0, T, 5);

                if (j > 0) {
                    mbar_wait(bars.pv_done, (T - 1) & 1);     // O and the P buffer are free again
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
                // (j == 0: the P buffer is free -- the epilogue of the previous item waited for its last P V)
                PFWD_TS(0, T, 6);
                tmem_st32(tm_p + 0, reinterpret_cast<const uint32_t(&)[32]>(pk[0]));
                tmem_st32(tm_p + 32, reinterpret_cast<const uint32_t(&)[32]>(pk[32]));
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars.p_full);
                PFWD_TS(0, T, 7);
            }

            // ---- epilogue: O / l -> global (row-contiguous 16-byte stores), LSE ----
            const bool row_ok = grow < p.M;
            uint8_t* o_row = reinterpret_cast<uint8_t*>(p.o) +
                             2 * ((int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)grow * p.o_sm);
            if (num_tiles > 0) {
                mbar_wait(bars.pv_done, (T - 1) & 1);
                tc_fence_after();
#if B200T5_EXP2_POLY > 0
                // the polynomial clamps exp2(-inf) to 2^-125 instead of 0: a row with no visible key is recognised by its max
                const float inv_l = (l_sum > 0.f && m_ref != -INFINITY) ? 1.f / l_sum : 0.f;
#else
                const float inv_l = l_sum > 0.f ? 1.f / l_sum : 0.f;
#endif
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
    }

    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc<L::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kBiasMode, bool kCausal>
static cudaError_t launch_pfwd_inst(const AttnFwdKernelParams& kp, cudaStream_t stream) {
    using L = PFwdSmem<kD>;
    auto kern = attn_fwd_persist_kernel<kD, kBf16, kBiasMode, kCausal>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    // SM count of the current device (the C ABI has made the caller's device current); queried per launch: the call costs
    // well under a microsecond and keeps the launcher stateless across devices and threads
    int dev = 0, num_sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const long long total = (long long)kp.B * kp.H * kp.num_m_blocks;
    const long long slots = (long long)num_sms * L::kCtasPerSm;
    const int grid = static_cast<int>(total < slots ? total : slots);
#ifdef B200T5_FWD_TIMING
    {
        int zero = 0;
        cudaMemcpyToSymbol(g_pfwd_ts_slots, &zero, sizeof(int));
    }
#endif
    kern<<<grid, 256, L::kTotal, stream>>>(kp, static_cast<int>(total));
    count_launch();
#ifdef B200T5_FWD_TIMING
    {
        cudaDeviceSynchronize();
        static long long ts[4][2][kPTimedTiles][8];
        int bids[4], n = 0;
        cudaMemcpyFromSymbol(ts, g_pfwd_ts, sizeof(ts));
        cudaMemcpyFromSymbol(bids, g_pfwd_ts_bid, sizeof(bids));
        cudaMemcpyFromSymbol(&n, g_pfwd_ts_slots, sizeof(int));
        if (n > 4) n = 4;
        long long t0 = n > 0 ? ts[0][0][0][0] : 0;
        for (int s = 0; s < n; ++s)
            if (ts[s][0][0][0] < t0) t0 = ts[s][0][0][0];
        printf("PFWD_TIMING sm %d: %d resident CTAs (grid %d); softmax thread 0 per score tile: [tile start, S ready, S in regs, "
               "bias+max done, -, exp done, P V(t-1) done, P stored]; mma: [S(t) issue start, S(t) issued, P(t) ready, PV(t) issued]\n",
               kPTimedSm, n, grid);
        for (int s = 0; s < n; ++s) {
            printf(" CTA slot %d (block %d)\n", s, bids[s]);
            for (int j = 0; j < kPTimedTiles; ++j) {
                printf("  t=%d sm:", j);
                for (int q = 0; q < 8; ++q) printf(" %7lld", q == 4 ? 0LL : ts[s][0][j][q] - t0);
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
static cudaError_t launch_pfwd_d(const AttnFwdKernelParams& kp, int bias_mode, bool causal, cudaStream_t stream) {
    switch (bias_mode * 2 + (causal ? 1 : 0)) {
        case 0: return launch_pfwd_inst<kD, kBf16, 0, false>(kp, stream);
        case 1: return launch_pfwd_inst<kD, kBf16, 0, true>(kp, stream);
        case 2: return launch_pfwd_inst<kD, kBf16, 1, false>(kp, stream);
        case 3: return launch_pfwd_inst<kD, kBf16, 1, true>(kp, stream);
        case 4: return launch_pfwd_inst<kD, kBf16, 2, false>(kp, stream);
        case 5: return launch_pfwd_inst<kD, kBf16, 2, true>(kp, stream);
        case 6: return launch_pfwd_inst<kD, kBf16, 3, false>(kp, stream);
        case 7: return launch_pfwd_inst<kD, kBf16, 3, true>(kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_attn_fwd_persist(const AttnFwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                                    cudaStream_t stream) {
#ifdef B200T5_HEADLINE_ONLY
    if (D == 64 && bf16) return launch_pfwd_d<64, true>(kp, bias_mode, causal, stream);
    return cudaErrorInvalidValue;
#else
#define B200T5_PFWD_CASE(DD)                                                             \
    case DD:                                                                             \
        return bf16 ? launch_pfwd_d<DD, true>(kp, bias_mode, causal, stream)             \
                    : launch_pfwd_d<DD, false>(kp, bias_mode, causal, stream);
    switch (D) {
        B200T5_PFWD_CASE(16)
        B200T5_PFWD_CASE(32)
        B200T5_PFWD_CASE(64)
        B200T5_PFWD_CASE(128)
        default: return cudaErrorInvalidValue;
    }
#undef B200T5_PFWD_CASE
#endif
}

}  // namespace b200t5
