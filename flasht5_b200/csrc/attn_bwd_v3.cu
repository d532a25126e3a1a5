// FlashAttention-2 backward with additive (T5) bias for sm_100a, head dims 16 / 32 / 64 -- transposed formulation.
//
// Replaces /root/reference/src/model/ops/flash_attention_v2_bias.py:
//   _bwd_kv_kernel :559-745 and _bwd_q_kernel :748-905 (one kernel, 5 tensor-core contractions per tile).
//
// Why this shape (measured, round 2): the previous kernel (lanes = query rows, P / dS handed to the dV / dK / dQ MMAs
// through shared memory) moved ~430 KB per 128x128 tile through the 128 B/clk shared-memory port (3 400 of its 3 900
// cycles per tile), and its stages ran back to back.  Here the CTA computes the TRANSPOSED score tile,
//     S^T = K Q^T   and   dP^T = V dO^T          (TMEM lanes = keys, columns = queries),
// with K and V held in TMEM for the whole CTA as the A operands (TMA -> shared memory -> tcgen05.st, once), so that
//     dV += P^T dO   and   dK += dS^T Q
// take P^T / dS^T straight FROM TMEM (the compute warps write them as packed 16-bit pairs into columns of their own), and
// only dS goes to shared memory -- once, in the layout the dQ MMA reads as an MN-major A operand and the TMA unit
// reduces into the (transposed) dBias surface.  Shared-memory traffic per tile drops from ~430 KB to ~210 KB (+64 KB
// with a dense bias).
//
// One CTA = one (batch, head, 128-key block); it walks the query sequence in 128-row tiles, each processed as four
// 32-query sub-tiles t = 4k + j.  Compute warpgroup j owns sub-tile j of every tile; S^T / dP^T of sub-tile t land in TMEM
// buffer t & 1 and are released as soon as the warpgroup has them in registers:
//
//   TMA producer    K, V once (requested before the setup barrier); ring of 64-query slots = Q rows + dO rows + one record of
//                   row statistics [-L log2e | -delta] written by the pre-kernel: three operations per slot
//   MMA warps A0-A3 one per sub-tile j: wait slot, wait buffer t & 1 free -> S^T(t), dP^T(t)   (8 MMAs, N = 32, TS)
//                   (four warps: a blocking barrier probe costs ~100 cycles even when the phase has long completed, and one
//                    warp doing two or three of them per sub-tile set the pace of the whole CTA at ~550 cycles per sub-tile)
//   compute WG j    wait S,dP(t) -> registers, release the buffer -> P, dS (packed f32x2 math) -> P^T, dS^T -> its own TMEM
//                   columns, dS^T -> its shared-memory box -> per-thread arrive (no barrier inside the warpgroup)
//   MMA warps B0/B1 even / odd t: wait P,dS(t), wait the other warp's token (dV / dK accumulate in sub-tile order: bitwise
//                   reproducible) -> dV,dK(t) (4 MMAs, N = D, TS) -> free the warpgroup's columns and the Q / dO slot
//   warp C          per tile: dS^T boxes -> dBias surface (TMA reduce-add / store; skipped for constant relative-position
//                   sub-tiles), then dQ(k) = dS K (8 MMAs, SS)
//   drain WG        prologue: K, V shared memory -> TMEM; per tile dQ(k): TMEM -> 16-bit staging tile -> TMA reduce-add into the
//                   dQ group surface
//
// (Measured on the way here -- profiles/r2b_*, r2c_bwd_v3_ablations.txt: (1) two 64-query half tiles over two buffers with P^T
//  aliased over S^T left every warpgroup idle ~900 cycles per tile: S^T(t+2) could not be issued before dV,dK(t) had consumed
//  P^T(t); (2) two warpgroups and one MMA warp: the single MMA thread's blocking waits set the pace at 3 300 cycles per tile
//  against 1 424 cycles of tensor work; (3) a ring of eleven 32-query slots: the producer lane's four TMA operations per
//  sub-tile (~650 cycles) set the pace; (4) K, V from global memory straight to TMEM: 4 000 cycles of prologue per CTA.
//  tools/micro/mma_rate.cu shows TS-mode MMAs run at the ideal rate down to N = 16 while SS-mode ones are bound by the
//  128 B/clk shared-memory port.)
//
//   warps 0-15  : compute warpgroups 0-3 (thread = key row)     warps 16-19 : K, V -> TMEM (prologue), dQ drain
//   warp 20 : TMA producer                  warps 21, 24, 26, 27 : MMA warps A (S^T, dP^T of sub-tile j = 0..3 of every tile)
//   warps 22, 25 : MMA warps B (dV, dK of even / odd sub-tiles)   warp 23 : warp C (dS out, dQ)
//
// TMEM columns: S^T 2 x [32] in [0,64) | dP^T 2 x [32] in [64,128) | P^T 4 x [16] in [128,192) | dS^T 4 x [16] in [192,256) |
//               dV [256,256+D) | dK [256+D,256+2D) | dQ [256+2D,256+3D) | K [448,448+D/2) | V [480,480+D/2).
//
// Bias modes: 0 none | 1 dense bias, read from a REPACKED copy that the pre-kernel makes in the workspace: for every (key
// block, 32-query sub-tile) the 128 x 32 values are stored as 4 x [128 keys][8 queries], so that thread = key row fetches
// its 32 bias values with four fully coalesced 16-byte cp.async copies one tile ahead (into shared memory: in registers
// they would be live across a whole sub-tile) | 3 T5 relative-position bias from the band in shared memory.
// dS leaves the CTA transposed: the surface is (G, H, N, M) and the post-kernel transposes while it reduces.
#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kBM = 128;            // queries per tile
constexpr int kBN = 128;            // keys per CTA
constexpr int kSub = 32;            // queries per sub-tile
constexpr int kNSub = kBM / kSub;   // 4
constexpr int kBoxBytes = 128 * kSub * 2;   // [128 keys][32 queries] 16-bit, 64B-swizzled rows
constexpr float kLog2e = 1.4426950408889634f;

template <int kD, int kBiasMode>
struct Bwd3Cfg {
    static_assert(kD == 16 || kD == 32 || kD == 64, "v3 backward covers head dims 16, 32, 64");
    static constexpr int kRowBytes = kD * 2;
    static constexpr int kTileBytes = 128 * kD * 2;
    static constexpr int kSubTileBytes = kSub * kD * 2;
    static constexpr uint32_t kSwizzle = kRowBytes == 128 ? kSwz128 : (kRowBytes == 64 ? kSwz64 : kSwz32);
    // Q / dO ring: one slot = the 64 query rows of two sub-tiles (Q rows + dO rows + their row statistics: three TMA operations).
    // A slot is released when the dV / dK MMAs of both its sub-tiles complete -- after the compute warps are through with them, so
    // a slot lives ~5 700 cycles (TMA issue + flight 1 500, S^T MMAs 500, math 2 500, dV / dK MMAs 700) and the ring depth sets the
    // ring must hold > 2 tiles.  5 slots; 7 without a bias, where the 32 KB of bias staging / band are not needed (measured: no
    // faster -- the ring is not the pacemaker any more; staging the bias inside the dS^T boxes to get 7 slots with a bias
    // cost 100 us: the boxes then have to be free a whole tile earlier).  (A ring of
    // two whole tiles left the first sub-tile of every other tile waiting ~2 400 cycles for its TMA; a ring of eleven 32-row
    // slots made the single producer lane the pacemaker of the whole CTA: 4 operations + a probe = ~650 cycles per sub-tile,
    // 2 600 per tile against 1 500 of math -- profiles/r2c_bwd_v3_timeline_*.)
    static constexpr int kSlots = kBiasMode == 0 ? 6 : 5;
    static constexpr int kSlotRows = 2 * kSub;
    static constexpr int kSlotBytes = 2 * kSubTileBytes;
    static constexpr int kK = 0;
    static constexpr int kQ = kK + 2 * kTileBytes;                     // K is double-buffered over the work items of the CTA
    static constexpr int kDO = kQ + kSlots * kSlotBytes;
    static constexpr int kDS = kDO + kSlots * kSlotBytes;              // [2 tiles][4 boxes] dS^T
    static constexpr int kDQ = kDS + 2 * kNSub * kBoxBytes;            // dQ staging tile [128][D] io dtype (prologue: the V tile)
    static constexpr int kBand = kDQ + kTileBytes;                     // relative-position band (mode 3) / bias staging (mode 1), 16 KB
    static constexpr int kStats = kBand + (kBiasMode == 0 ? 0 : 16384);                     // [kSlots][2][64] fp32: -L*log2e, -delta of the slot's queries
    static constexpr int kBars = kStats + kSlots * 2 * kSlotRows * 4;
    static constexpr int kNumBars = 2 + 2 * kSlots + 5 * kNSub + 3 + 2 + 2 + 1 + 2;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static_assert(kTotal <= 232448, "shared memory budget");
    static_assert(kTileBytes % 1024 == 0 && kSubTileBytes % 1024 == 0, "swizzled tiles need 1024-byte alignment");
    // S^T / dP^T: two 32-column buffers each (sub-tile t uses buffer t & 1; released as soon as the compute warps have loaded
    // them); P^T / dS^T: four packed 16-column buffers each (one per compute warpgroup; released by the dV / dK MMAs)
    static constexpr int kColS = 0;        // + (t & 1) * 32
    static constexpr int kColDP = 64;      // + (t & 1) * 32
    static constexpr int kColP = 128;      // + j * 16
    static constexpr int kColDS = 192;     // + j * 16
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

// P^T and dS^T for 16 query columns of one key row (half a sub-tile).
//   sr / dr : S^T, dP^T accumulators (fp32 bits), column c <-> query m0 + c
//   nl / nd : shared-memory rows of -L * log2e and -delta for those 16 queries (broadcast 16-byte loads; the TMA producer
//             copies them next to the Q / dO rows of the sub-tile)
//   bias    : kBiasMode 1: bw[c / 2] holds the packed 16-bit bias values of columns c, c + 1
//             kBiasMode 3: band pointer such that bias(c) = bp[-c]  (or the constant bconst, already times log2e, when kConst)
//   vis_lo  : [kMask] columns c < vis_lo are masked (causal: query before the key; whole row when the key is out of range)
template <bool kBf16, int kBiasMode, bool kMask, bool kConst, bool kSum>
__device__ __forceinline__ void v3_chunk(const uint32_t (&sr)[16], const uint32_t (&dr)[16], const float* nl, const float* nd,
                                         const uint32_t* bw, const float* bp, float bconst, float scale_log2, int vis_lo,
                                         uint32_t* pp, uint32_t* dd, float& ds_sum) {
#ifdef B200T5_DBG_SKIP_MATH
    for (int c = 0; c < 8; ++c) { pp[c] = sr[2 * c]; dd[c] = dr[2 * c + 1]; }
    return;
#endif
    const f32x2 sc2 = f2_pack(scale_log2, scale_log2);
    const f32x2 l2e2 = f2_pack(kLog2e, kLog2e);
    f32x2 sum2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 2; ++g) {                       // 8 columns per group
        const float4 nla = *reinterpret_cast<const float4*>(nl + g * 8), nlb = *reinterpret_cast<const float4*>(nl + g * 8 + 4);
        const float4 nda = *reinterpret_cast<const float4*>(nd + g * 8), ndb = *reinterpret_cast<const float4*>(nd + g * 8 + 4);
        const float nlv[8] = {nla.x, nla.y, nla.z, nla.w, nlb.x, nlb.y, nlb.z, nlb.w};
        const float ndv[8] = {nda.x, nda.y, nda.z, nda.w, ndb.x, ndb.y, ndb.z, ndb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = g * 8 + 2 * e;
            f32x2 t = f2_pack(nlv[2 * e], nlv[2 * e + 1]);
            if (kBiasMode == 1) {
                const float2 bf = unpack2<kBf16>(bw[c / 2]);
                t = f2_fma(f2_pack(bf.x, bf.y), l2e2, t);
            } else if (kBiasMode == 3) {
                if (kConst) t = f2_add(t, f2_pack(bconst, bconst));
                else t = f2_fma(f2_pack(bp[-c], bp[-c - 1]), l2e2, t);
            }
            float a0, a1;
            f2_unpack(f2_fma(f2_pack_bits(sr[c], sr[c + 1]), sc2, t), a0, a1);
#ifdef B200T5_DBG_SKIP_EXP
            float p0 = a0 * 1e-3f, p1 = a1 * 1e-3f;
#else
            float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
#endif
            if (kMask) {
                if (c < vis_lo) p0 = 0.f;
                if (c + 1 < vis_lo) p1 = 0.f;
            }
            const f32x2 p2 = f2_pack(p0, p1);
            const f32x2 ds2 = f2_mul(p2, f2_add(f2_pack_bits(dr[c], dr[c + 1]), f2_pack(ndv[2 * e], ndv[2 * e + 1])));
            float g0, g1;
            f2_unpack(ds2, g0, g1);
#ifdef B200T5_DBG_TRUNC_PACK
            pp[c / 2] = __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
            dd[c / 2] = __byte_perm(__float_as_uint(g0), __float_as_uint(g1), 0x7632);
#else
            pp[c / 2] = pack2<kBf16>(p0, p1);
            dd[c / 2] = pack2<kBf16>(g0, g1);
#endif
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
__device__ long long g_bwd3_item_ts[8][20];    // [CTA 0..7][item boundary]: clock64 of compute thread 0 at the start of every item (+ the end)
__device__ long long g_bwd3_ts[8][32][8];      // [role: wg0..wg3, mma B, drain, mma A, producer][iteration][slot]
#define BWD3_TS(role, k, slot)                                                          \
    do {                                                                                \
        if (blockIdx.x == 7 && it == 1 && (k) < 32) g_bwd3_ts[role][k][slot] = clock64();   \
    } while (0)
#else
#define BWD3_TS(role, k, slot) do { } while (0)
#endif

constexpr int kThreads = 896;        // 28 warps (warps are allocated four at a time: 25 would cost as much)
constexpr int kWarpDrain0 = 16, kWarpTma = 20, kWarpMmaA = 21, kWarpMmaB = 22, kWarpC = 23, kWarpMmaA1 = 24, kWarpMmaB1 = 25, kWarpMmaA2 = 26,
              kWarpMmaA3 = 27;

template <int kD, bool kBf16, int kBiasMode, bool kCausal>
__global__ void __launch_bounds__(kThreads, 1)     // 72 registers per thread at launch: a pool of 896 x 72 = 64 512 for setmaxnreg
attn_bwd_kernel_v3(const __grid_constant__ AttnBwdKernelParams p) {
    using C = Bwd3Cfg<kD, kBiasMode>;
    extern __shared__ __align__(1024) uint8_t smem[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- PERSISTENT: the CTA walks work items (batch, head, key block) item = blockIdx.x, + gridDim.x, ...; barriers, TMEM and
    //      the ring live across items, every role keeps two running counters (it: non-empty items done, kb: tiles done) from
    //      which all barrier parities follow.  What this buys over one CTA per item: no per-item launch / setup / teardown, and
    //      the tail of item i (last dQ drain, dK / dV write-out) overlaps the prologue of item i + 1 (K, V -> TMEM, first S^T).
    //      Work decode: batch fastest (bias tiles shared in L2), long key blocks first when causal ----
    const int nnb = p.num_n_blocks;
    const int pseq = p.N - p.M;
    const int n_items = p.B * p.H * nnb;
    struct Item { int b, nb, h, col0, i_start, n_iter; };
    auto decode = [&](int item) {
        Item w;
        int bid = item;
        w.b = bid % p.B;
        bid /= p.B;
        w.nb = kCausal ? (bid % nnb) : (nnb - 1 - bid % nnb);
        w.h = bid / nnb;
        w.col0 = w.nb * kBN;
        w.i_start = 0;
        if (kCausal) {
            const int first_row = w.col0 - pseq;                // first query row that sees key col0
            w.i_start = first_row <= 0 ? 0 : first_row / kBM;
        }
        w.n_iter = p.num_m_blocks > w.i_start ? p.num_m_blocks - w.i_start : 0;
        return w;
    };
    int it = 0, kb = 0;                                         // non-empty items / tiles this CTA has been through
#define B200T5_ITEM_BEGIN                                                                                                   \
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {                                                        \
        const Item w_ = decode(item);                                                                                       \
        const int b = w_.b, nb = w_.nb, h = w_.h, col0 = w_.col0, i_start = w_.i_start, n_iter = w_.n_iter, T = kNSub * n_iter; \
        (void)b; (void)nb; (void)h; (void)col0; (void)i_start; (void)T;
#define B200T5_ITEM_END                                                                                                     \
        kb += n_iter;                                                                                                       \
        ++it;                                                                                                               \
    }
    const Item first_item = decode(blockIdx.x);
    auto next_item = [&](int item) {                                  // the next item of this CTA that has work (n_iter == 0: none)
        Item wn;
        wn.n_iter = 0;
        for (int nxt = item + static_cast<int>(gridDim.x); nxt < n_items; nxt += gridDim.x) {
            wn = decode(nxt);
            if (wn.n_iter > 0) break;
        }
        return wn;
    };

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
    uint64_t* k_full = bars;
    uint64_t* kt_ready = bars + 1;
    uint64_t* qdo_full = bars + 2;
    uint64_t* qdo_empty = qdo_full + C::kSlots;
    uint64_t* sdp_full = qdo_empty + C::kSlots;         // [4] S^T / dP^T of buffer j written          (MMA A -> compute j)
    // [2][4] P^T / dS^T (and the dS^T box) of sub-tile j written, one set per tile parity (compute j -> MMA B, warp C).  An
    // mbarrier probe only tells phases of opposite parity apart, so no waiter may fall two phases behind.  The B warps are tied
    // to the warpgroups through pds_free, but warp C is not: a fast warpgroup signalled tiles 0 and 1 before C had looked at
    // tile 0, and C waited for ever (found on B200 with the in-kernel bias, whose constant sub-tiles make the warpgroups
    // uneven).  With one barrier set per tile parity a warpgroup would have to be two tiles ahead of C, which box_free forbids.
    uint64_t* pds_full = sdp_full + kNSub;
    uint64_t* pds_free = pds_full + 2 * kNSub;              // [4] dV, dK MMAs reading P^T / dS^T of j completed (MMA B -> compute j)
    uint64_t* s_empty = pds_free + kNSub;               // [4] compute j has S^T and dP^T of its sub-tile in registers  (compute j -> MMA A)
    uint64_t* dq_full = s_empty + kNSub;
    uint64_t* dq_empty = dq_full + 1;
    uint64_t* all_done = dq_empty + 1;                  // every MMA of the CTA completed (single phase: the epilogue's gate; B and C commit)
    uint64_t* b_turn = all_done + 3;                    // [2] B0 <-> B1 token: the dV / dK MMAs are issued in sub-tile order (bitwise reproducible sums)
    uint64_t* kbuf_free = all_done + 6;                 // [2] every dQ MMA that reads K buffer i completed (warp C -> producer)
    uint64_t* sdp_done = all_done + 5;                  // every S^T / dP^T MMA of the item completed (four A warps commit; warpgroup 0 waits)
    uint64_t* box_free = all_done + 1;                  // [2] dS^T boxes of tile parity: dQ MMAs done + TMA reduce reads done (C -> compute)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kTmemSlot);

    if (warp == kWarpTma && lane == 0) {
        // K and V are on the critical path of the prologue (TMA -> shared memory -> TMEM -> first S^T): their barrier is
        // initialised by the producer lane itself and the loads leave before the CTA-wide setup barrier
        // (V lands in the dQ staging tile, which is idle until the first dQ is drained: the drain warpgroup copies both
        //  tiles into TMEM)
        mbar_init(k_full, 1);
        mbar_init(qdo_full + 0, 1);
        fence_mbar_init();
        fence_proxy_async_smem();          // the barrier words were written through the generic proxy; TMA completes on them through the async one
        tma_prefetch_desc(&p.map_k);
        tma_prefetch_desc(&p.map_v);
        if (first_item.n_iter > 0) {
            mbar_arrive_expect_tx(k_full, C::kTileBytes);
            tma_load_4d(smem + C::kK, &p.map_k, k_full, 0, first_item.col0, first_item.h, first_item.b);
            // V travels through the Q / dO ring like a half tile (ring position 0 = slot 0): keys 0-63 in the slot's Q rows, keys
            // 64-127 in its dO rows
            mbar_arrive_expect_tx(qdo_full + 0, 2 * C::kSlotBytes);
            tma_load_4d(smem + C::kQ, &p.map_v, qdo_full + 0, 0, first_item.col0, first_item.h, first_item.b);
            tma_load_4d(smem + C::kDO, &p.map_v, qdo_full + 0, 0, first_item.col0 + C::kSlotRows, first_item.h, first_item.b);
        }
    }
    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        mbar_init(kt_ready, 4);
        for (int i = 0; i < C::kSlots; ++i) {
            if (i > 0) mbar_init(qdo_full + i, 1);        // (slot 0: initialised by the producer lane above)
            mbar_init(qdo_empty + i, 2);                  // one commit from each B warp (the slot's even and odd sub-tile); V: two arrivals of warpgroup 0
        }
        mbar_init(all_done, 3);                           // the two B warps and warp C
        mbar_init(box_free + 0, 2);
        mbar_init(box_free + 1, 2);
        mbar_init(b_turn + 0, 1);
        mbar_init(b_turn + 1, 1);
        mbar_init(sdp_done, 4);
        mbar_init(kbuf_free + 0, 1);
        mbar_init(kbuf_free + 1, 1);
        for (int i = 0; i < kNSub; ++i) {
            mbar_init(sdp_full + i, 1);
            mbar_init(pds_full + i, 128);                 // every thread of compute warpgroup i arrives by itself
            mbar_init(pds_full + kNSub + i, 128);
            mbar_init(pds_free + i, 1);
            mbar_init(s_empty + i, 4);
        }
        mbar_init(dq_full, 1);
        mbar_init(dq_empty, 4);
        fence_mbar_init();
    }
    if (warp == kWarpMmaA) tmem_alloc<512>(tmem_slot);
    if (warp == kWarpTma && lane == 0) {
        tma_prefetch_desc(&p.map_q);
        tma_prefetch_desc(&p.map_do);
        tma_prefetch_desc(&p.map_dq);
        if (kBiasMode != 0) tma_prefetch_desc(&p.map_ds);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= kWarpTma) {
        // =============================== control warps (20..27) ===============================
        setmaxnreg_dec<32>();
        constexpr uint32_t sbo = 8 * C::kRowBytes;
        constexpr uint32_t hi_op = sdesc_hi(sbo, C::kSwizzle);       // Q, dO, K tiles (either major)
        if (warp == kWarpTma && lane == 0) {
            // ---- per item: K, V (the first item's were requested before the setup barrier); then the Q / dO ring ----
            B200T5_ITEM_BEGIN
            if (n_iter == 0) continue;
            // Ring positions: item `it` takes position 2 kb + it for its V tile and 2 kb + it + 1 + u for its half tile u.  K and V
            // of an item are requested right after the last half tile of the item before it (below), so that warpgroup 0 can put
            // them into TMEM while that item is still finishing; only the CTA's first item asks for them here (if the setup code
            // above has not already).
            if (it == 0 && item != static_cast<int>(blockIdx.x)) {
                mbar_arrive_expect_tx(qdo_full + 0, 2 * C::kSlotBytes);
                tma_load_4d(smem + C::kQ, &p.map_v, qdo_full + 0, 0, col0, h, b);
                tma_load_4d(smem + C::kDO, &p.map_v, qdo_full + 0, 0, col0 + C::kSlotRows, h, b);
                mbar_arrive_expect_tx(k_full, C::kTileBytes);
                tma_load_4d(smem + C::kK, &p.map_k, k_full, 0, col0, h, b);
            }
            const float* stat_bh = p.nl + ((int64_t)b * p.H + h) * (2 * (int64_t)p.m_pad);
            for (int u = 0; u < 2 * n_iter; ++u) {
                const int ug = 2 * kb + it + 1 + u;                     // ring position over the whole CTA
                const int s = ug % C::kSlots;
                const int m0 = (i_start + (u >> 1)) * kBM + (u & 1) * C::kSlotRows;
                BWD3_TS(7, u, 0);
                mbar_wait_producer(qdo_empty + s, ((ug / C::kSlots) & 1) ^ 1);
                BWD3_TS(7, u, 1);
                mbar_arrive_expect_tx(qdo_full + s, 2 * C::kSlotBytes + 2 * C::kSlotRows * 4);
                tma_load_4d(smem + C::kQ + s * C::kSlotBytes, &p.map_q, qdo_full + s, 0, m0, h, b);
                tma_load_4d(smem + C::kDO + s * C::kSlotBytes, &p.map_do, qdo_full + s, 0, m0, h, b);
                // row statistics of the 64 queries: [-L * log2e | -delta], one 512-byte record per 64 padded rows
                bulk_load_1d(reinterpret_cast<float*>(smem + C::kStats) + s * (2 * C::kSlotRows), stat_bh + (m0 / C::kSlotRows) * (2 * C::kSlotRows),
                             2 * C::kSlotRows * 4, qdo_full + s);
            }
            {
                // ---- K and V of the CTA's next item.  V: the ring position after this item's last half tile (as soon as a slot
                //      is free).  K: the other shared-memory buffer, free since every dQ MMA of the item BEFORE this one completed.
                const Item wn = next_item(item);
                if (wn.n_iter > 0) {
                    const int pv = 2 * (kb + n_iter) + it + 1, sv = pv % C::kSlots;
                    mbar_wait_producer(qdo_empty + sv, ((pv / C::kSlots) & 1) ^ 1);
                    mbar_arrive_expect_tx(qdo_full + sv, 2 * C::kSlotBytes);
                    tma_load_4d(smem + C::kQ + sv * C::kSlotBytes, &p.map_v, qdo_full + sv, 0, wn.col0, wn.h, wn.b);
                    tma_load_4d(smem + C::kDO + sv * C::kSlotBytes, &p.map_v, qdo_full + sv, 0, wn.col0 + C::kSlotRows, wn.h, wn.b);
                    // (a barrier per K buffer, not all_done: with short items this lane can be late enough for all_done to have
                    //  completed TWO further phases, which a parity probe cannot tell from none -- found as a deadlock in the
                    //  causal full-shape test)
                    if (it > 0) mbar_wait(kbuf_free + ((it + 1) & 1), ((it - 1) >> 1) & 1);
                    // k_full takes ONE arrival per phase: K of this item (requested an item ago) must have landed before the
                    // next phase is armed.  With a one-tile item this lane gets here within a few hundred cycles of that
                    // request; the second arrival on the open phase was a hardware fault (mbarrier arrival-count underflow),
                    // seen as sporadic "unspecified launch failure" on the causal no-bias shapes.
                    mbar_wait(k_full, it & 1);
                    // ... and every warp of warpgroup 0 must have SEEN that phase (they arrive on kt_ready after their copy): a
                    // warp that loses the scheduler before its k_full probe would otherwise find the barrier two phases on
                    // and wait for ever (found by the protocol model, tests/test_bwd_protocol_model.py; long complete here)
                    mbar_wait(kt_ready, it & 1);
                    mbar_arrive_expect_tx(k_full, C::kTileBytes);
                    tma_load_4d(smem + C::kK + ((it + 1) & 1) * C::kTileBytes, &p.map_k, k_full, 0, wn.col0, wn.h, wn.b);
                }
            }
            B200T5_ITEM_END
        } else if (warp == kWarpMmaA || warp == kWarpMmaA1 || warp == kWarpMmaA2 || warp == kWarpMmaA3) {
            // ---- MMA warp A: S^T = K Q^T and dP^T = V dO^T of every sub-tile (A operands K, V in TMEM) ----
            // (the whole warp runs the loop so that descriptor arithmetic stays warp-uniform; one elected lane issues)
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, 128, kSub, false, false);
            const uint32_t q_lo0 = sdesc_lo(smem_u32(smem + C::kQ), 16);
            const uint32_t do_lo0 = sdesc_lo(smem_u32(smem + C::kDO), 16);
            const uint32_t tm_kt = tmem_base + C::kColKt;
            const uint32_t tm_vt = tmem_base + C::kColVt;
            const int j_mine = warp == kWarpMmaA ? 0 : (warp == kWarpMmaA1 ? 1 : (warp == kWarpMmaA2 ? 2 : 3));
            B200T5_ITEM_BEGIN
            if (n_iter == 0) continue;
            mbar_wait(kt_ready, it & 1);
            for (int t = j_mine; t < T; t += kNSub) {
                const int j = t & 3;
                const int tg = 4 * kb + t;                              // sub-tile count over the whole CTA
                if (lane == 0) BWD3_TS(6, t, 0);
                const int u = (tg >> 1) + it + 1;                       // ring position of the half tile
                mbar_wait(qdo_full + (u % C::kSlots), (u / C::kSlots) & 1);
                if (lane == 0) BWD3_TS(6, t, 1);
                if (tg >= 2) {
                    // buffer t & 1 held sub-tile tg - 2 (warpgroup (j + 2) & 3): in registers by now?
                    const int jp = (j + 2) & 3;
                    const uint32_t par = ((tg - 2) >> 2) & 1;
                    mbar_wait(s_empty + jp, par);           // (S^T and dP^T are loaded together: one barrier)
                }
                tc_fence_after();
                if (lane == 0) BWD3_TS(6, t, 2);
                const uint32_t so = ((u % C::kSlots) * 2 + (t & 1)) * (C::kSubTileBytes >> 4);
                const uint32_t tm_s = tmem_base + C::kColS + (t & 1) * kSub;
                const uint32_t tm_dp = tmem_base + C::kColDP + (t & 1) * kSub;
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_ts2(tm_s, tm_kt + kk * 8, q_lo0 + so + kk * 2, hi_op, idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)
                        umma_ts2(tm_dp, tm_vt + kk * 8, do_lo0 + so + kk * 2, hi_op, idesc_s, kk > 0 ? 1u : 0u);
                    umma_commit(sdp_full + j);
                    if (t + kNSub >= T) umma_commit(sdp_done);      // this warp's last sub-tile of the item: K, V in TMEM are done with
                }
                __syncwarp();
                if (lane == 0) BWD3_TS(6, t, 3);
            }
            B200T5_ITEM_END
        } else if (warp == kWarpMmaB || warp == kWarpMmaB1) {
            // ---- MMA warps B0 / B1: dV += P^T dO, dK += dS^T Q of the even / odd sub-tiles (A operands from TMEM).  The
            //      accumulators were zero-filled by the drain warpgroup (kt_ready): the two warps' first MMAs come in any order ----
            const bool leader = elect_one();
            constexpr uint32_t idesc_dkv = make_idesc(kBf16, 128, kD, false, true);    // A tmem, B MN-major
            const uint32_t q_mn_lo0 = sdesc_lo(smem_u32(smem + C::kQ), C::kSubTileBytes);
            const uint32_t do_mn_lo0 = sdesc_lo(smem_u32(smem + C::kDO), C::kSubTileBytes);
            const uint32_t tm_dv = tmem_base + C::kColDV;
            const uint32_t tm_dk = tmem_base + C::kColDK;
            B200T5_ITEM_BEGIN
            if (n_iter == 0) continue;
            for (int t = (warp == kWarpMmaB ? 0 : 1); t < T; t += 2) {
                const int j = t & 3;
                const int tg = 4 * kb + t, kg = kb + (t >> 2);          // counts over the whole CTA
                if (lane == 0) BWD3_TS(4, t, 0);
                mbar_wait(pds_full + (kg & 1) * kNSub + j, (kg >> 1) & 1);
                tc_fence_after();
                if (lane == 0) BWD3_TS(4, t, 1);
                const int slot = ((tg >> 1) + it + 1) % C::kSlots;      // ring position of the half tile
                const uint32_t so = (slot * 2 + (t & 1)) * (C::kSubTileBytes >> 4);
                const uint32_t tm_p = tmem_base + C::kColP + j * 16;         // P^T  (packed 16-bit pairs)
                const uint32_t tm_ds = tmem_base + C::kColDS + j * 16;       // dS^T (packed 16-bit pairs)
                // the other B warp has issued sub-tile t - 1: both accumulate into the same dV / dK columns, and a fixed order
                // of the fp32 additions makes dK, dV bitwise reproducible (strict alternation: nobody lags a phase)
                if (tg > 0) mbar_wait(b_turn + (tg & 1), ((tg - 1) >> 1) & 1);
                if (leader) {
                    // (K dimension = the 32 queries of this sub-tile)
#pragma unroll
                    for (int kk = 0; kk < kSub / 16; ++kk)
                        umma_ts2(tm_dv, tm_p + kk * 8, do_mn_lo0 + so + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dkv, 1u);
#pragma unroll
                    for (int kk = 0; kk < kSub / 16; ++kk)
                        umma_ts2(tm_dk, tm_ds + kk * 8, q_mn_lo0 + so + ((kk * 16 * C::kRowBytes) >> 4), hi_op, idesc_dkv, 1u);
                    umma_commit(pds_free + j);
                    // Q / dO slot: the S^T / dP^T MMAs of warp A that read it completed before the compute warps could produce
                    // the P^T this warp just consumed, so this commit covers every reader of the slot
                    umma_commit(qdo_empty + slot);
                    if (t >= T - 2) umma_commit(all_done);          // the last sub-tile of this warp in this item (T is a multiple of 4)
                    mbar_arrive(b_turn + ((tg + 1) & 1));
                }
                __syncwarp();
                if (lane == 0) BWD3_TS(4, t, 3);
            }
            B200T5_ITEM_END
        } else if (warp == kWarpC) {
            // ---- warp C: dS^T boxes -> dBias surface (TMA reduce-add / store), and dQ = dS K once per tile ----
            const bool leader = elect_one();
            constexpr uint32_t idesc_dq = make_idesc(kBf16, 128, kD, true, true);      // A, B MN-major
            constexpr uint32_t hi_ds = sdesc_hi(8 * 64, kSwz64);                       // dS^T boxes: 64-byte rows
            const uint32_t k_mn_lo = sdesc_lo(smem_u32(smem + C::kK), C::kTileBytes);
            const uint32_t ds_mn_lo0 = sdesc_lo(smem_u32(smem + C::kDS), kBoxBytes);   // A = dS (M = queries, 4 boxes of 32)
            const uint32_t tm_dq = tmem_base + C::kColDQ;
            const bool rpe_skip = kBiasMode == 3 && p.rpe.dconst != nullptr;
            B200T5_ITEM_BEGIN
            if (n_iter == 0) continue;
            const int g_ds = b % p.ds_groups;
            for (int k = 0; k < n_iter; ++k) {
                const int kg = kb + k;                                  // tile count over the whole CTA
                // second arrival for the boxes of the previous tile: its TMA group (committed a moment ago) has read them.  Done
                // first thing, so that the boxes are free again one whole tile before their next use.
                if (kg > 0 && leader) {
                    bulk_wait_group_read<0>();
                    mbar_arrive(box_free + ((kg - 1) & 1));
                }
                __syncwarp();
                for (int j = 0; j < kNSub; ++j) {
                    mbar_wait(pds_full + (kg & 1) * kNSub + j, (kg >> 1) & 1);
                    if (kBiasMode != 0 && leader) {
                        const int m0 = (i_start + k) * kBM + j * kSub;
                        bool skip = false;
                        if (kBiasMode == 3 && rpe_skip) {          // sub-tiles beyond a constant end of the table keep dS in the kernel
                            const int rel_min = col0 - (m0 + kSub - 1), rel_max = col0 + (kBN - 1) - m0;
                            skip = rel_max <= p.rpe.const_lo || rel_min >= p.rpe.const_hi;
                        }
#ifdef B200T5_DBG_SKIP_DS_TMA
                        skip = true;
#endif
                        if (!skip) {
                            const uint8_t* box = smem + C::kDS + ((kg & 1) * kNSub + j) * kBoxBytes;
                            if (p.ds_use_reduce) tma_reduce_add_4d(&p.map_ds, box, m0, col0, h, g_ds);
                            else tma_store_4d(&p.map_ds, box, m0, col0, h, g_ds);
                        }
                    }
                }
                if (leader) bulk_commit_group();                    // one group per tile (possibly empty)
                // (K of this item is in shared memory: warpgroup 0 copied it into TMEM before the first S^T of the item could be
                //  issued, and this tile's dS^T exists.  No wait on k_full here: with short items the NEXT item's K may already
                //  have completed a further phase of that barrier, and a parity probe cannot tell phases two apart.)
                if (kg > 0) mbar_wait(dq_empty, (kg - 1) & 1);     // the previous dQ tile has been drained out of TMEM
                tc_fence_after();
                if (lane == 0) BWD3_TS(4, 4 * k + 3, 2);
                const uint32_t ds_mn_lo = ds_mn_lo0 + (kg & 1) * ((kNSub * kBoxBytes) >> 4);
                if (leader) {
                    // dQ_tile = dS K   (K dimension = the 128 keys of this CTA; A = dS^T boxes read MN-major)
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk)
                        umma_ss2(tm_dq, ds_mn_lo + kk * ((16 * 64) >> 4), hi_ds, k_mn_lo + (it & 1) * (C::kTileBytes >> 4) + ((kk * 16 * C::kRowBytes) >> 4), hi_op,
                                 idesc_dq, kk > 0 ? 1u : 0u);
                    umma_commit(dq_full);
                    umma_commit(box_free + (kg & 1));               // first of the two arrivals: the MMAs have read the boxes
                    if (k == n_iter - 1) {
                        umma_commit(all_done);
                        umma_commit(kbuf_free + (it & 1));          // the item's K buffer is done with
                    }
                }
                __syncwarp();
            }
            B200T5_ITEM_END
            if (leader) bulk_wait_group_read<0>();                  // shared memory must outlive the reads; the writes complete by themselves
        }
    } else if (warp >= kWarpDrain0) {
        // =============================== drain warpgroup (16..19) ===============================
        setmaxnreg_dec<56>();
        const int r = (warp & 3) * 32 + lane;                 // TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tm_dq = tmem_base + lane_off + C::kColDQ;
        uint8_t* stage_row = smem + C::kDQ + r * (kD * 2);
        B200T5_ITEM_BEGIN
        if (n_iter == 0) continue;
        const int dq_c3 = (nb % p.dq_groups) * p.B + b;       // (group, batch) slice of the dQ surface
        for (int k = 0; k < n_iter; ++k) {
            if (r == 0) BWD3_TS(5, k, 0);
            mbar_wait(dq_full, (kb + k) & 1);
            tc_fence_after();
            if (r == 0) BWD3_TS(5, k, 1);
            // (56 registers per thread here: 16 columns at a time, keep only the packed words)
            uint32_t qp[kD / 2];
#pragma unroll
            for (int c0 = 0; c0 < kD; c0 += 16) {
                uint32_t q[16];
                tmem_ld16(tm_dq + c0, q);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 2) qp[(c0 + i) / 2] = pack2<kBf16>(__uint_as_float(q[i]), __uint_as_float(q[i + 1]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);             // the dQ columns may be overwritten by tile k + 1
            if (r == 0) bulk_wait_group_read<0>();            // staging tile: the previous reduce has read it
            named_bar_sync(5, 128);
            if (r == 0) BWD3_TS(5, k, 3);
#pragma unroll
            for (int i = 0; i < kD; i += 8) {
                const int c16 = i / 8;
                const int off = (kD == 64) ? ((c16 ^ (r & 7)) << 4) : (c16 << 4);
                *reinterpret_cast<uint4*>(stage_row + off) = make_uint4(qp[i / 2], qp[i / 2 + 1], qp[i / 2 + 2], qp[i / 2 + 3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(5, 128);
            if (r == 0) {
#ifndef B200T5_DBG_SKIP_DQ_TMA
                tma_reduce_add_4d(&p.map_dq, smem + C::kDQ, 0, (i_start + k) * kBM, h, dq_c3);
#endif
                bulk_commit_group();
            }
            if (r == 0) BWD3_TS(5, k, 2);
        }
        B200T5_ITEM_END
        if (r == 0) bulk_wait_group_read<0>();                // shared memory must outlive the reads
    } else {
        // =============================== compute warpgroups 0..3 (warps 0..15) ===============================
        setmaxnreg_inc<96>();     // the CTA owns 896 x 72 = 64 512 registers (its launch allocation): 512 x 96 + 128 x 56 (drain) + 256 x 32 (control)
        const int wg = warp >> 2;                             // == j: the sub-tile of every tile / the buffer this warpgroup owns
        const int r = (warp & 3) * 32 + lane;                 // key row in the block == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tm_s = tmem_base + lane_off + C::kColS + (wg & 1) * kSub;      // sub-tile t = 4k + wg uses buffer t & 1
        const uint32_t tm_dp = tmem_base + lane_off + C::kColDP + (wg & 1) * kSub;
        const uint32_t tm_p = tmem_base + lane_off + C::kColP + wg * 16;
        const uint32_t tm_ds = tmem_base + lane_off + C::kColDS + wg * 16;
        const int rx = (r >> 1) & 3;                          // 64-byte swizzle: 16-byte chunk ^= (row / 2) % 4
        const float scale_log2 = p.sm_scale * kLog2e;

        const float* band = reinterpret_cast<const float*>(smem + C::kBand);   // [bias mode 3]
        int band_head = -1;
        bool store_pending = false;                           // this warpgroup's dK / dV TMA store of the previous item may still be reading its box
        const bool rpe_skip = kBiasMode == 3 && p.rpe.dconst != nullptr;
        // The 64 bytes of bias this thread needs for its next sub-tile are copied global -> shared with cp.async (no registers:
        // they would be live across a whole sub-tile) into the band's shared memory (mode 3 is the other user): [wg][chunk][row].
        // Half a sub-tile (16 queries = 32 bytes per thread) at a time, [wg][2 chunks][128 rows]: the copy of a half is issued right
        // after the half before it has been read, i.e. ~900 cycles of arithmetic ahead of its use (a whole sub-tile ahead would
        // need 32 KB, which is the fifth ring slot).
        uint4* const bias_stage = reinterpret_cast<uint4*>(smem + C::kBand) + wg * (2 * 128) + r;

        // K and V rows of a key block, shared memory (TMA, swizzled rows) -> TMEM (the A operands of S^T and dP^T); warpgroup 0.
        // `item_no` = the (non-empty) item they belong to: K sits in buffer item_no & 1 (k_full phase item_no), V in the ring
        // slot of position `pv`.  (From global memory straight to TMEM this cost ~4 000 cycles per item.)
        auto kv_to_tmem = [&](int item_no, int pv) {
            const int sv = pv % C::kSlots;
            mbar_wait(k_full, item_no & 1);
            mbar_wait(qdo_full + sv, (pv / C::kSlots) & 1);
            const int sw = kD == 64 ? (r & 7) : (kD == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));   // 16-byte chunk ^= f(row)
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const uint8_t* row = which == 0 ? smem + C::kK + (item_no & 1) * C::kTileBytes + r * C::kRowBytes
                                                : smem + (r < C::kSlotRows ? C::kQ : C::kDO) + sv * C::kSlotBytes + (r & (C::kSlotRows - 1)) * C::kRowBytes;
                const uint32_t tm_dst = tmem_base + lane_off + (which == 0 ? C::kColKt : C::kColVt);
#pragma unroll
                for (int i = 0; i < kD / 16; ++i) {                    // 8 words (16 elements) at a time
                    const uint4 a = *reinterpret_cast<const uint4*>(row + (((2 * i) ^ sw) << 4));
                    const uint4 c = *reinterpret_cast<const uint4*>(row + (((2 * i + 1) ^ sw) << 4));
                    const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
                    tmem_st8(tm_dst + 8 * i, w);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async_smem();                                  // the slot goes back to the TMA producer
            __syncwarp();
            if (lane == 0) mbar_arrive(kt_ready);
            named_bar_sync(7, 128);                                    // every row of V has been read
            if (r == 0) {
                mbar_arrive(qdo_empty + sv);
                mbar_arrive(qdo_empty + sv);
            }
        };

        B200T5_ITEM_BEGIN
#ifdef B200T5_BWD_TIMING
        if (threadIdx.x == 0 && blockIdx.x < 8 && it < 19) g_bwd3_item_ts[blockIdx.x][it] = clock64();
#endif
        const int gn = col0 + r;                              // global key index
        const bool key_ok = gn < p.N;
        if (kBiasMode == 3 && n_iter > 0 && h != band_head) {
            // (every compute warp left the previous item through the barrier at its end: nobody reads the old band any more)
            float* dst = reinterpret_cast<float*>(smem + C::kBand);
            const float* src = p.rpe.band + (int64_t)h * p.rpe.band_len;
            for (int i = threadIdx.x; i < p.rpe.band_len; i += 512) dst[i] = __ldg(src + i);
            named_bar_sync(6, 512);
            band_head = h;
        }
        float ds_const_lo = 0.f, ds_const_hi = 0.f;

        // dense bias: repacked copy [bh][key block][32-query block][4][128 keys][8 queries] (16-bit)
        const uint4* bias_blk = nullptr;
        if (kBiasMode == 1) {
            const int hb_n = p.bias_h_bcast ? 1 : p.H;
            const int64_t bh = (int64_t)(p.bias_b_bcast ? 0 : b) * hb_n + (p.bias_h_bcast ? 0 : h);
            const int64_t n_mb = kNSub * p.num_m_blocks;              // the copy covers whole 128-query tiles
            bias_blk = reinterpret_cast<const uint4*>(p.bias) + ((bh * nnb + nb) * n_mb) * (4 * 128) + r;
        }
        auto load_bias = [&](int kk, int half) {                // kk: tile (iteration), half: 16-query half of the sub-tile
            const uint4* src = bias_blk + (int64_t)(((i_start + kk) * kBM + wg * kSub) / kSub) * (4 * 128) + half * (2 * 128);
#pragma unroll
            for (int c = 0; c < 2; ++c) cp_async_16(bias_stage + c * 128, src + c * 128);
        };
        if (n_iter > 0 && kBiasMode == 1) load_bias(0, 0);
        if (wg == 0 && n_iter > 0) {
            // dV, dK accumulators start at zero: every dV / dK MMA accumulates (two warps issue them).  Ordered before the first
            // of those MMAs by this warpgroup's first pds_full arrival (tcgen05.wait::st + fence come first) and the B0 -> B1 token.
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < 2 * kD; c += 8) tmem_st8(tmem_base + lane_off + C::kColDV + c, z);
            // (the CTA's first item: K, V -> TMEM here; later items: at the end of the item before, see below)
            if (it == 0) kv_to_tmem(0, 0);
        }

        for (int k = 0; k < n_iter; ++k) {
            const int kg = kb + k;                            // tile count over the whole CTA: every barrier parity follows from it
            const int m0 = (i_start + k) * kBM + wg * kSub;
            uint8_t* const sDS = smem + C::kDS + ((kg & 1) * kNSub + wg) * kBoxBytes + r * 64;

            // masks: key tail (whole row) and causal (query m sees key n iff n <= m + pseq, i.e. c >= gn - pseq - m0)
            const bool need_mask = (col0 + kBN > p.N) || (kCausal && (col0 + kBN - 1 - pseq > m0));
            int vis_lo = 0;
            if (kCausal) vis_lo = gn - pseq - m0;
            if (!key_ok) vis_lo = kSub;
            // bias mode 3: relative positions n - m of this sub-tile
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

            if (r == 0) BWD3_TS(wg, k, 0);
            mbar_wait(sdp_full + wg, kg & 1);
            tc_fence_after();
            if (r == 0) BWD3_TS(wg, k, 1);

            // ---------------- S^T (whole sub-tile) and dP^T (half by half) into registers; the buffers go back to warp A at
            //                  once.  P^T, dS^T of each half of 16 queries go to this warpgroup's own TMEM columns (A operands
            //                  of dV, dK) and to the shared-memory box (dQ, dBias) ----------------
            uint32_t sr[32], drr[32];
            tmem_ld32(tm_s, sr);
            tmem_ld32(tm_dp, drr);
            tmem_ld_wait();
            if (r == 0) BWD3_TS(wg, k, 4);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + wg);
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
                uint32_t pp[8], dd[8], bias_h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (kBiasMode == 1) {
                    cp_async_wait_all();                           // this thread's own copies: nobody else reads them
                    const uint4 u0 = bias_stage[0], u1 = bias_stage[128];
                    bias_h[0] = u0.x; bias_h[1] = u0.y; bias_h[2] = u0.z; bias_h[3] = u0.w;
                    bias_h[4] = u1.x; bias_h[5] = u1.y; bias_h[6] = u1.z; bias_h[7] = u1.w;
                    // the next half's copy goes into the same 32 bytes (the two LDS above precede it in the LSU)
                    if (hc == 0) load_bias(k, 1);
                    else if (k + 1 < n_iter) load_bias(k + 1, 0);
                }
                const uint32_t(&dr)[16] = *reinterpret_cast<const uint32_t(*)[16]>(drr + hc * 16);
                // statistics of the slot: the S^T MMAs of this sub-tile were issued after warp A saw the slot's barrier complete
                const float* nl = reinterpret_cast<const float*>(smem + C::kStats) + ((((4 * kg + wg) >> 1) + it + 1) % C::kSlots) * (2 * C::kSlotRows) + (wg & 1) * kSub + hc * 16;
                const float* nd = nl + C::kSlotRows;
                // bias mode 3, element c of this half: band[(gn - (m0 + 16 hc + c)) - band_lo]
                const float* bp = band + (gn - m0 - hc * 16 - p.rpe.band_lo);
                const int vl = vis_lo - hc * 16;
                const uint32_t* bw = bias_h;
                const uint32_t(&srh)[16] = *reinterpret_cast<const uint32_t(*)[16]>(sr + hc * 16);
                if (kBiasMode == 3 && rpe_const) {
                    float* acc = const_is_lo ? &ds_const_lo : &ds_const_hi;
                    if (rpe_skip) {
                        if (need_mask) v3_chunk<kBf16, 3, true, true, true>(srh, dr, nl, nd, bw, bp, rpe_cval, scale_log2, vl, pp, dd, *acc);
                        else v3_chunk<kBf16, 3, false, true, true>(srh, dr, nl, nd, bw, bp, rpe_cval, scale_log2, 0, pp, dd, *acc);
                    } else {
                        if (need_mask) v3_chunk<kBf16, 3, true, true, false>(srh, dr, nl, nd, bw, bp, rpe_cval, scale_log2, vl, pp, dd, *acc);
                        else v3_chunk<kBf16, 3, false, true, false>(srh, dr, nl, nd, bw, bp, rpe_cval, scale_log2, 0, pp, dd, *acc);
                    }
                } else {
                    float dummy = 0.f;
                    if (need_mask) v3_chunk<kBf16, kBiasMode, true, false, false>(srh, dr, nl, nd, bw, bp, 0.f, scale_log2, vl, pp, dd, dummy);
                    else v3_chunk<kBf16, kBiasMode, false, false, false>(srh, dr, nl, nd, bw, bp, 0.f, scale_log2, 0, pp, dd, dummy);
                }
                if (r == 0) BWD3_TS(wg, k, 5 + hc);                  // (5: first half computed, 6: second half computed)
                if (hc == 0 && kg > 0) {
                    mbar_wait(pds_free + wg, (kg - 1) & 1);       // dV,dK of this warpgroup's previous sub-tile have read P^T / dS^T
                    // the dS^T box (tile parity, j) was last used two tiles ago: its dQ MMAs and its TMA reduce have read it
                    if (kg >= 2) mbar_wait(box_free + (kg & 1), ((kg >> 1) - 1) & 1);
                    tc_fence_after();
                }
                if (hc == 0 && store_pending) {
                    // (first tile of an item: the box also staged this warpgroup's dK / dV of the item before)
                    if (r == 0) bulk_wait_group_read<0>();
                    named_bar_sync(8 + wg, 128);
                    store_pending = false;
                }
                if (r == 0 && hc == 0) BWD3_TS(wg, k, 7);             // (7: P^T / dS^T buffers and the box are free)
                tmem_st8(tm_p + hc * 8, pp);
                tmem_st8(tm_ds + hc * 8, dd);
                *reinterpret_cast<uint4*>(sDS + (((2 * hc) ^ rx) << 4)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
                *reinterpret_cast<uint4*>(sDS + (((2 * hc + 1) ^ rx) << 4)) = make_uint4(dd[4], dd[5], dd[6], dd[7]);
            }
            if (r == 0) BWD3_TS(wg, k, 2);
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(pds_full + (kg & 1) * kNSub + wg);
            if (r == 0) BWD3_TS(wg, k, 3);
        }
        if (r == 0) BWD3_TS(wg, 8, 0);                        // (row 8 of the timeline: the item boundary)
        if (wg == 0 && n_iter > 0 && next_item(item).n_iter > 0) {
            // ---- K, V of the CTA's next item -> TMEM, while the other warpgroups, the dV / dK / dQ MMAs and the dQ drain of this
            //      item are still at work: the S^T / dP^T MMAs of the next item then start at once (the ring runs on).  The TMEM
            //      copies of K and V are free as soon as the LAST S^T / dP^T MMAs of this item completed (sdp_done).
            // (A barrier of its own, one phase per item, committed by the four A warps after their last MMAs of the item: a
            //  non-consumer cannot wait on sdp_full[j] -- a parity probe cannot tell phase kg_last - 1 from kg_last + 1.)
            mbar_wait(sdp_done, it & 1);
            tc_fence_after();
            kv_to_tmem(it + 1, 2 * (kb + n_iter) + it + 1);
        }
        if (r == 0) BWD3_TS(wg, 8, 1);                        // next K, V in TMEM

        // ---- tail: constant-tile sums; then dV (warpgroups 0, 1: D/2 columns each) and dK * sm_scale (warpgroups 2, 3) ----
        if (kBiasMode == 3 && rpe_skip) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ds_const_lo += __shfl_xor_sync(0xffffffffu, ds_const_lo, off);
                ds_const_hi += __shfl_xor_sync(0xffffffffu, ds_const_hi, off);
            }
            if (lane == 0) {
                if (ds_const_lo != 0.f) atomicAdd(p.rpe.dconst + h * 2 + 0, ds_const_lo);
                if (ds_const_hi != 0.f) atomicAdd(p.rpe.dconst + h * 2 + 1, ds_const_hi);
            }
        }
        {
            const bool is_dv = wg < 2;
            constexpr int kColsPer = kD / 2;                          // columns of the accumulator this warpgroup writes
            const int c_first = (wg & 1) * kColsPer;
            uint8_t* out_row = is_dv
                ? reinterpret_cast<uint8_t*>(p.dv) + 2 * ((int64_t)b * p.dv_sb + (int64_t)h * p.dv_sh + (int64_t)gn * p.dv_sn)
                : reinterpret_cast<uint8_t*>(p.dk) + 2 * ((int64_t)b * p.dk_sb + (int64_t)h * p.dk_sh + (int64_t)gn * p.dk_sn);
            const float sc = is_dv ? 1.f : p.sm_scale;
            if (n_iter > 0) {
                // every MMA of this item completed -- warp B's are the last writers of dV / dK.  (A dedicated barrier, one phase
                // per item: the compute warps do not follow dq_full tile by tile, whose parity could match a phase two tiles old.)
                mbar_wait(all_done, it & 1);
                tc_fence_after();
                if (r == 0) BWD3_TS(wg, 8, 2);                // all MMAs of the item done
                const uint32_t tm_acc = tmem_base + lane_off + (is_dv ? C::kColDV : C::kColDK) + c_first;
                // (all of the warpgroup's columns with one load + one wait, then the stores: kColsPer = 8, 16 or 32)
                uint32_t acc[kColsPer];
                if constexpr (kColsPer == 32) tmem_ld32(tm_acc, acc);
                else if constexpr (kColsPer == 16) tmem_ld16(tm_acc, acc);
                else
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
                                 : "r"(tm_acc)
                                 : "memory");
                tmem_ld_wait();
                if constexpr (kD == 64) {
                    // Through shared memory and one TMA store per warpgroup: the warpgroup's own dS^T box of the tile parity that
                    // the NEXT tile will use (8 KB = 128 keys x 32 columns; free: its dQ MMAs and its TMA reduce are long done).
                    // The same thread waits for the store to have read the box before the warpgroup writes dS^T there again
                    // (first tile of the next item).  Rows beyond N are clipped by the tensor map.
                    uint8_t* stage = smem + C::kDS + ((((kb + n_iter) & 1) * kNSub) + wg) * kBoxBytes;
#pragma unroll
                    for (int c0 = 0; c0 < kColsPer; c0 += 8) {
                        const uint32_t* a = acc + c0;
                        uint4 out;
                        out.x = pack2<kBf16>(__uint_as_float(a[0]) * sc, __uint_as_float(a[1]) * sc);
                        out.y = pack2<kBf16>(__uint_as_float(a[2]) * sc, __uint_as_float(a[3]) * sc);
                        out.z = pack2<kBf16>(__uint_as_float(a[4]) * sc, __uint_as_float(a[5]) * sc);
                        out.w = pack2<kBf16>(__uint_as_float(a[6]) * sc, __uint_as_float(a[7]) * sc);
                        *reinterpret_cast<uint4*>(stage + r * 64 + (((c0 / 8) ^ rx) << 4)) = out;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(8 + wg, 128);
                    if (r == 0) {
                        tma_store_4d(is_dv ? &p.map_dv_st : &p.map_dk_st, stage, c_first, col0, h, b);
                        bulk_commit_group();
                    }
                    store_pending = true;
                } else if (key_ok) {
#pragma unroll
                    for (int c0 = 0; c0 < kColsPer; c0 += 8) {
                        const uint32_t* a = acc + c0;
                        uint4 out;
                        out.x = pack2<kBf16>(__uint_as_float(a[0]) * sc, __uint_as_float(a[1]) * sc);
                        out.y = pack2<kBf16>(__uint_as_float(a[2]) * sc, __uint_as_float(a[3]) * sc);
                        out.z = pack2<kBf16>(__uint_as_float(a[4]) * sc, __uint_as_float(a[5]) * sc);
                        out.w = pack2<kBf16>(__uint_as_float(a[6]) * sc, __uint_as_float(a[7]) * sc);
                        *reinterpret_cast<uint4*>(out_row + 2 * (c_first + c0)) = out;
                    }
                }
                tc_fence_before();
                // every compute warp has read dV / dK (and the band): warpgroup 0 may zero the accumulators for the next item
                if (r == 0) BWD3_TS(wg, 8, 3);                // dK / dV written
                named_bar_sync(6, 512);
                tc_fence_after();
                if (r == 0) BWD3_TS(wg, 8, 4);                // every compute warp through its epilogue
            } else if (key_ok) {
#pragma unroll
                for (int c = 0; c < kColsPer; c += 8) *reinterpret_cast<uint4*>(out_row + 2 * (c_first + c)) = make_uint4(0, 0, 0, 0);
            }
        }
        if (n_iter == 0) continue;                            // (an item no query sees: nothing was counted)
        B200T5_ITEM_END
#ifdef B200T5_BWD_TIMING
        if (threadIdx.x == 0 && blockIdx.x < 8 && it < 20) g_bwd3_item_ts[blockIdx.x][it] = clock64();
#endif
        if (store_pending && r == 0) bulk_wait_group_read<0>();   // shared memory must outlive the last dK / dV store
    }
#undef B200T5_ITEM_BEGIN
#undef B200T5_ITEM_END

    __syncthreads();
    if (warp == kWarpMmaA) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kBiasMode, bool kCausal>
static cudaError_t launch_bwd3_inst(const AttnBwdKernelParams& kp, cudaStream_t stream) {
    using C = Bwd3Cfg<kD, kBiasMode>;
    auto kern = attn_bwd_kernel_v3<kD, kBf16, kBiasMode, kCausal>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
    if (e != cudaSuccess) return e;
    // persistent: one CTA per SM walks the work items (batch, head, key block) round-robin
    static int sm_count[64] = {0};
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (sm_count[dev] == 0) {
        int n = 0;
        if ((e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        sm_count[dev] = n;
    }
    const long long n_items = (long long)kp.B * kp.H * kp.num_n_blocks;
    const int grid = static_cast<int>(n_items < sm_count[dev] ? n_items : sm_count[dev]);
    kern<<<grid, kThreads, C::kTotal, stream>>>(kp);
    count_launch();
#ifdef B200T5_BWD_TIMING
    {
        cudaDeviceSynchronize();
        static long long ts[8][32][8];
        cudaMemcpyFromSymbol(ts, g_bwd3_ts, sizeof(ts));
        const long long t0 = ts[0][0][0];
        const char* names[8] = {"wg0 [wait S, S ready, math done, stored + signalled | S in registers, half 0 computed, half 1 computed, buffers free]  (per tile)", "wg1", "wg2", "wg3",
                                "mma B [wait P/dS(t), ready, dq gate (warp C), issued]  (per sub-tile)", "drain [wait dQ, dQ ready, reduce issued, staging free]  (per tile)",
                                "mma A [wait Q/dO slot, slot ready, buffers free, issued]  (per sub-tile)", "producer [wait slot empty, empty]  (per sub-tile)"};
        for (int role = 0; role < 8; ++role) {
            printf("BWD3_TIMING %s\n", names[role]);
            for (int k = 0; k < ((role == 4 || role >= 6) ? 32 : (role < 4 ? 9 : 8)); ++k) {
                printf("  %2d:", k);
                for (int j = 0; j < (role < 4 ? 8 : 4); ++j) printf(" %7lld", ts[role][k][j] ? ts[role][k][j] - t0 : 0);
                printf("\n");
            }
        }
        static long long its[8][20];
        cudaMemcpyFromSymbol(its, g_bwd3_item_ts, sizeof(its));
        for (int c = 0; c < 8; ++c) {
            printf("BWD3_ITEMS cta %d: cycles per item:", c);
            for (int i = 0; i + 1 < 20 && its[c][i + 1] > its[c][i] && its[c][i] > 0; ++i) printf(" %lld", its[c][i + 1] - its[c][i]);
            printf("\n");
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
