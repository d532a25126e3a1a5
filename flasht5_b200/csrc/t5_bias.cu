// T5 relative-position bias producer for sm_100a: the step immediately upstream of the attention operator
// (SURVEY.md section 8, row f1).
//
// Replaces /root/reference/src/utils/positional_encoding.py:73-102 (`compute_bias`: bucket -> nn.Embedding gather
// -> permute -> (1, H, M, N)) and the backward of that gather (scatter-add of dBias into the (num_buckets, H) table).
// The relative-position -> bucket map is passed in as a lookup table over the relative distance (built by the
// caller with the reference's own formula, :25-71, so the buckets are bit-identical); these kernels do the two
// HBM-bound parts: the dense gather (writes H*M*N elements once, directly in the attention dtype -- no fp32
// (1,H,M,N) intermediate) and the segmented sum.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

template <int kDt>
__device__ __forceinline__ float ld_elem(const void* p, int64_t i) {
    if constexpr (kDt == 2) return __ldg(static_cast<const float*>(p) + i);
    else return to_float16bit<kDt == 1>(__ldg(static_cast<const uint16_t*>(p) + i));
}

constexpr int kMaxBuckets = 256;

// grid = (ceil(M / kRows), H); block = 256 threads; each thread produces 8 consecutive key positions of one row
constexpr int kRowsPerBlock = 8;

template <int kTabDt, int kOutDt>
__global__ void __launch_bounds__(256) t5_bias_fwd_kernel(const void* __restrict__ table, const int32_t* __restrict__ lut,
                                                          int lut_zero, int lut_len, const int32_t* __restrict__ ctx_pos,
                                                          const int32_t* __restrict__ mem_pos, void* __restrict__ bias,
                                                          int H, int M, int N, int num_buckets) {
    __shared__ float s_tab[kMaxBuckets];
    const int h = blockIdx.y;
    for (int i = threadIdx.x; i < num_buckets; i += blockDim.x) s_tab[i] = ld_elem<kTabDt>(table, (int64_t)i * H + h);
    __syncthreads();
    const int m0 = blockIdx.x * kRowsPerBlock;
    for (int mi = 0; mi < kRowsPerBlock; ++mi) {
        const int m = m0 + mi;
        if (m >= M) break;
        const int cp = ctx_pos ? __ldg(ctx_pos + m) : m;
        const int64_t row = ((int64_t)h * M + m) * N;
        for (int n0 = threadIdx.x * 8; n0 < N; n0 += blockDim.x * 8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int n = n0 + e;
                float val = 0.f;
                if (n < N) {
                    int idx = (mem_pos ? __ldg(mem_pos + n) : n) - cp + lut_zero;
                    idx = idx < 0 ? 0 : (idx >= lut_len ? lut_len - 1 : idx);
                    val = s_tab[__ldg(lut + idx)];
                }
                v[e] = val;
            }
            if constexpr (kOutDt == 2) {
                float* o = static_cast<float*>(bias) + row + n0;
                if (n0 + 8 <= N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                    reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
                    reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (n0 + e < N) o[e] = v[e];
                }
            } else {
                uint16_t* o = static_cast<uint16_t*>(bias) + row + n0;
                if (n0 + 8 <= N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                    uint4 u;
                    u.x = pack2<kOutDt == 1>(v[0], v[1]);
                    u.y = pack2<kOutDt == 1>(v[2], v[3]);
                    u.z = pack2<kOutDt == 1>(v[4], v[5]);
                    u.w = pack2<kOutDt == 1>(v[6], v[7]);
                    *reinterpret_cast<uint4*>(o) = u;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (n0 + e < N) o[e] = static_cast<uint16_t>(pack2<kOutDt == 1>(v[e], 0.f) & 0xFFFFu);
                }
            }
        }
    }
}

// dTable[bucket, h] += sum of dBias[h, m, n] over the (m, n) that map to the bucket.  Per-warp shared-memory
// histograms (runs of equal buckets are summed in registers first: neighbouring keys mostly share a bucket),
// then one global atomicAdd per (block, bucket).  dtable (num_buckets, H) fp32 must be zero on entry.
template <int kInDt>
__global__ void __launch_bounds__(256) t5_bias_bwd_kernel(const void* __restrict__ dbias, const int32_t* __restrict__ lut,
                                                          int lut_zero, int lut_len, const int32_t* __restrict__ ctx_pos,
                                                          const int32_t* __restrict__ mem_pos, float* __restrict__ dtable,
                                                          int H, int M, int N, int num_buckets, int rows_per_block) {
    extern __shared__ float s_hist[];                 // [8 warps][num_buckets]
    const int warp = threadIdx.x >> 5;
    float* my = s_hist + warp * num_buckets;
    for (int i = threadIdx.x; i < 8 * num_buckets; i += blockDim.x) s_hist[i] = 0.f;
    __syncthreads();
    const int h = blockIdx.y;
    const int m0 = blockIdx.x * rows_per_block;
    for (int mi = 0; mi < rows_per_block; ++mi) {
        const int m = m0 + mi;
        if (m >= M) break;
        const int cp = ctx_pos ? __ldg(ctx_pos + m) : m;
        const int64_t row = ((int64_t)h * M + m) * N;
        for (int n0 = threadIdx.x * 8; n0 < N; n0 += blockDim.x * 8) {
            float run = 0.f;
            int run_b = -1;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int n = n0 + e;
                if (n < N) {
                    int idx = (mem_pos ? __ldg(mem_pos + n) : n) - cp + lut_zero;
                    idx = idx < 0 ? 0 : (idx >= lut_len ? lut_len - 1 : idx);
                    const int bkt = __ldg(lut + idx);
                    const float g = ld_elem<kInDt>(dbias, row + n);
                    if (bkt != run_b) {
                        if (run_b >= 0) atomicAdd(my + run_b, run);
                        run_b = bkt;
                        run = 0.f;
                    }
                    run += g;
                }
            }
            if (run_b >= 0) atomicAdd(my + run_b, run);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < num_buckets; b += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_hist[w * num_buckets + b];
        if (acc != 0.f) atomicAdd(dtable + (int64_t)b * H + h, acc);
    }
}

// ---------------------------------------------------------------------------------------------------
// Toeplitz fast paths (default positions 0..M-1 / 0..N-1): bias[h, m, n] depends on n - m only.
// One block = one head x kTRows consecutive rows; the diagonal values those rows touch (N + kTRows - 1 of them)
// are staged once in shared memory, so the per-element work is one conflict-free shared load (forward) or one
// coalesced global load (backward) -- no per-element bucket lookup.
// ---------------------------------------------------------------------------------------------------
constexpr int kTRows = 32;

template <int kTabDt, int kOutDt>
__global__ void __launch_bounds__(256) t5_bias_fwd_toeplitz_kernel(const void* __restrict__ table,
                                                                   const int32_t* __restrict__ lut, int lut_zero,
                                                                   int lut_len, void* __restrict__ bias, int H, int M,
                                                                   int N, int num_buckets) {
    extern __shared__ float s_dyn[];              // [num_buckets] table column, then [N + kTRows - 1] diagonal values
    float* s_tab = s_dyn;
    float* s_vec = s_dyn + kMaxBuckets;
    const int h = blockIdx.y;
    const int m0 = blockIdx.x * kTRows;
    for (int i = threadIdx.x; i < num_buckets; i += blockDim.x) s_tab[i] = ld_elem<kTabDt>(table, (int64_t)i * H + h);
    __syncthreads();
    // s_vec[j] = value of relative position  j - (m0 + kTRows - 1)
    const int seg = N + kTRows - 1;
    for (int j = threadIdx.x; j < seg; j += blockDim.x) {
        int idx = j - (m0 + kTRows - 1) + lut_zero;
        idx = idx < 0 ? 0 : (idx >= lut_len ? lut_len - 1 : idx);
        s_vec[j] = s_tab[__ldg(lut + idx)];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int mi = warp; mi < kTRows; mi += 8) {
        const int m = m0 + mi;
        if (m >= M) break;
        const float* src = s_vec + (kTRows - 1 - mi);          // src[n] = bias[h, m, n]
        const int64_t row = ((int64_t)h * M + m) * N;
        if constexpr (kOutDt == 2) {
            float* o = static_cast<float*>(bias) + row;
            for (int n = lane; n < N; n += 32) o[n] = src[n];
        } else {
            uint16_t* o = static_cast<uint16_t*>(bias) + row;
            if ((row & 1) == 0) {                               // 4-byte aligned row: packed pairs, 128 B per warp store
                for (int n = 2 * lane; n < N; n += 64) {
                    if (n + 1 < N) *reinterpret_cast<uint32_t*>(o + n) = pack2<kOutDt == 1>(src[n], src[n + 1]);
                    else o[n] = static_cast<uint16_t>(pack2<kOutDt == 1>(src[n], 0.f) & 0xFFFFu);
                }
            } else {
                for (int n = lane; n < N; n += 32) o[n] = static_cast<uint16_t>(pack2<kOutDt == 1>(src[n], 0.f) & 0xFFFFu);
            }
        }
    }
}

// Backward: one block = one head x 32 rows; each warp takes 32x32 tiles.  Row r of a tile is rotated by r lanes
// with a shuffle, after which lane l holds an element of diagonal l (no wrap) or l - 32 (wrapped): two running
// sums per lane collect the tile's 63 diagonals with one shuffle + one add per row and NO atomics; the tile then
// adds 2 values per lane to the block's shared diagonal segment, which is folded into buckets at the end.
template <int kInDt>
__global__ void __launch_bounds__(256) t5_bias_bwd_toeplitz_kernel(const void* __restrict__ dbias,
                                                                   const int32_t* __restrict__ lut, int lut_zero,
                                                                   int lut_len, float* __restrict__ dtable, int H, int M,
                                                                   int N, int num_buckets) {
    extern __shared__ float s_dyn[];              // [N + 63] diagonal sums of this row group, then [8 warps][num_buckets]
    const int seg = N + 2 * kTRows - 1;
    float* s_hist = s_dyn + seg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y;
    const int m0 = blockIdx.x * kTRows;
    for (int i = threadIdx.x; i < seg + 8 * num_buckets; i += blockDim.x) s_dyn[i] = 0.f;
    __syncthreads();
    // s_dyn[j] <-> relative position j - (kTRows - 1) - m0 ... i.e. element (m0 + r, n) lands at j = n - r + (kTRows - 1)
    const int ntiles = (N + 31) / 32;
    for (int t = warp; t < ntiles; t += 8) {
        const int n = t * 32 + lane;
        float acc_lo = 0.f, acc_hi = 0.f;          // diagonals (lane) and (lane - 32) of this tile
        float v[kTRows];
#pragma unroll
        for (int r = 0; r < kTRows; ++r) {         // all 32 row loads in flight before the first shuffle
            const int m = m0 + r;
            v[r] = (m < M && n < N) ? ld_elem<kInDt>(dbias, ((int64_t)h * M + m) * N + n) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kTRows; ++r) {
            // lane l receives column (l + r) % 32 of row r: its diagonal is (l + r) % 32 - r = l  or  l - 32
            const float w = __shfl_sync(0xffffffffu, v[r], (lane + r) & 31);
            if (lane + r < 32) acc_lo += w; else acc_hi += w;
        }
        // tile-local diagonal c - r = lane (acc_lo) / lane - 32 (acc_hi); global j = t*32 + diag + (kTRows - 1)
        atomicAdd(s_dyn + t * 32 + lane + (kTRows - 1), acc_lo);
        if (lane > 0) atomicAdd(s_dyn + t * 32 + lane - 32 + (kTRows - 1), acc_hi);
    }
    __syncthreads();
    // fold the diagonal sums into buckets.  Neighbouring diagonals mostly share a bucket and shared-memory float
    // atomics on ONE address serialise (CAS loop), so each warp first sums its lanes per distinct bucket with
    // shuffles and only the group leader touches the warp's private histogram row -- no atomics.
    float* my_hist = s_hist + warp * num_buckets;
    for (int j0 = warp * 32; j0 < seg; j0 += 256) {
        const int j = j0 + lane;
        float acc = 0.f;
        int bkt = -1;
        if (j < seg) {
            acc = s_dyn[j];
            int idx = j - (kTRows - 1) - m0 + lut_zero;
            idx = idx < 0 ? 0 : (idx >= lut_len ? lut_len - 1 : idx);
            bkt = __ldg(lut + idx);
        }
        unsigned todo = __ballot_sync(0xffffffffu, bkt >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int b = __shfl_sync(0xffffffffu, bkt, leader);
            const unsigned grp = __ballot_sync(0xffffffffu, bkt == b);
            float v = bkt == b ? acc : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == leader) my_hist[b] += v;
            todo &= ~grp;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int b = threadIdx.x; b < num_buckets; b += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_hist[w * num_buckets + b];
        if (acc != 0.f) atomicAdd(dtable + (int64_t)b * H + h, acc);
    }
}

}  // namespace

cudaError_t launch_t5_bias_fwd(const void* table, const int32_t* lut, int lut_zero, int lut_len, const int32_t* ctx_pos,
                               const int32_t* mem_pos, void* bias, int H, int M, int N, int num_buckets, int table_dtype,
                               int bias_dtype, cudaStream_t stream) {
    if (num_buckets > kMaxBuckets) return cudaErrorInvalidValue;
    if (ctx_pos == nullptr && mem_pos == nullptr && (size_t)(kMaxBuckets + N + kTRows) * 4 <= 200 * 1024) {
        const dim3 tgrid((M + kTRows - 1) / kTRows, H);
        const size_t smem = (size_t)(kMaxBuckets + N + kTRows - 1) * sizeof(float);
#define B200T5_T5FT(TD, OD)                                                                                         \
    {                                                                                                               \
        if (smem > 48 * 1024)                                                                                       \
            cudaFuncSetAttribute(t5_bias_fwd_toeplitz_kernel<TD, OD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        t5_bias_fwd_toeplitz_kernel<TD, OD><<<tgrid, 256, smem, stream>>>(table, lut, lut_zero, lut_len, bias, H, M, N, num_buckets); \
    }
        switch (table_dtype * 3 + bias_dtype) {
            case 0: B200T5_T5FT(0, 0); break;
            case 1: B200T5_T5FT(0, 1); break;
            case 2: B200T5_T5FT(0, 2); break;
            case 3: B200T5_T5FT(1, 0); break;
            case 4: B200T5_T5FT(1, 1); break;
            case 5: B200T5_T5FT(1, 2); break;
            case 6: B200T5_T5FT(2, 0); break;
            case 7: B200T5_T5FT(2, 1); break;
            default: B200T5_T5FT(2, 2); break;
        }
#undef B200T5_T5FT
        count_launch();
        return cudaGetLastError();
    }
    const dim3 grid((M + kRowsPerBlock - 1) / kRowsPerBlock, H);
#define B200T5_T5F(TD, OD)                                                                                          \
    t5_bias_fwd_kernel<TD, OD><<<grid, 256, 0, stream>>>(table, lut, lut_zero, lut_len, ctx_pos, mem_pos, bias, H, M, N, \
                                                         num_buckets)
    switch (table_dtype * 3 + bias_dtype) {
        case 0: B200T5_T5F(0, 0); break;
        case 1: B200T5_T5F(0, 1); break;
        case 2: B200T5_T5F(0, 2); break;
        case 3: B200T5_T5F(1, 0); break;
        case 4: B200T5_T5F(1, 1); break;
        case 5: B200T5_T5F(1, 2); break;
        case 6: B200T5_T5F(2, 0); break;
        case 7: B200T5_T5F(2, 1); break;
        default: B200T5_T5F(2, 2); break;
    }
#undef B200T5_T5F
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// band of bias values over relative positions for the in-kernel bias mode of the attention kernels
// (kernels.h: RpeBand):  band[h][j] = io_round(table[lut[clamp(band_lo + j + lut_zero)], h])  as fp32
// ------------------------------------------------------------------------------------------
template <int kTabDt>
__global__ void __launch_bounds__(256) rpe_band_kernel(const void* __restrict__ table, int64_t stride_b, int64_t stride_h,
                                                       const int32_t* __restrict__ lut, int lut_zero, int lut_len,
                                                       float* __restrict__ band, int band_lo, int band_len, int io_dtype) {
    const int h = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= band_len) return;
    int idx = band_lo + j + lut_zero;
    idx = idx < 0 ? 0 : (idx >= lut_len ? lut_len - 1 : idx);
    const int bucket = __ldg(lut + idx);
    float v = ld_elem<kTabDt>(table, (int64_t)bucket * stride_b + (int64_t)h * stride_h);
    // the dense path hands the kernels a bias already cast to the attention dtype: round the same way
    if (io_dtype == 1) v = to_float16bit<true>(static_cast<uint16_t>(pack2<true>(v, 0.f) & 0xFFFFu));
    else if (io_dtype == 0) v = to_float16bit<false>(static_cast<uint16_t>(pack2<false>(v, 0.f) & 0xFFFFu));
    band[(int64_t)h * band_len + j] = v;
}

cudaError_t launch_rpe_band(const void* table, int64_t stride_b, int64_t stride_h, int table_dtype, const int32_t* lut,
                            int lut_zero, int lut_len, float* band, int H, int band_lo, int band_len, int io_dtype,
                            cudaStream_t stream) {
    const dim3 grid((band_len + 255) / 256, H);
    switch (table_dtype) {
        case 0: rpe_band_kernel<0><<<grid, 256, 0, stream>>>(table, stride_b, stride_h, lut, lut_zero, lut_len, band, band_lo, band_len, io_dtype); break;
        case 1: rpe_band_kernel<1><<<grid, 256, 0, stream>>>(table, stride_b, stride_h, lut, lut_zero, lut_len, band, band_lo, band_len, io_dtype); break;
        default: rpe_band_kernel<2><<<grid, 256, 0, stream>>>(table, stride_b, stride_h, lut, lut_zero, lut_len, band, band_lo, band_len, io_dtype); break;
    }
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Table gradient of the in-kernel relative-position bias straight from the 16-bit dS group surface (the D <= 64 backward of
// the in-kernel operator, B200T5_RPE_SKIP_LEVEL 2): only the tiles that are NOT entirely beyond a constant end of the bucket table carry
// data (the attention backward keeps the others' dS in registers), so only those are read -- 3 of 8 tiles per query
// block at S = 1024 -- and neither the dense (1,H,M,N) dBias scratch nor the producer's segmented sum is needed.
// One CTA = one (key block, query block, head, group slice).  Thread t walks the wrapped diagonal w = t % 128 of its half of
// the rows: lanes read consecutive 16-bit elements; a wrapped diagonal is two true diagonals (d = w and d = w - 128).
// ------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void __launch_bounds__(256) rpe_dtable_band_kernel(const uint16_t* __restrict__ ws, int pitch, int G, int H, int R,
                                                              int Cn, int pseq, int sign, const int32_t* __restrict__ lut,
                                                              int lut_zero, int lut_len, int const_lo, int const_hi,
                                                              float* __restrict__ dtable, int num_buckets, int causal) {
    // The surface holds one (R x Cn) matrix per (group, head): rows = queries, columns = keys (sign = +1) or the transposed
    // surface of the v3 attention kernel, rows = keys, columns = queries (sign = -1).  Relative position of element
    // (row0 + r, col0 + c): rel = n - m = sign * (col0 - row0 + (c - r)).
    const int col0 = blockIdx.x * 128, row0 = blockIdx.y * 128;
    const int h = blockIdx.z / G, g = blockIdx.z % G;
    const int base = col0 - row0;
    const int rel_min = sign > 0 ? base - 127 : -(base + 127), rel_max = sign > 0 ? base + 127 : -(base - 127);
    if (rel_max <= const_lo || rel_min >= const_hi) return;        // constant tile: summed inside the attention kernel
    if (causal) {                                                  // entirely masked (n > m + pseq everywhere): never written
        if (sign > 0 ? (col0 > row0 + 127 + pseq) : (row0 > col0 + 127 + pseq)) return;
    }
    __shared__ float sdiag[256];                                    // index d + 128, d = c - r in [-127, 127]
    __shared__ float sbucket[kMaxBuckets];
    sdiag[threadIdx.x] = 0.f;
    for (int i = threadIdx.x; i < num_buckets; i += 256) sbucket[i] = 0.f;
    __syncthreads();
    const int w = threadIdx.x & 127, half = threadIdx.x >> 7;
    const uint16_t* src = ws + ((int64_t)g * H + h) * (int64_t)R * pitch;
    float s_pos = 0.f, s_neg = 0.f;
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) {
        const int r = half * 64 + rr;
        const int row = row0 + r;
        int c = r + w;
        const bool wrapped = c >= 128;
        c &= 127;
        const int col = col0 + c;
        if (row < R && col < Cn) {
            const float v = to_float16bit<kBf16>(__ldg(src + (int64_t)row * pitch + col));
            if (wrapped) s_neg += v;
            else s_pos += v;
        }
    }
    atomicAdd(&sdiag[w + 128], s_pos);          // two adders per slot (the two halves of the rows)
    atomicAdd(&sdiag[w], s_neg);                // d = w - 128 (slot 0 = d -128: no elements, stays 0)
    __syncthreads();
    {
        const int d = static_cast<int>(threadIdx.x) - 128;
        const float v = sdiag[threadIdx.x];
        if (v != 0.f) {
            int idx = sign * (base + d) + lut_zero;
            idx = idx < 0 ? 0 : (idx >= lut_len ? lut_len - 1 : idx);
            atomicAdd(&sbucket[__ldg(lut + idx)], v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < num_buckets; i += 256) {
        const float v = sbucket[i];
        if (v != 0.f) atomicAdd(dtable + (int64_t)i * H + h, v);
    }
}

cudaError_t launch_rpe_dtable_band(const void* ds_ws, int pitch, int G, int H, int M, int N, const int32_t* lut, int lut_zero,
                                   int lut_len, int const_lo, int const_hi, float* dtable, int num_buckets, bool causal,
                                   bool bf16, bool transposed, cudaStream_t stream) {
    if (num_buckets > kMaxBuckets) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(dtable, 0, (size_t)num_buckets * H * sizeof(float), stream);
    if (e != cudaSuccess) return e;
    const int R = transposed ? N : M, Cn = transposed ? M : N;
    const dim3 grid((Cn + 127) / 128, (R + 127) / 128, H * G);
    if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
    if (bf16)
        rpe_dtable_band_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(ds_ws), pitch, G, H, R, Cn, N - M, transposed ? -1 : 1,
                                                               lut, lut_zero, lut_len, const_lo, const_hi, dtable, num_buckets, causal ? 1 : 0);
    else
        rpe_dtable_band_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(ds_ws), pitch, G, H, R, Cn, N - M, transposed ? -1 : 1,
                                                                lut, lut_zero, lut_len, const_lo, const_hi, dtable, num_buckets, causal ? 1 : 0);
    count_launch();
    return cudaGetLastError();
}

__global__ void rpe_dtable_add_const_kernel(float* __restrict__ dtable, const float* __restrict__ dconst,
                                            const int32_t* __restrict__ lut, int lut_zero, int lut_len, int const_lo,
                                            int const_hi, int H) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    int i0 = const_lo + lut_zero, i1 = const_hi + lut_zero;
    i0 = i0 < 0 ? 0 : (i0 >= lut_len ? lut_len - 1 : i0);
    i1 = i1 < 0 ? 0 : (i1 >= lut_len ? lut_len - 1 : i1);
    const int b0 = __ldg(lut + i0), b1 = __ldg(lut + i1);
    // one thread per head; the two buckets may coincide (a table with a single bucket): add one after the other
    dtable[(int64_t)b0 * H + h] += dconst[h * 2 + 0];
    dtable[(int64_t)b1 * H + h] += dconst[h * 2 + 1];
}

cudaError_t launch_rpe_dtable_add_const(float* dtable, const float* dconst, const int32_t* lut, int lut_zero, int lut_len,
                                        int const_lo, int const_hi, int H, cudaStream_t stream) {
    rpe_dtable_add_const_kernel<<<(H + 127) / 128, 128, 0, stream>>>(dtable, dconst, lut, lut_zero, lut_len, const_lo,
                                                                    const_hi, H);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_t5_bias_bwd(const void* dbias, const int32_t* lut, int lut_zero, int lut_len, const int32_t* ctx_pos,
                               const int32_t* mem_pos, float* dtable, int H, int M, int N, int num_buckets, int dbias_dtype,
                               cudaStream_t stream) {
    if (num_buckets > kMaxBuckets) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(dtable, 0, (size_t)num_buckets * H * sizeof(float), stream);
    if (e != cudaSuccess) return e;
    {
        const size_t smem = ((size_t)N + 2 * kTRows - 1 + 8 * (size_t)num_buckets) * sizeof(float);
        if (ctx_pos == nullptr && mem_pos == nullptr && smem <= 200 * 1024) {
            const dim3 tgrid((M + kTRows - 1) / kTRows, H);
#define B200T5_T5BT(DT)                                                                                             \
    {                                                                                                               \
        if (smem > 48 * 1024)                                                                                       \
            cudaFuncSetAttribute(t5_bias_bwd_toeplitz_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        t5_bias_bwd_toeplitz_kernel<DT><<<tgrid, 256, smem, stream>>>(dbias, lut, lut_zero, lut_len, dtable, H, M, N, num_buckets); \
    }
            switch (dbias_dtype) {
                case 0: B200T5_T5BT(0); break;
                case 1: B200T5_T5BT(1); break;
                default: B200T5_T5BT(2); break;
            }
#undef B200T5_T5BT
            count_launch();
            return cudaGetLastError();
        }
    }
    // enough blocks to fill the chip a few times over, few enough that the global atomics stay negligible
    int rows_per_block = std::max(1, (M * H + 148 * 8 - 1) / (148 * 8));
    rows_per_block = std::min(rows_per_block, 64);
    const dim3 grid((M + rows_per_block - 1) / rows_per_block, H);
    const size_t smem = 8 * (size_t)num_buckets * sizeof(float);
    switch (dbias_dtype) {
        case 0: t5_bias_bwd_kernel<0><<<grid, 256, smem, stream>>>(dbias, lut, lut_zero, lut_len, ctx_pos, mem_pos, dtable, H, M, N, num_buckets, rows_per_block); break;
        case 1: t5_bias_bwd_kernel<1><<<grid, 256, smem, stream>>>(dbias, lut, lut_zero, lut_len, ctx_pos, mem_pos, dtable, H, M, N, num_buckets, rows_per_block); break;
        default: t5_bias_bwd_kernel<2><<<grid, 256, smem, stream>>>(dbias, lut, lut_zero, lut_len, ctx_pos, mem_pos, dtable, H, M, N, num_buckets, rows_per_block); break;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200t5
