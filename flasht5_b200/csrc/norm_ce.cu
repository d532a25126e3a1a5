// RMSNorm and cross-entropy(+z-loss) for sm_100a.  Both are HBM-bound streaming kernels: 128-bit
// coalesced loads, fp32 math, one pass over the data where the row fits in registers.
//
// Replaces /root/reference/src/model/ops/rms_norm.py:25-131 and
//          /root/reference/src/model/ops/cross_entropy_loss.py:40-162.
#include <algorithm>
#include <cstdlib>

#include "../../include/b200t5.h"
#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

// ------------------------------------------------------------------------------------------
// 8-element (or fewer) typed loads / stores.  kDt: 0 = fp16, 1 = bf16, 2 = fp32.
// ------------------------------------------------------------------------------------------
template <int kDt>
struct ElemBytes {
    static constexpr int value = kDt == 2 ? 4 : 2;
};

template <int kDt>
__device__ __forceinline__ void load8(const void* base, int64_t idx, float (&v)[8]) {
    if constexpr (kDt == 2) {
        const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
        const float4 a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + idx));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<kDt == 1>(w[e]);
            v[2 * e] = f.x;
            v[2 * e + 1] = f.y;
        }
    }
}

template <int kDt>
__device__ __forceinline__ void store8(void* base, int64_t idx, const float (&v)[8]) {
    if constexpr (kDt == 2) {
        float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + idx);
        p[0] = make_float4(v[0], v[1], v[2], v[3]);
        p[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint4 u;
        u.x = pack2<kDt == 1>(v[0], v[1]);
        u.y = pack2<kDt == 1>(v[2], v[3]);
        u.z = pack2<kDt == 1>(v[4], v[5]);
        u.w = pack2<kDt == 1>(v[6], v[7]);
        *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + idx) = u;
    }
}

template <int kDt>
__device__ __forceinline__ float load1(const void* base, int64_t idx) {
    if constexpr (kDt == 2) return __ldg(static_cast<const float*>(base) + idx);
    else return to_float16bit<kDt == 1>(__ldg(static_cast<const uint16_t*>(base) + idx));
}

template <int kDt>
__device__ __forceinline__ void store1(void* base, int64_t idx, float v) {
    if constexpr (kDt == 2) static_cast<float*>(base)[idx] = v;
    else static_cast<uint16_t*>(base)[idx] = static_cast<uint16_t>(pack2<kDt == 1>(v, 0.f) & 0xFFFFu);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum over kWarps warps; every thread gets the result.  `red` holds kWarps floats.
template <int kWarps>
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    if constexpr (kWarps == 1) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t += red[i];
    return t;
}
template <int kWarps>
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    if constexpr (kWarps == 1) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0];
#pragma unroll
    for (int i = 1; i < kWarps; ++i) t = fmaxf(t, red[i]);
    return t;
}

// ==========================================================================================
// RMSNorm
// ==========================================================================================
// Vectorised path: a group of kT threads (one warp, or the whole 256-thread block) owns a row and keeps
// it in registers (kChunks x 8 elements per thread); persistent over rows.  Requires n % 8 == 0,
// 16-byte aligned rows and n <= kT * 8 * kChunks.
template <int kT>
__device__ __forceinline__ float group_sum(float v, float* red) {
    if constexpr (kT == 32) return warp_sum(v);
    else return block_sum<8>(v, red);
}

template <int kXDt, int kWDt, int kChunks, int kT>
__global__ void __launch_bounds__(256) rmsnorm_fwd_vec_kernel(const void* __restrict__ x, const void* __restrict__ w,
                                                              void* __restrict__ y, float* __restrict__ rstd_out,
                                                              int rows, int n, int64_t xs, int64_t ys, float eps) {
    __shared__ float red[8];
    constexpr int kGroups = 256 / kT;
    const int tig = threadIdx.x % kT;
    const int group_global = blockIdx.x * kGroups + threadIdx.x / kT;
    const int groups_total = gridDim.x * kGroups;
    float wv[kChunks][8];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int col = (c * kT + tig) * 8;
        if (col < n) load8<kWDt>(w, col, wv[c]);
    }
    // (kT == 256: every thread of the block runs the same number of iterations, so the block-wide
    //  reductions inside the loop are safe)
    for (int row = group_global; row < rows; row += groups_total) {
        float xv[kChunks][8];
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int col = (c * kT + tig) * 8;
            if (col < n) {
                load8<kXDt>(x, (int64_t)row * xs + col, xv[c]);
#pragma unroll
                for (int e = 0; e < 8; ++e) ss = fmaf(xv[c][e], xv[c][e], ss);
            }
        }
        ss = group_sum<kT>(ss, red);
        const float rstd = rsqrtf(ss / static_cast<float>(n) + eps);
        if (tig == 0) rstd_out[row] = rstd;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int col = (c * kT + tig) * 8;
            if (col < n) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = xv[c][e] * rstd * wv[c][e];
                store8<kXDt>(y, (int64_t)row * ys + col, o);
            }
        }
    }
}

// Generic path: one block per row, scalar accesses, any n / alignment.
template <int kXDt, int kWDt>
__global__ void __launch_bounds__(256) rmsnorm_fwd_generic_kernel(const void* __restrict__ x, const void* __restrict__ w,
                                                                  void* __restrict__ y, float* __restrict__ rstd_out,
                                                                  int rows, int n, int64_t xs, int64_t ys, float eps) {
    __shared__ float red[8];
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        float ss = 0.f;
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            const float v = load1<kXDt>(x, (int64_t)row * xs + c);
            ss = fmaf(v, v, ss);
        }
        ss = block_sum<8>(ss, red);
        const float rstd = rsqrtf(ss / static_cast<float>(n) + eps);
        if (threadIdx.x == 0) rstd_out[row] = rstd;
        for (int c = threadIdx.x; c < n; c += blockDim.x)
            store1<kXDt>(y, (int64_t)row * ys + c, load1<kXDt>(x, (int64_t)row * xs + c) * rstd * load1<kWDt>(w, c));
        __syncthreads();
    }
}

// Backward, vectorised: same row ownership as the forward; every thread keeps a private fp32 dW partial
// in registers, the groups of a block are summed through shared memory, one partial row per block.
template <int kXDt, int kWDt, int kChunks, int kT>
__global__ void __launch_bounds__(256, 2) rmsnorm_bwd_vec_kernel(const void* __restrict__ dy, const void* __restrict__ x,
                                                                 const void* __restrict__ w, const float* __restrict__ rstd_in,
                                                                 void* __restrict__ dx, float* __restrict__ dw_partial,
                                                                 int rows, int n, int64_t dys, int64_t xs, int64_t dxs) {
    extern __shared__ float sdyn[];   // [n] dW block partial, then [n] fp32 copy of w
    float* sdw = sdyn;
    float* sw = sdyn + n;
    __shared__ float red[8];
    constexpr int kGroups = 256 / kT;
    const int tig = threadIdx.x % kT;
    const int group_global = blockIdx.x * kGroups + threadIdx.x / kT;
    const int groups_total = gridDim.x * kGroups;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        sdw[c] = 0.f;
        sw[c] = load1<kWDt>(w, c);
    }
    __syncthreads();

    // w stays in shared memory (not registers): the kernel then fits 2 blocks per SM, and occupancy -- not
    // arithmetic -- is what this streaming kernel needs
    float dwv[kChunks][8];
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) dwv[c][e] = 0.f;
    const float inv_n = 1.f / static_cast<float>(n);
    for (int row = group_global; row < rows; row += groups_total) {
        const float rstd = __ldg(rstd_in + row);
        float xh[kChunks][8], wdy[kChunks][8];
        float c1 = 0.f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int col = (c * kT + tig) * 8;
            if (col < n) {
                float dyv[8];
                load8<kXDt>(x, (int64_t)row * xs + col, xh[c]);
                load8<kXDt>(dy, (int64_t)row * dys + col, dyv);
                const float4 w0 = *reinterpret_cast<const float4*>(sw + col);
                const float4 w1 = *reinterpret_cast<const float4*>(sw + col + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    xh[c][e] *= rstd;
                    wdy[c][e] = wv[e] * dyv[e];
                    c1 = fmaf(xh[c][e], wdy[c][e], c1);
                    dwv[c][e] = fmaf(dyv[e], xh[c][e], dwv[c][e]);
                }
            }
        }
        c1 = group_sum<kT>(c1, red) * inv_n;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int col = (c * kT + tig) * 8;
            if (col < n) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (wdy[c][e] - xh[c][e] * c1) * rstd;
                store8<kXDt>(dx, (int64_t)row * dxs + col, o);
            }
        }
    }
    // block-level reduction of the per-thread partials
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int col = (c * kT + tig) * 8;
        if (col < n) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&sdw[col + e], dwv[c][e]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) dw_partial[(int64_t)blockIdx.x * n + c] = sdw[c];
}

template <int kXDt, int kWDt>
__global__ void __launch_bounds__(256) rmsnorm_bwd_generic_kernel(const void* __restrict__ dy, const void* __restrict__ x,
                                                                  const void* __restrict__ w, const float* __restrict__ rstd_in,
                                                                  void* __restrict__ dx, float* __restrict__ dw_partial,
                                                                  int rows, int n, int64_t dys, int64_t xs, int64_t dxs) {
    __shared__ float red[8];
    // dw_partial row of this block is accumulated directly in global memory (only this block touches it)
    float* mydw = dw_partial + (int64_t)blockIdx.x * n;
    for (int c = threadIdx.x; c < n; c += blockDim.x) mydw[c] = 0.f;
    const float inv_n = 1.f / static_cast<float>(n);
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const float rstd = __ldg(rstd_in + row);
        float c1 = 0.f;
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            const float xh = load1<kXDt>(x, (int64_t)row * xs + c) * rstd;
            c1 = fmaf(xh, load1<kWDt>(w, c) * load1<kXDt>(dy, (int64_t)row * dys + c), c1);
        }
        c1 = block_sum<8>(c1, red) * inv_n;
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            const float xh = load1<kXDt>(x, (int64_t)row * xs + c) * rstd;
            const float dyv = load1<kXDt>(dy, (int64_t)row * dys + c);
            store1<kXDt>(dx, (int64_t)row * dxs + c, (load1<kWDt>(w, c) * dyv - xh * c1) * rstd);
            mydw[c] += dyv * xh;     // same thread owns column c for every row
        }
        __syncthreads();
    }
}

template <int kWDt>
__global__ void rmsnorm_dw_finalize_kernel(const float* __restrict__ partial, int num_partials, void* __restrict__ dw, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float acc = 0.f;
    for (int p = 0; p < num_partials; ++p) acc += partial[(int64_t)p * n + c];
    store1<kWDt>(dw, c, acc);
}

// ==========================================================================================
// Cross-entropy + z-loss
// ==========================================================================================
constexpr int kCeThreads = 512;

// One block per row.  Pass 1: online max / sum-exp / sum over the row (vectorised when aligned).
template <int kDt, bool kVec>
__global__ void __launch_bounds__(kCeThreads) ce_fwd_kernel(const void* __restrict__ logits, const int64_t* __restrict__ labels,
                                                            float* __restrict__ losses, float* __restrict__ z_losses,
                                                            float* __restrict__ lse_out, int lse_is_input, int vocab,
                                                            int64_t row_stride, float smoothing, float logit_scale,
                                                            float lse_square_scale, int64_t ignore_index) {
    __shared__ float red[kCeThreads / 32];
    const int row = blockIdx.x;
    const int64_t base = (int64_t)row * row_stride;
    if (lse_is_input) {          // precomputed log-sum-exp (smoothing == 0, logit_scale == 1 enforced by the ABI)
        if (threadIdx.x == 0) {
            const float lse = lse_out[row];
            const int64_t label = labels[row];
            float loss = 0.f, z = 0.f;
            if (label != ignore_index) {
                z = lse_square_scale * lse * lse;
                // a label outside [0, vocab) contributes no cross-entropy term (reference cross_entropy_loss.py:87-100)
                loss = (label >= 0 && label < vocab ? lse - load1<kDt>(logits, base + label) : 0.f) + z;
            }
            losses[row] = loss;
            z_losses[row] = z;
        }
        return;
    }
    float m = -INFINITY, s = 0.f, tot = 0.f;
    const float sl2 = logit_scale * 1.4426950408889634f;
    if constexpr (kVec) {
        for (int c0 = threadIdx.x * 8; c0 < vocab; c0 += kCeThreads * 8) {
            float v[8];
            load8<kDt>(logits, base + c0, v);
            float cm = v[0];
#pragma unroll
            for (int e = 1; e < 8; ++e) cm = fmaxf(cm, v[e]);
            cm *= logit_scale;
            // logit_scale may be negative: the max of scaled values is then min*scale; handle generally
            if (logit_scale < 0.f) {
                cm = v[0] * logit_scale;
#pragma unroll
                for (int e = 1; e < 8; ++e) cm = fmaxf(cm, v[e] * logit_scale);
            }
            const float nm = fmaxf(m, cm);
#pragma unroll
            for (int e = 0; e < 8; ++e) tot += v[e];
            if (nm > -INFINITY) {
                float add = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) add += ex2_approx(fmaf(v[e], sl2, -nm * 1.4426950408889634f));
                s = s * ex2_approx((m - nm) * 1.4426950408889634f) + add;
                m = nm;
            }
        }
    } else {
        for (int c = threadIdx.x; c < vocab; c += kCeThreads) {
            const float raw = load1<kDt>(logits, base + c);
            const float xv = raw * logit_scale;
            const float nm = fmaxf(m, xv);
            if (nm > -INFINITY) {
                s = s * __expf(m - nm) + __expf(xv - nm);
                m = nm;
            }
            tot += raw;
        }
    }
    // combine (m, s) across the block
    const float bm = block_max<kCeThreads / 32>(m, red);
    const float contrib = (m == -INFINITY) ? 0.f : s * __expf(m - bm);
    const float bs = block_sum<kCeThreads / 32>(contrib, red);
    const float btot = smoothing > 0.f ? block_sum<kCeThreads / 32>(tot, red) : 0.f;
    if (threadIdx.x == 0) {
        const float lse = bm + __logf(bs);
        lse_out[row] = lse;
        const int64_t label = labels[row];
        float loss = 0.f, z = 0.f;
        if (label != ignore_index) {
            if (label >= 0 && label < vocab) {
                const float xl = load1<kDt>(logits, base + label) * logit_scale;
                if (smoothing > 0.f)
                    loss = lse - smoothing * (btot * logit_scale) / static_cast<float>(vocab) - (1.f - smoothing) * xl;
                else
                    loss = lse - xl;
            } else {
                // label out of range: no cross-entropy term, only the smoothing term (reference cross_entropy_loss.py:96-100)
                loss = smoothing > 0.f ? smoothing * (lse - (btot * logit_scale) / static_cast<float>(vocab)) : 0.f;
            }
            z = lse_square_scale * lse * lse;
            loss += z;
        }
        losses[row] = loss;
        z_losses[row] = z;
    }
}

// grid = (rows, chunks); each block handles a kCeThreads*8*kIters-wide slice of one row
constexpr int kCeBwdIters = 4;
template <int kDt, bool kVec>
__global__ void __launch_bounds__(kCeThreads) ce_bwd_kernel(const void* logits, const int64_t* __restrict__ labels,
                                                            const float* __restrict__ lse_in, const float* __restrict__ dlosses,
                                                            int64_t dloss_stride, void* dlogits, int vocab,
                                                            int64_t row_stride, int64_t dl_row_stride, float smoothing,
                                                            float logit_scale, float lse_square_scale,
                                                            int64_t ignore_index) {
    const int row = blockIdx.x;
    const int64_t label = labels[row];
    const bool ignored = label == ignore_index;
    const float lse = lse_in[row];
    const float dl = ignored ? 0.f : dlosses[(int64_t)row * dloss_stride] * logit_scale;
    const float zf = 1.f + 2.f * lse_square_scale * lse;
    const float sm_v = smoothing / static_cast<float>(vocab);
    const float lse_l2 = lse * 1.4426950408889634f;
    const float sl2 = logit_scale * 1.4426950408889634f;
    const int64_t ib = (int64_t)row * row_stride;
    const int64_t ob = (int64_t)row * dl_row_stride;
    const int slice0 = blockIdx.y * (kCeThreads * 8 * kCeBwdIters);
    if constexpr (kVec) {
#pragma unroll
        for (int it = 0; it < kCeBwdIters; ++it) {
            const int c0 = slice0 + (it * kCeThreads + threadIdx.x) * 8;
            if (c0 < vocab) {
                float v[8], o[8];
                if (!ignored) load8<kDt>(logits, ib + c0, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (ignored) {
                        o[e] = 0.f;
                    } else {
                        float pr = ex2_approx(fmaf(v[e], sl2, -lse_l2)) * zf;
                        if (c0 + e == label) pr -= (1.f - smoothing);
                        o[e] = dl * (pr - sm_v);
                    }
                }
                store8<kDt>(dlogits, ob + c0, o);
            }
        }
    } else {
        for (int c = slice0 + threadIdx.x; c < vocab && c < slice0 + kCeThreads * 8 * kCeBwdIters; c += kCeThreads) {
            float o = 0.f;
            if (!ignored) {
                float pr = __expf(load1<kDt>(logits, ib + c) * logit_scale - lse) * zf;
                if (c == label) pr -= (1.f - smoothing);
                o = dl * (pr - sm_v);
            }
            store1<kDt>(dlogits, ob + c, o);
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
#define B200T5_DISPATCH_DT(dt, NAME, ...)              \
    switch (dt) {                                      \
        case 0: { constexpr int NAME = 0; __VA_ARGS__; break; } \
        case 1: { constexpr int NAME = 1; __VA_ARGS__; break; } \
        default: { constexpr int NAME = 2; __VA_ARGS__; break; } \
    }

// row ownership for the vectorised RMSNorm kernels: warp per row up to n = 1024, block per row up to 8192
struct RmsPlan {
    bool vec;
    int threads_per_row;   // 32 or 256
    int chunks;            // 1, 2 or 4
};
static RmsPlan rms_plan(int n, bool aligned) {
    RmsPlan pl{false, 32, 1};
    if (!aligned || n % 8 != 0 || n > 8192) return pl;
    pl.vec = true;
    pl.threads_per_row = n <= 1024 ? 32 : 256;
    const int per = pl.threads_per_row * 8;
    const int need = (n + per - 1) / per;
    pl.chunks = need;          // 1..4
    return pl;
}

#define B200T5_RMS_PLAN_DISPATCH(pl, MACRO)                 \
    if (pl.threads_per_row == 32) {                         \
        switch (pl.chunks) {                                \
            case 1: MACRO(1, 32); break;                    \
            case 2: MACRO(2, 32); break;                    \
            case 3: MACRO(3, 32); break;                    \
            default: MACRO(4, 32); break;                   \
        }                                                   \
    } else {                                                \
        switch (pl.chunks) {                                \
            case 1: MACRO(1, 256); break;                   \
            case 2: MACRO(2, 256); break;                   \
            case 3: MACRO(3, 256); break;                   \
            default: MACRO(4, 256); break;                  \
        }                                                   \
    }

cudaError_t launch_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int n, int64_t xs,
                               int64_t ys, float eps, int x_dtype, int w_dtype, cudaStream_t stream) {
    const int xb = x_dtype == 2 ? 4 : 2;
    const RmsPlan pl = rms_plan(n, aligned16(x) && aligned16(y) && aligned16(w) && (xs * xb) % 16 == 0 && (ys * xb) % 16 == 0);
    if (pl.vec) {
        const int groups = 256 / pl.threads_per_row;
        const int blocks = std::min((rows + groups - 1) / groups, 148 * 8);
#define B200T5_RMS_FWD(CH, T)                                                                                   \
    B200T5_DISPATCH_DT(x_dtype, XD, B200T5_DISPATCH_DT(w_dtype, WD,                                            \
        (rmsnorm_fwd_vec_kernel<XD, WD, CH, T><<<blocks, 256, 0, stream>>>(x, w, y, rstd, rows, n, xs, ys, eps))))
        B200T5_RMS_PLAN_DISPATCH(pl, B200T5_RMS_FWD)
#undef B200T5_RMS_FWD
    } else {
        const int blocks = std::min(rows, 148 * 8);
        B200T5_DISPATCH_DT(x_dtype, XD, B200T5_DISPATCH_DT(w_dtype, WD,
            (rmsnorm_fwd_generic_kernel<XD, WD><<<blocks, 256, 0, stream>>>(x, w, y, rstd, rows, n, xs, ys, eps))));
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, void* dw,
                               float* dw_partial, int rows, int n, int64_t dys, int64_t xs, int64_t dxs, int x_dtype,
                               int w_dtype, cudaStream_t stream) {
    const int xb = x_dtype == 2 ? 4 : 2;
    const RmsPlan pl = rms_plan(n, aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(w) && (xs * xb) % 16 == 0 &&
                                       (dys * xb) % 16 == 0 && (dxs * xb) % 16 == 0);
    int blocks = 0;
    if (rows > 0) {
        if (pl.vec) {
            const int groups = 256 / pl.threads_per_row;
            blocks = std::min((rows + groups - 1) / groups, kRmsnormMaxPartials);
            const size_t smem = 2 * (size_t)n * sizeof(float);
#define B200T5_RMS_BWD(CH, T)                                                                                   \
    B200T5_DISPATCH_DT(x_dtype, XD, B200T5_DISPATCH_DT(w_dtype, WD, {                                          \
        if (smem > 48 * 1024)                                                                                   \
            cudaFuncSetAttribute(rmsnorm_bwd_vec_kernel<XD, WD, CH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        rmsnorm_bwd_vec_kernel<XD, WD, CH, T><<<blocks, 256, smem, stream>>>(dy, x, w, rstd, dx, dw_partial, rows, n, dys, xs, dxs); }))
            B200T5_RMS_PLAN_DISPATCH(pl, B200T5_RMS_BWD)
#undef B200T5_RMS_BWD
        } else {
            blocks = std::min(rows, kRmsnormMaxPartials);
            B200T5_DISPATCH_DT(x_dtype, XD, B200T5_DISPATCH_DT(w_dtype, WD,
                (rmsnorm_bwd_generic_kernel<XD, WD><<<blocks, 256, 0, stream>>>(dy, x, w, rstd, dx, dw_partial, rows, n, dys, xs, dxs))));
        }
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const int fb = (n + 255) / 256;
    B200T5_DISPATCH_DT(w_dtype, WD, (rmsnorm_dw_finalize_kernel<WD><<<fb, 256, 0, stream>>>(dw_partial, blocks, dw, n)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_ce_fwd(const void* logits, const int64_t* labels, float* losses, float* z_losses, float* lse,
                          bool lse_is_input, int rows, int vocab, int64_t row_stride, float smoothing, float logit_scale,
                          float lse_square_scale, int64_t ignore_index, int dtype, cudaStream_t stream) {
    const int eb = dtype == 2 ? 4 : 2;
    const bool vec = (vocab % 8 == 0) && aligned16(logits) && (row_stride * eb) % 16 == 0;
    if (vec) {
        B200T5_DISPATCH_DT(dtype, DT, (ce_fwd_kernel<DT, true><<<rows, kCeThreads, 0, stream>>>(
            logits, labels, losses, z_losses, lse, lse_is_input ? 1 : 0, vocab, row_stride, smoothing, logit_scale, lse_square_scale, ignore_index)));
    } else {
        B200T5_DISPATCH_DT(dtype, DT, (ce_fwd_kernel<DT, false><<<rows, kCeThreads, 0, stream>>>(
            logits, labels, losses, z_losses, lse, lse_is_input ? 1 : 0, vocab, row_stride, smoothing, logit_scale, lse_square_scale, ignore_index)));
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* dlosses,
                          int64_t dloss_stride, void* dlogits, int rows, int vocab, int64_t row_stride,
                          int64_t dl_row_stride, float smoothing, float logit_scale, float lse_square_scale,
                          int64_t ignore_index, int dtype, cudaStream_t stream) {
    const int eb = dtype == 2 ? 4 : 2;
    const bool vec = (vocab % 8 == 0) && aligned16(logits) && aligned16(dlogits) && (row_stride * eb) % 16 == 0 &&
                     (dl_row_stride * eb) % 16 == 0;
    const int slice = kCeThreads * 8 * kCeBwdIters;
    const dim3 grid(rows, (vocab + slice - 1) / slice);
    if (vec) {
        B200T5_DISPATCH_DT(dtype, DT, (ce_bwd_kernel<DT, true><<<grid, kCeThreads, 0, stream>>>(
            logits, labels, lse, dlosses, dloss_stride, dlogits, vocab, row_stride, dl_row_stride, smoothing, logit_scale,
            lse_square_scale, ignore_index)));
    } else {
        B200T5_DISPATCH_DT(dtype, DT, (ce_bwd_kernel<DT, false><<<grid, kCeThreads, 0, stream>>>(
            logits, labels, lse, dlosses, dloss_stride, dlogits, vocab, row_stride, dl_row_stride, smoothing, logit_scale,
            lse_square_scale, ignore_index)));
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200t5
