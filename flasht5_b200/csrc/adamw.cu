// Fused multi-tensor AdamWScale step for sm_100a.
//
// Replaces /root/reference/src/utils/adamw_scaled.py:154-211 (`_adamwscaled`) and :213-281 (`_foreach_adamwscaled`):
// Adam moments, step size scaled by max(1e-3, rms(parameter)), optional Kahan compensation for 16-bit parameters,
// decoupled weight decay.  The reference's foreach path makes ~15 passes over every tensor of a dtype group and calls
// .item() once per tensor; here one step is three launches for ALL tensors of a group and no host synchronisation:
//
//   1. adamw_sumsq_kernel   : sum of squares of every 4096-element chunk of every parameter  (reads p once)
//   2. adamw_rms_kernel     : one warp per tensor adds its chunk partials in a fixed order -> step size of the tensor
//   3. adamw_update_kernel  : m, v, (Kahan term,) p updated in one pass
//
// HBM-bound.  Algorithmic bytes per element: pass 1 reads p; pass 3 reads p, g, m, v (+c) and writes p, m, v (+c):
// fp32 everything = 32 B, bf16 + Kahan = 20 B.
//
// Rounding follows the reference's in-place tensor ops (every op rounds to the tensor's dtype; the rms and, without
// bias correction, the step size of a 16-bit parameter are 16-bit tensors there): see oracle/adamw_ref.py.
#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

namespace {

constexpr int kChunk = 4096;          // elements per block: 256 threads x 4 elements x 4 iterations

template <int kDt>
__device__ __forceinline__ float ldf(const void* p, int64_t i) {
    if constexpr (kDt == 2) return static_cast<const float*>(p)[i];
    else return to_float16bit<kDt == 1>(static_cast<const uint16_t*>(p)[i]);
}
template <int kDt>
__device__ __forceinline__ float rnd(float x) {      // round to the storage dtype, return as fp32
    if constexpr (kDt == 2) return x;
    else return to_float16bit<kDt == 1>(static_cast<uint16_t>(pack2<kDt == 1>(x, 0.f) & 0xFFFFu));
}
template <int kDt>
__device__ __forceinline__ void stf(void* p, int64_t i, float x) {   // x is already representable in the dtype
    if constexpr (kDt == 2) static_cast<float*>(p)[i] = x;
    else static_cast<uint16_t*>(p)[i] = static_cast<uint16_t>(pack2<kDt == 1>(x, 0.f) & 0xFFFFu);
}
template <int kDt>
__device__ __forceinline__ void ld4(const void* p, int64_t i, float (&o)[4]) {      // i % 4 == 0, 16/8-byte aligned
    if constexpr (kDt == 2) {
        const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(p) + i);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    } else {
        const uint2 v = *reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(p) + i);
        const float2 a = unpack2<kDt == 1>(v.x), b = unpack2<kDt == 1>(v.y);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
}
template <int kDt>
__device__ __forceinline__ void st4(void* p, int64_t i, const float (&o)[4]) {
    if constexpr (kDt == 2) {
        *reinterpret_cast<float4*>(static_cast<float*>(p) + i) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        uint2 v;
        v.x = pack2<kDt == 1>(o[0], o[1]);
        v.y = pack2<kDt == 1>(o[2], o[3]);
        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(p) + i) = v;
    }
}

__device__ __forceinline__ float block_sum256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 8) t = red[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;       // valid in thread 0
}

// ---- pass 1: sum of squares per chunk ----
template <int kPDt>
__global__ void __launch_bounds__(256) adamw_sumsq_kernel(const AdamwTensor* __restrict__ tensors,
                                                          const int32_t* __restrict__ chunk_tensor,
                                                          float* __restrict__ chunk_sumsq) {
    __shared__ float red[8];
    const AdamwTensor t = tensors[chunk_tensor[blockIdx.x]];
    const int64_t base = (int64_t)(blockIdx.x - t.first_chunk) * kChunk;
    const int64_t end = base + kChunk < t.numel ? base + kChunk : t.numel;
    const bool vec = (reinterpret_cast<uintptr_t>(t.p) & 15) == 0;
    float s = 0.f;
    if (vec) {
        for (int64_t i = base + threadIdx.x * 4; i + 3 < end; i += 1024) {
            float x[4];
            ld4<kPDt>(t.p, i, x);
            s += x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
        }
        for (int64_t i = base + ((end - base) & ~int64_t(3)) + threadIdx.x; i < end; i += 256) {
            const float x = ldf<kPDt>(t.p, i);
            s += x * x;
        }
    } else {
        for (int64_t i = base + threadIdx.x; i < end; i += 256) {
            const float x = ldf<kPDt>(t.p, i);
            s += x * x;
        }
    }
    s = block_sum256(s, red);
    if (threadIdx.x == 0) chunk_sumsq[blockIdx.x] = s;
}

// ---- pass 2: one warp per tensor: partials in a fixed order -> rms -> the value handed to addcdiv ----
template <int kPDt>
__global__ void __launch_bounds__(32) adamw_rms_kernel(const AdamwTensor* __restrict__ tensors,
                                                       const float* __restrict__ chunk_sumsq, float* __restrict__ neg_step,
                                                       int round_step_to_p) {
    const AdamwTensor t = tensors[blockIdx.x];
    const int n_chunks = static_cast<int>((t.numel + kChunk - 1) / kChunk);
    float s = 0.f;
    for (int c = threadIdx.x; c < n_chunks; c += 32) s += chunk_sumsq[t.first_chunk + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
        // tensor.norm(2) / numel ** 0.5: both results are tensors of p's dtype (adamw_scaled.py:66-68)
        const float norm = rnd<kPDt>(sqrtf(s));
        const float rms = rnd<kPDt>(norm / t.sqrt_numel);
        float ss;
        if (rms > 1e-3f) {
            ss = t.ss_base * rms;                       // fp32 tensor x rms tensor (with bias correction) ...
            if (round_step_to_p) ss = rnd<kPDt>(ss);    // ... or python float x 16-bit tensor -> 16-bit tensor (without)
        } else {
            ss = t.ss_floor;                            // step size x the python float 1e-3
        }
        neg_step[blockIdx.x] = -ss;
    }
}

// ---- pass 3: the update ----
template <int kPDt, int kSDt, bool kKahan>
__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, float& c, float beta1, float om_beta1,
                                           float beta2, float om_beta2, float eps, float value, float neg_lr_wd, bool wd) {
    m = rnd<kSDt>(m * beta1);                           // exp_avg.mul_(beta1)
    m = rnd<kSDt>(m + om_beta1 * g);                    //        .add_(grad, alpha=1-beta1)
    v = rnd<kSDt>(v * beta2);                           // exp_avg_sq.mul_(beta2)
    v = rnd<kSDt>(v + om_beta2 * (g * g));              //           .addcmul_(grad, grad, value=1-beta2)
    float d = rnd<kSDt>(sqrtf(v));                      // denom = exp_avg_sq.sqrt()
    d = rnd<kSDt>(d + eps);                             //        .add_(eps)
    const float q = value * (m / d);
    if constexpr (kKahan) {
        c = rnd<kPDt>(c + q);                           // kahan_comp.addcdiv_(exp_avg, denom, value=-step_size)
        const float old = p;
        p = rnd<kPDt>(p + c);                           // p.add_(kahan_comp)
        const float lost = rnd<kPDt>(old - p);          // grad.copy_(p_old).sub_(p)
        c = rnd<kPDt>(c + lost);                        // kahan_comp.add_(grad)
    } else {
        p = rnd<kPDt>(p + q);                           // p.addcdiv_(exp_avg, denom, value=-step_size)
    }
    if (wd) p = rnd<kPDt>(p + neg_lr_wd * p);           // p.add_(p, alpha=-lr*weight_decay)
}

template <int kPDt, int kSDt, bool kKahan>
__global__ void __launch_bounds__(256) adamw_update_kernel(const AdamwTensor* __restrict__ tensors,
                                                           const int32_t* __restrict__ chunk_tensor,
                                                           const float* __restrict__ neg_step, float beta1, float om_beta1,
                                                           float beta2, float om_beta2, float eps) {
    const int ti = chunk_tensor[blockIdx.x];
    const AdamwTensor t = tensors[ti];
    const float value = neg_step[ti];
    const bool wd = t.neg_lr_wd != 0.f;
    const int64_t base = (int64_t)(blockIdx.x - t.first_chunk) * kChunk;
    const int64_t end = base + kChunk < t.numel ? base + kChunk : t.numel;
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                       reinterpret_cast<uintptr_t>(t.v) | (kKahan ? reinterpret_cast<uintptr_t>(t.comp) : 0)) & 15) == 0;
    int64_t scalar_from = base;
    if (vec) {
        for (int64_t i = base + threadIdx.x * 4; i + 3 < end; i += 1024) {
            float p[4], g[4], m[4], v[4], c[4] = {0.f, 0.f, 0.f, 0.f};
            ld4<kPDt>(t.p, i, p);
            ld4<kPDt>(t.g, i, g);
            ld4<kSDt>(t.m, i, m);
            ld4<kSDt>(t.v, i, v);
            if constexpr (kKahan) ld4<kPDt>(t.comp, i, c);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                adamw_elem<kPDt, kSDt, kKahan>(p[e], g[e], m[e], v[e], c[e], beta1, om_beta1, beta2, om_beta2, eps, value,
                                               t.neg_lr_wd, wd);
            st4<kPDt>(t.p, i, p);
            st4<kSDt>(t.m, i, m);
            st4<kSDt>(t.v, i, v);
            if constexpr (kKahan) st4<kPDt>(t.comp, i, c);
        }
        scalar_from = base + ((end - base) & ~int64_t(3));
    }
    for (int64_t i = scalar_from + threadIdx.x; i < end; i += 256) {
        float p = ldf<kPDt>(t.p, i), m = ldf<kSDt>(t.m, i), v = ldf<kSDt>(t.v, i), c = 0.f;
        const float g = ldf<kPDt>(t.g, i);
        if constexpr (kKahan) c = ldf<kPDt>(t.comp, i);
        adamw_elem<kPDt, kSDt, kKahan>(p, g, m, v, c, beta1, om_beta1, beta2, om_beta2, eps, value, t.neg_lr_wd, wd);
        stf<kPDt>(t.p, i, p);
        stf<kSDt>(t.m, i, m);
        stf<kSDt>(t.v, i, v);
        if constexpr (kKahan) stf<kPDt>(t.comp, i, c);
    }
}

}  // namespace

int adamw_chunk_elems() { return kChunk; }

cudaError_t launch_adamw_step(const AdamwTensor* tensors, int n_tensors, const int32_t* chunk_tensor, int n_chunks,
                              float* chunk_sumsq, float* neg_step, int p_dtype, int state_dtype, bool kahan, float beta1,
                              float beta2, float eps, bool round_step_to_p, cudaStream_t stream) {
    if (n_tensors <= 0 || n_chunks <= 0) return cudaSuccess;
    const float om1 = static_cast<float>(1.0 - static_cast<double>(beta1));
    const float om2 = static_cast<float>(1.0 - static_cast<double>(beta2));
#define B200T5_ADAMW_P(PD)                                                                                           \
    adamw_sumsq_kernel<PD><<<n_chunks, 256, 0, stream>>>(tensors, chunk_tensor, chunk_sumsq);                        \
    adamw_rms_kernel<PD><<<n_tensors, 32, 0, stream>>>(tensors, chunk_sumsq, neg_step, round_step_to_p ? 1 : 0)
#define B200T5_ADAMW_U(PD, SD, KH)                                                                                   \
    adamw_update_kernel<PD, SD, KH><<<n_chunks, 256, 0, stream>>>(tensors, chunk_tensor, neg_step, beta1, om1, beta2, om2, eps)
    // parameter dtype: 0 fp16, 1 bf16, 2 fp32; state dtype = parameter dtype, or 16-bit states under fp32 parameters
    if (kahan && p_dtype == 2) return cudaErrorInvalidValue;           // the reference only compensates 16-bit parameters
    if (p_dtype != 2 && state_dtype != p_dtype) return cudaErrorInvalidValue;
    switch (p_dtype) {
        case 0: B200T5_ADAMW_P(0); break;
        case 1: B200T5_ADAMW_P(1); break;
        case 2: B200T5_ADAMW_P(2); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (p_dtype == 2) {
        switch (state_dtype) {
            case 0: B200T5_ADAMW_U(2, 0, false); break;
            case 1: B200T5_ADAMW_U(2, 1, false); break;
            case 2: B200T5_ADAMW_U(2, 2, false); break;
            default: return cudaErrorInvalidValue;
        }
    } else if (p_dtype == 1) {
        if (kahan) B200T5_ADAMW_U(1, 1, true); else B200T5_ADAMW_U(1, 1, false);
    } else {
        if (kahan) B200T5_ADAMW_U(0, 0, true); else B200T5_ADAMW_U(0, 0, false);
    }
#undef B200T5_ADAMW_P
#undef B200T5_ADAMW_U
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200t5
