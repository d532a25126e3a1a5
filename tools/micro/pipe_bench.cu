// Microbenchmark (developer tool): per-SM throughput of the instructions the softmax loops are made of.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

template <int kMode>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float a[16];
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f;
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (kMode == 0 || kMode == 3 || kMode == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (kMode == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(0.001f));
            if (kMode == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(0.001f));
            // mixed-precision add (sm_100: FHADD.BF16): fp32 += one bf16 half of a packed register -- the dense-bias
            // add when sm_scale == 1 without a separate unpack
            if (kMode == 5) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.bf16 %0, lo, %0;}" : "+f"(a[i]) : "r"(0x3c003c00u + i));
            // the same add with an explicit unpack (what the kernels do today: shift + fma)
            if (kMode == 6) {
                float b = __uint_as_float((0x3c003c00u + i + threadIdx.x) << 16);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(b));
            }
        }
        if (kMode == 7) {
            // packed fp32x2 fma (FFMA2)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, m = 0x3f7fbe773f7fbe77ull, c = 0x3a83126f3a83126full;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(m), "l"(c));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(v));
            }
        }
        if (kMode == 8) {
            // three-input max (FMNMX3)
#pragma unroll
            for (int i = 0; i < 16; i += 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[i + 1]), "f"(seed));
        }
        if (kMode == 2 || kMode == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[2 * i + 1]), "f"(a[2 * i]));
                pk[i] ^= r;
                if (kMode == 2) { a[2 * i] += 1.0f; }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float(pk[i] & 0x3f800000);
    if (s == 123.456f) out[0] = s;
}

template <int kMode>
void run(const char* name, double ops_per_thread_iter) {
    float* out; cudaMalloc(&out, 4);
    const int iters = 4096, grid = 148 * 8, block = 256;
    k<kMode><<<grid, block>>>(out, 16, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<kMode><<<grid, block>>>(out, iters, 0.5f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ops = (double)grid * block * iters * ops_per_thread_iter;
    printf("%-34s %8.3f ms  %8.1f Gop/s  = %6.2f ops/clk/SM at %d MHz (max clock)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}

int main() {
    run<0>("ex2.approx (16/iter)", 16);
    run<1>("fma.rn.f32 (16/iter)", 16);
    run<2>("cvt.rn.bf16x2.f32 (8/iter) + 8 fadd", 8);
    run<3>("ex2 + fma interleaved (32/iter)", 32);
    run<4>("ex2 (16) + cvt pack (8)", 24);
    run<5>("add.rn.f32.bf16 (16/iter)", 16);
    run<6>("shl + fma.rn.f32 (16+16/iter)", 16);
    run<7>("fma.rn.f32x2 (8/iter = 16 lanes)", 16);
    run<8>("max.f32 3-input (8/iter)", 8);
    return 0;
}
