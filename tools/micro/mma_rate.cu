// Microbenchmark (developer tool, round 2): cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N and of the
// operand sources (A from shared memory / from TMEM; B K-major / MN-major), issued back to back by one thread.  Two numbers per
// case: how long the issuing thread needs per instruction, and how long the tensor pipe needs (issue start -> commit arrival).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I flasht5_b200/csrc -o tools/micro/mma_rate tools/micro/mma_rate.cu
#include "common.cuh"

using namespace b200t5;

__device__ long long g_out[148][2];

template <int kN, bool kTS, bool kBMn, bool kAMn>
__global__ void __launch_bounds__(128, 1) mma_rate(int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 98304 + 16);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (warp == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(true, 128, kN, kAMn, kBMn);
        constexpr uint32_t hi = sdesc_hi(1024, kSwz128);
        const uint32_t a_lo = sdesc_lo(smem_u32(smem), kAMn ? 16384 : 16);
        const uint32_t b_lo = sdesc_lo(smem_u32(smem + 32768), kBMn ? 16384 : 16);
        long long t0 = 0, t1 = 0, t2 = 0;
        for (int rep = 0; rep < 2; ++rep) {          // first pass warms up
            t0 = clock64();
            if (leader) {
                for (int i = 0; i < reps; ++i) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        // A, B advance by one K-step (32 bytes K-major, 16 rows = 2048 bytes MN-major) like the real kernels
                        const uint32_t ao = kAMn ? kk * (2048 >> 4) : kk * 2;
                        const uint32_t bo = kBMn ? kk * (2048 >> 4) : kk * 2;
                        if (kTS) umma_ts2(tm + (i & 1) * 128, tm + 256 + kk * 8, b_lo + bo, hi, idesc, kk > 0 ? 1u : 0u);
                        else umma_ss2(tm + (i & 1) * 128, a_lo + ao, hi, b_lo + bo, hi, idesc, kk > 0 ? 1u : 0u);
                    }
                }
                umma_commit(bar);
            }
            __syncwarp();
            t1 = clock64();
            mbar_wait(bar, rep & 1);
            t2 = clock64();
        }
        if ((threadIdx.x & 31) == 0) {
            g_out[blockIdx.x][0] = t1 - t0;
            g_out[blockIdx.x][1] = t2 - t0;
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tm);
    }
}

template <int kN, bool kTS, bool kBMn, bool kAMn>
static void run(const char* name) {
    const int reps = 64;
    auto kern = mma_rate<kN, kTS, kBMn, kAMn>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 64);
    kern<<<148, 128, 98304 + 64>>>(reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148][2];
    cudaMemcpyFromSymbol(h, g_out, sizeof(h));
    double iss = 0, tot = 0;
    for (int i = 0; i < 148; ++i) {
        iss += (double)h[i][0];
        tot += (double)h[i][1];
    }
    const double n = 148.0 * reps * 4;
    printf("%-52s N %3d  issue %6.1f cyc/mma   complete %6.1f cyc/mma   (ideal %5.1f)%s\n", name, kN, iss / n, tot / n, kN / 2.0,
           e == cudaSuccess ? "" : "  [CUDA ERROR]");
}

int main() {
    run<128, false, false, false>("SS  A K-major smem, B K-major");
    run<64, false, false, false>("SS  A K-major smem, B K-major");
    run<32, false, false, false>("SS  A K-major smem, B K-major");
    run<64, false, true, true>("SS  A MN-major smem, B MN-major (dQ, old dV/dK)");
    run<64, false, true, false>("SS  A K-major smem, B MN-major");
    run<128, true, false, false>("TS  A tmem, B K-major");
    run<64, true, false, false>("TS  A tmem, B K-major (S^T, dP^T half tiles)");
    run<32, true, false, false>("TS  A tmem, B K-major (S^T, dP^T 32-query sub-tiles)");
    run<16, true, false, false>("TS  A tmem, B K-major");
    run<64, true, true, false>("TS  A tmem, B MN-major (dV, dK; forward PV)");
    run<32, true, true, false>("TS  A tmem, B MN-major");
    run<128, true, true, false>("TS  A tmem, B MN-major");
    run<256, false, false, false>("SS  A K-major smem, B K-major");
    run<256, true, false, false>("TS  A tmem, B K-major");
    return 0;
}
