// Microbenchmark (developer tool): throughput of cp.reduce.async.bulk (smem -> global add) into an
// L2-resident surface, as a function of chunk size and element type.  Decides the dBias / dQ strategy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_reduce_bench tma_reduce_bench.cu && ./tma_reduce_bench
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: add.f32   mode 1: add.noftz.bf16   mode 2: plain bulk store (no reduce)   mode 3: red.global.add.v4.f32 per thread
template <int kMode>
__global__ void __launch_bounds__(128) bench_kernel(uint8_t* dst, size_t surface_bytes, int chunk_bytes, int iters, int tile_bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    for (int i = threadIdx.x; i < tile_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t ntiles = surface_bytes / tile_bytes;
    for (int it = 0; it < iters; ++it) {
        // every CTA walks the surface with a different phase so that, like the attention backward, many CTAs
        // hit the same addresses at about the same time (32 batch CTAs share one dBias tile)
        const size_t tile = ((size_t)(blockIdx.x / 32) * 131 + it * 7) % ntiles;
        uint8_t* g = dst + tile * tile_bytes;
        if (kMode == 3) {
            float4* gp = reinterpret_cast<float4*>(g);
            for (int i = threadIdx.x; i < tile_bytes / 16; i += blockDim.x)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp + i), "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f) : "memory");
        } else if (threadIdx.x == 0) {
            for (int off = 0; off < tile_bytes; off += chunk_bytes) {
                if (kMode == 0)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g + off), "r"(smem_u32(smem + off)), "r"(chunk_bytes) : "memory");
                else if (kMode == 1)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;" ::"l"(g + off), "r"(smem_u32(smem + off)), "r"(chunk_bytes) : "memory");
                else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + off), "r"(smem_u32(smem + off)), "r"(chunk_bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int kMode>
float run(uint8_t* dst, size_t surface, int chunk, int iters, int tile, int grid) {
    cudaFuncSetAttribute(bench_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    bench_kernel<kMode><<<grid, 128, tile>>>(dst, surface, chunk, 4, tile);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    bench_kernel<kMode><<<grid, 128, tile>>>(dst, surface, chunk, iters, tile);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    return ms;
}

int main() {
    const size_t surface = 32u << 20;           // 32 MiB: L2 resident (like the fp32 dBias of the headline shape)
    uint8_t* dst; cudaMalloc(&dst, surface); cudaMemset(dst, 0, surface);
    const int iters = 64;
    for (int grid : {37 * 4, 74 * 4, 148, 148 * 2, 148 * 4, 148 * 8}) {
    const char* names[4] = {"bulk add.f32", "bulk add.bf16", "bulk store", "red.v4.f32"};
    for (int tile : {16384}) {
        for (int chunk : {16384}) {
            if (chunk > tile) continue;
            float ms[4];
            ms[0] = run<0>(dst, surface, chunk, iters, tile, grid);
            ms[1] = run<1>(dst, surface, chunk, iters, tile, grid);
            ms[2] = run<2>(dst, surface, chunk, iters, tile, grid);
            ms[3] = run<3>(dst, surface, chunk, iters, tile, grid);
            const double bytes = (double)grid * iters * tile;
            printf("tile %6d B chunk %6d B grid %d:", tile, chunk, grid);
            for (int m = 0; m < 4; ++m) printf("  %s %.0f GB/s", names[m], bytes / ms[m] / 1e6);
            printf("\n");
        }
    }
    }
    return 0;
}
