// Microbenchmark (developer tool): aggregate L2 -> SM bandwidth of TMA bulk loads from an L2-resident
// buffer, unicast vs cluster multicast.  Decides whether bias / K / V tiles should be multicast.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned; barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int kStages = 4;
constexpr int kChunk = 16384;

// kCluster = 1: every CTA loads its own chunks.  kCluster = 2 / 4: the CTAs of a cluster want the SAME chunks
// (like CTAs sharing a bias tile); CTA r issues chunk j when j % kCluster == r, multicast to all.
template <int kCluster>
__global__ void __launch_bounds__(128) bench(const uint8_t* src, size_t bytes, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kChunk);
    const uint32_t rank = kCluster > 1 ? cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(full + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (kCluster > 1) cluster_sync();
    const size_t nchunks = bytes / kChunk;
    const size_t cluster_id = blockIdx.x / kCluster;
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters + kStages; ++it) {
            if (it >= kStages) mbar_wait(full + (it % kStages), ((it / kStages) - 1) & 1);   // previous use of the stage landed
            if (it < iters) {
                const int s = it % kStages;
                const size_t chunk = (cluster_id * 977 + (size_t)it * 131) % nchunks;
                mbar_expect(full + s, kChunk);
                if (kCluster == 1) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(smem + s * kChunk)), "l"(src + chunk * kChunk), "r"(kChunk), "r"(smem_u32(full + s)) : "memory");
                } else if ((uint32_t)(it % kCluster) == rank) {
                    const uint16_t mask = (1u << kCluster) - 1;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                                 ::"r"(smem_u32(smem + s * kChunk)), "l"(src + chunk * kChunk), "r"(kChunk), "r"(smem_u32(full + s)), "h"(mask) : "memory");
                }
            }
        }
    }
    __syncthreads();
    if (kCluster > 1) cluster_sync();
}

template <int kCluster>
float run(const uint8_t* src, size_t bytes, int iters, int grid) {
    const int smem = kStages * kChunk + 64;
    cudaFuncSetAttribute(bench<kCluster>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaLaunchKernelEx(&cfg, bench<kCluster>, src, bytes, 8);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    cudaLaunchKernelEx(&cfg, bench<kCluster>, src, bytes, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    return ms;
}

int main() {
    const int iters = 256;
    for (size_t mb : {32, 512}) {
        size_t bytes = mb << 20;
        uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
        for (int occ : {1, 2}) {
            const int grid = 148 * occ;
            const double delivered = (double)grid * iters * kChunk;
            float m1 = run<1>(src, bytes, iters, grid), m2 = run<2>(src, bytes, iters, grid), m4 = run<4>(src, bytes, iters, grid);
            printf("buffer %4zu MB, %d CTA/SM: delivered-to-smem GB/s: unicast %.0f | multicast x2 %.0f | multicast x4 %.0f\n", mb, occ,
                   delivered / m1 / 1e6, delivered / m2 / 1e6, delivered / m4 / 1e6);
        }
        cudaFree(src);
    }
    return 0;
}
