// Microbenchmark: tcgen05.ld throughput per SM vs number of warps and loads in flight (profiles/r01_notes.md).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DEPTH>
__global__ void __launch_bounds__(512, 1) k(int iters, int nwarps, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 % 448);
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t v[DEPTH][32];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[d][0]), "=r"(v[d][1]), "=r"(v[d][2]), "=r"(v[d][3]), "=r"(v[d][4]), "=r"(v[d][5]), "=r"(v[d][6]), "=r"(v[d][7]),
                      "=r"(v[d][8]), "=r"(v[d][9]), "=r"(v[d][10]), "=r"(v[d][11]), "=r"(v[d][12]), "=r"(v[d][13]), "=r"(v[d][14]), "=r"(v[d][15]),
                      "=r"(v[d][16]), "=r"(v[d][17]), "=r"(v[d][18]), "=r"(v[d][19]), "=r"(v[d][20]), "=r"(v[d][21]), "=r"(v[d][22]), "=r"(v[d][23]),
                      "=r"(v[d][24]), "=r"(v[d][25]), "=r"(v[d][26]), "=r"(v[d][27]), "=r"(v[d][28]), "=r"(v[d][29]), "=r"(v[d][30]), "=r"(v[d][31])
                    : "r"(base + (uint32_t)(d * 32 % 64)));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= v[d][j];
        }
        t1 = clock64();
    }
    if (threadIdx.x % 32 == 0 && warp < nwarps) cycles[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 16 * 8); cudaMalloc(&sink, 4);
    const int iters = 2000;
    for (int depth = 1; depth <= 2; ++depth)
        for (int nw : {1, 4, 8, 16}) {
            cudaMemset(cyc, 0, 148 * 16 * 8);
            if (depth == 1) k<1><<<148, 512>>>(iters, nw, cyc, sink); else k<2><<<148, 512>>>(iters, nw, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[16];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            const double bytes = (double)iters * depth * 4096.0 * nw;
            printf("depth %d warps %2d: %lld cycles, %.1f B/clk/SM, %.0f cycles per x32 load per warp (%s)\n", depth, nw, mx, bytes / mx,
                   (double)mx / (iters * depth), cudaGetErrorString(e));
        }
    return 0;
}
