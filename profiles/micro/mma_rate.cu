// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32) as a function of N, for cta_group::2 (M = 256 over a CTA
// pair, both operands in shared memory) and cta_group::1 (M = 128), with every SM busy (one CTA per SM), so shared-memory
// operand fetch and the pair's cross-SM B traffic are as in the real kernels but no TMA / epilogue competes.
// Question (profiles/r02_notes.md): do the N = 192 MMAs of dsam_fwd_kernel / ratio_front_kernel run at the N/2-cycle floor?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../rgb-d-instance-segmentation_b200/csrc -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "tc_ptx.cuh"

constexpr int kThreads = 128;

// PAIR: cta_group::2.  Each CTA: A tile 128 rows x 64 bf16 (16 KB, SW128 layout - contents irrelevant), B tile (N/2 | N) rows.
// LOAD: a third warp streams 16 KB bulk copies (global/L2 -> shared memory, 4 in flight) into a scratch region for as long
// as the MMAs run: do TMA writes landing in the same shared memory slow the tensor core's operand fetch?
template <bool PAIR, bool LOAD>
__global__ void __launch_bounds__(kThreads, 1) mma_rate_kernel(int N, int kblocks, int reps, long long* cycles, const uint8_t* src,
                                                               long long* loaded) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_a = smem;                         // 4 A tiles (64 KB) so consecutive K blocks read different addresses
    uint8_t* s_b = smem + 4 * 16384;             // 4 B tiles of up to 256 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint64_t lbar[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int done;
    uint8_t* s_scratch = smem + 4 * 16384 + 4 * 16384;      // LOAD runs with N <= 256 on pairs: B tiles are <= 16 KB each
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
    if (warp == 0 && lane == 0) {
        tc::mbar_init(&bar, 1);
        for (int i = 0; i < 4; ++i) tc::mbar_init(&lbar[i], 1);
        done = 0;
        tc::fence_barrier_init();
    }
    if (PAIR) tc::cluster_sync_all(); else __syncthreads();
    if (warp == 1) { if (PAIR) tc::tmem_alloc_2cta(&tmem_base_s, 512); else tc::tmem_alloc(&tmem_base_s, 512); }
    tc::tc_fence_before();
    if (PAIR) tc::cluster_sync_all(); else __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const bool issuer = warp == 0 && (!PAIR || tc::cluster_ctarank() == 0);
    tc::fence_proxy_async();
    if (issuer) {
        const uint32_t idesc = tc::make_idesc_bf16(PAIR ? 256 : 128, (uint32_t)N);
        const int b_rows = PAIR ? N / 2 : N;
        uint32_t phase = 0;
        long long t0 = 0, t1 = 0;
        for (int r = 0; r < reps + 1; ++r) {
            if (r == 1) t0 = clock64();
            for (int kb = 0; kb < kblocks; ++kb) {
                const uint64_t adesc = tc::make_kmajor_desc(tc::smem_u32(s_a + (kb & 3) * 16384), 128);
                const uint64_t bdesc = tc::make_kmajor_desc(tc::smem_u32(s_b + (kb & 3) * (b_rows * 128)), 128);
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (PAIR) tc::umma_bf16_2cta(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                        else tc::umma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    }
                }
                __syncwarp();
            }
            if (tc::elect_one()) { if (PAIR) tc::umma_commit_2cta(&bar); else tc::umma_commit(&bar); }
            __syncwarp();
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
        }
        t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
        done = 1;
    } else if (PAIR && warp == 0) {
        // the peer CTA: the multicast commits also arrive on its barrier; follow them to know when the run is over
        uint32_t phase = 0;
        for (int r = 0; r < reps + 1; ++r) { tc::mbar_wait(&bar, phase); phase ^= 1; }
        done = 1;
    } else if (LOAD && warp == 2 && lane == 0) {
        long long n = 0;
        uint32_t ph[4] = {0, 0, 0, 0};
        for (int it = 0; !done; ++it) {
            const int sl = it & 3;
            if (it >= 4) { tc::mbar_wait(&lbar[sl], ph[sl]); ph[sl] ^= 1; ++n; }
            tc::mbar_expect_tx(&lbar[sl], 16384);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tc::smem_u32(s_scratch + sl * 16384)), "l"(src + (size_t)((it * 37 + blockIdx.x) & 63) * 16384), "r"(16384),
                           "r"(tc::smem_u32(&lbar[sl]))
                         : "memory");
        }
        loaded[blockIdx.x] = n * 16384;
        for (int sl = 0; sl < 4; ++sl) tc::mbar_wait(&lbar[sl], ph[sl]);      // drain before the CTA exits
    }
    tc::tc_fence_before();
    if (PAIR) tc::cluster_sync_all(); else __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); if (PAIR) tc::tmem_dealloc_2cta(tmem, 512); else tc::tmem_dealloc(tmem, 512); }
}

template <bool PAIR, bool LOAD = false>
double run(int N, int kblocks, int reps, long long* cyc, const uint8_t* src = nullptr, long long* loaded = nullptr,
           double* load_rate = nullptr) {
    const int smem = 1024 + 4 * 16384 + 4 * 32768;
    cudaFuncSetAttribute(mma_rate_kernel<PAIR, LOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(cyc, 0, 148 * 8);
    if (loaded) cudaMemset(loaded, 0, 148 * 8);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<PAIR, LOAD>, N, kblocks, reps, cyc, src, loaded);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return -1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return -1; }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0; int n = 0;
    for (int i = 0; i < 148; ++i) if (h[i] > 0) { sum += (double)h[i]; ++n; }
    if (load_rate) {
        long long l[148];
        cudaMemcpy(l, loaded, sizeof(l), cudaMemcpyDeviceToHost);
        double bytes = 0;
        for (int i = 0; i < 148; ++i) bytes += (double)l[i];
        *load_rate = bytes / 148 / (sum / n);            // bytes per cycle per SM while the MMAs ran
    }
    return sum / n / ((double)reps * kblocks * 4);
}

int main() {
    long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    const int kblocks = 20, reps = 200;
    printf("cycles per tcgen05.mma (K = 16), all 148 SMs busy, %d MMAs per commit\n", kblocks * 4);
    printf("%6s %22s %22s\n", "N", "cta_group::2 (M=256)", "cta_group::1 (M=128)");
    for (int N : {64, 96, 128, 160, 192, 224, 256}) {
        const double p = run<true>(N, kblocks, reps, cyc);
        const double s = run<false>(N, kblocks, reps, cyc);
        printf("%6d %12.1f (N/2=%3d) %12.1f (N/2=%3d)\n", N, p, N / 2, s, N / 2);
    }
    uint8_t* src;
    long long* loaded;
    cudaMalloc(&src, 64 * 16384);
    cudaMemset(src, 1, 64 * 16384);
    cudaMalloc(&loaded, 148 * 8);
    printf("\ncta_group::2 with a concurrent 16 KB bulk-copy stream (L2-resident source) into the same shared memory\n");
    for (int N : {128, 192, 256}) {
        double rate = 0;
        const double q = run<true, true>(N, kblocks, reps, cyc, src, loaded, &rate);
        printf("%6d %12.1f cycles per MMA with %5.1f B/clk/SM of bulk copies landing\n", N, q, rate);
    }
    return 0;
}
