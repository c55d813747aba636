// Probe: does a tiled TMA load accept a dimension whose stride (16 B) is smaller than the inner extent (64 B), i.e. an
// overlapping "sliding window" view (im2col along x for free)?  nvcc -arch=sm_100a -o tma_overlap tma_overlap.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__global__ void k(const __grid_constant__ CUtensorMap m, uint16_t* out) {
    __shared__ __align__(1024) uint16_t tile[64 * 32];
    __shared__ uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), t = (uint32_t)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(64 * 64) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(t), "l"(reinterpret_cast<uint64_t>(&m)), "r"(b), "r"(0), "r"(3) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.b32 %0,1,0,P;}" : "=r"(ok) : "r"(b) : "memory");
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) out[i] = tile[i];
}

int main() {
    const int n = 4096;
    uint16_t* d; uint16_t* o;
    cudaMalloc(&d, n * 2); cudaMalloc(&o, 64 * 32 * 2);
    uint16_t h[n]; for (int i = 0; i < n; ++i) h[i] = (uint16_t)i;
    cudaMemcpy(d, h, n * 2, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[2] = {32, 200}; cuuint64_t strides[1] = {16}; cuuint32_t box[2] = {32, 64}; cuuint32_t es[2] = {1, 1};
    cuInit(0);
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    if (r) return 0;
    k<<<1, 128>>>(m, o);
    printf("run: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    uint16_t ho[64 * 32]; cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 64; ++row) for (int c = 0; c < 32; ++c) bad += ho[row * 32 + c] != (uint16_t)((row + 3) * 8 + c);
    printf("row0: %d %d ... row1: %d ; mismatches %d (expect value = (row+3)*8 + col)\n", ho[0], ho[1], ho[32], bad);
    return 0;
}
