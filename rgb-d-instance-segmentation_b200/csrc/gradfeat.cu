// K0: gradient features on device -- the offline half of DGGM (calculate_gradient_features,
// reference mask2former/utils/data_process.py:1247-1305, called from dataloader.py:412-421):
//   gx, gy = Sobel3x3(depth) (BORDER_REFLECT_101), mag = sqrt(gx^2+gy^2), zero where depth is
//   invalid (==invalid_value or NaN), vmask = mag>0, norm = (mag-min_valid)/(max_all-min_valid)
//   if max>min else 0.
// Bit-exact with the reference for every input the reference can see (uint8 depth => all
// intermediate values are exact in float32; sqrt/div are IEEE-rounded).  This file is compiled
// with -fmad=false so no multiply-add is contracted.
//
// Two passes over a 1.2 MB image: pass 1 stages a (32+2)x(32+2) halo tile in shared memory,
// writes the un-normalised magnitude and the mask, and reduces the per-image min/max with warp
// shuffles + one atomic per warp; pass 2 normalises in place and replicates the channels.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kT = 32;

struct ImgStats {
    uint32_t min_valid;   // float bits of min over mag>0 (positive floats order as uint32)
    uint32_t max_all;     // float bits of max over all pixels (mag >= 0)
    uint32_t has_nan;
    uint32_t pad;
};

__global__ void gradfeat_init_kernel(ImgStats* st, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        st[i].min_valid = 0x7f800000u;
        st[i].max_all = 0u;
        st[i].has_nan = 0u;
        st[i].pad = 0u;
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

template <typename T>
__global__ void __launch_bounds__(kT* kT / 4) gradfeat_pass1_kernel(const T* __restrict__ depth, long long depth_bs,
                                                                  float* __restrict__ mag_out, long long mag_bs,
                                                                  float* __restrict__ vmask_out, long long vmask_bs,
                                                                  ImgStats* __restrict__ stats, int H, int W,
                                                                  float invalid) {
    __shared__ float tile[kT + 2][kT + 3];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
    const T* d = depth + (long long)b * depth_bs;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = tid; i < (kT + 2) * (kT + 2); i += blockDim.x * blockDim.y) {
        int ty = i / (kT + 2), tx = i - ty * (kT + 2);
        int gy = reflect101(y0 + ty - 1, H), gx = reflect101(x0 + tx - 1, W);
        gy = max(min(gy, H - 1), 0);   // positions past the image are never consumed
        gx = max(min(gx, W - 1), 0);
        tile[ty][tx] = (float)d[(long long)gy * W + gx];
    }
    __syncthreads();
    float lmin = __uint_as_float(0x7f800000u), lmax = 0.f;
    bool lnan = false;
    // each thread handles 4 rows of one column (blockDim = 32 x 8)
    for (int r = threadIdx.y; r < kT; r += blockDim.y) {
        int x = x0 + threadIdx.x, y = y0 + r;
        if (x < W && y < H) {
            int tx = threadIdx.x + 1, ty = r + 1;
            // separable: dx = p[x+1]-p[x-1];  sm = (p[x-1] + 2 p[x]) + p[x+1]
            float dxm = tile[ty - 1][tx + 1] - tile[ty - 1][tx - 1];
            float dx0 = tile[ty][tx + 1] - tile[ty][tx - 1];
            float dxp = tile[ty + 1][tx + 1] - tile[ty + 1][tx - 1];
            float smm = (tile[ty - 1][tx - 1] + 2.0f * tile[ty - 1][tx]) + tile[ty - 1][tx + 1];
            float smp = (tile[ty + 1][tx - 1] + 2.0f * tile[ty + 1][tx]) + tile[ty + 1][tx + 1];
            float gx = (dxm + 2.0f * dx0) + dxp;
            float gy = smp - smm;
            float c = tile[ty][tx];
            bool valid = (c != invalid) && !isnan(c);
            float mag = valid ? __fsqrt_rn(gx * gx + gy * gy) : 0.0f;
            bool pos = mag > 0.0f;
            mag_out[(long long)b * mag_bs + (long long)y * W + x] = mag;
            vmask_out[(long long)b * vmask_bs + (long long)y * W + x] = pos ? 1.0f : 0.0f;
            if (isnan(mag)) lnan = true;
            else {
                lmax = fmaxf(lmax, mag);
                if (pos) lmin = fminf(lmin, mag);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    unsigned any_nan = __ballot_sync(0xffffffffu, lnan);
    if (threadIdx.x == 0) {
        atomicMin(&stats[b].min_valid, __float_as_uint(lmin));
        atomicMax(&stats[b].max_all, __float_as_uint(lmax));
        if (any_nan) atomicOr(&stats[b].has_nan, 1u);
    }
}

__global__ void gradfeat_pass2_kernel(float* __restrict__ norm, long long norm_bs, int n_rep,
                                      const ImgStats* __restrict__ stats, int HW) {
    const int b = blockIdx.y;
    const float mn = __uint_as_float(stats[b].min_valid);
    const float mx = __uint_as_float(stats[b].max_all);
    // np.max propagates NaN -> `max_val > min_val` is False -> zeros; no valid pixel -> zeros
    const bool live = !stats[b].has_nan && (stats[b].min_valid != 0x7f800000u) && (mx > mn);
    const float den = mx - mn;
    float* base = norm + (long long)b * norm_bs;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float v = live ? __fdiv_rn(base[i] - mn, den) : 0.0f;
        for (int r = 0; r < n_rep; ++r) base[(long long)r * HW + i] = v;
    }
}

// uint8 RGB (B,H,W,3) + uint8 depth (B,H,W) -> channels 0:6 of pixel_values (B,10,H,W): the Hugging Face
// processor's rescale (float32(double(x) * rescale_factor)) and normalize ((x - mean) / std, float32) applied to the
// colour image and to the depth image replicated to three channels (DL:389-410).  Bit-exact with the numpy path.
__global__ void __launch_bounds__(256) pack_rgbd_kernel(const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ depth,
                                                        float* __restrict__ pv, long long pv_bs, int HW, double rescale,
                                                        float m0, float m1, float m2, float s0, float s1, float s2) {
    const int b = blockIdx.y;
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    float* out = pv + (long long)b * pv_bs;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const uint8_t* px = rgb + ((long long)b * HW + i) * 3;
        const float d = (float)((double)depth[(long long)b * HW + i] * rescale);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = (float)((double)px[c] * rescale);
            out[(long long)c * HW + i] = __fdiv_rn(v - mean[c], stdv[c]);
            out[(long long)(3 + c) * HW + i] = __fdiv_rn(d - mean[c], stdv[c]);
        }
    }
}

}  // namespace

extern "C" int rgbd_pack_pixel_values(const uint8_t* rgb_hwc, const uint8_t* depth, float* pixel_values,
                                      long long pv_batch_stride, int B, int H, int W, double rescale_factor,
                                      const float* mean3_host, const float* std3_host, float invalid_value, void* workspace,
                                      rgbd_stream_t stream) {
    RGBD_CHECK_ARG(rgb_hwc && depth && pixel_values && mean3_host && std3_host && workspace, "pack_pixel_values: null pointer");
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && pv_batch_stride >= 10ll * H * W, "pack_pixel_values: bad geometry");
    const int HW = H * W;
    dim3 grid(min(ceil_div(HW, 256), 1024), B);
    pack_rgbd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rgb_hwc, depth, pixel_values, pv_batch_stride, HW, rescale_factor,
                                                           mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0],
                                                           std3_host[1], std3_host[2]);
    RGBD_CHECK_LAUNCH();
    // channels 6:9 (normalised Sobel magnitude x3) and 9 (valid-gradient mask): K0 on the raw uint8 depth (DL:412-421)
    return rgbd_gradient_features(depth, RGBD_DTYPE_U8, HW, pixel_values + 6ll * HW, pv_batch_stride, 3,
                                  pixel_values + 9ll * HW, pv_batch_stride, B, H, W, invalid_value, workspace, stream);
}

extern "C" size_t rgbd_gradient_features_workspace_bytes(int B) { return sizeof(ImgStats) * (size_t)(B > 0 ? B : 0); }

extern "C" int rgbd_gradient_features(const void* depth, int depth_dtype, long long depth_batch_stride, float* norm_out,
                                      long long norm_batch_stride, int n_rep, float* vmask_out,
                                      long long vmask_batch_stride, int B, int H, int W, float invalid_value,
                                      void* workspace, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(depth && norm_out && vmask_out && workspace, "gradient_features: null pointer");
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && n_rep >= 1, "gradient_features: bad geometry");
    RGBD_CHECK_ARG(depth_dtype == RGBD_DTYPE_F32 || depth_dtype == RGBD_DTYPE_U8, "gradient_features: depth dtype must be f32 or u8");
    cudaStream_t s = (cudaStream_t)stream;
    ImgStats* st = (ImgStats*)workspace;
    gradfeat_init_kernel<<<ceil_div(B, 128), 128, 0, s>>>(st, B);
    RGBD_CHECK_LAUNCH();
    dim3 grid(ceil_div(W, kT), ceil_div(H, kT), B), block(kT, kT / 4);
    if (depth_dtype == RGBD_DTYPE_F32)
        gradfeat_pass1_kernel<float><<<grid, block, 0, s>>>((const float*)depth, depth_batch_stride, norm_out,
                                                           norm_batch_stride, vmask_out, vmask_batch_stride, st, H, W,
                                                           invalid_value);
    else
        gradfeat_pass1_kernel<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)depth, depth_batch_stride, norm_out,
                                                             norm_batch_stride, vmask_out, vmask_batch_stride, st, H, W,
                                                             invalid_value);
    RGBD_CHECK_LAUNCH();
    int HW = H * W;
    dim3 g2(min(ceil_div(HW, 256), 1024), B);
    gradfeat_pass2_kernel<<<g2, 256, 0, s>>>(norm_out, norm_batch_stride, n_rep, st, HW);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
