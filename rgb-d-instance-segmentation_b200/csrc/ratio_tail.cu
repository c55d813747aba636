// K4 tail: the small end of EnhancedDepthImageRatioPredictor.forward (reference
// mask2former/utils/custom_model.py:1473-1485) after the fused 3x3 conv + BN + ReLU + AdaptiveAvgPool2d(4):
//   pooled sums (64-bit fixed point, see rgbd_conv_gemm epi_mode 2) / cell size -> Conv3x3(256->512, pad 1) on the 4x4 map -> BN (folded) -> ReLU -> global
//   average pool -> Linear 512->128->64->32->1 with ReLU (Dropout is identity in eval) ->
//   ratio = 0.01 + 0.49 * sigmoid(raw).
// 19 MFLOP per image: fp32 CUDA cores, two launches (conv: CTAs of 16 output channels x 4 images so the weights are read
// once per 4 images; then one CTA per image for the MLP).
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kCin = 256, kCout = 512;
constexpr int kOcPerCta = 16, kImgPerCta = 4, kIcChunk = 32;
constexpr int kXImgStride = kIcChunk * 36 + 8;     // +8 floats: the 4 images of a warp land in different banks

// CTA = 16 output channels x 4 images: the 16 x 2304 weights (147 KB) are read once per 4 images, coalesced, through a
// shared-memory chunk of 32 input channels; thread = (oc, image, output row) computes the 4 outputs of its row.
// RAW (train-mode BatchNorm, CM:1418-1420 under .train()): the un-normalised conv outputs go to raw[b][oc][16]; the batch
// statistics need every image before anything can be normalised.
template <bool RAW>
__global__ void __launch_bounds__(256) ratio_tail_conv_kernel(const long long* __restrict__ pool, int pool_stride,
                                                              float inv_cell, const float* __restrict__ w,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, float* __restrict__ gap, int B) {
    __shared__ float xs[kImgPerCta * kXImgStride];          // [img][ic][6][6] zero-padded 4x4 maps
    __shared__ float ws[kOcPerCta * kIcChunk * 9];          // [oc][ic][3][3]
    const int oc0 = blockIdx.x * kOcPerCta, b0 = blockIdx.y * kImgPerCta;
    const int oy = threadIdx.x & 3, im = (threadIdx.x >> 2) & 3, ol = threadIdx.x >> 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ic0 = 0; ic0 < kCin; ic0 += kIcChunk) {
        __syncthreads();
        for (int i = threadIdx.x; i < kImgPerCta * kIcChunk * 36; i += blockDim.x) {
            const int c = i % 6, r = (i / 6) % 6, ic = (i / 36) % kIcChunk, bi = i / (36 * kIcChunk);
            float v = 0.f;
            if (r >= 1 && r <= 4 && c >= 1 && c <= 4 && b0 + bi < B)
                v = (float)((double)pool[((size_t)(b0 + bi) * 16 + (r - 1) * 4 + (c - 1)) * pool_stride + ic0 + ic] *
                            (1.0 / RGBD_POOL_FIXED_ONE)) * inv_cell;
            xs[bi * kXImgStride + ic * 36 + r * 6 + c] = v;
        }
        for (int i = threadIdx.x; i < kOcPerCta * kIcChunk * 9; i += blockDim.x) {
            const int o = i / (kIcChunk * 9), r = i % (kIcChunk * 9);
            ws[i] = __ldg(w + ((size_t)(oc0 + o) * kCin + ic0) * 9 + r);
        }
        __syncthreads();
        const float* xr = xs + im * kXImgStride + oy * 6;
        const float* wr = ws + ol * kIcChunk * 9;
#pragma unroll 4
        for (int ic = 0; ic < kIcChunk; ++ic) {
            float wk[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) wk[k] = wr[ic * 9 + k];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                float row[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) row[c] = xr[ic * 36 + ky * 6 + c];
#pragma unroll
                for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc[ox] = fmaf(wk[ky * 3 + kx], row[ox + kx], acc[ox]);
            }
        }
    }
    const int oc = oc0 + ol;
    if (RAW) {
        if (b0 + im < B) {
            float4* dst = reinterpret_cast<float4*>(gap + ((size_t)(b0 + im) * kCout + oc) * 16 + oy * 4);
            *dst = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        return;
    }
    const float sc = scale[oc], sh = shift[oc];
    float s = 0.f;
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) s += fmaxf(fmaf(acc[ox], sc, sh), 0.f);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (oy == 0 && b0 + im < B) gap[(size_t)(b0 + im) * kCout + oc] = s * (1.0f / 16.0f);
}

__device__ __forceinline__ void fc_layer(const float* __restrict__ w, const float* __restrict__ bias, const float* in,
                                         float* out, int n_in, int n_out, bool relu) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int o = warp; o < n_out; o += nw) {
        float s = 0.f;
        for (int i = lane; i < n_in; i += 32) s = fmaf(__ldg(w + (size_t)o * n_in + i), in[i], s);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
        if (lane == 0) {
            s += bias[o];
            out[o] = relu ? fmaxf(s, 0.f) : s;
        }
    }
    __syncthreads();
}

// Tensor-core variant of the tail conv (rgbd_ratio_tail_prepare + rgbd_conv_gemm + rgbd_ratio_tail_mlp_fx): the pooled 4x4 map
// as a bf16 channels-last operand, and the zeroed fixed-point GAP accumulator the GEMM's pooled epilogue adds into.
__global__ void __launch_bounds__(256) ratio_tail_prepare_kernel(const long long* __restrict__ pool, int pool_stride, float inv_cell,
                                                                 __nv_bfloat16* __restrict__ a, long long* __restrict__ gap_fx, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // (b, cell, c)
    if (i < B * 16 * kCin) {
        const int c = i % kCin, cell = (i / kCin) % 16, b = i / (kCin * 16);
        const float v = (float)((double)pool[((size_t)b * 16 + cell) * pool_stride + c] * (1.0 / RGBD_POOL_FIXED_ONE)) * inv_cell;
        a[i] = __float2bfloat16(v);
    }
    if (i < B * kCout) gap_fx[i] = 0ll;
}

// GAP_FX: `gap` holds the fixed-point sums over the 16 pixels (conv_gemm epilogue mode 2 with a 1x1 cell grid)
template <bool GAP_FX>
__global__ void __launch_bounds__(512) ratio_tail_mlp_kernel(const float* __restrict__ gap, const float* w0, const float* b0,
                                                             const float* w1, const float* b1, const float* w2,
                                                             const float* b2, const float* w3, const float* b3,
                                                             float out_min, float out_span, float* __restrict__ ratio,
                                                             const float* __restrict__ drop0, const float* __restrict__ drop1) {
    __shared__ float a[kCout], h0[128], h1[64], h2[32], raw[1];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < kCout; i += blockDim.x) {
        if (GAP_FX)
            a[i] = (float)((double)reinterpret_cast<const long long*>(gap)[(size_t)b * kCout + i] * (1.0 / RGBD_POOL_FIXED_ONE)) * (1.0f / 16.0f);
        else
            a[i] = gap[(size_t)b * kCout + i];
    }
    __syncthreads();
    fc_layer(w0, b0, a, h0, 512, 128, true);
    if (drop0) {        // nn.Dropout(0.3) in train mode (CM:1430): y = x * keep / (1 - p), the multiplier comes from the caller
        for (int i = threadIdx.x; i < 128; i += blockDim.x) h0[i] *= drop0[(size_t)b * 128 + i];
        __syncthreads();
    }
    fc_layer(w1, b1, h0, h1, 128, 64, true);
    if (drop1) {        // nn.Dropout(0.2) (CM:1433)
        for (int i = threadIdx.x; i < 64; i += blockDim.x) h1[i] *= drop1[(size_t)b * 64 + i];
        __syncthreads();
    }
    fc_layer(w2, b2, h1, h2, 64, 32, true);
    fc_layer(w3, b3, h2, raw, 32, 1, false);
    if (threadIdx.x == 0) {
        const float sg = 1.0f / (1.0f + expf(-raw[0]));
        ratio[b] = __fadd_rn(out_min, __fmul_rn(out_span, sg));
    }
}

// Train-mode BatchNorm2d(512) over the (B, 4, 4) samples of every channel (CM:1419 under .train()), then ReLU + GAP.
// One warp per channel: batch mean / biased variance of the raw conv outputs (the conv bias cancels in the normalised
// value; it only enters the running mean), running statistics updated in place with torch's rule (momentum * unbiased).
__global__ void __launch_bounds__(256) ratio_tail_bn_train_kernel(const float* __restrict__ raw, const float* __restrict__ conv_bias,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float eps, float momentum, float* __restrict__ running_mean,
                                                                  float* __restrict__ running_var, float* __restrict__ gap, int B) {
    const int oc = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n = B * 16;
    double s = 0.0, q = 0.0;
    for (int i = lane; i < n; i += 32) {
        const float v = raw[((size_t)(i >> 4) * kCout + oc) * 16 + (i & 15)];
        s += v;
        q += (double)v * v;
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, k);
        q += __shfl_xor_sync(0xffffffffu, q, k);
    }
    const double mean = s / n;
    double var = q / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float sc = gamma[oc] * (float)(1.0 / sqrt(var + (double)eps));
    const float sh = beta[oc] - (float)mean * sc;
    if (lane == 0 && running_mean) {
        running_mean[oc] = (1.f - momentum) * running_mean[oc] + momentum * ((float)mean + conv_bias[oc]);
        running_var[oc] = (1.f - momentum) * running_var[oc] + momentum * (float)(n > 1 ? var * n / (n - 1) : var);
    }
    for (int b = 0; b < B; ++b) {
        float v = lane < 16 ? fmaxf(fmaf(raw[((size_t)b * kCout + oc) * 16 + lane], sc, sh), 0.f) : 0.f;
#pragma unroll
        for (int k = 8; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
        if (lane == 0) gap[(size_t)b * kCout + oc] = v * (1.0f / 16.0f);
    }
}

// AdaptiveAvgPool2d(4) of a bf16 channels-last map (B,H,W,256) for sizes the fused pooled epilogue cannot take (H or W not
// divisible by 4: torch's windows [floor(i*H/4), ceil((i+1)*H/4)) then overlap and differ in size).  CTA = (cell, image,
// row chunk), thread = channel; every contribution is scaled by 1/window size so the result, read with cell_pixels = 1, is
// the window MEAN in the same 64-bit fixed point as the fused epilogue (order-independent integer atomics).
__global__ void __launch_bounds__(256) adaptive_pool4_kernel(const __nv_bfloat16* __restrict__ x, long long* __restrict__ pool,
                                                             int H, int W) {
    const int cell = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
    const int i = cell >> 2, j = cell & 3;
    const int y0 = (i * H) / 4, y1 = ((i + 1) * H + 3) / 4, x0 = (j * W) / 4, x1 = ((j + 1) * W + 3) / 4;
    const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
    float s = 0.f;
    for (int y = y0 + blockIdx.z; y < y1; y += gridDim.z) {
        const __nv_bfloat16* row = x + (((size_t)b * H + y) * W) * kCin + c;
        float r = 0.f;
        for (int xx = x0; xx < x1; ++xx) r += __bfloat162float(row[(size_t)xx * kCin]);
        s += r;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(pool + ((size_t)b * 16 + cell) * kCin + c),
              (unsigned long long)__double2ll_rn((double)(s * inv) * RGBD_POOL_FIXED_ONE));
}

}  // namespace

extern "C" int rgbd_adaptive_avg_pool4(const void* x_bf16, long long* pool_means, int B, int H, int W, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(x_bf16 && pool_means && B >= 1 && H >= 1 && W >= 1, "adaptive_avg_pool4: bad arguments");
    RGBD_CHECK_CUDA(cudaMemsetAsync(pool_means, 0, (size_t)B * 16 * kCin * sizeof(long long), (cudaStream_t)stream));
    const int chunks = H >= 64 ? 8 : 1;
    adaptive_pool4_kernel<<<dim3(16, B, chunks), kCin, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x_bf16),
                                                                                pool_means, H, W);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_ratio_tail_train(const long long* pool_sums, int pool_stride, int cell_pixels, const float* conv_w,
                                     const float* conv_bias, const float* bn_gamma, const float* bn_beta, float eps,
                                     float momentum, float* running_mean, float* running_var, const float* const* fc_w,
                                     const float* const* fc_b, const float* drop0, const float* drop1, float out_min,
                                     float out_max, float* raw_ws, float* gap_ws, float* ratio_out, int B,
                                     rgbd_stream_t stream) {
    RGBD_CHECK_ARG(pool_sums && conv_w && conv_bias && bn_gamma && bn_beta && fc_w && fc_b && raw_ws && gap_ws && ratio_out,
                   "ratio_tail_train: null pointer");
    RGBD_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "ratio_tail_train: pass both running statistics or neither");
    RGBD_CHECK_ARG(B >= 1 && cell_pixels >= 1 && pool_stride >= kCin, "ratio_tail_train: bad geometry");
    for (int i = 0; i < 4; ++i) RGBD_CHECK_ARG(fc_w[i] && fc_b[i], "ratio_tail_train: null fc layer %d", i);
    cudaStream_t s = (cudaStream_t)stream;
    ratio_tail_conv_kernel<true><<<dim3(kCout / kOcPerCta, ceil_div(B, kImgPerCta)), 256, 0, s>>>(
        pool_sums, pool_stride, 1.0f / (float)cell_pixels, conv_w, nullptr, nullptr, raw_ws, B);
    RGBD_CHECK_LAUNCH();
    ratio_tail_bn_train_kernel<<<kCout / 8, 256, 0, s>>>(raw_ws, conv_bias, bn_gamma, bn_beta, eps, momentum, running_mean,
                                                         running_var, gap_ws, B);
    RGBD_CHECK_LAUNCH();
    ratio_tail_mlp_kernel<false><<<B, 512, 0, s>>>(gap_ws, fc_w[0], fc_b[0], fc_w[1], fc_b[1], fc_w[2], fc_b[2], fc_w[3], fc_b[3],
                                            out_min, out_max - out_min, ratio_out, drop0, drop1);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_ratio_tail(const long long* pool_sums, int pool_stride, int cell_pixels, const float* conv_w,
                               const float* conv_scale, const float* conv_shift, const float* const* fc_w,
                               const float* const* fc_b, float out_min, float out_max, float* gap_ws, float* ratio_out,
                               int B, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(pool_sums && conv_w && conv_scale && conv_shift && fc_w && fc_b && gap_ws && ratio_out,
                   "ratio_tail: null pointer");
    RGBD_CHECK_ARG(B >= 1 && cell_pixels >= 1 && pool_stride >= kCin, "ratio_tail: bad geometry");
    for (int i = 0; i < 4; ++i) RGBD_CHECK_ARG(fc_w[i] && fc_b[i], "ratio_tail: null fc layer %d", i);
    cudaStream_t s = (cudaStream_t)stream;
    ratio_tail_conv_kernel<false><<<dim3(kCout / kOcPerCta, ceil_div(B, kImgPerCta)), 256, 0, s>>>(
        pool_sums, pool_stride, 1.0f / (float)cell_pixels, conv_w, conv_scale, conv_shift, gap_ws, B);
    RGBD_CHECK_LAUNCH();
    ratio_tail_mlp_kernel<false><<<B, 512, 0, s>>>(gap_ws, fc_w[0], fc_b[0], fc_w[1], fc_b[1], fc_w[2], fc_b[2], fc_w[3], fc_b[3],
                                            out_min, out_max - out_min, ratio_out, nullptr, nullptr);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_ratio_tail_prepare(const long long* pool_sums, int pool_stride, int cell_pixels, void* a_bf16, long long* gap_fx,
                                       int B, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(pool_sums && a_bf16 && gap_fx, "ratio_tail_prepare: null pointer");
    RGBD_CHECK_ARG(B >= 1 && cell_pixels >= 1 && pool_stride >= kCin, "ratio_tail_prepare: bad geometry");
    ratio_tail_prepare_kernel<<<ceil_div(B * 16 * kCin, 256), 256, 0, (cudaStream_t)stream>>>(
        pool_sums, pool_stride, 1.0f / (float)cell_pixels, reinterpret_cast<__nv_bfloat16*>(a_bf16), gap_fx, B);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_ratio_tail_mlp_fx(const long long* gap_fx, const float* const* fc_w, const float* const* fc_b, float out_min,
                                      float out_max, float* ratio_out, int B, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(gap_fx && fc_w && fc_b && ratio_out && B >= 1, "ratio_tail_mlp_fx: bad arguments");
    for (int i = 0; i < 4; ++i) RGBD_CHECK_ARG(fc_w[i] && fc_b[i], "ratio_tail_mlp_fx: null fc layer %d", i);
    ratio_tail_mlp_kernel<true><<<B, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(gap_fx), fc_w[0], fc_b[0], fc_w[1],
                                                                     fc_b[1], fc_w[2], fc_b[2], fc_w[3], fc_b[3], out_min,
                                                                     out_max - out_min, ratio_out, nullptr, nullptr);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
