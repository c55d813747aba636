// K6: the feature-based window-ratio predictor of the version 0.1.3 / 0.3.0 models (reference
// mask2former/utils/custom_model.py:823-898, `RatioPredictor.forward`): global average pool of every depth-encoder
// feature map, concatenation, Linear sum(C_i)->64 -> ReLU -> Linear 64->32 -> ReLU -> Linear 32->1,
// ratio = output_min + (output_max - output_min) * sigmoid(raw).
//   ratio_feat_gap_kernel   one warp per (image, channel) plane, 128-bit loads when the plane allows it (HBM-bound:
//                           the feature pyramid is read exactly once)
//   ratio_feat_mlp_kernel   one CTA per image, fp32 CUDA cores (0.1 MFLOP)
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kMaxLevels = 8;

struct GapParams {
    const float* feat[kMaxLevels];
    int C[kMaxLevels], HW[kMaxLevels], c_off[kMaxLevels + 1];
    int n_levels, B, c_total;
};

__global__ void __launch_bounds__(256) ratio_feat_gap_kernel(const __grid_constant__ GapParams p, float* __restrict__ pooled) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + warp, b = blockIdx.y;
    if (c >= p.c_total) return;
    int l = 0;
    while (l + 1 < p.n_levels && c >= p.c_off[l + 1]) ++l;
    const int hw = p.HW[l];
    const float* plane = p.feat[l] + ((size_t)b * p.C[l] + (c - p.c_off[l])) * hw;
    float s = 0.f;
    if ((hw & 3) == 0 && (reinterpret_cast<uintptr_t>(plane) & 15) == 0) {
        const float4* p4 = reinterpret_cast<const float4*>(plane);
        for (int i = lane; i < hw / 4; i += 32) {
            const float4 v = __ldg(p4 + i);
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (int i = lane; i < hw; i += 32) s += __ldg(plane + i);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (lane == 0) pooled[(size_t)b * p.c_total + c] = s / (float)hw;
}

__device__ __forceinline__ void fc(const float* __restrict__ w, const float* __restrict__ bias, const float* in, float* out,
                                   int n_in, int n_out, bool relu) {
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

__global__ void __launch_bounds__(256) ratio_feat_mlp_kernel(const float* __restrict__ pooled, int c_total, const float* w0,
                                                             const float* b0, const float* w1, const float* b1, const float* w2,
                                                             const float* b2, float out_min, float out_span,
                                                             float* __restrict__ ratio) {
    extern __shared__ float x[];                 // c_total pooled features
    __shared__ float h0[64], h1[32], raw[1];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < c_total; i += blockDim.x) x[i] = pooled[(size_t)b * c_total + i];
    __syncthreads();
    fc(w0, b0, x, h0, c_total, 64, true);
    fc(w1, b1, h0, h1, 64, 32, true);
    fc(w2, b2, h1, raw, 32, 1, false);
    if (threadIdx.x == 0) {
        const float sg = 1.0f / (1.0f + expf(-raw[0]));
        ratio[b] = __fadd_rn(out_min, __fmul_rn(out_span, sg));
    }
}

}  // namespace

extern "C" int rgbd_ratio_from_features(int n_levels, const float* const* feats_host, const int* C_host, const int* HW_host, int B,
                                        const float* const* fc_w_host, const float* const* fc_b_host, float out_min,
                                        float out_max, float* pooled_ws, float* ratio_out, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(feats_host && C_host && HW_host && fc_w_host && fc_b_host && pooled_ws && ratio_out, "ratio_from_features: null pointer");
    RGBD_CHECK_ARG(n_levels >= 1 && n_levels <= kMaxLevels && B >= 1, "ratio_from_features: 1..%d levels", kMaxLevels);
    GapParams p;
    p.n_levels = n_levels; p.B = B;
    int off = 0;
    for (int l = 0; l < n_levels; ++l) {
        RGBD_CHECK_ARG(feats_host[l] && C_host[l] >= 1 && HW_host[l] >= 1, "ratio_from_features: bad level %d", l);
        p.feat[l] = feats_host[l]; p.C[l] = C_host[l]; p.HW[l] = HW_host[l]; p.c_off[l] = off;
        off += C_host[l];
    }
    p.c_off[n_levels] = off;
    p.c_total = off;
    RGBD_CHECK_ARG(off * 4 <= 40 * 1024, "ratio_from_features: at most 10240 pooled features");
    for (int i = 0; i < 3; ++i) RGBD_CHECK_ARG(fc_w_host[i] && fc_b_host[i], "ratio_from_features: null fc layer %d", i);
    cudaStream_t s = (cudaStream_t)stream;
    ratio_feat_gap_kernel<<<dim3(ceil_div(off, 8), B), 256, 0, s>>>(p, pooled_ws);
    RGBD_CHECK_LAUNCH();
    ratio_feat_mlp_kernel<<<B, 256, (size_t)off * 4, s>>>(pooled_ws, off, fc_w_host[0], fc_b_host[0], fc_w_host[1], fc_b_host[1],
                                                         fc_w_host[2], fc_b_host[2], out_min, out_max - out_min, ratio_out);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
