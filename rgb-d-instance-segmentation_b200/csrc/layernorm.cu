// LayerNorm over the last dimension with the output dtype chosen by the caller (inference).  Used behind the PRE-norm LayerNorms
// of the stock Swin encoder (transformers SwinLayer.layernorm_before / layernorm_after; the backbone the reference calls at
// mask2former/utils/custom_model.py:330): under bf16 autocast torch computes them in float32, writes float32, and every consumer
// -- pad / roll / window partition, then nn.Linear -- first moves and finally casts that tensor to bf16 (14 bytes per element for
// LayerNorm + one cast, three casts for the q / k / v projections).  Writing the normalised row as bf16 directly gives the
// Linear layers bit-for-bit the operand autocast would have produced (same float32 arithmetic, one rounding) at 6 bytes per
// element, and halves the traffic of the data movement in between.  HBM-bound: one pass, the row lives in registers.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

// G lanes per row (8, 16 or 32), up to 8 float4 chunks per lane (C <= 1024, C % 4 == 0)
template <bool XBF16>
__global__ void __launch_bounds__(256) layer_norm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, void* __restrict__ out, long long rows, int C,
                                                         int G, float eps, int out_bf16) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & (G - 1);                            // lane within the row's group
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool live = row < rows;
    const long long rr = live ? row : rows - 1;               // keep whole warps in the shuffles
    const int n4 = C >> 2;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c4 = sub + i * G;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < n4) {
            if (XBF16) {                                       // 4 bf16 = 8 bytes; widened exactly, like autocast's cast to float32
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + rr * C) + c4);
                v[i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                                   __uint_as_float(u.y & 0xffff0000u));
            } else {
                v[i] = ld_stream_f4(reinterpret_cast<const float*>(x) + rr * C + 4 * c4);
            }
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, G);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (sub + i * G < n4) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o, G);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (!live) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c4 = sub + i * G;
        if (c4 < n4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
            const float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
            const float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
            if (out_bf16) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1), hi = __floats2bfloat162_rn(y2, y3);
                uint2 w;
                w.x = *reinterpret_cast<const uint32_t*>(&lo);
                w.y = *reinterpret_cast<const uint32_t*>(&hi);
                reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * C)[c4] = w;
            } else {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * C)[c4] = make_float4(y0, y1, y2, y3);
            }
        }
    }
}

}  // namespace

extern "C" int rgbd_layer_norm(const void* x, int x_dtype, const float* gamma, const float* beta, void* out, int out_dtype,
                               long long rows, int C, float eps, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(x && gamma && beta && out, "layer_norm: null pointer");
    RGBD_CHECK_ARG(out_dtype == RGBD_DTYPE_F32 || out_dtype == RGBD_DTYPE_BF16, "layer_norm: output is f32 or bf16");
    RGBD_CHECK_ARG(x_dtype == RGBD_DTYPE_F32 || x_dtype == RGBD_DTYPE_BF16, "layer_norm: input is f32 or bf16");
    RGBD_CHECK_ARG(rows >= 1 && C >= 4 && C <= 1024 && C % 4 == 0, "layer_norm: C must be a multiple of 4 in 4..1024 (got %d)", C);
    int G = 8;
    while (G < 32 && G * 4 < (C >> 2)) G <<= 1;               // ~3-4 chunks per lane; 8 at most (C = 1024, G = 32)
    const long long threads = rows * G;
    const long long blocks = (threads + 255) / 256;
    RGBD_CHECK_ARG(blocks <= 0x7fffffffLL, "layer_norm: too many rows");
    if (x_dtype == RGBD_DTYPE_BF16)
        layer_norm_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, out, rows, C, G, eps,
                                                                                 out_dtype == RGBD_DTYPE_BF16);
    else
        layer_norm_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, out, rows, C, G, eps,
                                                                                  out_dtype == RGBD_DTYPE_BF16);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
