// Swin window attention (inference), the inner op of the stock backbone that PRODUCES the hot path's input pyramid
// (reference call site mask2former/utils/custom_model.py:330 `self.encoder(rgb).feature_maps`; transformers
// SwinSelfAttention.forward): per (window, head)
//     out = softmax(q k^T / sqrt(d) + relative_position_bias[head] + shift_mask[window]) v          49 x 49 scores, d = 32
// The stock path runs it as bmm -> div -> add -> add -> softmax -> cast -> bmm -> permute-copy over a (windows*heads, 49, 49)
// score tensor (0.4 GB in fp32 at the first stage of a 32-frame batch): ~9 of the backbone's 35 ms.  Here one warp owns one
// (window, head): K and V are staged in shared memory as fp32, every lane owns one query row (two for rows >= 32) and runs an
// online softmax over 8-key chunks in registers; scores are never written.  58 GFLOP for the whole Swin-T at batch 32 -- CUDA-core
// work, bound by the FMA issue rate, not by memory.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kHd = 32;          // head dimension of every Swin variant (embed_dim / heads)
constexpr int kMaxN = 64;        // tokens per window (7 x 7 = 49; up to 8 x 8)
constexpr int kWarps = 4;

template <typename T> struct Ld;
template <> struct Ld<float> {
    static __device__ __forceinline__ void row(const float* p, float (&x)[kHd]) {
#pragma unroll
        for (int i = 0; i < kHd / 4; ++i) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
    }
    static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void store(float* p, const float (&x)[kHd]) {
#pragma unroll
        for (int i = 0; i < kHd / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    }
};
template <> struct Ld<__nv_bfloat16> {
    static __device__ __forceinline__ void row(const __nv_bfloat16* p, float (&x)[kHd]) {
#pragma unroll
        for (int i = 0; i < kHd / 8; ++i) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i);
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                x[8 * i + 2 * e] = __uint_as_float(u[e] << 16);
                x[8 * i + 2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u);
            }
        }
    }
    static __device__ __forceinline__ float one(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&x)[kHd]) {
#pragma unroll
        for (int i = 0; i < kHd / 8; ++i) {
            uint4 v;
            uint32_t* u = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 t = __floats2bfloat162_rn(x[8 * i + 2 * e], x[8 * i + 2 * e + 1]);
                u[e] = *reinterpret_cast<const uint32_t*>(&t);
            }
            reinterpret_cast<uint4*>(p)[i] = v;
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) window_attention_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                                     const T* __restrict__ v, const float* __restrict__ bias,
                                                                     const float* __restrict__ mask, T* __restrict__ out,
                                                                     long long n_units, int heads, int N, int nW, float sqrt_d) {
    extern __shared__ float s_kv[];                          // per warp: K [N][32], V [N][32] as fp32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * kWarps + warp;
    if (unit >= n_units) return;
    const long long win = unit / heads;                      // b_ = image * nW + window
    const int h = (int)(unit % heads);
    const int C = heads * kHd;
    float* sk = s_kv + (size_t)warp * 2 * N * kHd;
    float* sv = sk + N * kHd;
    const T* kb = k + (size_t)win * N * C + h * kHd;
    const T* vb = v + (size_t)win * N * C + h * kHd;
    for (int i = lane; i < N * kHd; i += 32) {
        const int j = i >> 5, d = i & 31;
        sk[i] = Ld<T>::one(kb + (size_t)j * C + d);
        sv[i] = Ld<T>::one(vb + (size_t)j * C + d);
    }
    __syncwarp();
    const float* bias_h = bias + (size_t)h * N * N;
    const float* mask_w = mask ? mask + (size_t)(win % nW) * N * N : nullptr;
    for (int r = lane; r < N; r += 32) {
        float qv[kHd], o[kHd];
        Ld<T>::row(q + ((size_t)win * N + r) * C + h * kHd, qv);
#pragma unroll
        for (int d = 0; d < kHd; ++d) o[d] = 0.f;
        float m = -INFINITY, l = 0.f;
        const float* brow = bias_h + (size_t)r * N;
        const float* mrow = mask_w ? mask_w + (size_t)r * N : nullptr;
        for (int j0 = 0; j0 < N; j0 += 8) {
            float s[8];
            float cmax = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = j0 + jj;
                s[jj] = -INFINITY;
                if (j < N) {
                    const float4* kr = reinterpret_cast<const float4*>(sk + j * kHd);
                    float acc = 0.f;
#pragma unroll
                    for (int d4 = 0; d4 < kHd / 4; ++d4) {
                        const float4 kk = kr[d4];
                        acc = fmaf(qv[4 * d4], kk.x, acc);
                        acc = fmaf(qv[4 * d4 + 1], kk.y, acc);
                        acc = fmaf(qv[4 * d4 + 2], kk.z, acc);
                        acc = fmaf(qv[4 * d4 + 3], kk.w, acc);
                    }
                    float sc = acc / sqrt_d + __ldg(brow + j);
                    if (mrow) sc += __ldg(mrow + j);
                    s[jj] = sc;
                    cmax = fmaxf(cmax, sc);
                }
            }
            const float m_new = fmaxf(m, cmax);
            const float corr = __expf(m - m_new);              // 0 on the first chunk (m = -inf)
            l *= corr;
#pragma unroll
            for (int d = 0; d < kHd; ++d) o[d] *= corr;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = j0 + jj;
                if (j < N) {
                    const float p = __expf(s[jj] - m_new);
                    l += p;
                    const float4* vr = reinterpret_cast<const float4*>(sv + j * kHd);
#pragma unroll
                    for (int d4 = 0; d4 < kHd / 4; ++d4) {
                        const float4 vv = vr[d4];
                        o[4 * d4] = fmaf(p, vv.x, o[4 * d4]);
                        o[4 * d4 + 1] = fmaf(p, vv.y, o[4 * d4 + 1]);
                        o[4 * d4 + 2] = fmaf(p, vv.z, o[4 * d4 + 2]);
                        o[4 * d4 + 3] = fmaf(p, vv.w, o[4 * d4 + 3]);
                    }
                }
            }
            m = m_new;
        }
        const float inv = 1.0f / l;
#pragma unroll
        for (int d = 0; d < kHd; ++d) o[d] *= inv;
        Ld<T>::store(out + ((size_t)win * N + r) * C + h * kHd, o);
    }
}

}  // namespace

extern "C" int rgbd_window_attention(const void* q, const void* k, const void* v, int dtype, const float* bias, const float* mask,
                                     void* out, long long n_windows, int N, int heads, int head_dim, int n_mask_windows,
                                     rgbd_stream_t stream) {
    RGBD_CHECK_ARG(q && k && v && bias && out, "window_attention: null pointer");
    RGBD_CHECK_ARG(dtype == RGBD_DTYPE_F32 || dtype == RGBD_DTYPE_BF16, "window_attention: q / k / v are f32 or bf16");
    RGBD_CHECK_ARG(head_dim == kHd, "window_attention: head dimension must be %d (got %d)", kHd, head_dim);
    RGBD_CHECK_ARG(N >= 1 && N <= kMaxN, "window_attention: 1..%d tokens per window (got %d)", kMaxN, N);
    RGBD_CHECK_ARG(n_windows >= 1 && heads >= 1, "window_attention: bad sizes");
    RGBD_CHECK_ARG(!mask || (n_mask_windows >= 1 && n_windows % n_mask_windows == 0),
                   "window_attention: the window count must be a multiple of the mask's window count");
    const long long n_units = n_windows * heads;
    const long long blocks = (n_units + kWarps - 1) / kWarps;
    RGBD_CHECK_ARG(blocks <= 0x7fffffffLL, "window_attention: too many windows");
    const size_t smem = (size_t)kWarps * 2 * N * kHd * sizeof(float);
    RgbdDeviceInfo di;
    if (int rc = rgbd_device_info(&di)) return rc;
    RGBD_ONCE_PER_DEVICE(di.device, {
        RGBD_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kWarps * 2 * kMaxN * kHd * (int)sizeof(float)));
        RGBD_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kWarps * 2 * kMaxN * kHd * (int)sizeof(float)));
    });
    cudaStream_t s = (cudaStream_t)stream;
    const float sqrt_d = sqrtf((float)head_dim);
    if (dtype == RGBD_DTYPE_BF16)
        window_attention_kernel<__nv_bfloat16><<<(unsigned)blocks, kWarps * 32, smem, s>>>(
            (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, bias, mask, (__nv_bfloat16*)out, n_units, heads, N,
            mask ? n_mask_windows : 1, sqrt_d);
    else
        window_attention_kernel<float><<<(unsigned)blocks, kWarps * 32, smem, s>>>((const float*)q, (const float*)k, (const float*)v, bias,
                                                                                  mask, (float*)out, n_units, heads, N,
                                                                                  mask ? n_mask_windows : 1, sqrt_d);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
