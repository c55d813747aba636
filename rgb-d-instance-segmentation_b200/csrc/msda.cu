// Multi-scale deformable attention sampling, forward (inference): the consumer side of the hot path.  The fused pyramid
// goes straight into Hugging Face's Mask2FormerPixelDecoder (reference call site mask2former/utils/custom_model.py:383),
// whose six encoder layers spend 80 ms per 32 frames in `multi_scale_deformable_attention` (grid_sample per level + stack +
// multiply + sum: ~10 feature-sized fp32 temporaries of 0.8 GB each).  This kernel computes the same sum in one pass:
//   out[b,q,h,:] = sum_{l,p} softmax_lp(logit[b,q,h,l,p]) * bilinear(value_l[b,:,h,:], loc[b,q,h,l,p])
// with grid_sample(align_corners=False, padding_mode="zeros") arithmetic, and optionally the two element-wise steps in front
// of it (softmax over the L*P logits, loc = reference_point + offset / (W_l, H_l)).  HBM/L2-bound gather: one lane owns 8
// channels of one (query, head) (a 16-byte bf16 load per bilinear corner, the 4 lanes of a head cover its 32 channels =
// one 64-byte segment), all corners of a level's points are in flight before the first FMA.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kMaxLevels = 8;

struct MsdaParams {
    const void* value;
    const void* offs;
    const float* ref;
    const void* attn;
    void* out;
    long long n_items;          // B * Q * H
    int S, Q, H, D, L, P, lph;  // lph: lanes per head = D / 8
    int offs_bf16, attn_bf16, out_bf16, softmax;
    int lvl_h[kMaxLevels], lvl_w[kMaxLevels], lvl_start[kMaxLevels];
};

__device__ __forceinline__ float ld_scalar(const void* p, long long i, int bf16) {
    return bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}

template <bool VBF16>
__device__ __forceinline__ void load8(const void* base, long long elem, bool ok, uint4& a, uint4& b) {
    a = make_uint4(0u, 0u, 0u, 0u);
    b = a;
    if (!ok) return;
    if (VBF16) {
        a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem));
    } else {
        const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + elem);
        a = __ldg(q);
        b = __ldg(q + 1);
    }
}

template <bool VBF16>
__device__ __forceinline__ void fma8(float (&acc)[8], float w, const uint4& a, const uint4& b) {
    if (VBF16) {
        const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[2 * i] = fmaf(w, __uint_as_float(u[i] << 16), acc[2 * i]);
            acc[2 * i + 1] = fmaf(w, __uint_as_float(u[i] & 0xffff0000u), acc[2 * i + 1]);
        }
    } else {
        const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, __uint_as_float(u[i]), acc[i]);
    }
}

// PT > 0: points per level known at compile time (loads of a whole level batched); PT == 0: runtime p.P, one point at a time
template <bool VBF16, int PT>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const MsdaParams p) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long item = g / p.lph;
    if (item >= p.n_items) return;
    const int c0 = (int)(g % p.lph) * 8;
    const int h = (int)(item % p.H);
    const long long bq = item / p.H;
    const long long b = bq / p.Q;
    const int P = PT > 0 ? PT : p.P;
    const int LP = p.L * P;

    // softmax statistics over the L*P logits of this (query, head) (torch: exp(x - max) / sum, float32)
    float mx = 0.f, inv_sum = 1.f;
    const long long a_base = item * LP;
    if (p.softmax) {
        mx = -INFINITY;
        for (int i = 0; i < LP; ++i) mx = fmaxf(mx, ld_scalar(p.attn, a_base + i, p.attn_bf16));
        float s = 0.f;
        for (int i = 0; i < LP; ++i) s += expf(ld_scalar(p.attn, a_base + i, p.attn_bf16) - mx);
        inv_sum = 1.0f / s;
    }

    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;

    const long long v_img = b * p.S;                       // first value row of image b
    const long long o_base = item * LP * 2;

    for (int l = 0; l < p.L; ++l) {
        const int Hl = p.lvl_h[l], Wl = p.lvl_w[l];
        const float Wf = (float)Wl, Hf = (float)Hl;
        float rx = 0.f, ry = 0.f;
        if (p.ref) {
            rx = p.ref[(bq * p.L + l) * 2];
            ry = p.ref[(bq * p.L + l) * 2 + 1];
        }
        const long long v_lvl = v_img + p.lvl_start[l];
        constexpr int U = PT > 0 ? (VBF16 ? PT : PT / 2) : 1;      // corner loads in flight: 16 (bf16: 16 B each) / 8 (fp32: 32 B)
        for (int pt0 = 0; pt0 < P; pt0 += U) {
            float wgt[U][4];
            uint4 va[U][4], vb[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int lp = l * P + pt0 + u;
                float x = ld_scalar(p.offs, o_base + 2 * lp, p.offs_bf16);
                float y = ld_scalar(p.offs, o_base + 2 * lp + 1, p.offs_bf16);
                if (p.ref) {
                    // sampling_offsets / offset_normalizer + reference_points: the quotient has the dtype of the offsets
                    x = x / Wf;
                    y = y / Hf;
                    if (p.offs_bf16) {
                        x = __bfloat162float(__float2bfloat16_rn(x));
                        y = __bfloat162float(__float2bfloat16_rn(y));
                    }
                    x = rx + x;
                    y = ry + y;
                }
                float aw = ld_scalar(p.attn, a_base + lp, p.attn_bf16);
                if (p.softmax) aw = expf(aw - mx) * inv_sum;
                // grid = 2 * loc - 1; grid_sample un-normalises with ((grid + 1) * size - 1) / 2 (align_corners=False)
                const float gx = 2.0f * x - 1.0f, gy = 2.0f * y - 1.0f;
                const float ix = ((gx + 1.0f) * Wf - 1.0f) * 0.5f;
                const float iy = ((gy + 1.0f) * Hf - 1.0f) * 0.5f;
                const float fx = floorf(ix), fy = floorf(iy);
                // ATen grid_sampler: nw = (x_se - ix)(y_se - iy), ne = (ix - x_sw)(y_sw - iy), sw = (x_ne - ix)(iy - y_ne), se = ...
                const float ax = ix - fx, ay = iy - fy, bx = (fx + 1.0f) - ix, by = (fy + 1.0f) - iy;
                // clamp before the int conversion: wildly out-of-range (or NaN) locations sample nothing, like zeros padding
                const int x0 = (int)fminf(fmaxf(fx, -2.0f), (float)Wl + 1.0f), y0 = (int)fminf(fmaxf(fy, -2.0f), (float)Hl + 1.0f);
                const bool fin = (ix == ix) && (iy == iy);
                const bool x0ok = fin && x0 >= 0 && x0 < Wl, x1ok = fin && x0 + 1 >= 0 && x0 + 1 < Wl;
                const bool y0ok = y0 >= 0 && y0 < Hl, y1ok = y0 + 1 >= 0 && y0 + 1 < Hl;
                wgt[u][0] = aw * (bx * by);
                wgt[u][1] = aw * (ax * by);
                wgt[u][2] = aw * (bx * ay);
                wgt[u][3] = aw * (ax * ay);
                const long long r00 = ((v_lvl + (long long)y0 * Wl + x0) * p.H + h) * p.D + c0;
                const long long dx = (long long)p.H * p.D, dy = dx * Wl;
                load8<VBF16>(p.value, r00, x0ok && y0ok, va[u][0], vb[u][0]);
                load8<VBF16>(p.value, r00 + dx, x1ok && y0ok, va[u][1], vb[u][1]);
                load8<VBF16>(p.value, r00 + dy, x0ok && y1ok, va[u][2], vb[u][2]);
                load8<VBF16>(p.value, r00 + dy + dx, x1ok && y1ok, va[u][3], vb[u][3]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int c = 0; c < 4; ++c) fma8<VBF16>(acc, wgt[u][c], va[u][c], vb[u][c]);
        }
    }

    const long long o = item * p.D + c0;
    if (p.out_bf16) {
        uint4 r;
        uint32_t* rr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
            rr[i] = *reinterpret_cast<const uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = r;
    } else {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// Fast path: D = 32 (4 lanes x 8 channels per head), 4 points per level, <= 4 levels, offsets and logits of one dtype.
// Lane j of a head's quad owns point j of EVERY level: it loads that point's offset pair and logit, takes part in the
// softmax through two quad shuffles, computes the bilinear corner set once and hands (base index, validity bits, four
// weights) to the other three lanes by shuffle -- the coordinate arithmetic is done once per point instead of once per lane,
// and all index arithmetic inside an image is 32-bit.  (The generic kernel below spent 3.4 k instructions per thread and was
// issue-bound at 0.93 ms per launch, batch 32; ncu: profiles/r02_ncu_msda.txt.)
template <bool VBF16, bool SBF16, int LT>
__global__ void __launch_bounds__(256) msda_fwd_quad_kernel(const MsdaParams p) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long item = g >> 2;
    const bool live = item < p.n_items;
    if (!live) item = p.n_items - 1;                       // keep the whole warp in the shuffles; the store is predicated
    const int j = (int)(g & 3);
    const int c0 = j * 8;
    const int h = (int)(item % p.H);
    const long long bq = item / p.H;
    const long long b = bq / p.Q;
    constexpr int LP = LT * 4;
    const unsigned FULL = 0xffffffffu;

    float aw[LT], ox[LT], oy[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) {
        const long long i = item * LP + l * 4 + j;
        if (SBF16) {
            aw[l] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.attn)[i]);
            const __nv_bfloat162 o = reinterpret_cast<const __nv_bfloat162*>(p.offs)[i];
            ox[l] = __low2float(o);
            oy[l] = __high2float(o);
        } else {
            aw[l] = reinterpret_cast<const float*>(p.attn)[i];
            const float2 o = reinterpret_cast<const float2*>(p.offs)[i];
            ox[l] = o.x;
            oy[l] = o.y;
        }
    }
    if (p.softmax) {
        float mx = aw[0];
#pragma unroll
        for (int l = 1; l < LT; ++l) mx = fmaxf(mx, aw[l]);
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 2));
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            aw[l] = __expf(aw[l] - mx);
            s += aw[l];
        }
        s += __shfl_xor_sync(FULL, s, 1);
        s += __shfl_xor_sync(FULL, s, 2);
        const float inv = 1.0f / s;
#pragma unroll
        for (int l = 0; l < LT; ++l) aw[l] *= inv;
    }

    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const int row = p.H * p.D;                               // elements per value row (all heads of one pixel)
    const char* v_img = reinterpret_cast<const char*>(p.value) + (size_t)b * p.S * row * (VBF16 ? 2 : 4);
    const int lane_off = h * p.D + c0;

#pragma unroll
    for (int l = 0; l < LT; ++l) {
        const int Hl = p.lvl_h[l], Wl = p.lvl_w[l];
        const float Wf = (float)Wl, Hf = (float)Hl;
        float x = ox[l], y = oy[l];
        if (p.ref) {
            x = x / Wf;
            y = y / Hf;
            if (SBF16) {
                x = __bfloat162float(__float2bfloat16_rn(x));
                y = __bfloat162float(__float2bfloat16_rn(y));
            }
            const float2 r = *reinterpret_cast<const float2*>(p.ref + (bq * LT + l) * 2);
            x = r.x + x;
            y = r.y + y;
        }
        const float gx = 2.0f * x - 1.0f, gy = 2.0f * y - 1.0f;
        const float ix = ((gx + 1.0f) * Wf - 1.0f) * 0.5f;
        const float iy = ((gy + 1.0f) * Hf - 1.0f) * 0.5f;
        const float fx = floorf(ix), fy = floorf(iy);
        const float ax = ix - fx, ay = iy - fy, bx = (fx + 1.0f) - ix, by = (fy + 1.0f) - iy;
        const int x0 = (int)fminf(fmaxf(fx, -2.0f), Wf + 1.0f), y0 = (int)fminf(fmaxf(fy, -2.0f), Hf + 1.0f);
        const bool fin = (ix == ix) && (iy == iy);
        const bool x0ok = fin && x0 >= 0 && x0 < Wl, x1ok = fin && x0 + 1 >= 0 && x0 + 1 < Wl;
        const bool y0ok = y0 >= 0 && y0 < Hl, y1ok = y0 + 1 >= 0 && y0 + 1 < Hl;
        const int my_mask = (x0ok && y0ok ? 1 : 0) | (x1ok && y0ok ? 2 : 0) | (x0ok && y1ok ? 4 : 0) | (x1ok && y1ok ? 8 : 0);
        const int my_base = (p.lvl_start[l] + y0 * Wl + x0) * row;       // element index of the north-west corner's row
        const float w0 = aw[l] * (bx * by), w1 = aw[l] * (ax * by), w2 = aw[l] * (bx * ay), w3 = aw[l] * (ax * ay);
        const int dyr = Wl * row;
        float wgt[4][4];
        uint4 va[4][4], vb[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int base = __shfl_sync(FULL, my_base, u, 4) + lane_off;
            const int m = __shfl_sync(FULL, my_mask, u, 4);
            wgt[u][0] = __shfl_sync(FULL, w0, u, 4);
            wgt[u][1] = __shfl_sync(FULL, w1, u, 4);
            wgt[u][2] = __shfl_sync(FULL, w2, u, 4);
            wgt[u][3] = __shfl_sync(FULL, w3, u, 4);
            const int e[4] = {base, base + row, base + dyr, base + dyr + row};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                va[u][c] = make_uint4(0u, 0u, 0u, 0u);
                vb[u][c] = va[u][c];
                if (m & (1 << c)) {
                    if (VBF16) {
                        va[u][c] = __ldg(reinterpret_cast<const uint4*>(v_img + (long long)e[c] * 2));
                    } else {
                        const uint4* q = reinterpret_cast<const uint4*>(v_img + (long long)e[c] * 4);
                        va[u][c] = __ldg(q);
                        vb[u][c] = __ldg(q + 1);
                    }
                }
            }
            if (!VBF16) {                                   // fp32 values: 8 x 32-byte loads in flight, then accumulate
#pragma unroll
                for (int c = 0; c < 4; ++c) fma8<VBF16>(acc, wgt[u][c], va[u][c], vb[u][c]);
            }
        }
        if (VBF16) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < 4; ++c) fma8<VBF16>(acc, wgt[u][c], va[u][c], vb[u][c]);
        }
    }

    if (!live) return;
    const long long o = item * p.D + c0;
    if (p.out_bf16) {
        uint4 r;
        uint32_t* rr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
            rr[i] = *reinterpret_cast<const uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = r;
    } else {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

template <bool VBF16, bool SBF16>
void launch_quad(const MsdaParams& p, unsigned blocks, cudaStream_t s) {
    switch (p.L) {
        case 1: msda_fwd_quad_kernel<VBF16, SBF16, 1><<<blocks, 256, 0, s>>>(p); break;
        case 2: msda_fwd_quad_kernel<VBF16, SBF16, 2><<<blocks, 256, 0, s>>>(p); break;
        case 3: msda_fwd_quad_kernel<VBF16, SBF16, 3><<<blocks, 256, 0, s>>>(p); break;
        default: msda_fwd_quad_kernel<VBF16, SBF16, 4><<<blocks, 256, 0, s>>>(p); break;
    }
}

}  // namespace

extern "C" int rgbd_msda_fwd(const void* value, int value_dtype, const int* level_hw_host, int n_levels, const void* offsets,
                             int offsets_dtype, const float* reference_points, const void* attn, int attn_dtype, int softmax,
                             void* out, int out_dtype, int B, int S, int Q, int H, int D, int P, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(value && level_hw_host && offsets && attn && out, "msda_fwd: null pointer");
    RGBD_CHECK_ARG(n_levels >= 1 && n_levels <= kMaxLevels, "msda_fwd: 1..%d levels", kMaxLevels);
    RGBD_CHECK_ARG(B >= 1 && Q >= 1 && H >= 1 && P >= 1 && P <= 64, "msda_fwd: bad sizes");
    RGBD_CHECK_ARG(D >= 8 && D % 8 == 0, "msda_fwd: head dimension must be a multiple of 8 (got %d)", D);
    for (int dt : {value_dtype, offsets_dtype, attn_dtype, out_dtype})
        RGBD_CHECK_ARG(dt == RGBD_DTYPE_F32 || dt == RGBD_DTYPE_BF16, "msda_fwd: dtypes are f32 or bf16");
    MsdaParams p;
    p.value = value; p.offs = offsets; p.ref = reference_points; p.attn = attn; p.out = out;
    p.n_items = (long long)B * Q * H;
    p.S = S; p.Q = Q; p.H = H; p.D = D; p.L = n_levels; p.P = P; p.lph = D / 8;
    p.offs_bf16 = offsets_dtype == RGBD_DTYPE_BF16; p.attn_bf16 = attn_dtype == RGBD_DTYPE_BF16;
    p.out_bf16 = out_dtype == RGBD_DTYPE_BF16; p.softmax = softmax != 0;
    long long start = 0;
    for (int l = 0; l < n_levels; ++l) {
        const int h = level_hw_host[2 * l], w = level_hw_host[2 * l + 1];
        RGBD_CHECK_ARG(h >= 1 && w >= 1, "msda_fwd: bad level %d", l);
        p.lvl_h[l] = h; p.lvl_w[l] = w; p.lvl_start[l] = (int)start;
        start += (long long)h * w;
    }
    RGBD_CHECK_ARG(start == S, "msda_fwd: the level sizes sum to %lld, the value sequence has %d rows", start, S);
    const long long threads = p.n_items * p.lph;
    const long long blocks = (threads + 255) / 256;
    RGBD_CHECK_ARG(blocks <= 0x7fffffffLL, "msda_fwd: too many work items");
    cudaStream_t s = (cudaStream_t)stream;
    const bool vb = value_dtype == RGBD_DTYPE_BF16;
    const long long img_elems = (long long)S * H * D;
    if (P == 4 && D == 32 && n_levels <= 4 && offsets_dtype == attn_dtype && img_elems < (1ll << 30)) {
        const bool sb = offsets_dtype == RGBD_DTYPE_BF16;
        if (vb && sb) launch_quad<true, true>(p, (unsigned)blocks, s);
        else if (vb) launch_quad<true, false>(p, (unsigned)blocks, s);
        else if (sb) launch_quad<false, true>(p, (unsigned)blocks, s);
        else launch_quad<false, false>(p, (unsigned)blocks, s);
    } else if (P == 4) {
        if (vb) msda_fwd_kernel<true, 4><<<(unsigned)blocks, 256, 0, s>>>(p);
        else msda_fwd_kernel<false, 4><<<(unsigned)blocks, 256, 0, s>>>(p);
    } else {
        if (vb) msda_fwd_kernel<true, 0><<<(unsigned)blocks, 256, 0, s>>>(p);
        else msda_fwd_kernel<false, 0><<<(unsigned)blocks, 256, 0, s>>>(p);
    }
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
