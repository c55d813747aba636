// K1: fused DGGM forward (DepthGradientInjectionResidual.forward, reference
// mask2former/utils/custom_model.py:1204-1269) for all pyramid scales in ONE launch:
//   gated_i = bilinear_down(grad)(y,x) * nearest_down(mask)(y,x)          (CM:1231-1246)
//   out_i   = color_i + ReLU(W_i . gated_i + b_i)                         (CM:1251-1255)
// and, when `branch1` is given, the v0.4.0 branch sum out_i = branch1_i + (color_i + enh_i)
// (CM:354-355) so the colour features are read once and the fused features written once.
//
// HBM-bound streaming kernel: a CTA owns (image, scale, 256-pixel run, channel chunk); the
// gated gradient for the run is staged in shared memory once and reused by every channel;
// colour planes are streamed with 128-bit L1-bypassing loads, 4 independent loads in flight
// per thread, and streaming (evict-first) 128-bit stores.
//
// K1b: dW_i / db_i of the 1x1 enhancement convs (the only gradients DGGM produces: colour
// features are detached and the gradient map is data).
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kMaxScales = 8;
constexpr int kMaxD = 8;
constexpr int kTP = 256;        // pixels per tile
constexpr int kThreads = 256;

struct DggmScale {
    const float* color;
    const float* branch1;
    float* out;
    const float* w;   // (C, D)
    const float* b;   // (C)
    int C, H, W, P;
    int tiles_p, tiles_c, TC;
    int tile_begin;
    int vec4;         // P % 4 == 0 and 16-byte aligned planes
    float sy, sx;     // in/out scale factors as torch computes them (float32 division)
};

struct DggmParams {
    DggmScale s[kMaxScales];
    int n_scales;
    const float* grad;   // (B, D, H, W) with batch stride grad_bs
    const float* mask;   // (B, 1, H, W) with batch stride mask_bs
    long long grad_bs, mask_bs;
    int B, D, H, W;
};

// torch area_pixel_compute_source_index(align_corners=False) + linear weights
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    src = fmaxf(src, 0.0f);
    i0 = (int)src;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = __fsub_rn(src, (float)i0);
}

__device__ __forceinline__ void stage_gated(const DggmParams& p, const DggmScale& S, int b, int p0, int np,
                                            float (*g_s)[kTP]) {
    const float* gb = p.grad + (long long)b * p.grad_bs;
    const float* mb = p.mask + (long long)b * p.mask_bs;
    const long long plane = (long long)p.H * p.W;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        int pix = p0 + i;
        int y = pix / S.W, x = pix - y * S.W;
        int y0, y1, x0, x1;
        float ly, lx;
        src_index(y, S.sy, p.H, y0, y1, ly);
        src_index(x, S.sx, p.W, x0, x1, lx);
        int my = min((int)floorf(__fmul_rn((float)y, S.sy)), p.H - 1);
        int mx = min((int)floorf(__fmul_rn((float)x, S.sx)), p.W - 1);
        float m = __ldg(mb + (long long)my * p.W + mx);
        float hy = 1.0f - ly, hx = 1.0f - lx;
        for (int d = 0; d < p.D; ++d) {
            const float* g = gb + d * plane;
            float v00 = __ldg(g + (long long)y0 * p.W + x0), v01 = __ldg(g + (long long)y0 * p.W + x1);
            float v10 = __ldg(g + (long long)y1 * p.W + x0), v11 = __ldg(g + (long long)y1 * p.W + x1);
            float top = v00 * hx + v01 * lx;
            float bot = v10 * hx + v11 * lx;
            g_s[d][i] = (top * hy + bot * ly) * m;
        }
    }
}

__device__ __forceinline__ const DggmScale& find_scale(const DggmParams& p, int& t) {
    int si = 0;
    while (si + 1 < p.n_scales && t >= p.s[si + 1].tile_begin) ++si;
    t -= p.s[si].tile_begin;
    return p.s[si];
}

template <int DMAX>
__global__ void __launch_bounds__(kThreads, DMAX <= 4 ? 4 : 2) dggm_fwd_kernel(const __grid_constant__ DggmParams p) {
    __shared__ __align__(16) float g_s[kMaxD][kTP];
    int t = blockIdx.x;
    const DggmScale& S = find_scale(p, t);
    const int ct = t % S.tiles_c;
    t /= S.tiles_c;
    const int pt = t % S.tiles_p;
    const int b = t / S.tiles_p;
    const int p0 = pt * kTP;
    const int np = min(kTP, S.P - p0);
    const int c0 = ct * S.TC;
    const int c1 = min(c0 + S.TC, S.C);
    const int D = p.D;

    stage_gated(p, S, b, p0, np, g_s);
    __syncthreads();

    const long long img_off = (long long)b * S.C * S.P + p0;
    if (S.vec4) {
        const int nf4 = np >> 2;
        const int f = threadIdx.x & 63;
        const int lane_c = threadIdx.x >> 6;           // 4 channel lanes
        if (f >= nf4) return;
        float4 g[DMAX];
#pragma unroll
        for (int d = 0; d < DMAX; ++d)
            if (d < D) g[d] = *reinterpret_cast<const float4*>(&g_s[d][f * 4]);
        constexpr int U = 4;
        for (int c = c0 + lane_c; c < c1; c += 4 * U) {
            float4 v[U], r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int cc = c + 4 * u;
                if (cc < c1) {
                    long long off = img_off + (long long)cc * S.P + f * 4;
                    v[u] = ld_stream_f4(S.color + off);
                    if (S.branch1) r[u] = ld_stream_f4(S.branch1 + off);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int cc = c + 4 * u;
                if (cc < c1) {
                    float bias = __ldg(S.b + cc);
                    float4 e = make_float4(bias, bias, bias, bias);
#pragma unroll
                    for (int d = 0; d < DMAX; ++d) {
                        if (d < D) {
                            float w = __ldg(S.w + cc * D + d);
                            e.x = fmaf(w, g[d].x, e.x);
                            e.y = fmaf(w, g[d].y, e.y);
                            e.z = fmaf(w, g[d].z, e.z);
                            e.w = fmaf(w, g[d].w, e.w);
                        }
                    }
                    float4 o;
                    o.x = v[u].x + fmaxf(e.x, 0.f);
                    o.y = v[u].y + fmaxf(e.y, 0.f);
                    o.z = v[u].z + fmaxf(e.z, 0.f);
                    o.w = v[u].w + fmaxf(e.w, 0.f);
                    if (S.branch1) {
                        o.x = r[u].x + o.x;
                        o.y = r[u].y + o.y;
                        o.z = r[u].z + o.z;
                        o.w = r[u].w + o.w;
                    }
                    st_stream_f4(S.out + img_off + (long long)cc * S.P + f * 4, o);
                }
            }
        }
    } else {
        // scalar path for planes that are not 16-byte aligned (tiny / ragged test shapes)
        const int nc = c1 - c0;
        for (int idx = threadIdx.x; idx < nc * np; idx += blockDim.x) {
            int cc = c0 + idx / np;
            int i = idx - (idx / np) * np;
            float e = __ldg(S.b + cc);
            for (int d = 0; d < D; ++d) e = fmaf(__ldg(S.w + cc * D + d), g_s[d][i], e);
            long long off = img_off + (long long)cc * S.P + i;
            float o = S.color[off] + fmaxf(e, 0.f);
            if (S.branch1) o = S.branch1[off] + o;
            S.out[off] = o;
        }
    }
}

// ---- K1b: parameter gradients ------------------------------------------------------------
// dW[c][d] = sum_{b,pix} 1[pre>0] * dOut[b,c,pix] * gated[b,d,pix];  db[c] = sum 1[pre>0] * dOut.
// A CTA owns (image, scale, pixel run, channel chunk); each warp reduces its channels over the
// run with shuffles and adds one atomic per (c, d) per tile.
struct DggmBwdScale {
    const float* dout;
    const float* w;
    const float* b;
    float* dw;
    float* db;
    int C, H, W, P;
    int tiles_p, tiles_c, TC;
    int tile_begin;
    float sy, sx;
};
struct DggmBwdParams {
    DggmBwdScale s[kMaxScales];
    DggmParams geom;   // only grad/mask/geometry fields are used
};

__global__ void __launch_bounds__(kThreads) dggm_bwd_params_kernel(const __grid_constant__ DggmBwdParams q) {
    __shared__ __align__(16) float g_s[kMaxD][kTP];
    const DggmParams& p = q.geom;
    int t = blockIdx.x;
    int si = 0;
    while (si + 1 < p.n_scales && t >= q.s[si + 1].tile_begin) ++si;
    t -= q.s[si].tile_begin;
    const DggmBwdScale& S = q.s[si];
    const int ct = t % S.tiles_c;
    t /= S.tiles_c;
    const int pt = t % S.tiles_p;
    const int b = t / S.tiles_p;
    const int p0 = pt * kTP;
    const int np = min(kTP, S.P - p0);
    const int c0 = ct * S.TC;
    const int c1 = min(c0 + S.TC, S.C);
    const int D = p.D;
    DggmScale G;
    G.W = S.W; G.H = S.H; G.sy = S.sy; G.sx = S.sx;
    stage_gated(p, G, b, p0, np, g_s);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long img_off = (long long)b * S.C * S.P + p0;
    for (int cc = c0 + warp; cc < c1; cc += kThreads / 32) {
        float wv[kMaxD];
#pragma unroll
        for (int d = 0; d < kMaxD; ++d) wv[d] = d < D ? __ldg(S.w + cc * D + d) : 0.f;
        const float bias = __ldg(S.b + cc);
        float acc[kMaxD];
#pragma unroll
        for (int d = 0; d < kMaxD; ++d) acc[d] = 0.f;
        float accb = 0.f;
        for (int i = lane; i < np; i += 32) {
            float e = bias;
#pragma unroll
            for (int d = 0; d < kMaxD; ++d)
                if (d < D) e = fmaf(wv[d], g_s[d][i], e);
            float go = e > 0.f ? ld_stream_f1(S.dout + img_off + (long long)cc * S.P + i) : 0.f;
            accb += go;
#pragma unroll
            for (int d = 0; d < kMaxD; ++d)
                if (d < D) acc[d] = fmaf(go, g_s[d][i], acc[d]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            accb += __shfl_xor_sync(0xffffffffu, accb, o);
#pragma unroll
            for (int d = 0; d < kMaxD; ++d)
                if (d < D) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
        }
        if (lane == 0) {
            atomicAdd(S.db + cc, accb);
#pragma unroll
            for (int d = 0; d < kMaxD; ++d)
                if (d < D) atomicAdd(S.dw + cc * D + d, acc[d]);
        }
    }
}

int pick_tc(int C) {
    if (C % 96 == 0) return 96;
    if (C % 128 == 0) return 128;
    if (C % 64 == 0) return 64;
    return C < 96 ? C : 96;
}

int fill_geometry(DggmParams& p, int n_scales, const int* C, const int* Hs, const int* Ws, const float* grad,
                  const float* mask, long long grad_bs, long long mask_bs, int B, int D, int H, int W) {
    RGBD_CHECK_ARG(n_scales >= 1 && n_scales <= kMaxScales, "dggm: n_scales %d out of [1,%d]", n_scales, kMaxScales);
    RGBD_CHECK_ARG(D >= 1 && D <= kMaxD, "dggm: depth_gradient_channels %d out of [1,%d]", D, kMaxD);
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "dggm: bad geometry B=%d H=%d W=%d", B, H, W);
    RGBD_CHECK_ARG(grad && mask, "dggm: grad/mask pointers are null (passthrough is handled by the caller)");
    p.n_scales = n_scales;
    p.grad = grad; p.mask = mask; p.grad_bs = grad_bs; p.mask_bs = mask_bs;
    p.B = B; p.D = D; p.H = H; p.W = W;
    for (int i = 0; i < n_scales; ++i)
        RGBD_CHECK_ARG(C[i] >= 1 && Hs[i] >= 1 && Ws[i] >= 1, "dggm: bad scale %d geometry", i);
    return RGBD_OK;
}

}  // namespace

extern "C" int rgbd_dggm_fwd(int n_scales, const float* const* color, const float* const* branch1, float* const* out,
                             const int* C, const int* Hs, const int* Ws, const float* const* weight,
                             const float* const* bias, const float* grad, long long grad_batch_stride,
                             const float* mask, long long mask_batch_stride, int B, int D, int H, int W,
                             rgbd_stream_t stream) {
    DggmParams p;
    int rc = fill_geometry(p, n_scales, C, Hs, Ws, grad, mask, grad_batch_stride, mask_batch_stride, B, D, H, W);
    if (rc) return rc;
    long long tiles = 0;
    for (int i = 0; i < n_scales; ++i) {
        DggmScale& S = p.s[i];
        RGBD_CHECK_ARG(color[i] && out[i] && weight[i] && bias[i], "dggm: null pointer at scale %d", i);
        S.color = color[i];
        S.branch1 = branch1 ? branch1[i] : nullptr;
        S.out = out[i];
        S.w = weight[i];
        S.b = bias[i];
        S.C = C[i]; S.H = Hs[i]; S.W = Ws[i]; S.P = Hs[i] * Ws[i];
        S.TC = pick_tc(S.C);
        S.tiles_c = ceil_div(S.C, S.TC);
        S.tiles_p = ceil_div(S.P, kTP);
        S.tile_begin = (int)tiles;
        S.sy = (float)H / (float)S.H;
        S.sx = (float)W / (float)S.W;
        bool aligned = ((uintptr_t)S.color % 16 == 0) && ((uintptr_t)S.out % 16 == 0) &&
                       (!S.branch1 || (uintptr_t)S.branch1 % 16 == 0);
        S.vec4 = (S.P % 4 == 0) && aligned;
        tiles += (long long)B * S.tiles_p * S.tiles_c;
    }
    RGBD_CHECK_ARG(tiles < (1ll << 31), "dggm: too many tiles");
    if (D <= 4) dggm_fwd_kernel<4><<<(unsigned)tiles, kThreads, 0, (cudaStream_t)stream>>>(p);
    else dggm_fwd_kernel<kMaxD><<<(unsigned)tiles, kThreads, 0, (cudaStream_t)stream>>>(p);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_dggm_bwd_params(int n_scales, const float* const* dout, const int* C, const int* Hs, const int* Ws,
                                    const float* const* weight, const float* const* bias, float* const* dweight,
                                    float* const* dbias, const float* grad, long long grad_batch_stride,
                                    const float* mask, long long mask_batch_stride, int B, int D, int H, int W,
                                    rgbd_stream_t stream) {
    DggmBwdParams q;
    int rc = fill_geometry(q.geom, n_scales, C, Hs, Ws, grad, mask, grad_batch_stride, mask_batch_stride, B, D, H, W);
    if (rc) return rc;
    long long tiles = 0;
    for (int i = 0; i < n_scales; ++i) {
        DggmBwdScale& S = q.s[i];
        RGBD_CHECK_ARG(dout[i] && weight[i] && bias[i] && dweight[i] && dbias[i], "dggm_bwd: null pointer at scale %d", i);
        S.dout = dout[i]; S.w = weight[i]; S.b = bias[i]; S.dw = dweight[i]; S.db = dbias[i];
        S.C = C[i]; S.H = Hs[i]; S.W = Ws[i]; S.P = Hs[i] * Ws[i];
        S.TC = 32;
        S.tiles_c = ceil_div(S.C, S.TC);
        S.tiles_p = ceil_div(S.P, kTP);
        S.tile_begin = (int)tiles;
        S.sy = (float)H / (float)S.H;
        S.sx = (float)W / (float)S.W;
        tiles += (long long)B * S.tiles_p * S.tiles_c;
        RGBD_CHECK_CUDA(cudaMemsetAsync(S.dw, 0, sizeof(float) * S.C * D, (cudaStream_t)stream));
        RGBD_CHECK_CUDA(cudaMemsetAsync(S.db, 0, sizeof(float) * S.C, (cudaStream_t)stream));
    }
    RGBD_CHECK_ARG(tiles < (1ll << 31), "dggm_bwd: too many tiles");
    dggm_bwd_params_kernel<<<(unsigned)tiles, kThreads, 0, (cudaStream_t)stream>>>(q);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
