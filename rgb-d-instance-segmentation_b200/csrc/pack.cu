// Operand packing for the tensor-core kernels (conv_gemm.cu).
//
// dsam_pack: NCHW fp32 features + pooled 4-bit region codes -> the K-concatenated, masked, bf16
//   channels-last operand of a DSAM stage (reference mask2former/utils/custom_model.py:683-696:
//   `masked_features = rgb_features * resized_mask` for each region, plus the unmasked copy for
//   rgb_projection).  For the stride-2 variant the pixels are split into 4 parity planes so every
//   3x3 stride-2 tap becomes a dense, stride-1 TMA box:
//       out[img][seg][py*2+px][y>>1][x>>1][c] = bf16(F[img][c][y][x] * bit(code, seg))   seg < n_seg-1
//       out[img][n_seg-1][...]                = bf16(F)                                   (projection copy)
// ratio_stem_pack: depth (B,3,H,W) -> row-im2col tensor R[img][H+6][W][64] with
//       R[img][r][x][(j*8+dx)*4+c] = depth[img][c][r-3+j][x+dx-3]   (0 outside, 0 for dx==7 or c==3)
//   so the 3x3/5x5/7x7 stem convs (CM:1458-1460) are one GEMM of 4 K-slices of 64 (taps dy = 2t+j).
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

// Variant for many segments per pixel (pre-masked five-fold layouts): shared-memory transpose, 128-byte coalesced rows.
__global__ void __launch_bounds__(256) dsam_pack_tile_kernel(const float* __restrict__ feat, const uint8_t* __restrict__ codes,
                                                        __nv_bfloat16* __restrict__ out, int C, int Cp, int H, int W,
                                                        int n_seg, int masked_segs, int split, int hi_lo) {
    // CTA: 32 pixels of one row x 64 channels.  Load NCHW coalesced along x, transpose through shared memory, then each
    // thread owns (pixel, 8 channels) and writes one 16-byte piece per segment: 8 lanes = one 128-byte channels-last row.
    __shared__ float tile[64][33];
    const int img = blockIdx.z;
    const int y = blockIdx.y;
    const int x0 = (blockIdx.x % ((W + 31) / 32)) * 32;
    const int c0 = (blockIdx.x / ((W + 31) / 32)) * 64;
    const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;   // 32 x 8
    const size_t plane = (size_t)H * W;
    for (int c = tyy; c < 64; c += 8) {
        const int cc = c0 + c, x = x0 + tx;
        tile[c][tx] = (cc < C && x < W) ? feat[((size_t)img * C + cc) * plane + (size_t)y * W + x] : 0.f;
    }
    __syncthreads();
    const int H2 = split ? (H + 1) / 2 : H, W2 = split ? (W + 1) / 2 : W;
    const int n_par = split ? 4 : 1;
    const int px = threadIdx.x >> 3, cg = threadIdx.x & 7;     // pixel 0..31, channel group of 8
    const int x = x0 + px;
    const int c = c0 + cg * 8;
    if (x >= W || c >= Cp) return;
    const unsigned code = codes[(size_t)img * plane + (size_t)y * W + x];
    const int par = split ? ((y & 1) * 2 + (x & 1)) : 0;
    const int yy = split ? (y >> 1) : y, xx = split ? (x >> 1) : x;
    // hi = bf16(v); lo = bf16(v - hi): with W = W_hi + W_lo the three products hi*W_hi + lo*W_hi + hi*W_lo carry ~16
    // mantissa bits through the bf16 tensor cores (the "fp32" precision mode of DSAModule)
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = tile[cg * 8 + k][px];
    uint4 v, vl;
    {
        uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
        uint32_t* pl = reinterpret_cast<uint32_t*>(&vl);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
            pv[k] = *reinterpret_cast<uint32_t*>(&h);
            const float2 hf = __bfloat1622float2(h);
            __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * k] - hf.x, f[2 * k + 1] - hf.y);
            pl[k] = *reinterpret_cast<uint32_t*>(&l);
        }
    }
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const int mult = hi_lo ? 2 : 1;
    for (int s = 0; s < n_seg; ++s) {
        const bool keep = s >= masked_segs || ((code >> s) & 1u);
        const size_t pl = ((size_t)img * n_seg * mult + s * mult) * n_par + par;
        const size_t off = ((pl * H2 + yy) * W2 + xx) * Cp + c;
        *reinterpret_cast<uint4*>(out + off) = keep ? v : zero;
        if (hi_lo) *reinterpret_cast<uint4*>(out + off + (size_t)n_par * H2 * W2 * Cp) = keep ? vl : zero;
    }
}

// CTA = 128 consecutive pixels (flattened y*W + x) x 64 channels; warp w owns channels 8w..8w+7, lane L pixels 4L..4L+3.
// Loads: 8 x 128-bit per thread, each fully coalesced along the NCHW plane (512 B per warp and channel).  Stores: every
// thread writes its pixel's 8 channels as one 16-byte piece per segment; the 8 warps of the CTA fill the 128-byte
// channels-last rows of the same pixels together, so the L2 merges them into whole lines.  No shared memory, no syncs.
template <bool VEC>
__global__ void __launch_bounds__(256) dsam_pack_kernel(const float* __restrict__ feat, const uint8_t* __restrict__ codes,
                                                        __nv_bfloat16* __restrict__ out, int C, int Cp, int H, int W,
                                                        int n_seg, int masked_segs, int split, int hi_lo) {
    const int img = blockIdx.z;
    const int c0 = blockIdx.y * 64 + (threadIdx.x >> 5) * 8;
    const int plane = H * W;
    const int p0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    if (p0 >= plane || c0 >= Cp) return;
    float f[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int cc = c0 + k;
        if (cc < C) {
            const float* src = feat + ((size_t)img * C + cc) * plane + p0;
            if (VEC) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                f[k][0] = v.x; f[k][1] = v.y; f[k][2] = v.z; f[k][3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) f[k][e] = p0 + e < plane ? __ldg(src + e) : 0.f;
            }
        } else {
            f[k][0] = f[k][1] = f[k][2] = f[k][3] = 0.f;
        }
    }
    const int H2 = split ? (H + 1) / 2 : H, W2 = split ? (W + 1) / 2 : W;
    const int n_par = split ? 4 : 1;
    const int mult = hi_lo ? 2 : 1;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int p = p0 + e;
        if (p >= plane) break;
        const int y = p / W, x = p - y * W;
        const unsigned code = masked_segs ? codes[(size_t)img * plane + p] : 0u;
        const int par = split ? ((y & 1) * 2 + (x & 1)) : 0;
        const int yy = split ? (y >> 1) : y, xx = split ? (x >> 1) : x;
        // hi = bf16(v); lo = bf16(v - hi): with W = W_hi + W_lo the three products hi*W_hi + lo*W_hi + hi*W_lo carry ~16
        // mantissa bits through the bf16 tensor cores (the "fp32" precision mode of DSAModule)
        uint4 v, vl;
        {
            uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
            uint32_t* pl = reinterpret_cast<uint32_t*>(&vl);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k][e], f[2 * k + 1][e]);
                pv[k] = *reinterpret_cast<uint32_t*>(&h);
                const float2 hf = __bfloat1622float2(h);
                __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * k][e] - hf.x, f[2 * k + 1][e] - hf.y);
                pl[k] = *reinterpret_cast<uint32_t*>(&l);
            }
        }
        for (int s = 0; s < n_seg; ++s) {
            const bool keep = s >= masked_segs || ((code >> s) & 1u);
            const size_t pln = ((size_t)img * n_seg * mult + s * mult) * n_par + par;
            const size_t off = ((pln * H2 + yy) * W2 + xx) * Cp + c0;
            *reinterpret_cast<uint4*>(out + off) = keep ? v : zero;
            if (hi_lo) *reinterpret_cast<uint4*>(out + off + (size_t)n_par * H2 * W2 * Cp) = keep ? vl : zero;
        }
    }
}

__global__ void __launch_bounds__(256) ratio_stem_pack_kernel(const float* __restrict__ depth, long long bs, long long cs,
                                                              __nv_bfloat16* __restrict__ out, int H, int W, int C) {
    // thread = (pixel, piece): piece p of 8 holds taps dx = 2*(p&3), 2*(p&3)+1 of row j = p>>2  (8 bf16 = 16 bytes);
    // the 8 lanes of a pixel write its 128-byte row contiguously, a warp writes 4 pixels = 512 contiguous bytes
    const int img = blockIdx.z;
    const int r = blockIdx.y;                                 // 0 .. H+5
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = gid >> 3, piece = gid & 7;
    if (x >= W) return;
    const int j = piece >> 2, dxp = piece & 3;
    const float* d = depth + (size_t)img * bs;
    const int y = r - 3 + j;
    const bool yok = y >= 0 && y < H;
    float v[8];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int dx = dxp * 2 + e;
        const int xs = x + dx - 3;
        const bool ok = yok && dx < 7 && xs >= 0 && xs < W;
#pragma unroll
        for (int c = 0; c < 4; ++c) v[e * 4 + c] = (ok && c < C) ? __ldg(d + c * cs + (size_t)y * W + xs) : 0.f;
    }
    uint4 w;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    w.x = *reinterpret_cast<uint32_t*>(&h0);
    w.y = *reinterpret_cast<uint32_t*>(&h1);
    w.z = *reinterpret_cast<uint32_t*>(&h2);
    w.w = *reinterpret_cast<uint32_t*>(&h3);
    reinterpret_cast<uint4*>(out + (((size_t)img * (H + 6) + r) * W + x) * 64)[piece] = w;
}

// Compact stem operand E[img][s][r][xx][4] = depth[img][c][r-3][xx+s-3] (0 outside / for c == 3): the depth image itself,
// channels-last with zero borders, in two copies shifted by one pixel (ratio_front.cu reads 64-byte sliding windows of it).
__global__ void __launch_bounds__(256) ratio_stem_pack_compact_kernel(const float* __restrict__ depth, long long bs, long long cs,
                                                                      __nv_bfloat16* __restrict__ out, int H, int W, int Wp, int C) {
    const int img = blockIdx.z, r = blockIdx.y;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= Wp) return;
    const float* d = depth + (size_t)img * bs;
    const int y = r - 3;
    const bool yok = y >= 0 && y < H;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int xs = xx + s - 3;
        const bool ok = yok && xs >= 0 && xs < W;
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = (ok && c < C) ? __ldg(d + c * cs + (size_t)y * W + xs) : 0.f;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        uint2 w;
        w.x = *reinterpret_cast<uint32_t*>(&h0);
        w.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(out + ((((size_t)img * 2 + s) * (H + 6) + r) * Wp + xx) * 4) = w;
    }
}

}  // namespace

extern "C" int rgbd_ratio_stem_compact_width(int W) { return W + 8; }

extern "C" int rgbd_ratio_stem_pack_compact(const float* depth3, long long batch_stride, long long channel_stride, void* out_bf16,
                                            int B, int C, int H, int W, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(depth3 && out_bf16, "ratio_stem_pack_compact: null pointer");
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && C <= 4, "ratio_stem_pack_compact: bad geometry (1..4 channels)");
    const int Wp = rgbd_ratio_stem_compact_width(W);
    dim3 grid(ceil_div(Wp, 256), H + 6, B);
    ratio_stem_pack_compact_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth3, batch_stride, channel_stride,
                                                                           (__nv_bfloat16*)out_bf16, H, W, Wp, C);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_dsam_pack(const float* feat, const uint8_t* codes, void* out_bf16, int B, int C, int C_pad, int H, int W,
                              int n_seg, int masked_segs, int parity_split, int hi_lo, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(feat && out_bf16 && (codes || masked_segs == 0), "dsam_pack: null pointer");
    RGBD_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1, "dsam_pack: bad geometry");
    RGBD_CHECK_ARG(C_pad >= C && C_pad % 32 == 0, "dsam_pack: C_pad %d must be a multiple of 32 and >= C", C_pad);
    RGBD_CHECK_ARG(n_seg >= 1 && n_seg <= 8 && masked_segs >= 0 && masked_segs <= n_seg && masked_segs <= 4,
                   "dsam_pack: bad segment counts");
    if (n_seg > 2) {
        dim3 grid_t(ceil_div(W, 32) * ceil_div(C_pad, 64), H, B);
        dsam_pack_tile_kernel<<<grid_t, 256, 0, (cudaStream_t)stream>>>(feat, codes, (__nv_bfloat16*)out_bf16, C, C_pad, H, W, n_seg,
                                                                       masked_segs, parity_split ? 1 : 0, hi_lo ? 1 : 0);
        RGBD_CHECK_LAUNCH();
        return RGBD_OK;
    }
    dim3 grid(ceil_div(H * W, 128), ceil_div(C_pad, 64), B);
    const bool vec = (H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
    if (vec)
        dsam_pack_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, codes, (__nv_bfloat16*)out_bf16, C, C_pad, H, W, n_seg,
                                                                      masked_segs, parity_split ? 1 : 0, hi_lo ? 1 : 0);
    else
        dsam_pack_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, codes, (__nv_bfloat16*)out_bf16, C, C_pad, H, W, n_seg,
                                                                       masked_segs, parity_split ? 1 : 0, hi_lo ? 1 : 0);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_ratio_stem_pack(const float* depth3, long long batch_stride, long long channel_stride, void* out_bf16,
                                    int B, int C, int H, int W, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(depth3 && out_bf16, "ratio_stem_pack: null pointer");
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && C <= 4, "ratio_stem_pack: bad geometry (1..4 channels)");
    dim3 grid(ceil_div(W * 8, 256), H + 6, B);
    ratio_stem_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth3, batch_stride, channel_stride,
                                                                   (__nv_bfloat16*)out_bf16, H, W, C);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

// ---- GroupNorm over an NCHW fp32 tensor, in place (the pixel decoder's input_projections are
// Conv2d(C_i, 256, 1) + GroupNorm(32, 256), transformers Mask2FormerPixelDecoder; the conv runs as rgbd_conv_gemm) ------------
namespace {

// one CTA per (image, group): the group's cpg x HW values (a few hundred KB at most) are read three times from L2 --
// mean, centred variance (numerically what torch's rowwise moments give), normalise + affine
__global__ void __launch_bounds__(256) group_norm_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int C, int HW, int groups, float eps) {
    __shared__ double red[8];
    __shared__ float s_mean, s_rstd;
    const int g = blockIdx.x % groups, b = blockIdx.x / groups;
    const int cpg = C / groups;
    float* base = x + ((size_t)b * C + (size_t)g * cpg) * HW;
    const int n = cpg * HW;
    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
        return t;
    };
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += base[i];
    const double mean = block_sum(s) / n;
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = (double)base[i] - mean;
        q += d * d;
    }
    const double var = block_sum(q) / n;
    if (threadIdx.x == 0) { s_mean = (float)mean; s_rstd = (float)(1.0 / sqrt(var + (double)eps)); }
    __syncthreads();
    const float m = s_mean, r = s_rstd;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = g * cpg + i / HW;
        base[i] = (base[i] - m) * r * gamma[c] + beta[c];
    }
}

}  // namespace

extern "C" int rgbd_group_norm_inplace(float* x, const float* gamma, const float* beta, int B, int C, int HW, int groups, float eps,
                                       rgbd_stream_t stream) {
    RGBD_CHECK_ARG(x && gamma && beta, "group_norm: null pointer");
    RGBD_CHECK_ARG(B >= 1 && C >= 1 && HW >= 1 && groups >= 1 && C % groups == 0, "group_norm: C must be divisible by groups");
    group_norm_kernel<<<B * groups, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, C, HW, groups, eps);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
