// K0r: the two uint8 image resizes of the reference's data mapper (map_10channel_case2, reference
// mask2former/utils/dataloader.py:405-414), bit-exact with the libraries the reference calls:
//   * Pillow `Image.resize((w, h), BILINEAR)` (HF Mask2FormerImageProcessor, PIL backend, `resample = 2`): libImaging's
//     ImagingResample for 8 bits per channel -- separable convolution with the triangle filter widened by the scale when
//     shrinking, double-precision coefficients normalised and rounded to 22 fractional bits, int32 accumulation from
//     1 << 21, clip to 0..255, horizontal pass first with a uint8 intermediate.  The coefficient tables are computed ON
//     THE DEVICE in fp64 with Pillow's own operation order (this file is compiled with -fmad=false).
//   * OpenCV `cv2.resize(depth, (w, h), interpolation=INTER_LINEAR)` for CV_8UC1: float source coordinates from
//     scale = 1 / ((double)dst / src), 11-bit weights `saturate_cast<short>(w * 2048)`, the x weights reset to (1, 0) where
//     the window leaves the image, the y ROW INDICES clipped instead (border rows blend a row with itself), vertical
//     combine (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// HBM-bound byte work: one thread per output sample, coalesced along x and the interleaved channels.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

// Pillow precompute_coeffs + normalize_coeffs_8bpc (bilinear filter, support 1.0), one thread per output index
__global__ void pil_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = (double)in_size / (double)out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale;
    const double center = 0.0 + (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = (x + xmin - center + 0.5) * ss;
        if (a < 0.0) a = -a;
        ww += a < 1.0 ? 1.0 - a : 0.0;
    }
    int* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) {
        double w = 0.0;
        if (x < xmax) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w = a < 1.0 ? 1.0 - a : 0.0;
            if (ww != 0.0) w /= ww;
        }
        k[x] = w < 0.0 ? (int)(-0.5 + w * (double)(1 << kPrecisionBits)) : (int)(0.5 + w * (double)(1 << kPrecisionBits));
    }
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;                 // arithmetic shift, like Pillow's lookup index
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// src (B, H, W, C) -> dst (B, H, w, C)
__global__ void __launch_bounds__(256) pil_resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int w,
                                                           int C, int ksize, const int* __restrict__ bounds,
                                                           const int* __restrict__ kk) {
    const int b = blockIdx.z, y = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;            // (xx, c)
    if (i >= w * C) return;
    const int xx = i / C, c = i - xx * C;
    const int x0 = bounds[xx * 2], n = bounds[xx * 2 + 1];
    const uint8_t* row = src + ((size_t)b * H + y) * W * C;
    const int* k = kk + (size_t)xx * ksize;
    int ss = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < n; ++x) ss += (int)row[(size_t)(x0 + x) * C + c] * k[x];
    dst[((size_t)b * H + y) * w * C + i] = clip8(ss);
}

// src (B, H, w, C) -> dst (B, h, w, C)
__global__ void __launch_bounds__(256) pil_resize_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int h, int wC,
                                                           int ksize, const int* __restrict__ bounds, const int* __restrict__ kk) {
    const int b = blockIdx.z, yy = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wC) return;
    const int y0 = bounds[yy * 2], n = bounds[yy * 2 + 1];
    const int* k = kk + (size_t)yy * ksize;
    const uint8_t* col = src + ((size_t)b * H + y0) * wC + i;
    int ss = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < n; ++y) ss += (int)col[(size_t)y * wC] * k[y];
    dst[((size_t)b * h + yy) * wC + i] = clip8(ss);
}

__global__ void __launch_bounds__(256) copy_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// OpenCV tables: ofs[d], wts[2d], wts[2d+1]; clamp_weights = 1 along x, 0 along y
__global__ void cv_tables_kernel(int in_size, int out_size, int clamp_weights, int* __restrict__ ofs, int* __restrict__ wts) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= out_size) return;
    const double inv_scale = (double)out_size / (double)in_size;
    const double scale = 1.0 / inv_scale;
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_weights) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= in_size - 1) { f = 0.f; s = in_size - 1; }
    }
    ofs[d] = s;
    wts[2 * d] = __float2int_rn((1.f - f) * 2048.f);       // saturate_cast<short>: cvRound, round half to even
    wts[2 * d + 1] = __float2int_rn(f * 2048.f);
}

// src (B, H, W) -> dst (B, h, w)
__global__ void __launch_bounds__(256) cv_resize_linear_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int h,
                                                               int w, const int* __restrict__ xofs, const int* __restrict__ alpha,
                                                               const int* __restrict__ yofs, const int* __restrict__ beta) {
    const int b = blockIdx.z, dy = blockIdx.y;
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    if (dx >= w) return;
    const int sx = xofs[dx], sx1 = sx + 1 < W ? sx + 1 : W - 1;
    const int a0 = alpha[2 * dx], a1 = alpha[2 * dx + 1];
    int sy0 = yofs[dy], sy1 = sy0 + 1;
    sy0 = sy0 < 0 ? 0 : (sy0 < H ? sy0 : H - 1);
    sy1 = sy1 < 0 ? 0 : (sy1 < H ? sy1 : H - 1);
    const uint8_t* img = src + (size_t)b * H * W;
    const int r0 = (int)img[(size_t)sy0 * W + sx] * a0 + (int)img[(size_t)sy0 * W + sx1] * a1;
    const int r1 = (int)img[(size_t)sy1 * W + sx] * a0 + (int)img[(size_t)sy1 * W + sx1] * a1;
    const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
    int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    dst[((size_t)b * h + dy) * w + dx] = (uint8_t)v;
}

int pil_ksize(int in_size, int out_size) {
    double filterscale = (double)in_size / (double)out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    return (int)ceil(1.0 * filterscale) * 2 + 1;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" size_t rgbd_resize_workspace_bytes(int B, int H, int W, int C, int h, int w) {
    if (B < 1 || H < 1 || W < 1 || C < 1 || h < 1 || w < 1) return 0;
    const size_t tab_x = align256((size_t)w * (2 + pil_ksize(W, w)) * 4), tab_y = align256((size_t)h * (2 + pil_ksize(H, h)) * 4);
    const size_t cv_tab = align256((size_t)(w + h) * 3 * 4);
    return tab_x + tab_y + cv_tab + align256((size_t)B * H * w * C) + 256;
}

extern "C" int rgbd_resize_pil_bilinear_u8(const uint8_t* src, uint8_t* dst, int B, int H, int W, int C, int h, int w, void* workspace,
                                           rgbd_stream_t stream) {
    RGBD_CHECK_ARG(src && dst && workspace, "resize_pil_bilinear: null pointer");
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && C <= 4 && h >= 1 && w >= 1, "resize_pil_bilinear: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const int kx = pil_ksize(W, w), ky = pil_ksize(H, h);
    int* bx = reinterpret_cast<int*>(ws);
    int* kkx = bx + 2 * w;
    ws += align256((size_t)w * (2 + kx) * 4);
    int* by = reinterpret_cast<int*>(ws);
    int* kky = by + 2 * h;
    ws += align256((size_t)h * (2 + ky) * 4);
    ws += align256((size_t)(w + h) * 3 * 4);
    uint8_t* tmp = ws;
    const uint8_t* cur = src;
    if (w == W && h == H) {                                     // Pillow returns a copy
        copy_u8_kernel<<<296, 256, 0, s>>>(src, dst, (size_t)B * H * W * C);
        RGBD_CHECK_LAUNCH();
        return RGBD_OK;
    }
    if (w != W) {
        pil_coeffs_kernel<<<ceil_div(w, 128), 128, 0, s>>>(W, w, kx, bx, kkx);
        RGBD_CHECK_LAUNCH();
        uint8_t* out = h != H ? tmp : dst;
        pil_resize_h_kernel<<<dim3(ceil_div(w * C, 256), H, B), 256, 0, s>>>(cur, out, H, W, w, C, kx, bx, kkx);
        RGBD_CHECK_LAUNCH();
        cur = out;
    }
    if (h != H) {
        pil_coeffs_kernel<<<ceil_div(h, 128), 128, 0, s>>>(H, h, ky, by, kky);
        RGBD_CHECK_LAUNCH();
        pil_resize_v_kernel<<<dim3(ceil_div(w * C, 256), h, B), 256, 0, s>>>(cur, dst, H, h, w * C, ky, by, kky);
        RGBD_CHECK_LAUNCH();
    }
    return RGBD_OK;
}

extern "C" int rgbd_resize_cv_linear_u8(const uint8_t* src, uint8_t* dst, int B, int H, int W, int h, int w, void* workspace,
                                        rgbd_stream_t stream) {
    RGBD_CHECK_ARG(src && dst && workspace, "resize_cv_linear: null pointer");
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && h >= 1 && w >= 1, "resize_cv_linear: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    if (w == W && h == H) {                                     // OpenCV copies
        copy_u8_kernel<<<296, 256, 0, s>>>(src, dst, (size_t)B * H * W);
        RGBD_CHECK_LAUNCH();
        return RGBD_OK;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    ws += align256((size_t)w * (2 + pil_ksize(W, w)) * 4) + align256((size_t)h * (2 + pil_ksize(H, h)) * 4);
    int* xofs = reinterpret_cast<int*>(ws);
    int* alpha = xofs + w;
    int* yofs = alpha + 2 * w;
    int* beta = yofs + h;
    cv_tables_kernel<<<ceil_div(w, 128), 128, 0, s>>>(W, w, 1, xofs, alpha);
    RGBD_CHECK_LAUNCH();
    cv_tables_kernel<<<ceil_div(h, 128), 128, 0, s>>>(H, h, 0, yofs, beta);
    RGBD_CHECK_LAUNCH();
    cv_resize_linear_kernel<<<dim3(ceil_div(w, 256), h, B), 256, 0, s>>>(src, dst, H, W, h, w, xofs, alpha, yofs, beta);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
