// K5: instance post-processing on the device.  Replaces the HuggingFace routine the reference calls on the
// model outputs, `image_processor.post_process_instance_segmentation(outputs, threshold, target_sizes,
// return_binary_maps=True)` (reference mask2former/utils/model_essential_part.py:86-91, mask2former/predictor.py:34-36,
// 701-703), for a batch of same-sized images and without a host round trip per query:
//   1. postproc_select_kernel   softmax over the classes, top-Q of the Q*C (query, label) scores (bitonic sort in
//                               shared memory; defined order: score descending, flat index ascending on ties)
//   2. postproc_grid_kernel     per candidate: bilinear 384x384 upsample of its mask logits evaluated on the fly in
//                               separable form (torch's align_corners=False source-index arithmetic, operation for
//                               operation), pixel count of (logit > 0), sum of sigmoid over those pixels (deterministic
//                               two-level sum) and the SIGN BITMAP of the grid (18 KB per candidate);
//      postproc_stats_kernel    point-wise fallback for very tall logit planes, and the non-emptiness of the
//                               nearest-resized target mask when the target is smaller than the grid
//   3. postproc_finalize_kernel mask score = sum / (count + 1e-6), score = class score * mask score, keep rule
//                               (non-empty and score >= threshold), compact slots in candidate order
//   4. postproc_paint_kernel    binary masks of the kept segments at the target size (nearest sample of the 384 grid's sign
//                               bitmap written by pass 2), and the painted segmentation map (id of the last kept segment
//                               covering a pixel)
// The 384x384 upsampled logits (59 MB per image for 100 queries) are never materialised.
// rgbd_mask_iou: pairwise mask IoU (the evaluator's torchmetrics `iou_type="segm"` core) with 16-byte loads + popcount.
// Compiled with -fmad=false: the interpolation is written as separate multiplies and adds like ATen's kernel.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kGrid = 384;          // HF: "Scale back to preprocessed image size - (384, 384) for all models"
constexpr int kChunk = 4096;        // grid points per CTA in the stats pass
static_assert((kGrid * kGrid) % kChunk == 0 && kChunk % 256 == 0, "whole chunks, whole warps");
constexpr int kMaxSort = 8192;      // Q*C candidates sorted in shared memory

struct PostGeom {
    int B, Q, C1, h, w, Ht, Wt;
    float rh, rw;                   // h / 384, w / 384          (torch: area_pixel_compute_scale)
    float sy, sx;                   // 384 / Ht, 384 / Wt        (nearest: floor(dst * scale))
};

// value of the bilinearly upsampled (384 x 384, align_corners=False) logit plane at grid point (gy, gx)
__device__ __forceinline__ float grid_logit(const float* __restrict__ plane, const PostGeom& g, int gy, int gx) {
    float fy = __fsub_rn(__fmul_rn(g.rh, __fadd_rn((float)gy, 0.5f)), 0.5f);
    float fx = __fsub_rn(__fmul_rn(g.rw, __fadd_rn((float)gx, 0.5f)), 0.5f);
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < g.h - 1 ? 1 : 0), x1 = x0 + (x0 < g.w - 1 ? 1 : 0);
    const float ly1 = __fsub_rn(fy, (float)y0), ly0 = __fsub_rn(1.f, ly1);
    const float lx1 = __fsub_rn(fx, (float)x0), lx0 = __fsub_rn(1.f, lx1);
    const float v00 = __ldg(plane + (size_t)y0 * g.w + x0), v01 = __ldg(plane + (size_t)y0 * g.w + x1);
    const float v10 = __ldg(plane + (size_t)y1 * g.w + x0), v11 = __ldg(plane + (size_t)y1 * g.w + x1);
    const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
    const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
    return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// target pixel -> grid point (torch "nearest": min(floor(dst * scale), in - 1))
__device__ __forceinline__ int nearest_src(int dst, float scale) {
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < kGrid - 1 ? s : kGrid - 1;
}

__global__ void __launch_bounds__(256) postproc_select_kernel(const float* __restrict__ cls, PostGeom g, int n_sort,
                                                              float* __restrict__ sel_score, int* __restrict__ sel_query,
                                                              int* __restrict__ sel_label) {
    extern __shared__ unsigned long long keys[];
    const int b = blockIdx.x, C = g.C1 - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < n_sort; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    for (int q = warp; q < g.Q; q += nw) {
        const float* row = cls + ((size_t)b * g.Q + q) * g.C1;
        float mx = -INFINITY;
        for (int c = lane; c < g.C1; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, k));
        float sum = 0.f;
        for (int c = lane; c < g.C1; c += 32) sum += expf(row[c] - mx);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, k);
        for (int c = lane; c < C; c += 32) {
            const float p = expf(row[c] - mx) / sum;
            const unsigned flat = (unsigned)(q * C + c);
            keys[flat] = ((unsigned long long)__float_as_uint(p) << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
        }
    }
    __syncthreads();
    // bitonic sort, descending
    for (int k = 2; k <= n_sort; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_sort; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], c2 = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < c2 : a > c2) {
                        keys[i] = c2;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < g.Q; j += blockDim.x) {
        const unsigned long long kk = keys[j];
        const unsigned flat = 0xFFFFFFFFu - (unsigned)(kk & 0xFFFFFFFFull);
        sel_score[(size_t)b * g.Q + j] = __uint_as_float((unsigned)(kk >> 32));
        sel_query[(size_t)b * g.Q + j] = (int)(flat / (unsigned)C);
        sel_label[(size_t)b * g.Q + j] = (int)(flat % (unsigned)C);
    }
}

// blockIdx.x < n_chunk_grid: statistics over a chunk of the 384x384 grid; the remaining blocks: any(mask) over a chunk of
// the target pixels
__global__ void __launch_bounds__(256) postproc_stats_kernel(const float* __restrict__ masks, PostGeom g,
                                                             const int* __restrict__ sel_query, int n_chunk_grid,
                                                             unsigned* __restrict__ part_cnt, float* __restrict__ part_sum,
                                                             unsigned* __restrict__ any_target, unsigned* __restrict__ bitmap) {
    const int j = blockIdx.y, b = blockIdx.z;
    const int q = sel_query[(size_t)b * g.Q + j];
    const float* plane = masks + ((size_t)b * g.Q + q) * g.h * g.w;
    __shared__ unsigned s_cnt[8];
    __shared__ float s_sum[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((int)blockIdx.x < n_chunk_grid) {
        unsigned cnt = 0;
        float sum = 0.f;
        const int base = blockIdx.x * kChunk;
        for (int i = threadIdx.x; i < kChunk; i += blockDim.x) {
            const int pt = base + i;                        // kGrid^2 is a multiple of kChunk: no partial chunk, warps stay whole
            const float v = grid_logit(plane, g, pt / kGrid, pt % kGrid);
            if (v > 0.f) {
                ++cnt;
                sum += 1.0f / (1.0f + expf(-v));
            }
            if (bitmap) {                                   // 32 consecutive grid points per warp -> one word of the sign bitmap
                const unsigned word = __ballot_sync(0xffffffffu, v > 0.f);
                if (lane == 0) bitmap[((size_t)b * g.Q + j) * (kGrid * (kGrid / 32)) + (pt >> 5)] = word;
            }
        }
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, k);
            sum += __shfl_xor_sync(0xffffffffu, sum, k);
        }
        if (lane == 0) {
            s_cnt[warp] = cnt;
            s_sum[warp] = sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned c = 0;
            float s = 0.f;
            for (int k = 0; k < 8; ++k) {
                c += s_cnt[k];
                s += s_sum[k];
            }
            const size_t o = ((size_t)b * g.Q + j) * n_chunk_grid + blockIdx.x;
            part_cnt[o] = c;
            part_sum[o] = s;
        }
    } else {
        const int base = ((int)blockIdx.x - n_chunk_grid) * kChunk;
        const int n_t = g.Ht * g.Wt;
        bool any = false;
        for (int i = threadIdx.x; i < kChunk; i += blockDim.x) {
            const int pt = base + i;
            if (pt >= n_t) break;
            const float v = grid_logit(plane, g, nearest_src(pt / g.Wt, g.sy), nearest_src(pt % g.Wt, g.sx));
            any |= v > 0.f;
        }
        if (__syncthreads_or(any ? 1 : 0) && threadIdx.x == 0) atomicOr(&any_target[(size_t)b * g.Q + j], 1u);
    }
}

// Grid pass, separable form (used when a band's source rows fit in shared memory).  ATen's bilinear value is
//   ly0 * (lx0 * v[y0][x0] + lx1 * v[y0][x1]) + ly1 * (lx0 * v[y1][x0] + lx1 * v[y1][x1]):
// the bracketed horizontal interpolations depend on (source row, grid column) only, so a CTA computes them ONCE for the source
// rows its band of 32 grid rows touches (12 rows for 120 -> 384) and every grid point costs two multiplies and an add -- the same
// operations in the same order as `grid_logit`, hence bit-identical.  Besides the (count, sigmoid sum) partials it writes the
// SIGN BITMAP of the grid (384 x 12 words per candidate), which the mask pass reads instead of re-evaluating the bilinear.
constexpr int kBand = 32;
constexpr int kBands = kGrid / kBand;
constexpr int kWordsPerRow = kGrid / 32;
__global__ void __launch_bounds__(256) postproc_grid_kernel(const float* __restrict__ masks, PostGeom g,
                                                            const int* __restrict__ sel_query, int rows_cap,
                                                            unsigned* __restrict__ part_cnt, float* __restrict__ part_sum,
                                                            unsigned* __restrict__ bitmap) {
    extern __shared__ float s_t[];                          // rows_cap x kGrid horizontal interpolations
    __shared__ int s_x0[kGrid], s_x1[kGrid];
    __shared__ float s_lx0[kGrid], s_lx1[kGrid];
    __shared__ unsigned s_cnt[8];
    __shared__ float s_sum[8];
    const int band = blockIdx.x, j = blockIdx.y, b = blockIdx.z;
    const int q = sel_query[(size_t)b * g.Q + j];
    const float* plane = masks + ((size_t)b * g.Q + q) * g.h * g.w;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int gx = threadIdx.x; gx < kGrid; gx += blockDim.x) {
        float fx = __fsub_rn(__fmul_rn(g.rw, __fadd_rn((float)gx, 0.5f)), 0.5f);
        fx = fx < 0.f ? 0.f : fx;
        const int x0 = (int)fx;
        s_x0[gx] = x0;
        s_x1[gx] = x0 + (x0 < g.w - 1 ? 1 : 0);
        const float lx1 = __fsub_rn(fx, (float)x0);
        s_lx1[gx] = lx1;
        s_lx0[gx] = __fsub_rn(1.f, lx1);
    }
    auto src_row = [&](int gy, int& y0, int& y1, float& ly0, float& ly1) {
        float fy = __fsub_rn(__fmul_rn(g.rh, __fadd_rn((float)gy, 0.5f)), 0.5f);
        fy = fy < 0.f ? 0.f : fy;
        y0 = (int)fy;
        y1 = y0 + (y0 < g.h - 1 ? 1 : 0);
        ly1 = __fsub_rn(fy, (float)y0);
        ly0 = __fsub_rn(1.f, ly1);
    };
    int y_lo, y_hi, t0, t1;
    float f0, f1;
    src_row(band * kBand, y_lo, t1, f0, f1);
    src_row(band * kBand + kBand - 1, t0, y_hi, f0, f1);
    const int rows = y_hi - y_lo + 1;                       // <= rows_cap (checked on the host)
    __syncthreads();
    for (int i = threadIdx.x; i < rows * kGrid; i += blockDim.x) {
        const int r = i / kGrid, gx = i - r * kGrid;
        const float* row = plane + (size_t)(y_lo + r) * g.w;
        s_t[i] = __fadd_rn(__fmul_rn(s_lx0[gx], __ldg(row + s_x0[gx])), __fmul_rn(s_lx1[gx], __ldg(row + s_x1[gx])));
    }
    __syncthreads();
    unsigned cnt = 0;
    float sum = 0.f;
    unsigned* bm = bitmap + ((size_t)b * g.Q + j) * (kGrid * kWordsPerRow);
    for (int gl = warp; gl < kBand; gl += 8) {
        const int gy = band * kBand + gl;
        int y0, y1;
        float ly0, ly1;
        src_row(gy, y0, y1, ly0, ly1);
        const float* top = s_t + (y0 - y_lo) * kGrid;
        const float* bot = s_t + (y1 - y_lo) * kGrid;
#pragma unroll 4
        for (int k = 0; k < kWordsPerRow; ++k) {
            const int gx = k * 32 + lane;
            const float v = __fadd_rn(__fmul_rn(ly0, top[gx]), __fmul_rn(ly1, bot[gx]));
            const bool pos = v > 0.f;
            if (pos) {
                ++cnt;
                sum += 1.0f / (1.0f + expf(-v));
            }
            const unsigned word = __ballot_sync(0xffffffffu, pos);
            if (lane == 0) bm[gy * kWordsPerRow + k] = word;
        }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, k);
        sum += __shfl_xor_sync(0xffffffffu, sum, k);
    }
    if (lane == 0) {
        s_cnt[warp] = cnt;
        s_sum[warp] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned c = 0;
        float sm = 0.f;
        for (int k = 0; k < 8; ++k) {
            c += s_cnt[k];
            sm += s_sum[k];
        }
        const size_t o = ((size_t)b * g.Q + j) * kBands + band;
        part_cnt[o] = c;
        part_sum[o] = sm;
    }
}

__global__ void __launch_bounds__(256) postproc_finalize_kernel(PostGeom g, float threshold, int n_chunk_grid,
                                                                const float* __restrict__ sel_score,
                                                                const int* __restrict__ sel_query,
                                                                const int* __restrict__ sel_label,
                                                                const unsigned* __restrict__ part_cnt,
                                                                const float* __restrict__ part_sum,
                                                                const unsigned* __restrict__ any_target, int* __restrict__ slot,
                                                                int* __restrict__ inv_slot, int* __restrict__ out_labels, float* __restrict__ out_scores,
                                                                int* __restrict__ out_query, int* __restrict__ out_count) {
    extern __shared__ float s_pred[];           // Q predicted scores, then Q keep flags
    int* s_keep = reinterpret_cast<int*>(s_pred + g.Q);
    const int b = blockIdx.x;
    for (int j = threadIdx.x; j < g.Q; j += blockDim.x) {
        const size_t o = ((size_t)b * g.Q + j) * n_chunk_grid;
        unsigned cnt = 0;
        float sum = 0.f;
        for (int k = 0; k < n_chunk_grid; ++k) {   // fixed order: deterministic
            cnt += part_cnt[o + k];
            sum += part_sum[o + k];
        }
        const float mask_score = sum / ((float)cnt + 1e-6f);
        const float pred = sel_score[(size_t)b * g.Q + j] * mask_score;
        s_pred[j] = pred;
        // nearest UPsampling visits every grid point, so the target mask is empty iff the grid mask is
        const bool nonempty = (g.Ht >= kGrid && g.Wt >= kGrid) ? cnt != 0u : any_target[(size_t)b * g.Q + j] != 0u;
        s_keep[j] = (nonempty && pred >= threshold) ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int j = 0; j < g.Q; ++j) {
            if (s_keep[j]) {
                slot[(size_t)b * g.Q + j] = n;
                inv_slot[(size_t)b * g.Q + n] = j;
                out_labels[(size_t)b * g.Q + n] = sel_label[(size_t)b * g.Q + j];
                out_scores[(size_t)b * g.Q + n] = s_pred[j];
                out_query[(size_t)b * g.Q + n] = sel_query[(size_t)b * g.Q + j];
                ++n;
            } else {
                slot[(size_t)b * g.Q + j] = -1;
            }
        }
        for (int k = n; k < g.Q; ++k) {
            out_labels[(size_t)b * g.Q + k] = -1;
            out_scores[(size_t)b * g.Q + k] = 0.f;
            out_query[(size_t)b * g.Q + k] = -1;
        }
        out_count[b] = n;
    }
}

// One thread = four consecutive target pixels of one image, for ALL kept segments: the pixel -> grid-point mapping is computed
// once, every segment costs a few bit extractions from its sign bitmap (L1/L2-resident: 18 KB per candidate) and one 4-byte
// store; the painted map keeps the highest slot whose bit is set ("last kept segment wins") in a register -- no atomics, no
// fill pass.  (Per-segment CTAs with an atomicMax per positive pixel took 3/4 of the post-processing time.)
__global__ void __launch_bounds__(256) postproc_paint_kernel(PostGeom g, const int* __restrict__ inv_slot,
                                                             const int* __restrict__ out_count,
                                                             const unsigned* __restrict__ bitmap, uint8_t* __restrict__ out_masks,
                                                             int* __restrict__ seg) {
    const int b = blockIdx.y;
    const int n_t = g.Ht * g.Wt;
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= n_t) return;
    const int n = n_t - p0 < 4 ? n_t - p0 : 4;
    int wi[4], sh[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int pt = p0 + (e < n ? e : 0);
        const int gy = nearest_src(pt / g.Wt, g.sy), gx = nearest_src(pt % g.Wt, g.sx);
        wi[e] = gy * (kGrid / 32) + (gx >> 5);
        sh[e] = gx & 31;
    }
    int painted[4] = {-1, -1, -1, -1};
    const int count = out_count[b];
    const bool vec = n == 4 && (n_t & 3) == 0;
    for (int sl = 0; sl < count; ++sl) {
        const unsigned* bm = bitmap + ((size_t)b * g.Q + inv_slot[(size_t)b * g.Q + sl]) * (kGrid * (kGrid / 32));
        unsigned packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned bit = (__ldg(bm + wi[e]) >> sh[e]) & 1u;
            packed |= bit << (8 * e);
            if (bit) painted[e] = sl;
        }
        uint8_t* dst = out_masks + ((size_t)b * g.Q + sl) * n_t + p0;
        if (vec) {
            *reinterpret_cast<unsigned*>(dst) = packed;
        } else {
            for (int e = 0; e < n; ++e) dst[e] = (uint8_t)((packed >> (8 * e)) & 1u);
        }
    }
    if (seg) {
        int* sp = seg + (size_t)b * n_t + p0;
        if (vec) {
            *reinterpret_cast<int4*>(sp) = make_int4(painted[0], painted[1], painted[2], painted[3]);
        } else {
            for (int e = 0; e < n; ++e) sp[e] = painted[e];
        }
    }
}

__global__ void __launch_bounds__(256) fill_int_kernel(int* __restrict__ p, size_t n, int v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// one CTA per (pred, gt) pair; masks are bytes holding 0 / 1
__global__ void __launch_bounds__(256) mask_iou_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                       long long n, float* __restrict__ iou, int G) {
    const int p = blockIdx.y, gi = blockIdx.x;
    const uint8_t* a = pred + (size_t)p * n;
    const uint8_t* c = gt + (size_t)gi * n;
    unsigned inter = 0, na = 0, nc = 0;
    const bool vec = (n % 16 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(c)) % 16 == 0);
    if (vec) {
        const uint4* a4 = reinterpret_cast<const uint4*>(a);
        const uint4* c4 = reinterpret_cast<const uint4*>(c);
        for (long long i = threadIdx.x; i < n / 16; i += blockDim.x) {
            const uint4 x = __ldg(a4 + i), y = __ldg(c4 + i);
            inter += __popc(x.x & y.x & 0x01010101u) + __popc(x.y & y.y & 0x01010101u) + __popc(x.z & y.z & 0x01010101u) +
                     __popc(x.w & y.w & 0x01010101u);
            na += __popc(x.x & 0x01010101u) + __popc(x.y & 0x01010101u) + __popc(x.z & 0x01010101u) + __popc(x.w & 0x01010101u);
            nc += __popc(y.x & 0x01010101u) + __popc(y.y & 0x01010101u) + __popc(y.z & 0x01010101u) + __popc(y.w & 0x01010101u);
        }
    } else {
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned x = a[i] & 1u, y = c[i] & 1u;
            inter += x & y;
            na += x;
            nc += y;
        }
    }
    __shared__ unsigned s[3][8];
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        inter += __shfl_xor_sync(0xffffffffu, inter, k);
        na += __shfl_xor_sync(0xffffffffu, na, k);
        nc += __shfl_xor_sync(0xffffffffu, nc, k);
    }
    if ((threadIdx.x & 31) == 0) {
        s[0][threadIdx.x >> 5] = inter;
        s[1][threadIdx.x >> 5] = na;
        s[2][threadIdx.x >> 5] = nc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned i2 = 0, a2 = 0, c2 = 0;
        for (int k = 0; k < 8; ++k) {
            i2 += s[0][k];
            a2 += s[1][k];
            c2 += s[2][k];
        }
        const unsigned uni = a2 + c2 - i2;
        iou[(size_t)p * G + gi] = uni ? (float)((double)i2 / (double)uni) : 0.f;
    }
}

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

struct WsLayout {
    size_t sel_score, sel_query, sel_label, part_cnt, part_sum, any_target, slot, inv_slot, bitmap, total;
};

WsLayout ws_layout(int B, int Q) {
    const int n_chunk = ceil_div(kGrid * kGrid, kChunk);
    WsLayout l;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += (bytes + 255) / 256 * 256;
        return at;
    };
    l.sel_score = take((size_t)B * Q * 4);
    l.sel_query = take((size_t)B * Q * 4);
    l.sel_label = take((size_t)B * Q * 4);
    l.part_cnt = take((size_t)B * Q * n_chunk * 4);
    l.part_sum = take((size_t)B * Q * n_chunk * 4);
    l.any_target = take((size_t)B * Q * 4);
    l.slot = take((size_t)B * Q * 4);
    l.inv_slot = take((size_t)B * Q * 4);
    l.bitmap = take((size_t)B * Q * kGrid * (kGrid / 32) * 4);
    l.total = o;
    return l;
}

}  // namespace

extern "C" size_t rgbd_postprocess_workspace_bytes(int B, int Q) {
    if (B < 1 || Q < 1) return 0;
    return ws_layout(B, Q).total;
}

extern "C" int rgbd_postprocess_instances(const float* class_logits, const float* mask_logits, int B, int Q, int C1, int h,
                                          int w, int Ht, int Wt, float threshold, void* workspace, uint8_t* out_masks,
                                          int* out_labels, float* out_scores, int* out_query, int* out_count,
                                          int* out_segmentation, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(class_logits && mask_logits && workspace && out_masks && out_labels && out_scores && out_query && out_count,
                   "postprocess: null pointer");
    RGBD_CHECK_ARG(B >= 1 && Q >= 1 && C1 >= 2 && h >= 1 && w >= 1 && Ht >= 1 && Wt >= 1, "postprocess: bad geometry");
    const int n_cand = Q * (C1 - 1);
    RGBD_CHECK_ARG(n_cand <= kMaxSort, "postprocess: Q*(C) = %d candidates exceed the in-shared-memory sort (%d)", n_cand, kMaxSort);
    RGBD_CHECK_ARG(Q <= 4096, "postprocess: at most 4096 queries");
    cudaStream_t s = (cudaStream_t)stream;
    PostGeom g;
    g.B = B; g.Q = Q; g.C1 = C1; g.h = h; g.w = w; g.Ht = Ht; g.Wt = Wt;
    g.rh = (float)h / (float)kGrid; g.rw = (float)w / (float)kGrid;
    g.sy = (float)kGrid / (float)Ht; g.sx = (float)kGrid / (float)Wt;
    const WsLayout l = ws_layout(B, Q);
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    float* sel_score = reinterpret_cast<float*>(ws + l.sel_score);
    int* sel_query = reinterpret_cast<int*>(ws + l.sel_query);
    int* sel_label = reinterpret_cast<int*>(ws + l.sel_label);
    unsigned* part_cnt = reinterpret_cast<unsigned*>(ws + l.part_cnt);
    float* part_sum = reinterpret_cast<float*>(ws + l.part_sum);
    unsigned* any_target = reinterpret_cast<unsigned*>(ws + l.any_target);
    int* slot = reinterpret_cast<int*>(ws + l.slot);
    int* inv_slot = reinterpret_cast<int*>(ws + l.inv_slot);
    unsigned* bitmap = reinterpret_cast<unsigned*>(ws + l.bitmap);

    const int n_sort = next_pow2(n_cand < Q ? Q : n_cand);
    RgbdDeviceInfo di;
    if (int rc = rgbd_device_info(&di)) return rc;
    RGBD_ONCE_PER_DEVICE(di.device, {
        RGBD_CHECK_CUDA(cudaFuncSetAttribute(postproc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSort * 8));
    });
    postproc_select_kernel<<<B, 256, (size_t)n_sort * 8, s>>>(class_logits, g, n_sort, sel_score, sel_query, sel_label);
    RGBD_CHECK_LAUNCH();
    RGBD_CHECK_CUDA(cudaMemsetAsync(any_target, 0, (size_t)B * Q * 4, s));
    int n_chunk_grid = ceil_div(kGrid * kGrid, kChunk);
    const int n_chunk_t = (Ht >= kGrid && Wt >= kGrid) ? 0 : ceil_div(Ht * Wt, kChunk);
    // source rows one band of grid rows can touch: floor(rh * 31) + 3 bounds y1(last) - y0(first) + 1
    const int rows_cap = (int)((double)g.rh * (kBand - 1)) + 3;
    const size_t band_smem = (size_t)rows_cap * kGrid * sizeof(float);
    if (band_smem <= 64 * 1024) {
        RGBD_ONCE_PER_DEVICE(di.device, {
            RGBD_CHECK_CUDA(cudaFuncSetAttribute(postproc_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        });
        postproc_grid_kernel<<<dim3(kBands, Q, B), 256, band_smem, s>>>(mask_logits, g, sel_query, rows_cap, part_cnt, part_sum, bitmap);
        RGBD_CHECK_LAUNCH();
        if (n_chunk_t) {          // downsampled targets: any(mask) over the target pixels (the grid chunks of this launch are empty)
            postproc_stats_kernel<<<dim3(n_chunk_t, Q, B), 256, 0, s>>>(mask_logits, g, sel_query, 0, part_cnt, part_sum, any_target,
                                                                       nullptr);
            RGBD_CHECK_LAUNCH();
        }
        n_chunk_grid = kBands;
    } else {                      // very tall logit planes: point-wise evaluation
        postproc_stats_kernel<<<dim3(n_chunk_grid + n_chunk_t, Q, B), 256, 0, s>>>(mask_logits, g, sel_query, n_chunk_grid, part_cnt,
                                                                                   part_sum, any_target, bitmap);
        RGBD_CHECK_LAUNCH();
    }
    postproc_finalize_kernel<<<B, 256, (size_t)Q * 8, s>>>(g, threshold, n_chunk_grid, sel_score, sel_query, sel_label, part_cnt,
                                                           part_sum, any_target, slot, inv_slot, out_labels, out_scores, out_query,
                                                           out_count);
    RGBD_CHECK_LAUNCH();
    RGBD_CHECK_ARG(B <= 65535, "postprocess: at most 65535 images per call");
    postproc_paint_kernel<<<dim3(ceil_div(Ht * Wt, 256 * 4), B), 256, 0, s>>>(g, inv_slot, out_count, bitmap, out_masks, out_segmentation);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_mask_iou(const uint8_t* pred_masks, const uint8_t* gt_masks, int P, int G, long long pixels, float* iou,
                             rgbd_stream_t stream) {
    RGBD_CHECK_ARG(iou && pixels >= 1 && P >= 0 && G >= 0, "mask_iou: bad arguments");
    if (P == 0 || G == 0) return RGBD_OK;
    RGBD_CHECK_ARG(pred_masks && gt_masks, "mask_iou: null pointer");
    RGBD_CHECK_ARG(P <= 65535, "mask_iou: at most 65535 predictions per call");
    mask_iou_kernel<<<dim3(G, P), 256, 0, (cudaStream_t)stream>>>(pred_masks, gt_masks, pixels, iou, G);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
