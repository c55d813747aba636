// K2: E-DSAM depth decomposition on device, bit-exact with the reference's host numpy/scipy path
// (mask2former/utils/custom_model.py):
//   to_grayscale                         CM:466-480   ((0.299f*r + 0.587f*g) + 0.114f*b, no FMA)
//   _calculate_depth_histogram           CM:701-718   np.histogram(bins=512, range=(nanmin,nanmax))
//   _select_depth_distribution_modes     CM:720-752   scipy find_peaks(prominence=0.01*max) + top-3
//   _define_depth_interval_windows       CM:754-772   float32 window arithmetic (NumPy-2 promotion)
//   _generate_depth_region_masks         CM:774-798   inclusive interval masks + complement of union
//   adaptive_max_pool2d of each mask     CM:687       to every feature resolution that needs it
// The reference does this per image per DSAM stage on the host (3 D2H copies, 12 H2D copies and
// 6 stream syncs per image per forward); here one batch-wide, sync-free sequence of small kernels
// produces, for every image, a 4-bit region code per pixel (bit t = region mask t in the
// reference's list order) at full resolution and max-pooled to each requested resolution.
// Compiled with -fmad=false: every float op is a single IEEE-rounded operation.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

constexpr int kBins = RGBD_HIST_BINS;   // 512

struct ImgState {
    uint32_t min_inv, max_enc;      // order-preserving encodings of nanmax and (bitwise inverted, so that an all-zero state is
                                    // the identity of both atomicMax reductions and a memset initialises it) of nanmin
    uint32_t n_finite, has_inf;
    float first, last, step, denom; // histogram range after the first==last widening
    int step_zero;                  // numpy's denormal special case (gh-5437)
    int n_modes;                    // surviving modes (0..3)
    int status;                     // RGBD_DECOMP_* flags
    int pad;
    float lo[3], hi[3];             // interval windows in mode order
    float centre[3];
    int peak_bin[3];
};

// block-wide min/max/count reduction of the per-thread partials, then ONE set of atomics per CTA
__device__ __forceinline__ void reduce_minmax(ImgState* st, uint32_t lmin, uint32_t lmax, uint32_t nfin, uint32_t ninf) {
    __shared__ uint32_t red[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        nfin += __shfl_xor_sync(0xffffffffu, nfin, o);
        ninf |= __shfl_xor_sync(0xffffffffu, ninf, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = lmin; red[1][w] = lmax; red[2][w] = nfin; red[3][w] = ninf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; ++k) {
            lmin = min(lmin, red[0][k]); lmax = max(lmax, red[1][k]); nfin += red[2][k]; ninf |= red[3][k];
        }
        if (nfin) {
            atomicMax(&st->min_inv, ~lmin);
            atomicMax(&st->max_enc, lmax);
            atomicAdd(&st->n_finite, nfin);
            if (ninf) atomicOr(&st->has_inf, 1u);
        }
    }
}

// gray + per-image nanmin/nanmax; 4 pixels per thread (128-bit loads / stores when VEC)
template <bool VEC>
__global__ void __launch_bounds__(256) decomp_gray_kernel(const float* __restrict__ depth3, long long bs, long long cs,
                                                          float* __restrict__ gray, ImgState* __restrict__ st, int HW) {
    const int b = blockIdx.y;
    const float* r = depth3 + (long long)b * bs;
    const float* g = r + cs;
    const float* bl = g + cs;
    float* out = gray + (long long)b * HW;
    uint32_t lmin = 0xffffffffu, lmax = 0u, nfin = 0u, ninf = 0u;
    for (int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i4 < HW; i4 += gridDim.x * blockDim.x * 4) {
        float rv[4], gv[4], bv[4], v[4];
        const int n = HW - i4 < 4 ? HW - i4 : 4;
        if (VEC) {
            const float4 a = *reinterpret_cast<const float4*>(r + i4), c = *reinterpret_cast<const float4*>(g + i4),
                         d = *reinterpret_cast<const float4*>(bl + i4);
            rv[0] = a.x; rv[1] = a.y; rv[2] = a.z; rv[3] = a.w;
            gv[0] = c.x; gv[1] = c.y; gv[2] = c.z; gv[3] = c.w;
            bv[0] = d.x; bv[1] = d.y; bv[2] = d.z; bv[3] = d.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                rv[e] = e < n ? r[i4 + e] : 0.f;
                gv[e] = e < n ? g[i4 + e] : 0.f;
                bv[e] = e < n ? bl[i4 + e] : 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = (0.299f * rv[e] + 0.587f * gv[e]) + 0.114f * bv[e];
            if (e < n && !isnan(v[e])) {
                uint32_t enc = f32_to_ordered(v[e]);
                lmin = min(lmin, enc);
                lmax = max(lmax, enc);
                nfin++;
                if (isinf(v[e])) ninf = 1u;
            }
        }
        if (VEC) {
            *reinterpret_cast<float4*>(out + i4) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int e = 0; e < n; ++e) out[i4 + e] = v[e];
        }
    }
    reduce_minmax(st + b, lmin, lmax, nfin, ninf);
}

// gray supplied by the caller (DSAModule.forward API): only the min/max reduction
__global__ void __launch_bounds__(256) decomp_minmax_kernel(const float* __restrict__ gray, ImgState* __restrict__ st, int HW) {
    const int b = blockIdx.y;
    const float* in = gray + (long long)b * HW;
    uint32_t lmin = 0xffffffffu, lmax = 0u, nfin = 0u, ninf = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float v = in[i];
        if (!isnan(v)) {
            uint32_t e = f32_to_ordered(v);
            lmin = min(lmin, e);
            lmax = max(lmax, e);
            nfin++;
            if (isinf(v)) ninf = 1u;
        }
    }
    reduce_minmax(st + b, lmin, lmax, nfin, ninf);
}

// histogram range of one image from its nanmin / nanmax (numpy _get_outer_edges + linspace step); every kernel that needs it
// recomputes it from the reduced state (deterministic), decomp_modes_kernel stores it
__device__ __forceinline__ void compute_range(ImgState& s) {
    if (s.n_finite == 0 || s.has_inf) {   // numpy raises "range ... is not finite"
        s.status |= RGBD_DECOMP_RANGE_NOT_FINITE;
        s.first = 0.f; s.last = 1.f;
    } else {
        s.first = ordered_to_f32(~s.min_inv);
        s.last = ordered_to_f32(s.max_enc);
        if (s.first == s.last) {            // _get_outer_edges: expand empty range
            s.first = s.first - 0.5f;
            s.last = s.last + 0.5f;
        }
    }
    s.denom = s.last - s.first;
    s.step = s.denom / (float)kBins;       // linspace: step = delta / div
    s.step_zero = (s.step == 0.0f);
}

// np.linspace(first, last, 513, dtype=float32)[i]
__device__ __forceinline__ float bin_edge(const ImgState& s, int i) {
    if (i == kBins) return s.last;
    float y = s.step_zero ? ((float)i / (float)kBins) * s.denom : (float)i * s.step;
    return y + s.first;
}

// the binning loop of one CTA.  STEP_ZERO (numpy's denormal-range special case, uniform per image) is a template parameter:
// as a run-time select inside bin_edge it costs two extra IEEE divisions per pixel (the kernel is issue-bound: ncu).
template <bool STEP_ZERO>
__device__ __forceinline__ void hist_accumulate(const float* __restrict__ in, int HW, const ImgState& s, unsigned int* hw) {
    auto edge = [&](int i) -> float {
        const float y = STEP_ZERO ? ((float)i / (float)kBins) * s.denom : (float)i * s.step;
        return i == kBins ? s.last : y + s.first;
    };
    // numpy's uniform-bin fast path, branch-free (the kernel is issue-bound and five divergent regions per pixel cost more than
    // the arithmetic): candidate bin by truncation, then the +-1 corrections against the float32 edges.  Equivalent to
    //   if (idx == 512) --idx;  if (x < edge(idx)) --idx;  if (x >= edge(idx + 1) && idx != 511) ++idx;
    // because after a decrement the third test compares x with the very edge it was just found to lie below.
    auto bin_of = [&](float x) -> int {
        const bool ok = x >= s.first && x <= s.last;          // drops NaN
        const float f = ((x - s.first) / s.denom) * (float)kBins;
        int idx = (int)f;                                      // astype(intp): truncation (garbage when !ok: clamped, discarded)
        idx = min(max(idx, 0), kBins - 1);
        const float e0 = edge(idx), e1 = edge(idx + 1);
        const bool below = x < e0;
        idx -= below ? 1 : 0;
        idx += (!below && x >= e1 && idx != kBins - 1) ? 1 : 0;
        return ok ? idx : -1;
    };
    // neighbouring pixels of a depth image mostly share a bin (8-bit depth: 256 distinct values): the lanes of a warp that hit
    // the same bin elect one leader that adds their count -- one shared-memory atomic per distinct bin instead of a
    // serialised 32-way conflict.  Four pixels per thread and iteration (one 128-bit load) with the next load issued before
    // the current one is binned.
    auto add = [&](int idx) {
        const unsigned peers = __match_any_sync(0xffffffffu, idx);
        if (idx >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hw[idx], (unsigned)__popc(peers));
    };
    const float nan = __int_as_float(0x7fc00000);
    const int stride = gridDim.x * blockDim.x;
    const bool vec = (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (vec) {
        const int n4 = HW >> 2;
        const int n_iter = (n4 + stride - 1) / stride;
        const float4 nan4 = make_float4(nan, nan, nan, nan);
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        float4 cur = i < n4 ? __ldg(reinterpret_cast<const float4*>(in) + i) : nan4;
        for (int it = 0; it < n_iter; ++it) {
            const int nx = i + stride;
            const float4 next = (it + 1 < n_iter && nx < n4) ? __ldg(reinterpret_cast<const float4*>(in) + nx) : nan4;
            add(bin_of(cur.x)); add(bin_of(cur.y)); add(bin_of(cur.z)); add(bin_of(cur.w));
            cur = next;
            i = nx;
        }
    } else {
        const int n_iter = (HW + stride - 1) / stride;
        for (int it = 0; it < n_iter; ++it) {
            const int i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
            add(bin_of(i < HW ? in[i] : nan));
        }
    }
}

__global__ void __launch_bounds__(256) decomp_hist_kernel(const float* __restrict__ gray, const ImgState* __restrict__ st,
                                                          unsigned long long* __restrict__ hist, int HW) {
    __shared__ unsigned int h[8][kBins];            // one sub-histogram per warp: no atomics contention between warps
    const int b = blockIdx.y;
    ImgState s = st[b];
    compute_range(s);
    for (int i = threadIdx.x; i < 8 * kBins; i += blockDim.x) (&h[0][0])[i] = 0u;
    __syncthreads();
    if (s.status & RGBD_DECOMP_RANGE_NOT_FINITE) return;
    const float* in = gray + (long long)b * HW;
    unsigned int* hw = h[threadIdx.x >> 5];
    if (s.step_zero) hist_accumulate<true>(in, HW, s, hw);
    else hist_accumulate<false>(in, HW, s, hw);
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
        unsigned int t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += h[w][i];
        if (t) atomicAdd(&hist[(long long)b * kBins + i], (unsigned long long)t);
    }
}

struct ExportPtrs {
    int* n_modes; int* peak_bins; float* centres; float* windows; int* status; int* bias_variant;
};

// One CTA (512 threads, one per bin) per image: scipy _local_maxima_1d + _peak_prominences(wlen=None),
// keep prominence >= 0.01*max (float64), order by (height, centre) descending, first 3; then windows.
// Parallel form: (1) every rising edge starts at most one plateau -> peak list; (2) one WARP per peak scans outwards 32 bins
// at a time for the first higher bin (ballot) and the minimum on the way (shuffle-min) -- uint8 depth gives ~200 peaks whose
// walks are hundreds of bins long, serial per thread before; (3) three block-wide arg-max rounds over (height, bin).
// edges_in != nullptr (helper API, CM:720-752): bin edges (B, 513) supplied by the caller instead of the image's range;
// ratio == nullptr: no windows (they belong to _define_depth_interval_windows).  range_from_minmax: the state holds only the
// reduced nanmin / nanmax (batched path); the histogram range is derived and stored here.
__global__ void __launch_bounds__(kBins) decomp_modes_kernel(const unsigned long long* __restrict__ hist,
                                                             ImgState* __restrict__ st, const float* __restrict__ ratio,
                                                             int num_modes, double prom_thr,
                                                             const float* __restrict__ edges_in, int range_from_minmax,
                                                             ExportPtrs ex) {
    __shared__ long long h[kBins];
    __shared__ int peak_bin[kBins / 2];
    __shared__ unsigned long long cand[kBins / 2];     // (height << 16) | bin of the peaks that pass the prominence filter
    __shared__ int n_peaks, n_cand;
    __shared__ unsigned long long red[kBins / 32];
    __shared__ unsigned long long picked[3];
    const int b = blockIdx.x;
    const int i = threadIdx.x;
    const int lane = i & 31, warp = i >> 5;
    ImgState& s = st[b];
    h[i] = (long long)hist[(long long)b * kBins + i];
    if (i == 0) { n_peaks = 0; n_cand = 0; }
    __syncthreads();
    // block max
    long long m = h[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if (lane == 0) red[warp] = (unsigned long long)m;
    __syncthreads();
    long long hmax = 0;
    for (int k = 0; k < kBins / 32; ++k) hmax = (long long)red[k] > hmax ? (long long)red[k] : hmax;
    const double pmin = prom_thr * (double)hmax;

    // (1) every rising edge starts at most one plateau -> independent per thread
    const int n = kBins, i_max = n - 1;
    if (i >= 1 && i < i_max && h[i - 1] < h[i]) {
        int ahead = i + 1;
        while (ahead < i_max && h[ahead] == h[i]) ++ahead;
        if (h[ahead] < h[i]) peak_bin[atomicAdd(&n_peaks, 1)] = (i + ahead - 1) / 2;
    }
    __syncthreads();
    // (2) prominence of peak p by warp (p mod 16)
    const int np = n_peaks;
    for (int pi = warp; pi < np; pi += kBins / 32) {
        const int peak = peak_bin[pi];
        const long long hp = h[peak];
        long long lmin = hp, rmin = hp;
        for (int base = peak; ; base -= 32) {            // bins peak, peak-1, ...: stop at the first bin higher than the peak
            const int k = base - lane;
            const long long v = k >= 0 ? h[k] : 0;
            const bool stop = k < 0 || v > hp;
            const unsigned sm = __ballot_sync(0xffffffffu, stop);
            const int first = sm ? __ffs(sm) - 1 : 32;
            long long mine = lane < first ? v : hp;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const long long other = __shfl_xor_sync(0xffffffffu, mine, o);
                mine = other < mine ? other : mine;
            }
            lmin = mine < lmin ? mine : lmin;
            if (sm) break;
        }
        for (int base = peak; ; base += 32) {
            const int k = base + lane;
            const long long v = k <= i_max ? h[k] : 0;
            const bool stop = k > i_max || v > hp;
            const unsigned sm = __ballot_sync(0xffffffffu, stop);
            const int first = sm ? __ffs(sm) - 1 : 32;
            long long mine = lane < first ? v : hp;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const long long other = __shfl_xor_sync(0xffffffffu, mine, o);
                mine = other < mine ? other : mine;
            }
            rmin = mine < rmin ? mine : rmin;
            if (sm) break;
        }
        const long long prom = hp - (lmin > rmin ? lmin : rmin);
        if (lane == 0 && pmin <= (double)prom) cand[atomicAdd(&n_cand, 1)] = ((unsigned long long)hp << 16) | (unsigned)peak;
    }
    __syncthreads();
    // (3) selection by (height desc, centre desc); centres are non-decreasing in the bin index, so (height, bin) descending
    // gives the same centres in the same order.  Keys are unique (one per bin); 0 never is a key (peaks sit at bins >= 1).
    const int nc = n_cand;
    unsigned long long key = i < nc ? cand[i] : 0ull;
    for (int pick = 0; pick < 3; ++pick) {
        unsigned long long best = key;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) red[warp] = best;
        __syncthreads();
        if (i == 0) {
            unsigned long long t = 0;
            for (int k = 0; k < kBins / 32; ++k) t = red[k] > t ? red[k] : t;
            picked[pick] = t;
        }
        __syncthreads();
        if (key == picked[pick]) key = 0ull;
    }
    if (i == 0) {
        ImgState loc = s;                 // one read of the per-image state; the serial part below works on registers
        if (range_from_minmax) compute_range(loc);
        const float rt = ratio ? ratio[b] : 0.f;
        int nm = 0;
        if (!(loc.status & RGBD_DECOMP_RANGE_NOT_FINITE)) {
            for (int pick = 0; pick < num_modes && pick < nc; ++pick) {
                const int pb = (int)(picked[pick] & 0xffffu);
                float e0, e1;
                if (edges_in) {
                    e0 = edges_in[(long long)b * (kBins + 1) + pb];
                    e1 = edges_in[(long long)b * (kBins + 1) + pb + 1];
                } else {
                    e0 = bin_edge(loc, pb);
                    e1 = bin_edge(loc, pb + 1);
                }
                float centre = e0 + (e1 - e0) / 2.0f;
                float hw = (centre * rt) / 2.0f;
                float lo = centre - hw;
                lo = lo > 0.0f ? lo : 0.0f;          // python max(0, x)
                loc.peak_bin[nm] = pb;
                loc.centre[nm] = centre;
                loc.lo[nm] = lo;
                loc.hi[nm] = centre + hw;
                ++nm;
            }
        }
        loc.n_modes = nm;
        for (int k = nm; k < 3; ++k) { loc.peak_bin[k] = -1; loc.centre[k] = 0.f; loc.lo[k] = 0.f; loc.hi[k] = 0.f; }
        s = loc;
        if (ex.n_modes) ex.n_modes[b] = nm;
        if (ex.status) ex.status[b] = loc.status;
        // number of conv biases the reference adds: m+1 used regions, or all R+1 when no mode survives (CM:676-691)
        if (ex.bias_variant) ex.bias_variant[b] = nm == 0 ? num_modes + 1 : nm + 1;
        for (int k = 0; k < 3; ++k) {
            if (ex.peak_bins) ex.peak_bins[b * 3 + k] = loc.peak_bin[k];
            if (ex.centres) ex.centres[b * 3 + k] = loc.centre[k];
            if (ex.windows) { ex.windows[(b * 3 + k) * 2] = loc.lo[k]; ex.windows[(b * 3 + k) * 2 + 1] = loc.hi[k]; }
        }
    }
}

// region code of one pixel; the windows are passed in registers (indexing ImgState's arrays with a run-time t would put the
// struct in local memory)
struct Windows { float lo0, hi0, lo1, hi1, lo2, hi2; int n; };
__device__ __forceinline__ Windows load_windows(const ImgState& s) {
    Windows w = {s.lo[0], s.hi[0], s.lo[1], s.hi[1], s.lo[2], s.hi[2], s.n_modes};
    return w;
}
__device__ __forceinline__ unsigned region_code(const Windows& w, float g) {
    unsigned c = 0;
    if (w.n > 0 && g >= w.lo0 && g <= w.hi0) c |= 1u;
    if (w.n > 1 && g >= w.lo1 && g <= w.hi1) c |= 2u;
    if (w.n > 2 && g >= w.lo2 && g <= w.hi2) c |= 4u;
    if (w.n > 0 && c == 0) c = 1u << w.n;
    return c;
}

// Region codes AND their OR-pooled copies at the three pyramid levels in one pass, for the usual geometry where the levels are
// exactly (H/4, W/4), (H/8, W/8), (H/16, W/16) (Swin strides 4/8/16; then adaptive_max_pool2d's windows are the aligned 4x4 /
// 8x8 / 16x16 blocks).  A warp owns a 32x16-pixel block: lane (lx 0..7, ly 0..3) reads its 4x4 cell as four float4, ORs its
// 16 codes (level 0), and the coarser levels are butterfly ORs over the lane bits (lx bit0, ly bit0 | lx bit1, ly bit1).
__global__ void __launch_bounds__(256) decomp_codes_pool3_kernel(const float* __restrict__ gray, const ImgState* __restrict__ st,
                                                                 uint8_t* __restrict__ codes, uint8_t* __restrict__ p0,
                                                                 uint8_t* __restrict__ p1, uint8_t* __restrict__ p2, int H, int W) {
    const int b = blockIdx.y;
    const Windows s = load_windows(st[b]);
    const int lane = threadIdx.x & 31, lx = lane & 7, ly = lane >> 3;
    const int bw = (W + 31) >> 5, bh = H >> 4;                     // warp blocks per image
    const float* in = gray + (long long)b * H * W;
    for (int wb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wb < bw * bh; wb += gridDim.x * (blockDim.x >> 5)) {
        const int x = (wb % bw) * 32 + lx * 4, y = (wb / bw) * 16 + ly * 4;
        const bool on = x < W;                                      // W % 16 == 0: a lane's cell is entirely in or out
        unsigned c0 = 0;
        if (on) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 g = *reinterpret_cast<const float4*>(in + (long long)(y + r) * W + x);
                const unsigned q0 = region_code(s, g.x), q1 = region_code(s, g.y), q2 = region_code(s, g.z), q3 = region_code(s, g.w);
                if (codes) *reinterpret_cast<unsigned*>(codes + ((long long)b * H + y + r) * W + x) = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
                c0 |= q0 | q1 | q2 | q3;
            }
            p0[((long long)b * (H >> 2) + (y >> 2)) * (W >> 2) + (x >> 2)] = (uint8_t)c0;
        }
        unsigned c1 = c0 | __shfl_xor_sync(0xffffffffu, c0, 1);
        c1 |= __shfl_xor_sync(0xffffffffu, c1, 8);
        if (on && !(lx & 1) && !(ly & 1)) p1[((long long)b * (H >> 3) + (y >> 3)) * (W >> 3) + (x >> 3)] = (uint8_t)c1;
        unsigned c2 = c1 | __shfl_xor_sync(0xffffffffu, c1, 2);
        c2 |= __shfl_xor_sync(0xffffffffu, c2, 16);
        if (on && !(lx & 3) && !(ly & 3)) p2[((long long)b * (H >> 4) + (y >> 4)) * (W >> 4) + (x >> 4)] = (uint8_t)c2;
    }
}

// 4-bit region code per pixel: bit t (t<m) = lo_t <= g <= hi_t ; bit m = none of them; m==0 -> 0
__global__ void __launch_bounds__(256) decomp_codes_kernel(const float* __restrict__ gray, const ImgState* __restrict__ st,
                                                           uint8_t* __restrict__ codes, int HW) {
    const int b = blockIdx.y;
    const Windows wn = load_windows(st[b]);
    const float* in = gray + (long long)b * HW;
    uint8_t* out = codes + (long long)b * HW;
    const bool vec = (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
    for (int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i4 < HW; i4 += gridDim.x * blockDim.x * 4) {
        const int n = HW - i4 < 4 ? HW - i4 : 4;
        float g[4];
        if (vec) {
            const float4 a = *reinterpret_cast<const float4*>(in + i4);
            g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w;
        } else {
            for (int e = 0; e < 4; ++e) g[e] = e < n ? in[i4 + e] : 0.f;
        }
        unsigned packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) packed |= region_code(wn, g[e]) << (8 * e);
        if (vec) {
            *reinterpret_cast<unsigned*>(out + i4) = packed;
        } else {
            for (int e = 0; e < n; ++e) out[i4 + e] = (uint8_t)((packed >> (8 * e)) & 0xffu);
        }
    }
}

// adaptive_max_pool2d of each region mask == bitwise OR of the codes over the window
// [floor(i*H/h), ceil((i+1)*H/h)) x [floor(j*W/w), ceil((j+1)*W/w))
struct PoolLevels {
    uint8_t* out[8];
    int h[8], w[8];
};

// all pyramid levels in one launch (blockIdx.z = level)
__global__ void __launch_bounds__(256) decomp_pool_levels_kernel(const uint8_t* __restrict__ codes, const __grid_constant__ PoolLevels lv,
                                                                 int H, int W) {
    const int b = blockIdx.y, l = blockIdx.z;
    const int h = lv.h[l], w = lv.w[l];
    const uint8_t* in = codes + (long long)b * H * W;
    uint8_t* out = lv.out[l] + (long long)b * h * w;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < h * w; o += gridDim.x * blockDim.x) {
        int i = o / w, j = o - i * w;
        int y0 = (int)(((long long)i * H) / h), y1 = (int)((((long long)i + 1) * H + h - 1) / h);
        int x0 = (int)(((long long)j * W) / w), x1 = (int)((((long long)j + 1) * W + w - 1) / w);
        unsigned c = 0;
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) c |= in[(long long)y * W + x];
        out[o] = (uint8_t)c;
    }
}

__global__ void decomp_edges_kernel(const ImgState* __restrict__ st, float* __restrict__ edges) {
    const int b = blockIdx.x;
    const ImgState s = st[b];
    for (int i = threadIdx.x; i <= kBins; i += blockDim.x) edges[(long long)b * (kBins + 1) + i] = bin_edge(s, i);
}

}  // namespace

extern "C" size_t rgbd_depth_decompose_workspace_bytes(int B) {
    if (B < 0) B = 0;
    return (size_t)B * (sizeof(ImgState) + sizeof(unsigned long long) * kBins) + 256;
}

extern "C" int rgbd_depth_decompose(const float* depth3, long long depth_batch_stride, long long depth_channel_stride,
                                    const float* gray_in, const float* ratio, int B, int H, int W, int num_modes,
                                    float* gray_out, long long* hist_out, float* edges_out, int* n_modes_out,
                                    int* peak_bins_out, float* centres_out, float* windows_out, int* status_out,
                                    int* bias_variant_out, uint8_t* codes_out, int n_levels, const int* level_h, const int* level_w,
                                    uint8_t* const* pooled_out, void* workspace, rgbd_stream_t stream) {
    RGBD_CHECK_ARG((depth3 != nullptr) != (gray_in != nullptr), "depth_decompose: pass exactly one of depth3 / gray_in");
    RGBD_CHECK_ARG(ratio && workspace, "depth_decompose: null ratio / workspace");
    RGBD_CHECK_ARG(depth3 == nullptr || gray_out != nullptr, "depth_decompose: gray_out is required with depth3");
    RGBD_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "depth_decompose: bad geometry");
    RGBD_CHECK_ARG(num_modes >= 1 && num_modes <= 3, "depth_decompose: num_modes must be in [1,3]");
    RGBD_CHECK_ARG(n_levels >= 0 && n_levels <= 8, "depth_decompose: n_levels out of range");
    cudaStream_t s = (cudaStream_t)stream;
    const int HW = H * W;
    uintptr_t wp = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    unsigned long long* hist = (unsigned long long*)wp;
    ImgState* st = (ImgState*)(hist + (size_t)B * kBins);
    // hist + state are one block of the workspace: an all-zero state is the identity of the min/max reductions
    RGBD_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)B * (sizeof(unsigned long long) * kBins + sizeof(ImgState)), s));
    // streaming kernels: ~4 CTAs per SM in total, each CTA reduces to one set of atomics
    const int per_img = max(1, min(ceil_div(HW, 256 * 4), ceil_div(4 * 148, B)));
    dim3 grid(per_img, B);
    const float* gray = gray_in;
    if (depth3) {
        const bool vec = HW % 4 == 0 && depth_batch_stride % 4 == 0 && depth_channel_stride % 4 == 0 &&
                         ((reinterpret_cast<uintptr_t>(depth3) | reinterpret_cast<uintptr_t>(gray_out)) & 15) == 0;
        if (vec) decomp_gray_kernel<true><<<grid, 256, 0, s>>>(depth3, depth_batch_stride, depth_channel_stride, gray_out, st, HW);
        else decomp_gray_kernel<false><<<grid, 256, 0, s>>>(depth3, depth_batch_stride, depth_channel_stride, gray_out, st, HW);
        gray = gray_out;
    } else {
        decomp_minmax_kernel<<<grid, 256, 0, s>>>(gray_in, st, HW);
    }
    RGBD_CHECK_LAUNCH();
    decomp_hist_kernel<<<grid, 256, 0, s>>>(gray, st, hist, HW);
    RGBD_CHECK_LAUNCH();
    ExportPtrs ex = {n_modes_out, peak_bins_out, centres_out, windows_out, status_out, bias_variant_out};
    decomp_modes_kernel<<<B, kBins, 0, s>>>(hist, st, ratio, num_modes, 0.01, nullptr, 1, ex);
    RGBD_CHECK_LAUNCH();
    for (int l = 0; l < n_levels; ++l)
        RGBD_CHECK_ARG(level_h[l] >= 1 && level_w[l] >= 1 && pooled_out[l], "depth_decompose: bad level %d", l);
    bool pyramid = n_levels == 3 && H % 16 == 0 && W % 16 == 0 && (reinterpret_cast<uintptr_t>(gray) & 15) == 0 &&
                   (codes_out == nullptr || (reinterpret_cast<uintptr_t>(codes_out) & 3) == 0);
    for (int l = 0; l < 3 && pyramid; ++l) pyramid = level_h[l] == (H >> (2 + l)) && level_w[l] == (W >> (2 + l));
    if (pyramid) {
        // codes + the three OR-pooled levels in one pass over the gray image
        const int blocks = ceil_div(ceil_div(W, 32) * (H / 16), 8);
        decomp_codes_pool3_kernel<<<dim3(min(blocks, ceil_div(8 * 148, B)), B), 256, 0, s>>>(gray, st, codes_out, pooled_out[0],
                                                                                          pooled_out[1], pooled_out[2], H, W);
        RGBD_CHECK_LAUNCH();
    } else {
        RGBD_CHECK_ARG(codes_out, "depth_decompose: codes_out may only be omitted for a (H/4, H/8, H/16) pyramid");
        decomp_codes_kernel<<<grid, 256, 0, s>>>(gray, st, codes_out, HW);
        RGBD_CHECK_LAUNCH();
        if (n_levels > 0) {
            PoolLevels lv;
            int max_px = 1;
            for (int l = 0; l < n_levels; ++l) {
                lv.out[l] = pooled_out[l]; lv.h[l] = level_h[l]; lv.w[l] = level_w[l];
                max_px = max(max_px, level_h[l] * level_w[l]);
            }
            dim3 g(min(ceil_div(max_px, 256), 296), B, n_levels);
            decomp_pool_levels_kernel<<<g, 256, 0, s>>>(codes_out, lv, H, W);
            RGBD_CHECK_LAUNCH();
        }
    }
    if (hist_out)
        RGBD_CHECK_CUDA(cudaMemcpyAsync(hist_out, hist, sizeof(long long) * (size_t)B * kBins, cudaMemcpyDeviceToDevice, s));
    if (edges_out) {
        decomp_edges_kernel<<<B, 256, 0, s>>>(st, edges_out);
        RGBD_CHECK_LAUNCH();
    }
    return RGBD_OK;
}

// ---- reference helper API on caller-supplied intermediates (DSAModule._select_depth_distribution_modes CM:720-752 and
// ---- _generate_depth_region_masks CM:774-798); the batched path above never needs them ----------------------------------
namespace {

__global__ void helper_state_init_kernel(ImgState* st, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) st[i] = ImgState{};
}

__global__ void helper_state_windows_kernel(ImgState* st, int B, const float* __restrict__ windows, const int* __restrict__ n_windows) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ImgState s = {};
    s.n_modes = n_windows[b];
    for (int k = 0; k < 3; ++k) {
        s.lo[k] = k < s.n_modes ? windows[(b * 3 + k) * 2] : 0.f;
        s.hi[k] = k < s.n_modes ? windows[(b * 3 + k) * 2 + 1] : 0.f;
    }
    st[b] = s;
}

}  // namespace

extern "C" size_t rgbd_depth_helper_workspace_bytes(int B) { return sizeof(ImgState) * (size_t)(B > 0 ? B : 0) + 256; }

extern "C" int rgbd_depth_select_modes(const long long* hist, const float* edges, int B, int num_modes, double prominence_threshold,
                                       int* n_modes_out, int* peak_bins_out, float* centres_out, void* workspace,
                                       rgbd_stream_t stream) {
    RGBD_CHECK_ARG(hist && edges && n_modes_out && peak_bins_out && centres_out && workspace, "depth_select_modes: null pointer");
    RGBD_CHECK_ARG(B >= 1 && num_modes >= 1 && num_modes <= 3, "depth_select_modes: num_modes must be in [1,3]");
    cudaStream_t s = (cudaStream_t)stream;
    ImgState* st = (ImgState*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    helper_state_init_kernel<<<ceil_div(B, 128), 128, 0, s>>>(st, B);
    RGBD_CHECK_LAUNCH();
    ExportPtrs ex = {n_modes_out, peak_bins_out, centres_out, nullptr, nullptr, nullptr};
    decomp_modes_kernel<<<B, kBins, 0, s>>>(reinterpret_cast<const unsigned long long*>(hist), st, nullptr, num_modes,
                                            prominence_threshold, edges, 0, ex);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

extern "C" int rgbd_depth_region_codes(const float* gray, const float* windows, const int* n_windows, int B, long long pixels,
                                       uint8_t* codes_out, void* workspace, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(gray && windows && n_windows && codes_out && workspace, "depth_region_codes: null pointer");
    RGBD_CHECK_ARG(B >= 1 && pixels >= 1 && pixels < (1ll << 31), "depth_region_codes: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    ImgState* st = (ImgState*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    helper_state_windows_kernel<<<ceil_div(B, 128), 128, 0, s>>>(st, B, windows, n_windows);
    RGBD_CHECK_LAUNCH();
    dim3 grid(min(ceil_div((int)pixels, 256 * 4), 296), B);
    decomp_codes_kernel<<<grid, 256, 0, s>>>(gray, st, codes_out, (int)pixels);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}

// CustomMask2FormerPixelLevelModule.to_grayscale (CM:392-502) for 3-channel float32 tensors: gray = (0.299 r + 0.587 g) + 0.114 b
// per pixel (no FMA contraction), B images with arbitrary batch / channel strides.  workspace: rgbd_depth_helper_workspace_bytes(B).
extern "C" int rgbd_to_grayscale(const float* rgb3, long long batch_stride, long long channel_stride, float* gray_out, int B,
                                 long long pixels, void* workspace, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(rgb3 && gray_out && workspace, "to_grayscale: null pointer");
    RGBD_CHECK_ARG(B >= 1 && pixels >= 1 && pixels < (1ll << 31), "to_grayscale: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    ImgState* st = (ImgState*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    helper_state_init_kernel<<<ceil_div(B, 128), 128, 0, s>>>(st, B);
    RGBD_CHECK_LAUNCH();
    const int HW = (int)pixels;
    dim3 grid(min(ceil_div(HW, 256 * 4), 296), B);
    const bool vec = HW % 4 == 0 && batch_stride % 4 == 0 && channel_stride % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(rgb3) | reinterpret_cast<uintptr_t>(gray_out)) & 15) == 0;
    if (vec) decomp_gray_kernel<true><<<grid, 256, 0, s>>>(rgb3, batch_stride, channel_stride, gray_out, st, HW);
    else decomp_gray_kernel<false><<<grid, 256, 0, s>>>(rgb3, batch_stride, channel_stride, gray_out, st, HW);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
