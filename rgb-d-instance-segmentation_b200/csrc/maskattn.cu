// Masked-attention mask of the Mask2Former transformer decoder (inference), as Hugging Face's Mask2FormerMaskPredictor
// builds it from the per-query mask logits the model also returns (the consumer of the hot path's fused pyramid; reference
// call chain mask2former/utils/custom_model.py:383 -> Mask2FormerModel.forward -> transformer_module):
//   m = interpolate(mask_logits (B,Q,h,w), size=(th,tw), mode="bilinear", align_corners=False)
//   attention_mask = (m.sigmoid().flatten(2).unsqueeze(1).repeat(1, heads, 1, 1).flatten(0, 1) < 0.5)     (B*heads, Q, th*tw) bool
// ATen's upsample_bilinear2d kernel parallelises over the th*tw OUTPUT pixels only and loops over B*Q planes inside a thread
// (300 threads for a 15x20 target: 1.8 ms per call, ten calls per forward); here one thread makes one (plane, pixel) decision
// and writes its `heads` copies.  Arithmetic follows ATen operation by operation (source index = scale*(dst+0.5)-0.5 clamped
// at 0, lambda weights, bf16 inputs rounded to bf16 after the interpolation and after the sigmoid) -- compiled without FMA
// contraction because the result is a threshold decision.
#include "common.cuh"
#include "rgbd_b200.h"

namespace {

template <bool BF16>
__device__ __forceinline__ float ld_logit(const void* p, long long i) {
    if (BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    return __ldg(reinterpret_cast<const float*>(p) + i);
}

template <bool BF16>
__global__ void __launch_bounds__(256) attention_mask_kernel(const void* __restrict__ logits, long long planes, int Q, int h, int w,
                                                             int th, int tw, int heads, float rh, float rw,
                                                             uint8_t* __restrict__ out) {
    const long long n_out = planes * th * tw;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int N = th * tw;
    const long long plane = i / N;
    const int pix = (int)(i % N), ty = pix / tw, tx = pix % tw;
    float fy = __fsub_rn(__fmul_rn(rh, __fadd_rn((float)ty, 0.5f)), 0.5f);
    float fx = __fsub_rn(__fmul_rn(rw, __fadd_rn((float)tx, 0.5f)), 0.5f);
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly1 = __fsub_rn(fy, (float)y0), ly0 = __fsub_rn(1.f, ly1);
    const float lx1 = __fsub_rn(fx, (float)x0), lx0 = __fsub_rn(1.f, lx1);
    const long long base = plane * h * w;
    const float v00 = ld_logit<BF16>(logits, base + (long long)y0 * w + x0), v01 = ld_logit<BF16>(logits, base + (long long)y0 * w + x1);
    const float v10 = ld_logit<BF16>(logits, base + (long long)y1 * w + x0), v11 = ld_logit<BF16>(logits, base + (long long)y1 * w + x1);
    const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
    const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
    float m = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    if (BF16) m = __bfloat162float(__float2bfloat16_rn(m));
    float sg = 1.0f / (1.0f + expf(-m));                        // torch.sigmoid in float32
    if (BF16) sg = __bfloat162float(__float2bfloat16_rn(sg));
    const uint8_t flag = sg < 0.5f ? 1 : 0;
    const long long b = plane / Q;
    const int q = (int)(plane % Q);
    for (int hd = 0; hd < heads; ++hd) out[((b * heads + hd) * Q + q) * N + pix] = flag;
}

}  // namespace

extern "C" int rgbd_attention_mask(const void* mask_logits, int dtype, int B, int Q, int h, int w, int th, int tw, int heads,
                                   uint8_t* out, rgbd_stream_t stream) {
    RGBD_CHECK_ARG(mask_logits && out, "attention_mask: null pointer");
    RGBD_CHECK_ARG(B >= 1 && Q >= 1 && h >= 1 && w >= 1 && th >= 1 && tw >= 1 && heads >= 1, "attention_mask: bad sizes");
    RGBD_CHECK_ARG(dtype == RGBD_DTYPE_F32 || dtype == RGBD_DTYPE_BF16, "attention_mask: logits are f32 or bf16");
    const long long planes = (long long)B * Q, n_out = planes * th * tw;
    const long long blocks = (n_out + 255) / 256;
    RGBD_CHECK_ARG(blocks <= 0x7fffffffLL, "attention_mask: too many outputs");
    const float rh = (float)h / (float)th, rw = (float)w / (float)tw;     // area_pixel_compute_scale without scale factors
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == RGBD_DTYPE_BF16)
        attention_mask_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(mask_logits, planes, Q, h, w, th, tw, heads, rh, rw, out);
    else
        attention_mask_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(mask_logits, planes, Q, h, w, th, tw, heads, rh, rw, out);
    RGBD_CHECK_LAUNCH();
    return RGBD_OK;
}
