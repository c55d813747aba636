"""Timing driver for the two decoder_ops kernels at the benchmarked geometry (batch 32, 480x640 Swin-T: levels 15x20 / 30x40 /
60x80, 8 heads x 32 channels, 4 points; mask logits (32,100,120,160)).  CUDA events, best / median of 5 x 10 launches.
Usage: python profiles/msda_time.py [batch]          (ncu: add `-k regex:msda_fwd|attention_mask -c 4`)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rgbd_b200  # noqa: F401
from rgbd_b200 import functional as Fn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sp = [(15, 20), (30, 40), (60, 80)]
S = sum(h * w for h, w in sp)
H, D, L, P = 8, 32, 3, 4
g = torch.Generator(device="cuda").manual_seed(0)
value = torch.randn(B, S, H, D, device="cuda", generator=g).bfloat16()
offs = (torch.randn(B, S, H, L, P, 2, device="cuda", generator=g) * 2.0).bfloat16()
logit = torch.randn(B, S, H, L * P, device="cuda", generator=g).bfloat16()
ref = torch.rand(B, S, L, 2, device="cuda", generator=g)


def timeit(fn, reps=5, n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


best, med = timeit(lambda: Fn.msda_forward(value, sp, offs, logit, reference_points=ref, softmax=True, out_dtype=torch.bfloat16))
alg = value.numel() * 2 + offs.numel() * 2 + logit.numel() * 2 + ref.numel() * 4 + B * S * H * D * 2
gather = B * S * H * L * P * 4 * D * 2
print("msda_fwd bf16 batch %d: best %.1f us median %.1f us; algorithmic HBM bytes %.1f MB -> %.0f GB/s; L2 gather bytes %.2f GB -> %.2f TB/s"
      % (B, best * 1e3, med * 1e3, alg / 1e6, alg / best / 1e6, gather / 1e9, gather / best / 1e9))
vf = value.float(); of = offs.float(); lf = logit.float()
best, med = timeit(lambda: Fn.msda_forward(vf, sp, of, lf, reference_points=ref, softmax=True))
print("msda_fwd f32  batch %d: best %.1f us median %.1f us" % (B, best * 1e3, med * 1e3))
for dt in (torch.bfloat16, torch.float32):
    ml = torch.randn(B, 100, 120, 160, device="cuda", generator=g).to(dt)
    for tgt in sp:
        best, med = timeit(lambda: Fn.attention_mask(ml, tgt, 8))
        a = torch.nn.functional.interpolate(ml, size=tgt, mode="bilinear", align_corners=False)
        stock = lambda: (torch.nn.functional.interpolate(ml, size=tgt, mode="bilinear", align_corners=False).sigmoid()
                         .flatten(2).unsqueeze(1).repeat(1, 8, 1, 1).flatten(0, 1) < 0.5).bool()
        sb, sm = timeit(stock, reps=3, n=3)
        print("attention_mask %s -> %dx%d: best %.1f us median %.1f us (stock ATen chain: %.1f us)" % (str(dt)[6:], tgt[0], tgt[1], best * 1e3, med * 1e3, sb * 1e3))

# Swin window attention at the four stage geometries of Swin-T on 480x640 (padded to 126x161 -> 18x23 windows at stage 1)
import math
for (hw, heads) in (((126, 161), 3), ((63, 84), 6), ((35, 42), 12), ((21, 21), 24)):
    nw = (hw[0] // 7) * (hw[1] // 7)
    n_win = B * nw
    C = heads * 32
    q, k, v = (torch.randn(n_win, 49, C, device="cuda", generator=g).bfloat16() for _ in range(3))
    bias = torch.randn(heads, 49, 49, device="cuda", generator=g)
    mask = torch.where(torch.rand(nw, 49, 49, device="cuda", generator=g) < 0.2, -100.0, 0.0)
    best, med = timeit(lambda: Fn.window_attention(q, k, v, bias, mask, heads))

    def stock():
        sh = (n_win, 49, heads, 32)
        ql, kl, vl = (t.view(sh).transpose(1, 2) for t in (q, k, v))
        s = torch.matmul(ql, kl.transpose(-1, -2)) / math.sqrt(32) + bias.unsqueeze(0)
        s = (s.view(n_win // nw, nw, heads, 49, 49) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, 49, 49)
        p = torch.softmax(s, -1).to(vl.dtype)
        return torch.matmul(p, vl).permute(0, 2, 1, 3).contiguous()
    sb, _ = timeit(stock, reps=3, n=3)
    flop = n_win * heads * 2 * 2 * 49 * 49 * 32
    print("window_attention bf16 %d windows x %d heads: best %.1f us median %.1f us = %.1f TFLOP/s fp32 FMA (stock op chain: %.1f us)"
          % (n_win, heads, best * 1e3, med * 1e3, flop / best / 1e9, sb * 1e3))

# masked cross-attention of the transformer decoder: 100 queries x S keys, 8 heads x 32, boolean mask
for S in (4800, 1200, 300):
    q = torch.randn(100, B, 256, device="cuda", generator=g).bfloat16()
    k = torch.randn(S, B, 256, device="cuda", generator=g).bfloat16()
    v = torch.randn(S, B, 256, device="cuda", generator=g).bfloat16()
    mask = torch.rand(B * 8, 100, S, device="cuda", generator=g) < 0.5
    best, med = timeit(lambda: Fn.masked_cross_attention(q, k, v, mask, 8))

    def stock():
        ql, kl, vl = (t.view(t.shape[0], B * 8, 32).transpose(0, 1) for t in (q, k, v))
        am = torch.zeros_like(mask, dtype=q.dtype).masked_fill_(mask, float("-inf"))
        w = torch.baddbmm(am, ql * (32 ** -0.5), kl.transpose(-2, -1))
        w = torch.softmax(w.float(), -1).to(q.dtype)
        return torch.bmm(w, vl).transpose(0, 1).contiguous().view(100, B, 256)
    sb, _ = timeit(stock, reps=3, n=3)
    print("masked_cross_attention bf16 S=%d: best %.1f us median %.1f us (stock op chain of nn.MultiheadAttention: %.1f us)" % (S, best * 1e3, med * 1e3, sb * 1e3))

# LayerNorm: this library's one-pass kernel vs ATen at the shapes of the whole model (float32 in / out = autocast semantics)
for rows, C in ((B * 19200, 96), (B * 6300, 256), (B * 4800, 192), (B * 1200, 384), (B * 300, 768)):
    x = torch.randn(rows, C, device="cuda", generator=g)
    w = torch.randn(C, device="cuda", generator=g); bb = torch.randn(C, device="cuda", generator=g)
    mine, _ = timeit(lambda: Fn.layer_norm(x, w, bb, 1e-5))
    mine16, _ = timeit(lambda: Fn.layer_norm(x, w, bb, 1e-5, out_dtype=torch.bfloat16))
    aten, _ = timeit(lambda: torch.nn.functional.layer_norm(x, (C,), w, bb, 1e-5))
    print("layer_norm (%d, %d) f32->f32: %.1f us = %.0f GB/s; f32->bf16: %.1f us; ATen f32->f32: %.1f us" % (rows, C, mine * 1e3, rows * C * 8 / mine / 1e6, mine16 * 1e3, aten * 1e3))
