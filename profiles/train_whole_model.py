"""Whole-model fine-tuning step (BASELINE configs[3] as finetuning.py runs it, mask2former/finetuning.py:98-113): RGB-D
Mask2Former in train mode, bf16 autocast, batch 8, synthetic rectangle labels -> HF Mask2FormerLoss (Hungarian matching on the
host, as in the reference) -> backward (stock autograd + this library's DSAM / DGGM backward kernels) -> AdamW step.
Experiment driver (never a bench number): prints ms per phase and which parameter groups received gradients."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rgbd_b200  # noqa: F401
from rgbd_b200 import functional as Fn, synthetic, synthetic_weights as SW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 480, 640
model, _ = SW.build_synthetic_rgbd_mask2former()
model.cuda().train()
frames = [synthetic.synth_rgbd_u8(100 + j, H, W) for j in range(B)]
rgb = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
pv = Fn.pack_pixel_values(rgb, depth)
rs = np.random.RandomState(0)
mask_labels, class_labels = [], []
for _ in range(B):
    k = rs.randint(3, 21)
    m = torch.zeros(k, H, W)
    for j in range(k):
        h, w = rs.randint(30, 240), rs.randint(30, 320)
        y, x = rs.randint(0, H - h), rs.randint(0, W - w)
        m[j, y:y + h, x:x + w] = 1
    mask_labels.append(m.cuda())
    class_labels.append(torch.from_numpy(rs.randint(0, 48, size=k)).cuda())
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-5)


def step():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(pixel_values=pv, mask_labels=mask_labels, class_labels=class_labels)
    ev[1].record()
    out.loss.backward()
    ev[2].record()
    opt.step()
    opt.zero_grad(set_to_none=True)
    ev[3].record()
    torch.cuda.synchronize()
    return float(out.loss), [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]


loss, _ = step()
plm = model.model.pixel_level_module
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = model(pixel_values=pv, mask_labels=mask_labels, class_labels=class_labels)
out.loss.backward()
groups = {"encoder": plm.encoder, "ratio_predictor": plm.ratio_predictor, "dsam0": plm.dsam0, "dsam1": plm.dsam1, "dsam2": plm.dsam2,
          "dggm": plm.depth_gradient_injection, "pixel_decoder": plm.decoder, "transformer": model.model.transformer_module}
for name, mod in groups.items():
    ps = list(mod.parameters())
    print(f"{name}: {sum(p.grad is not None and bool(p.grad.abs().sum() > 0) for p in ps)} of {len(ps)} parameters with non-zero gradient")
opt.zero_grad(set_to_none=True)
ts = []
for _ in range(5):
    loss, t = step()
    ts.append(t)
t = np.median(np.array(ts), axis=0)
print(f"batch {B}: loss {loss:.4f}; forward+loss {t[0]:.1f} ms, backward {t[1]:.1f} ms, AdamW {t[2]:.1f} ms -> {B / sum(t) * 1e3:.1f} frames/s")
