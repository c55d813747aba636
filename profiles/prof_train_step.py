import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import rgbd_b200
from rgbd_b200 import functional as Fn, synthetic, synthetic_weights as SW
B, H, W = 8, 480, 640
model, _ = SW.build_synthetic_rgbd_mask2former()
model.cuda().train()
plm = model.model.pixel_level_module
for p in (*plm.encoder.parameters(), *plm.ratio_predictor.parameters()):
    p.requires_grad_(False)
frames = [synthetic.synth_rgbd_u8(100 + j, H, W) for j in range(B)]
pv = Fn.pack_pixel_values(torch.from_numpy(np.stack([f[0] for f in frames])).cuda(), torch.from_numpy(np.stack([f[1] for f in frames])).cuda())
rs = np.random.RandomState(0)
ml, cl = [], []
for _ in range(B):
    k = rs.randint(3, 21)
    m = torch.zeros(k, H, W)
    for j in range(k):
        h, w = rs.randint(30, 240), rs.randint(30, 320)
        y, x = rs.randint(0, H - h), rs.randint(0, W - w)
        m[j, y:y + h, x:x + w] = 1
    ml.append(m.cuda()); cl.append(torch.from_numpy(rs.randint(0, 48, size=k)).cuda())
def T(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def fwd_nolabel():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return model(pixel_values=pv)
def fwd_label():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return model(pixel_values=pv, mask_labels=ml, class_labels=cl)
print("train-mode forward without labels: %.1f ms" % T(fwd_nolabel))
print("train-mode forward with loss: %.1f ms" % T(fwd_label))
def fb():
    o = fwd_label(); o.loss.backward(); model.zero_grad(set_to_none=True)
print("forward + loss + backward: %.1f ms" % T(fb))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
