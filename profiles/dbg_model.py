import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
import rgbd_b200
from rgbd_b200 import pixel_level, synthetic, functional as Fn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = pixel_level.swin_tiny_mask2former_config()
model = pixel_level.build_rgbd_mask2former(cfg).eval().cuda()
pv = torch.randn(B, 10, 480, 640, device="cuda")
pv[:, 9] = (pv[:, 9] > 0).float()
pv[:, 6:9] = pv[:, 6:9].abs().clamp(max=1)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2):
        out = model(pixel_values=pv)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = model(pixel_values=pv)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
print("whole model bf16 autocast: %.1f ms / batch %d -> %.1f frames/s" % (dt * 1e3, B, B / dt), out.masks_queries_logits.shape)
