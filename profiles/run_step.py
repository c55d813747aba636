"""One warm-up + N hot-path steps at the bench workload (for ncu captures; never a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import rgbd_b200
from rgbd_b200 import functional as Fn, modules, synthetic
from oracle import weights as OW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
m = modules.DepthGuidance(bench.CHANS)
m.load_state_dict(OW.guidance_weights(seed=42, channels=bench.CHANS))
m.to(dev).eval()
rgb, depth = bench.make_frames(min(B, 4))
pv = torch.empty(B, 10, bench.H, bench.W, device=dev)
for j in range(B):
    k = j % len(rgb)
    pv[j, 0:3] = torch.from_numpy(synthetic.normalise_u8(rgb[k])).to(dev)
    pv[j, 3:6] = torch.from_numpy(synthetic.normalise_u8(np.repeat(depth[k][:, :, None], 3, axis=2))).to(dev)
dd = torch.from_numpy(depth).to(dev)[torch.arange(B) % len(rgb)].contiguous()
Fn.gradient_features(dd, norm_out=pv[:, 6:9], vmask_out=pv[:, 9:10])
feats = bench.make_features(B, 7, dev)
with torch.no_grad():
    for _ in range(1 + steps):
        out = m(pv, feats)
torch.cuda.synchronize()
print("ok", float(out[0].abs().mean()))
