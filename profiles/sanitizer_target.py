"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
forward hot path (smoke), DSAM + DGGM backward, post-processing.  Run as
    compute-sanitizer --tool memcheck python profiles/sanitizer_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as G
import rgbd_b200  # noqa: F401
from rgbd_b200 import modules, synthetic, functional as Fn, postprocess
from oracle import weights as OW

G.smoke()
chans, (H, W), B = (32, 64, 96, 160), (64, 96), 2
m = modules.DepthGuidance(chans)
m.load_state_dict(OW.guidance_weights(seed=3, channels=chans))
m.cuda().train()
rgbs, ds = zip(*[synthetic.synth_rgbd_u8(j, H, W, "nyu") for j in range(B)])
pv = Fn.pack_pixel_values(torch.from_numpy(np.stack(rgbs)).cuda(), torch.from_numpy(np.stack(ds)).cuda())
feats = [torch.randn(B, c, H // s, W // s, device="cuda") for c, s in zip(chans, (4, 8, 16, 32))]
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    out = m(pv, feats)
torch.autograd.backward(out, [torch.randn_like(o) for o in out])
torch.cuda.synchronize()
cls = torch.randn(2, 20, 9, device="cuda") * 3
msk = torch.randn(2, 20, 16, 24, device="cuda") * 3
r = Fn.post_process_instances(cls, msk, 0.0, (64, 96))
torch.cuda.synchronize()
print("[sanitizer target] done")
