"""N whole-model serving steps at the bench workload (batch 32, 480x640; serving.RgbdInstanceSegmenter with decoder_ops), for
ncu launch lists of the e2e step: which kernels run, own vs stock share (never a bench number).
Usage: python profiles/serve_step.py [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import rgbd_b200  # noqa: F401
from rgbd_b200 import serving

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model, _ = bench.build_whole_model()
model.cuda()
rgb, depth = bench.make_frames(B)
seg = serving.RgbdInstanceSegmenter(model, B, (bench.H, bench.W), threshold=bench.POST_THRESHOLD)
rgb, depth = torch.from_numpy(rgb), torch.from_numpy(depth)
for _ in range(1 + steps):
    b = seg.submit(rgb, depth)
seg.drain()
torch.cuda.synchronize()
print("ok", int(seg.out_host[b]["count"].sum()))
