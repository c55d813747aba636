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
if os.environ.get("FUSE_PROJ"):          # SURVEY 8f-2 path: input projections + GroupNorm on this library's kernels
    model.model.pixel_level_module.fuse_input_projections = True
for _ in range(1 + steps):
    b = seg.submit(rgb, depth)
seg.drain()
torch.cuda.synchronize()
print("ok", int(seg.out_host[b]["count"].sum()))
if os.environ.get("TIME"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = int(os.environ["TIME"])
    e0.record()
    for _ in range(n):
        seg.submit()
    seg.drain()
    e1.record()
    torch.cuda.synchronize()
    print("fuse_input_projections=%s: %.2f ms per step of %d frames" % (bool(os.environ.get("FUSE_PROJ")), e0.elapsed_time(e1) / n, B))
