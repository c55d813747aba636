"""Time the device post-processing at Mask2Former output sizes (100 queries, 48 classes, 120x160 logits -> 480x640) and
the oracle (torch-CPU restatement of the HF routine) on one image beside it; never a bench number."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_b200
from rgbd_b200 import functional as Fn
from oracle import postproc as OP
from oracle.make_golden_postproc import synth_outputs

B = int(os.environ.get("B", 32))
cls, masks = synth_outputs(5, B, 100, 48, 120, 160)
c, m = cls.cuda(), masks.cuda()
for _ in range(2):
    r = Fn.post_process_instances(c, m, 0.05, (480, 640))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    r = Fn.post_process_instances(c, m, 0.05, (480, 640))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
kept = int(r.count.sum())
print(f"device: B={B} {ms:.3f} ms/batch = {B/ms*1e3:.0f} images/s, {kept} segments kept")
t0 = time.perf_counter()
OP.post_process_image(cls[0], masks[0], 0.05, (480, 640))
t1 = time.perf_counter()
print(f"oracle (torch CPU, {torch.get_num_threads()} threads): {(t1-t0)*1e3:.1f} ms/image")
