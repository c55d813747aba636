"""Time rgbd_ratio_front with the sliding-window (compact) stem operand alone at the bench shape; never a bench number."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_b200
from rgbd_b200 import functional as Fn
from rgbd_b200.modules import EnhancedDepthImageRatioPredictor, _best_box
from oracle import weights as OW
B, H, W = 32, 480, 640
m = EnhancedDepthImageRatioPredictor(3); m.load_state_dict(OW.ratio_weights(seed=1)); m.cuda().eval()
pk = m._refresh()
x = torch.randn(B, 3, H, W, device="cuda")
r = torch.empty(B, 2, H + 6, Fn.ratio_stem_compact_width(W), 4, device="cuda", dtype=torch.bfloat16)
out = torch.empty(B, H, W, 128, device="cuda", dtype=torch.bfloat16)
Fn.ratio_stem_pack_compact(x, r)
args = (r, pk["w1c"], pk["w2"], pk["w3"], pk["w4"], pk["sh1"], pk["sh2"], pk["sh3"], pk["sh4"], out, (128, 1))
for _ in range(3): Fn.ratio_front(*args)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): Fn.ratio_front(*args)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 10)
print(f"ratio_front compact: best {min(ts)*1e3:.1f} us median {sorted(ts)[2]*1e3:.1f} us")
