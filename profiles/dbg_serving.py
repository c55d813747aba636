"""Debug: where does serving.RgbdInstanceSegmenter differ from the direct call sequence?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import functional as Fn, postprocess, serving, synthetic, synthetic_weights as SW
model, _ = SW.build_synthetic_rgbd_mask2former(decisive=True, num_labels=8)
model.cuda()
B, H, W = 2, 128, 160
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rgbs, ds = zip(*[synthetic.synth_rgbd_u8(700 + j, H, W, "nyu") for j in range(B)])
rgb, d = torch.from_numpy(np.stack(rgbs)), torch.from_numpy(np.stack(ds))
outs = []
with torch.no_grad():
    for _ in range(3):
        pv = Fn.pack_pixel_values(rgb.cuda(), d.cuda())
        o = model(pixel_values=pv)
        outs.append((pv.clone(), o.class_queries_logits.clone(), o.masks_queries_logits.clone()))
print("direct x3: pv equal", torch.equal(outs[0][0], outs[1][0]), "cls equal", torch.equal(outs[0][1], outs[1][1]), torch.equal(outs[1][1], outs[2][1]),
      "msk equal", torch.equal(outs[0][2], outs[1][2]), "max diff", float((outs[0][2] - outs[1][2]).abs().max()))
r1 = Fn.post_process_instances(outs[0][1].float().contiguous(), outs[0][2].float().contiguous(), 0.5, (H, W))
r2 = Fn.post_process_instances(outs[0][1].float().contiguous(), outs[0][2].float().contiguous(), 0.5, (H, W))
print("postproc x2 on same logits: seg equal", torch.equal(r1.segmentation, r2.segmentation), "count", r1.count.tolist(), r2.count.tolist(),
      "labels equal", torch.equal(r1.labels, r2.labels), "scores equal", torch.equal(r1.scores, r2.scores))
n = int(r1.count[0])
print("scores[0][:n] sorted desc?", bool((r1.scores[0, :n][:-1] >= r1.scores[0, :n][1:]).all()), r1.scores[0, :8].tolist())
seg = serving.RgbdInstanceSegmenter(model, B, (H, W), threshold=0.5, autocast_dtype=None)
res = seg(rgb, d)
s_direct = r1.segmentation.cpu()
s_pipe = torch.stack([x["segmentation"] for x in res])
print("pipe vs direct: differing pixels", int((s_direct != s_pipe).sum()), "of", s_pipe.numel())
print("pipe pv equal direct pv", torch.equal(seg.pv[0], outs[0][0]))
