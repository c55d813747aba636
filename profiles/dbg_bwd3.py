import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import functional as Fn
ci, co, H, W = [int(v) for v in sys.argv[1:5]]
B, n_seg = 2, 4
x = torch.randn(B, ci, H, W, device="cuda")
codes = torch.randint(0, 16, (B, H, W), device="cuda", dtype=torch.uint8)
g = torch.randn(B, co, H, W, device="cuda")
wp = (W + 7) // 8 * 8
gp = Fn.cast_bf16_pitched(g, wp)
xt = torch.zeros(B, n_seg, 1, ci, H, wp, device="cuda", dtype=torch.bfloat16)
Fn.dsam_pack_t(x, codes, xt, ci, wp, n_seg, 4, False)
# check pack_t
for s in range(n_seg):
    m = ((codes >> s) & 1).float()[:, None]
    e = float((xt[:, s, 0, :, :, :W].float() - (x * m).to(torch.bfloat16).float()).abs().max())
    assert e == 0, ("pack_t", s, e)
assert float((gp[..., :W].float() - g.to(torch.bfloat16).float()).abs().max()) == 0
dw = Fn.dsam_wgrad(gp, xt, co, ci, (H, W), n_seg, False); torch.cuda.synchronize()
gb = g.to(torch.bfloat16).float()
for s in range(n_seg):
    m = ((codes >> s) & 1).float()[:, None]
    xm = (x * m).to(torch.bfloat16).float()
    ref = torch.einsum("bnyx,bcyx->nc", gb, xm)
    print("seg", s, "rel err", float((dw[:, s, 0] - ref).abs().max() / ref.abs().max()), float(dw[:, s, 0].norm()), float(ref.norm()))
