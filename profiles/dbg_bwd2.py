import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import functional as Fn, modules
ci, co, H, W = [int(v) for v in sys.argv[1:5]]
B = 2
m = modules.DSAModule(ci, co, 3).cuda()
x = torch.randn(B, ci, H, W, device="cuda")
codes = torch.randint(0, 16, (B, H, W), device="cuda", dtype=torch.uint8)
g = torch.randn(B, co, H // 2, W // 2, device="cuda")
c_pad, kb, n_pad, n_seg = m._geometry()
wp = (W // 2 + 7) // 8 * 8
wpx = wp
gp = Fn.cast_bf16_pitched(g, wp)
xt = torch.zeros(B, n_seg, 6, c_pad, H // 2, wpx, device="cuda", dtype=torch.bfloat16)
Fn.dsam_pack_t(x, codes, xt, c_pad, wpx, n_seg, 4, True)
try:
    dw = Fn.dsam_wgrad(gp, xt, co, c_pad, (H // 2, W // 2), n_seg, True); torch.cuda.synchronize()
except Exception as e:
    print("FAIL", sys.argv[1:], str(e)[-120:]); sys.exit(0)
import torch.nn.functional as F
ref = []
for s in range(n_seg):
    mask = ((codes >> s) & 1).float()[:, None] if s < 4 else torch.ones(B, 1, H, W, device="cuda")
    xm = (x * mask).to(torch.bfloat16).float(); gb = g.to(torch.bfloat16).float()
    wgt = torch.zeros(co, ci, 3, 3, device="cuda", requires_grad=True)
    (F.conv2d(xm, wgt, None, stride=2, padding=1) * gb).sum().backward()
    ref.append(wgt.grad.permute(0, 2, 3, 1).reshape(co, 9, ci))
ref = torch.stack(ref, 1)
print("ok", sys.argv[1:], "wgrad rel err", float((dw[..., :ci] - ref).abs().max() / ref.abs().max()))
d = dw[..., :ci]
print("norms", float(d.norm()), float(ref.norm()))
for s in range(n_seg):
    errs = []
    for t in range(9):
        # which reference tap matches best?
        best = min(range(9), key=lambda u: float((d[:, s, t] - ref[:, s, u]).norm()))
        errs.append((t, best, round(float((d[:, s, t] - ref[:, s, best]).norm() / ref[:, s, best].norm()), 3)))
    print("seg", s, errs)
