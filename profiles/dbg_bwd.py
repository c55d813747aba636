import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import functional as Fn, modules
rs = np.random.RandomState(0)
B, ci, co, H, W = 2, 32, 64, 24, 32
m = modules.DSAModule(ci, co, 3).cuda()
x = torch.randn(B, ci, H, W, device="cuda")
codes = torch.randint(0, 16, (B, H, W), device="cuda", dtype=torch.uint8)
variant = torch.full((B,), 4, device="cuda", dtype=torch.int32)
g = torch.randn(B, co, H // 2, W // 2, device="cuda")
def step(name, f):
    try:
        r = f(); torch.cuda.synchronize(); print("ok  ", name); return r
    except Exception as e:
        print("FAIL", name, str(e)[:200]); sys.exit(1)
db = step("dbias", lambda: Fn.dsam_dbias(g, variant, 4))
gp = step("cast", lambda: Fn.cast_bf16_pitched(g, 16))
c_pad, kb, n_pad, n_seg = m._geometry()
xt = torch.zeros(B, n_seg, 4, c_pad, H // 2, 16, device="cuda", dtype=torch.bfloat16)
step("pack_t", lambda: Fn.dsam_pack_t(x, codes, xt, c_pad, 16, n_seg, 4, True))
dw = step("wgrad", lambda: Fn.dsam_wgrad(gp, xt, co, c_pad, (H // 2, W // 2), n_seg, True))
# reference wgrad in torch
import torch.nn.functional as F
ref = []
for s in range(n_seg):
    mask = ((codes >> s) & 1).float()[:, None] if s < 4 else torch.ones(B, 1, H, W, device="cuda")
    xm = (x * mask).to(torch.bfloat16).float()
    gb = g.to(torch.bfloat16).float()
    wgt = torch.zeros(co, ci, 3, 3, device="cuda", requires_grad=True)
    out = F.conv2d(xm, wgt, None, stride=2, padding=1)
    (out * gb).sum().backward()
    ref.append(wgt.grad.permute(0, 2, 3, 1).reshape(co, 9, ci))
ref = torch.stack(ref, 1)
print("wgrad rel err", float((dw[..., :ci] - ref).abs().max() / ref.abs().max()))
dx, grads = step("backward_impl", lambda: m._stage_backward_impl(x, codes, variant, g, True))
print("done", dx.shape)
