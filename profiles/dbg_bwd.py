import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import functional as Fn, modules
from rgbd_b200.modules import _best_box, _round_up
B, ci, co, H, W = 2, 32, 64, 24, 32
m = modules.DSAModule(ci, co, 3).cuda()
x = torch.randn(B, ci, H, W, device="cuda")
codes = torch.randint(0, 16, (B, H, W), device="cuda", dtype=torch.uint8)
g = torch.randn(B, co, H // 2, W // 2, device="cuda")
Ho, Wo = H // 2, W // 2
def step(name, f):
    try:
        r = f(); torch.cuda.synchronize(); print("ok  ", name); return r
    except Exception as e:
        print("FAIL", name, str(e)[-200:]); sys.exit(0)
pk = m._refresh_bwd(); np64 = pk["np64"]; n_seg = pk["n_seg"]
gcl = torch.empty(B, Ho, Wo, np64, device="cuda", dtype=torch.bfloat16)
step("pack g", lambda: Fn.dsam_pack(g, codes.reshape(-1)[:B * Ho * Wo].reshape(B, Ho, Wo), gcl, np64, 1, 0, False))
print("gcl err", float((gcl[..., :co].float() - g.permute(0, 2, 3, 1).to(torch.bfloat16).float()).abs().max()))
dx = torch.zeros_like(x)
H2, W2 = (H + 1) // 2, (W + 1) // 2
cl = pk["classes"][0]
o1 = torch.zeros(B, 160, H2, W2, device="cuda")
sh = torch.zeros(160, device="cuda")
step("mode1 N=160", lambda: Fn.conv_gemm(gcl, (B, Ho, Wo, np64), 1, cl["w"], cl["slices"], 64, B, (H2, W2), _best_box(H2, W2), 160, sh,
     epi_mode=1, out=o1, block_n=160))
ref1 = torch.einsum("byxk,nk->bnyx", gcl.float(), cl["w"].float())
print("mode1 err", float((o1 - ref1).abs().max()))
for cl in pk["classes"]:
    print(cl["py"], cl["px"], cl["w"].shape, cl["slices"].tolist())
    step("dgrad class", lambda: Fn.conv_gemm(gcl, (B, Ho, Wo, np64), 1, cl["w"], cl["slices"], 64, B, (H2, W2), _best_box(H2, W2), ci, None,
         epi_mode=3, out=dx, residual=None, codes=codes, in_hw=(H, W), parity=(cl["py"], cl["px"]), m3_stride=2, m3_masked_segs=4,
         m3_n_seg=n_seg, block_n=32 * n_seg))
# reference
import torch.nn.functional as F
xr = x.clone().requires_grad_(True)
out = 0
for s in range(n_seg):
    mask = ((codes >> s) & 1).float()[:, None] if s < 4 else torch.ones(B, 1, H, W, device="cuda")
    w = (m.conv_layers[s].weight if s < 4 else m.rgb_projection.weight).detach().to(torch.bfloat16).float()
    out = out + F.conv2d(xr * mask, w, None, stride=2, padding=1)
(out * g.to(torch.bfloat16).float()).sum().backward()
print("dgrad rel err", float((dx - xr.grad).abs().max() / xr.grad.abs().max()))
