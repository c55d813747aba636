import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import functional as fn
from rgbd_b200.modules import _best_box
rs = np.random.RandomState(1)
kb, c, n, (H, W), B = 64, 128, 64, (4, 128), 2
a = torch.from_numpy(rs.randn(B, H, W, c).astype(np.float32)).cuda().to(torch.bfloat16)
w = torch.from_numpy((rs.randn(n, c) / np.sqrt(c)).astype(np.float32)).cuda().to(torch.bfloat16)
shift = torch.from_numpy(rs.randn(n).astype(np.float32)).cuda()
scale = torch.from_numpy(rs.uniform(0.5, 1.5, n).astype(np.float32)).cuda()
slices = torch.tensor([(kb * i, 0, 0, 0) for i in range(c // kb)], dtype=torch.int32).cuda()
out = torch.zeros(B, H, W, n, device="cuda", dtype=torch.bfloat16)
fn.conv_gemm(a, (B, H, W, c), 1, w, slices, kb, B, (H, W), _best_box(H, W), n, shift, scale=scale, act=1, out=out)
torch.cuda.synchronize()
ref = torch.relu((a.float() @ w.float().T) * scale + shift)
print("err", float((out.float() - ref).abs().max()))
