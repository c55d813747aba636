"""[The RGBD_DBG_SHIFT / RGBD_DBG_BO measurement hook this script drives lived in csrc/conv_gemm.cu up to commit 47717bd and was
removed afterwards; check that commit out to re-run the experiment.  Result: profiles/r01_notes.md.]
Experiment: does tcgen05.mma accept a 128B-swizzled K-major A tile whose descriptor start address is advanced by
whole 128-byte rows (not 1024-byte aligned)?  Needed to reuse one shared-memory tile for the 3 dx taps of a 3x3 conv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import functional as fn
rs = np.random.RandomState(1)
kb, c, n, (H, W), B = 64, 128, 64, (2, 256), 1
a = torch.from_numpy(rs.randn(B, H, W, c).astype(np.float32)).cuda().to(torch.bfloat16)
w = torch.from_numpy((rs.randn(n, c) / np.sqrt(c)).astype(np.float32)).cuda().to(torch.bfloat16)
shift = torch.zeros(n, device="cuda")
slices = torch.tensor([(kb * i, 0, 0, 0) for i in range(c // kb)], dtype=torch.int32).cuda()
ref = (a.float() @ w.float().T)
for s in (0, 1, 2, 3, 8, 9):
    for bo in (0, 1):
        os.environ["RGBD_DBG_SHIFT"] = str(s); os.environ["RGBD_DBG_BO"] = str(bo)
        out = torch.zeros(B, H, W, n, device="cuda", dtype=torch.bfloat16)
        fn.conv_gemm(a, (B, H, W, c), 1, w, slices, kb, B, (H, W), (128, 1), n, shift, out=out)
        torch.cuda.synchronize()
        # rows i < 128 - s of every tile are computed from loaded data
        o = out.float().reshape(B, H, W // 128, 128, n)[:, :, :, :128 - s]
        r = ref.reshape(B, H, W // 128, 128, n)[:, :, :, :128 - s]
        print(f"shift={s} base_offset_mode={bo}: max err {float((o - r).abs().max()):.4f} (ref max {float(r.abs().max()):.2f})")
