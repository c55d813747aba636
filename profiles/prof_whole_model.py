"""Where the whole-model step goes: torch.profiler kernel table of one batch through the RGB-D Mask2Former
(stock HF Swin-T / pixel decoder / transformer decoder + the CUDA depth-guidance hot path), bf16 autocast.
Usage: python profiles/prof_whole_model.py [batch] [fast] > gpurun_out/prof_whole_model.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity, record_function
import rgbd_b200
from rgbd_b200 import pixel_level

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
FAST = "fast" in sys.argv[2:]                 # rgbd_b200.decoder_ops: MSDA + attention-mask kernels
cfg = pixel_level.swin_tiny_mask2former_config()
model = pixel_level.build_rgbd_mask2former(cfg).eval().cuda()
if FAST:
    from rgbd_b200 import decoder_ops
    decoder_ops.install_fast_decoder_ops(model)
pv = torch.randn(B, 10, 480, 640, device="cuda")
pv[:, 9] = (pv[:, 9] > 0).float()
pv[:, 6:9] = pv[:, 6:9].abs().clamp(max=1)

m2f = model.model
hooks = []
def wrap(mod, name):
    orig = mod.forward
    def f(*a, **k):
        with record_function("MOD::" + name):
            return orig(*a, **k)
    mod.forward = f
wrap(m2f.pixel_level_module.encoder, "swin_encoder")
wrap(m2f.pixel_level_module.decoder, "pixel_decoder")
wrap(m2f.pixel_level_module.decoder.encoder, "pixel_decoder.msdeform_encoder")
wrap(m2f.transformer_module, "transformer_module")

with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2):
        model(pixel_values=pv)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model(pixel_values=pv)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=90))
# per-module device time
ev = [e for e in prof.key_averages() if e.key.startswith("MOD::")]
for e in ev:
    print(e.key, "device ms", e.device_time_total / 1e3, "cpu ms", e.cpu_time_total / 1e3)
