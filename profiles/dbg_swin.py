import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import rgbd_b200
from rgbd_b200 import decoder_ops, functional as Fn, synthetic, synthetic_weights as SW
def rel(a,b): return float((a.double()-b.double()).norm()/b.double().norm())
model = SW.build_synthetic_rgbd_mask2former()[0].eval().cuda()
enc = model.model.pixel_level_module.encoder
for (H,W) in ((128,160),(480,640)):
    x = torch.randn(2,3,H,W,device="cuda")
    with torch.no_grad():
        a = enc(x).feature_maps
        decoder_ops.install_fast_decoder_ops(enc)
        b = enc(x).feature_maps
        decoder_ops.uninstall_fast_decoder_ops(enc)
        c = enc(x).feature_maps
    print(H,W,[tuple(t.shape) for t in a],"fast vs stock",[rel(q,p) for p,q in zip(a,b)],"stock rerun",[rel(q,p) for p,q in zip(a,c)])
# per-layer: hook SwinSelfAttention inputs
from transformers.models.swin.modeling_swin import SwinSelfAttention
mods=[m for m in enc.modules() if isinstance(m,SwinSelfAttention)]
x = torch.randn(2,3,128,160,device="cuda")
rec=[]
hs=[m.register_forward_hook(lambda m,i,o: rec.append((i[0].detach().clone(), None if len(i)<2 or i[1] is None else i[1].detach().clone(), o[0].detach().clone()))) for m in mods]
with torch.no_grad(): enc(x)
for h in hs: h.remove()
decoder_ops.install_fast_decoder_ops(enc)
with torch.no_grad():
    for m,(inp,mask,out) in zip(mods,rec):
        got = m(inp, mask)[0]
        print(tuple(inp.shape), None if mask is None else tuple(mask.shape), m.num_attention_heads, "rel", rel(got,out))
