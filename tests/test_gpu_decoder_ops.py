"""GPU parity of the opt-in kernels inside the stock modules on either side of the hot path (csrc/msda.cu, csrc/maskattn.cu:
reference call site mask2former/utils/custom_model.py:383 and the transformer module behind it; csrc/winattn.cu: the Swin
encoder called at CM:330).  The checker is Hugging Face's own
pure-PyTorch code (transformers/models/mask2former/modeling_mask2former.py: `multi_scale_deformable_attention`, the
deformable-attention module, `Mask2FormerMaskPredictor`) run in float32 on the CPU or, for the module-level checks, the
stock forward of the same module object on the GPU.  Tolerances: float32 1e-5 (summation order only), bf16 1e-2."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fn():
    import rgbd_b200  # noqa: F401
    from rgbd_b200 import functional
    functional._lib.load()
    return functional


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def hf():
    from transformers.models.mask2former import modeling_mask2former as m
    return m


SHAPES = {
    "swin_t_480x640": [(15, 20), (30, 40), (60, 80)],
    "odd": [(7, 5), (13, 9), (3, 11)],
    "one_level": [(16, 16)],
}


@pytest.mark.parametrize("shapes", list(SHAPES), ids=list(SHAPES))
@pytest.mark.parametrize("P", [4, 3])
@pytest.mark.parametrize("vdtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_msda_function_matches_hf(fn, shapes, P, vdtype):
    """`multi_scale_deformable_attention` signature: locations + weights given.  Locations reach outside [0,1] (zeros
    padding), sit exactly on pixel centres and on the border."""
    sp = SHAPES[shapes]
    g = torch.Generator().manual_seed(11 + P)
    B, H, D, L = 2, 8, 32, len(sp)
    S = sum(h * w for h, w in sp)
    Q = S
    value = torch.randn(B, S, H, D, generator=g)
    if vdtype == torch.bfloat16:
        value = value.bfloat16().float()                   # exactly representable: the bf16 kernel path sees the same numbers
    loc = torch.rand(B, Q, H, L, P, 2, generator=g) * 1.3 - 0.15
    loc[0, 0] = 0.0
    loc[0, 1] = 1.0
    loc[0, 2] = 0.5
    for l, (h, w) in enumerate(sp):                        # exact pixel centres of each level
        loc[1, 3, :, l, :, 0] = (torch.randint(0, w, (H, P), generator=g) + 0.5) / w
        loc[1, 3, :, l, :, 1] = (torch.randint(0, h, (H, P), generator=g) + 0.5) / h
    loc[1, 4] = 5.0                                        # far outside: samples nothing (zeros padding)
    attw = torch.softmax(torch.randn(B, Q, H, L * P, generator=g), -1).view(B, Q, H, L, P)
    ref = hf().multi_scale_deformable_attention(value, sp, loc, attw)
    got = fn.msda_forward(value.cuda().to(vdtype), sp, loc.cuda(), attw.cuda())
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert rel_l2(got.cpu(), ref) < 1e-5
    assert float((got.cpu() - ref).abs().max()) < 2e-5
    assert float(got[1, 4].abs().max()) == 0.0


def test_msda_fused_softmax_and_locations(fn):
    """Fused mode: raw offsets + reference points + logits, against the module's own arithmetic in float32."""
    sp = SHAPES["swin_t_480x640"]
    g = torch.Generator().manual_seed(5)
    B, H, D, L, P = 2, 8, 32, 3, 4
    S = sum(h * w for h, w in sp)
    value = torch.randn(B, S, H, D, generator=g)
    off = torch.randn(B, S, H, L, P, 2, generator=g) * 3.0
    logit = torch.randn(B, S, H, L * P, generator=g) * 2.0
    ref_pts = torch.rand(B, S, L, 2, generator=g)
    norm = torch.tensor([[w, h] for h, w in sp], dtype=torch.long)
    loc = ref_pts[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    attw = torch.softmax(logit, -1).view(B, S, H, L, P)
    ref = hf().multi_scale_deformable_attention(value, sp, loc, attw)
    got = fn.msda_forward(value.cuda(), sp, off.cuda(), logit.cuda(), reference_points=ref_pts.cuda(), softmax=True)
    assert rel_l2(got.cpu(), ref) < 1e-5
    got16 = fn.msda_forward(value.cuda(), sp, off.cuda(), logit.cuda(), reference_points=ref_pts.cuda(), softmax=True,
                            out_dtype=torch.bfloat16)
    assert got16.dtype == torch.bfloat16 and torch.equal(got16, got.bfloat16())


def test_msda_argument_errors(fn):
    from rgbd_b200._lib import RgbdB200Error
    v = torch.zeros(1, 10, 2, 32, device="cuda")
    loc = torch.zeros(1, 10, 2, 1, 4, 2, device="cuda")
    w = torch.zeros(1, 10, 2, 1, 4, device="cuda")
    with pytest.raises(RgbdB200Error):
        fn.msda_forward(v, [(3, 3)], loc, w)                           # 9 != 10 rows
    with pytest.raises(RgbdB200Error):
        fn.msda_forward(v.cpu(), [(2, 5)], loc, w)                     # no CPU path
    with pytest.raises(RgbdB200Error):
        fn.msda_forward(torch.zeros(1, 10, 2, 12, device="cuda"), [(2, 5)], loc, w)   # head dim not a multiple of 8
    with pytest.raises(RgbdB200Error):
        fn.msda_forward(v.half(), [(2, 5)], loc, w)


@pytest.mark.parametrize("autocast", [False, True], ids=["fp32", "bf16_autocast"])
def test_deformable_attention_module_matches_stock_forward(fn, autocast):
    """The rebound forward of HF's deformable-attention module against its stock forward, same module object and weights, at
    the benchmarked geometry (480x640 Swin-T levels, 8 heads x 32 channels)."""
    from rgbd_b200 import decoder_ops
    m = hf()
    torch.manual_seed(3)
    mod = m.Mask2FormerPixelDecoderEncoderMultiscaleDeformableAttention(256, 8, 3, 4).cuda().eval()
    for p in mod.parameters():
        torch.nn.init.normal_(p, std=0.08)
    sp = SHAPES["swin_t_480x640"]
    B, S = 2, sum(h * w for h, w in sp)
    x = torch.randn(B, S, 256, device="cuda")
    pos = torch.randn(B, S, 256, device="cuda") * 0.3
    ref_pts = torch.rand(B, S, 3, 2, device="cuda")
    mask = torch.zeros(B, S, dtype=torch.bool, device="cuda")
    mask[1, -50:] = True
    kw = dict(attention_mask=mask, encoder_hidden_states=x, position_embeddings=pos, reference_points=ref_pts,
              spatial_shapes_list=sp)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        stock, _ = mod(x, **kw)
        decoder_ops.install_fast_decoder_ops(mod)
        before = fn.LAUNCHES
        fast, _ = mod(x, **kw)
        assert fn.LAUNCHES == before + 1
        decoder_ops.uninstall_fast_decoder_ops(mod)
        again, _ = mod(x, **kw)
    assert torch.equal(again, stock)
    assert fast.dtype == stock.dtype
    assert rel_l2(fast.float(), stock.float()) < (1e-2 if autocast else 1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("target", [(15, 20), (30, 40), (60, 80), (17, 23)], ids=lambda t: "%dx%d" % t)
def test_attention_mask_matches_hf_mask_predictor(fn, dtype, target):
    """`Mask2FormerMaskPredictor`'s attention mask (ATen bilinear resize + sigmoid + threshold + repeat), float32 /
    bfloat16 ATen (CUDA) arithmetic as the checker."""
    g = torch.Generator().manual_seed(target[0])
    B, Q, h, w, heads = 2, 100, 120, 160, 8
    logits = (torch.randn(B, Q, h, w, generator=g) * 4.0).to(dtype)
    logits[0, 0] = 0.0                                               # sigmoid(0) = 0.5 is NOT < 0.5
    logits[0, 1] = -1e-9 if dtype == torch.float32 else -1e-3        # rounds to 0.5 after the sigmoid
    logits = logits.cuda()                                           # the checker is ATen's CUDA path, the one being replaced
    a = torch.nn.functional.interpolate(logits, size=target, mode="bilinear", align_corners=False)
    a = a.sigmoid().flatten(2).unsqueeze(1).repeat(1, heads, 1, 1)
    ref = (a.flatten(0, 1) < 0.5).bool()
    got = fn.attention_mask(logits, target, heads)
    assert got.dtype == torch.bool and got.shape == ref.shape
    diff = int((got != ref).sum())
    assert diff == 0, f"{diff} of {ref.numel()} mask bits differ"
    assert not bool(got[0:heads, 0].any()) and not bool(got[0:heads, 1].any())


def test_whole_model_with_fast_decoder_ops(fn):
    """RGB-D Mask2Former with and without the rebound forwards: same logits (float32).  The model is made decisive first
    (synthetic_weights.make_decisive): out of the box a random-init Mask2Former has near-zero mask logits, so the 1e-7 differences
    between two float32 implementations flip attention-mask bits and the outputs of BOTH drift apart chaotically (measured: Swin
    feature maps equal to 1e-6, mask logits 2.6e-2 apart)."""
    import numpy as np
    from rgbd_b200 import decoder_ops, synthetic, synthetic_weights
    model = synthetic_weights.build_synthetic_rgbd_mask2former(decisive=True)[0].eval().cuda()
    frames = [synthetic.synth_rgbd_u8(40 + j, 128, 160) for j in range(2)]
    rgb = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
    depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
    pv = fn.pack_pixel_values(rgb, depth)
    with torch.no_grad():
        stock = model(pixel_values=pv)
        decoder_ops.install_fast_decoder_ops(model)
        before = fn.LAUNCHES
        fast = model(pixel_values=pv)
        used = fn.LAUNCHES - before
        decoder_ops.uninstall_fast_decoder_ops(model)
        used_stock_before = fn.LAUNCHES
        model(pixel_values=pv)
        hot_path_launches = fn.LAUNCHES - used_stock_before
    assert used - hot_path_launches >= 6 + 10 + 2 * 12               # (+ one launch per generic LayerNorm call) six encoder layers, ten mask-predictor calls, twelve Swin blocks (table + attention); float32: the decoder cross-attention and the pre-norms stay stock
    assert rel_l2(fast.masks_queries_logits, stock.masks_queries_logits) < 2e-3
    assert rel_l2(fast.class_queries_logits, stock.class_queries_logits) < 2e-3


def _window_attention_reference(q, k, v, bias, mask, heads):
    """transformers SwinSelfAttention.forward's inner arithmetic in float32."""
    n_win, N, C = q.shape
    sh = (n_win, N, heads, 32)
    ql, kl, vl = (t.float().view(sh).transpose(1, 2) for t in (q, k, v))
    s = torch.matmul(ql, kl.transpose(-1, -2)) / (32 ** 0.5) + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        s = (s.view(n_win // nw, nw, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    ctx = torch.matmul(torch.softmax(s, -1), vl)
    return ctx.permute(0, 2, 1, 3).contiguous().view(n_win, N, C)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("N,heads,masked", [(49, 3, True), (49, 6, False), (64, 4, True), (16, 1, True), (33, 24, False)])
def test_window_attention_matches_torch(fn, dtype, N, heads, masked):
    g = torch.Generator(device="cuda").manual_seed(N + heads)
    nw = 6
    n_win = nw * 5 + (0 if masked else 1)
    q, k, v = (torch.randn(n_win, N, heads * 32, device="cuda", generator=g).to(dtype) * 1.5 for _ in range(3))
    bias = torch.randn(heads, N, N, device="cuda", generator=g)
    mask = None
    if masked:
        mask = torch.where(torch.rand(nw, N, N, device="cuda", generator=g) < 0.3, -100.0, 0.0)
        mask[:, torch.arange(N), torch.arange(N)] = 0.0           # a token always sees itself (as in Swin's shift masks)
    want = _window_attention_reference(q, k, v, bias, mask, heads)
    got = fn.window_attention(q, k, v, bias, mask, heads)
    assert got.dtype == dtype and got.shape == want.shape
    assert rel_l2(got.float(), want) < (1e-5 if dtype == torch.float32 else 3e-3)   # bf16: output rounding only


def test_window_attention_argument_errors(fn):
    from rgbd_b200._lib import RgbdB200Error
    q = torch.zeros(4, 49, 96, device="cuda")
    bias = torch.zeros(3, 49, 49, device="cuda")
    with pytest.raises(RgbdB200Error):
        fn.window_attention(q, q, q, bias, None, 4)                                  # 96 != 4 * 32
    with pytest.raises(RgbdB200Error):
        fn.window_attention(q, q, q, bias, torch.zeros(3, 49, 49, device="cuda"), 3)  # 4 windows, 3 mask windows
    with pytest.raises(RgbdB200Error):
        fn.window_attention(q.cpu(), q.cpu(), q.cpu(), bias.cpu(), None, 3)
    with pytest.raises(RgbdB200Error):
        fn.window_attention(torch.zeros(4, 81, 96, device="cuda"), q, q, bias, None, 3)


@pytest.mark.parametrize("autocast", [False, True], ids=["fp32", "bf16_autocast"])
def test_swin_self_attention_module_matches_stock_forward(fn, autocast):
    """HF's SwinSelfAttention with its rebound forward against its stock forward (same module object and weights), shifted-window
    mask included.  Under autocast both are compared with the float32 stock result: the kernel keeps the scores in float32 where
    the stock path rounds them to bfloat16, so it must not be further away than the stock path itself."""
    from transformers import SwinConfig
    from transformers.models.swin.modeling_swin import SwinSelfAttention
    from rgbd_b200 import decoder_ops
    torch.manual_seed(1)
    mod = SwinSelfAttention(SwinConfig(), dim=192, num_heads=6, window_size=7).cuda().eval()
    torch.nn.init.normal_(mod.relative_position_bias_table, std=0.5)
    nw, batch = 12, 3
    x = torch.randn(batch * nw, 49, 192, device="cuda")
    mask = torch.where(torch.rand(nw, 49, 49, device="cuda") < 0.25, -100.0, 0.0)
    mask[:, torch.arange(49), torch.arange(49)] = 0.0
    with torch.no_grad():
        exact = mod(x, mask)[0]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            stock = mod(x, mask)[0]
            decoder_ops.install_fast_decoder_ops(mod)
            before = fn.LAUNCHES
            fast = mod(x, mask)[0]
            fast_nomask = mod(x, None)[0]
            assert fn.LAUNCHES == before + 4
            decoder_ops.uninstall_fast_decoder_ops(mod)
            stock_nomask = mod(x, None)[0]
    assert fast.dtype == stock.dtype and fast.shape == stock.shape
    if autocast:
        assert rel_l2(fast.float(), exact) <= max(rel_l2(stock.float(), exact), 4e-3) * 1.05
        assert rel_l2(fast_nomask.float(), stock_nomask.float()) < 1e-2
    else:
        assert rel_l2(fast, stock) < 1e-5 and rel_l2(fast_nomask, stock_nomask) < 1e-5


@pytest.mark.parametrize("C", [96, 192, 384, 768, 1024, 4, 100])
def test_layer_norm_matches_torch(fn, C):
    g = torch.Generator(device="cuda").manual_seed(C)
    x = torch.randn(3, 1001, C, device="cuda", generator=g) * 2.0 + 0.5
    w = torch.randn(C, device="cuda", generator=g)
    b = torch.randn(C, device="cuda", generator=g)
    want = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
    got = fn.layer_norm(x, w, b, 1e-5)
    assert got.dtype == torch.float32 and rel_l2(got, want) < 1e-6
    got16 = fn.layer_norm(x, w, b, 1e-5, out_dtype=torch.bfloat16)
    assert got16.dtype == torch.bfloat16
    # what nn.Linear would see under autocast: the float32 result rounded to bf16 -- equal up to last-bit ties of the fp32 arithmetic
    same = float((got16 == want.bfloat16()).float().mean())
    assert same > 0.999 and rel_l2(got16.float(), want) < 4e-3, same
    xb = x.bfloat16()                                          # bf16 residual stream (Swin stages 2-4): widened exactly, like autocast
    want_b = torch.nn.functional.layer_norm(xb.float(), (C,), w, b, 1e-5)
    got_b = fn.layer_norm(xb, w, b, 1e-5, out_dtype=torch.bfloat16)
    assert float((got_b == want_b.bfloat16()).float().mean()) > 0.999 and rel_l2(got_b.float(), want_b) < 4e-3


def test_swin_encoder_with_bf16_prenorms_and_window_attention(fn):
    """The stock Swin-T encoder under bf16 autocast with the rebound pre-norm LayerNorms (bf16 output) and window attention against
    the untouched encoder, both measured against the float32 forward: the rebound version must not be further away."""
    from rgbd_b200 import decoder_ops, synthetic_weights
    enc = synthetic_weights.build_synthetic_rgbd_mask2former()[0].model.pixel_level_module.encoder.eval().cuda()
    x = torch.randn(2, 3, 224, 288, device="cuda")
    with torch.no_grad():
        exact = enc(x).feature_maps
        with torch.autocast("cuda", dtype=torch.bfloat16):
            stock = enc(x).feature_maps
            decoder_ops.install_fast_decoder_ops(enc, window_attention=False, layer_norm=False)
            before = fn.LAUNCHES
            ln_only = enc(x).feature_maps
            assert fn.LAUNCHES == before + 24                      # 12 blocks x (layernorm_before, layernorm_after)
            decoder_ops.uninstall_fast_decoder_ops(enc)
            decoder_ops.install_fast_decoder_ops(enc)
            both = enc(x).feature_maps
            decoder_ops.uninstall_fast_decoder_ops(enc)
    for i in range(4):
        e_stock = rel_l2(stock[i].float(), exact[i])
        assert stock[i].dtype == ln_only[i].dtype == both[i].dtype
        # bf16 pre-norm outputs are what autocast hands the Linear layers anyway: same error level as the stock autocast run
        assert rel_l2(ln_only[i].float(), exact[i]) <= max(e_stock * 1.25, 1e-3), (i, e_stock)
        assert rel_l2(both[i].float(), exact[i]) <= max(e_stock * 1.25, 1e-3), (i, e_stock)


@pytest.mark.parametrize("S", [300, 1200, 4800, 70])
def test_masked_cross_attention_module_matches_stock_forward(fn, S):
    """nn.MultiheadAttention as the Mask2Former decoder layer calls it (boolean mask, batch_first=False) with its rebound forward
    under bf16 autocast, against the stock forward and the float32 result."""
    from rgbd_b200 import decoder_ops
    torch.manual_seed(S)
    mha = torch.nn.MultiheadAttention(256, 8, 0.0).cuda().eval()
    torch.nn.init.normal_(mha.in_proj_bias, std=0.2)
    B, L = 3, 100
    query = torch.randn(L, B, 256, device="cuda")
    key = torch.randn(S, B, 256, device="cuda")
    value = torch.randn(S, B, 256, device="cuda")
    mask = torch.rand(B * 8, L, S, device="cuda") < 0.6
    mask[:, :, 5] = False                                      # no fully masked row (the decoder's `where` fix guarantees it)
    mask[:, 7, :] = True
    mask[:, 7, S - 1] = False                                  # a row with a single visible key, in the last tile
    with torch.no_grad():                                      # install() looks for decoder layers; rebind by hand here
        exact = mha(query, key, value, attn_mask=mask)[0]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            stock = mha(query, key, value, attn_mask=mask)[0]
            import types
            mha._rgbd_stock_forward = mha.forward
            mha.forward = types.MethodType(decoder_ops._mha_cross_attention_forward, mha)
            before = fn.LAUNCHES
            fast, w = mha(query, key, value, attn_mask=mask)
            assert fn.LAUNCHES == before + 1 and w is None
        fp32_call = mha(query, key, value, attn_mask=mask)[0]     # no autocast: stock path
    assert fast.dtype == stock.dtype and fast.shape == stock.shape
    assert torch.equal(fp32_call, exact)
    e_stock, e_fast = rel_l2(stock.float(), exact), rel_l2(fast.float(), exact)
    assert e_fast <= max(e_stock * 1.1, 4e-3), (e_fast, e_stock)
    assert rel_l2(fast.float(), stock.float()) < 1e-2


def test_masked_cross_attention_fully_masked_rows_give_zeros(fn):
    q = torch.randn(100, 2, 256, device="cuda").bfloat16()
    k = torch.randn(300, 2, 256, device="cuda").bfloat16()
    v = torch.randn(300, 2, 256, device="cuda").bfloat16()
    mask = torch.zeros(16, 100, 300, dtype=torch.bool, device="cuda")
    mask[:, 3] = True
    out = fn.masked_cross_attention(q, k, v, mask, 8)
    assert bool((out[3] == 0).all()) and bool(torch.isfinite(out.float()).all())
    want = torch.nn.functional.scaled_dot_product_attention(q.float().view(100, 16, 32).transpose(0, 1), k.float().view(300, 16, 32).transpose(0, 1),
                                                            v.float().view(300, 16, 32).transpose(0, 1))
    want = want.transpose(0, 1).reshape(100, 2, 256)
    keep = [i for i in range(100) if i != 3]
    assert rel_l2(out[keep].float(), want[keep]) < 4e-3


def test_whole_model_under_autocast_with_all_fast_ops(fn):
    """The serving configuration: bf16 autocast with every decoder_ops kernel active (bf16 pre-norms, window attention, deformable
    attention, attention masks, masked cross-attention) against the untouched model under the same autocast, both measured against
    the float32 forward of the stock model: the rebound model must not be further from float32 than the stock autocast run."""
    import numpy as np
    from rgbd_b200 import decoder_ops, synthetic, synthetic_weights
    model = synthetic_weights.build_synthetic_rgbd_mask2former(decisive=True)[0].eval().cuda()
    frames = [synthetic.synth_rgbd_u8(60 + j, 192, 256) for j in range(2)]
    rgb = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
    depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
    pv = fn.pack_pixel_values(rgb, depth)
    with torch.no_grad():
        exact = model(pixel_values=pv)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            stock = model(pixel_values=pv)
            decoder_ops.install_fast_decoder_ops(model)
            before = fn.LAUNCHES
            fast = model(pixel_values=pv)
            used = fn.LAUNCHES - before
            decoder_ops.uninstall_fast_decoder_ops(model)
            before = fn.LAUNCHES
            model(pixel_values=pv)
            hot = fn.LAUNCHES - before
    assert used - hot >= 6 + 10 + 2 * 12 + 24 + 9            # + 24 pre-norm LayerNorms + 9 masked cross-attentions (+ one per generic LayerNorm call)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _parity_report import report
    for name in ("masks_queries_logits", "class_queries_logits"):
        e_stock = rel_l2(getattr(stock, name).float(), getattr(exact, name))
        e_fast = rel_l2(getattr(fast, name).float(), getattr(exact, name))
        report("decoder_ops_autocast_whole_model", tensor=name, stock_autocast_vs_fp32=e_stock, decoder_ops_autocast_vs_fp32=e_fast)
        assert e_fast <= max(1.5 * e_stock, 1e-2), (name, e_fast, e_stock)
    # per-pixel winner among the queries: the decision the post-processing takes
    agree_stock = float((stock.masks_queries_logits.argmax(1) == exact.masks_queries_logits.argmax(1)).float().mean())
    agree_fast = float((fast.masks_queries_logits.argmax(1) == exact.masks_queries_logits.argmax(1)).float().mean())
    report("decoder_ops_autocast_whole_model", pixel_winner_agreement_with_fp32_stock=agree_stock, decoder_ops=agree_fast)
    assert agree_fast >= agree_stock - 0.02, (agree_fast, agree_stock)


def test_generic_layer_norm_rebind_keeps_torch_semantics(fn):
    """Every nn.LayerNorm that is not a Swin pre-norm: float32 result for float32 input and for bf16 input under autocast, stock
    path otherwise (autograd, odd widths)."""
    from rgbd_b200 import decoder_ops
    holder = torch.nn.Sequential(torch.nn.LayerNorm(256), torch.nn.LayerNorm(30)).cuda().eval()
    for ln in holder:
        torch.nn.init.normal_(ln.weight, std=1.0)
        torch.nn.init.normal_(ln.bias, std=1.0)
    x = torch.randn(7, 33, 256, device="cuda")
    y = torch.randn(5, 30, device="cuda")
    with torch.no_grad():
        want, want_odd = holder[0](x), holder[1](y)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            want_auto = holder[0](x.bfloat16())
        decoder_ops.install_fast_decoder_ops(holder)
        before = fn.LAUNCHES
        got, got_odd = holder[0](x), holder[1](y)
        assert fn.LAUNCHES == before + 1                       # width 30: stock path
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got_auto = holder[0](x.bfloat16())
        assert fn.LAUNCHES == before + 2
    assert got.dtype == want.dtype == torch.float32 and rel_l2(got, want) < 1e-6 and torch.equal(got_odd, want_odd)
    assert got_auto.dtype == want_auto.dtype == torch.float32 and rel_l2(got_auto, want_auto) < 1e-6
    xg = x.clone().requires_grad_(True)
    out = holder[0](xg)                                        # autograd: stock forward
    out.sum().backward()
    assert xg.grad is not None
    decoder_ops.uninstall_fast_decoder_ops(holder)
