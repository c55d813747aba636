"""Opt-in inference kernels inside the stock Hugging Face modules on either side of the hot path: the pixel decoder and the
transformer decoder that CONSUME the fused pyramid (reference call site mask2former/utils/custom_model.py:383
``self.decoder(backbone_features)`` and the transformer module behind it) and the Swin encoder that PRODUCES its input
(CM:330).  The module tree, parameter names and state_dict stay Hugging Face's: ``install_fast_decoder_ops`` only rebinds the
``forward`` of

* ``Mask2FormerPixelDecoderEncoderMultiscaleDeformableAttention`` (six encoder layers): the linear layers stay
  ``nn.Linear`` (cuBLAS); softmax + sampling locations + ``multi_scale_deformable_attention`` (per level grid_sample, stack,
  multiply, sum: ~80 ms per 32 frames of 480x640) run as ONE kernel, ``rgbd_msda_fwd``;
* ``Mask2FormerMaskPredictor`` (ten calls per forward): the attention mask (bilinear resize of the (B,Q,120,160) logits, sigmoid,
  threshold, repeat per head: ~2 ms per call in ATen's plane-serial upsample kernel) is ``rgbd_attention_mask``;
* ``Mask2FormerMaskedAttentionDecoderLayer.cross_attn`` (``nn.MultiheadAttention`` with a boolean mask, nine layers) under bf16
  autocast: the additive -inf mask, the (batch*heads, 100, S) float32 score tensor, softmax, cast and second bmm are
  ``rgbd_masked_cross_attention`` (FlashAttention-2 style, the boolean mask read directly);
* ``SwinSelfAttention`` (12 blocks of Swin-T): the three ``nn.Linear`` projections stay; bmm -> div -> + relative position bias ->
  + shift mask -> softmax -> cast -> bmm -> permute-copy over a (windows*heads, 49, 49) score tensor is ``rgbd_window_attention``
  (one warp per (window, head), softmax in registers);
* ``SwinLayer.layernorm_before`` / ``layernorm_after`` under bf16 autocast: ``rgbd_layer_norm`` writes the bf16 tensor the
  following ``nn.Linear`` layers would have cast the float32 result to (same values, 6 instead of 14+ bytes per element, and the
  pad / roll / window-partition copies in between move half the bytes);
* every other ``nn.LayerNorm`` (same output dtype as torch): the one-pass ``rgbd_layer_norm`` is 3-8x faster than ATen's kernel
  at these shapes (many rows, 96-256 channels).

All of them fall back to the stock forward when autograd is recording (these are inference kernels: run under
``torch.no_grad()``), when the tensors are not on CUDA, or when a shape / dtype is outside what the kernel covers.
"""
from __future__ import annotations

import types

import torch
from torch import nn

from . import functional as Fn


def _msda_attention_forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                            position_embeddings=None, reference_points=None, spatial_shapes_list=None, level_start_index=None,
                            output_attentions: bool = False):
    if torch.is_grad_enabled() or not hidden_states.is_cuda \
            or output_attentions or reference_points is None or reference_points.shape[-1] != 2:
        return self._rgbd_stock_forward(hidden_states, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                                        encoder_attention_mask=encoder_attention_mask, position_embeddings=position_embeddings,
                                        reference_points=reference_points, spatial_shapes_list=spatial_shapes_list,
                                        level_start_index=level_start_index, output_attentions=output_attentions)
    if position_embeddings is not None:
        hidden_states = hidden_states + position_embeddings
    if hidden_states.dtype == torch.float32 and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        # the two nn.Linear below would each cast this tensor to bf16 (autocast caches weight casts only): cast it once
        hidden_states = hidden_states.to(torch.bfloat16)
    batch_size, num_queries, _ = hidden_states.shape
    _, sequence_length, _ = encoder_hidden_states.shape
    if sum(h * w for h, w in spatial_shapes_list) != sequence_length:
        raise ValueError("Make sure to align the spatial shapes with the sequence length of the encoder hidden states")
    value = self.value_proj(encoder_hidden_states)
    if attention_mask is not None:
        value = value.masked_fill(attention_mask[..., None], float(0))
    value = value.view(batch_size, sequence_length, self.n_heads, self.d_model // self.n_heads)
    offsets = self.sampling_offsets(hidden_states).view(batch_size, num_queries, self.n_heads, self.n_levels, self.n_points, 2)
    logits = self.attention_weights(hidden_states).view(batch_size, num_queries, self.n_heads, self.n_levels * self.n_points)
    # the kernel emits what output_proj would read: under autocast nn.Linear casts its input to the autocast dtype anyway
    out_dtype = value.dtype if value.dtype == torch.bfloat16 else torch.float32
    if value.dtype not in (torch.float32, torch.bfloat16):
        value, offsets, logits = value.float(), offsets.float(), logits.float()
    output = Fn.msda_forward(value.contiguous(), [tuple(s) for s in spatial_shapes_list], offsets.contiguous(), logits.contiguous(),
                             reference_points=reference_points.float().contiguous(), softmax=True, out_dtype=out_dtype)
    return self.output_proj(output), None


def _mask_predictor_forward(self, outputs, pixel_embeddings, attention_mask_target_size=None):
    if torch.is_grad_enabled() or not outputs.is_cuda or attention_mask_target_size is None:
        return self._rgbd_stock_forward(outputs, pixel_embeddings, attention_mask_target_size)
    mask_embeddings = self.mask_embedder(outputs.transpose(0, 1))
    outputs_mask = torch.einsum("bqc, bchw -> bqhw", mask_embeddings, pixel_embeddings)
    logits = outputs_mask if outputs_mask.dtype in (torch.float32, torch.bfloat16) else outputs_mask.float()
    size = attention_mask_target_size
    size = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    return outputs_mask, Fn.attention_mask(logits.contiguous(), size, self.num_heads)


def _swin_self_attention_forward(self, hidden_states, attention_mask=None, output_attentions=False):
    n_tok = hidden_states.shape[1]
    if torch.is_grad_enabled() or not hidden_states.is_cuda or output_attentions or self.attention_head_size != 32 or n_tok > 64 \
            or (self.training and self.dropout.p > 0):
        return self._rgbd_stock_forward(hidden_states, attention_mask, output_attentions)
    q, k, v = self.query(hidden_states), self.key(hidden_states), self.value(hidden_states)
    if q.dtype not in (torch.float32, torch.bfloat16):
        q, k, v = q.float(), k.float(), v.float()
    table = self.relative_position_bias_table
    key = (table.data_ptr(), table._version, table.device)
    cache = self.__dict__.get("_rgbd_bias_cache")
    if cache is None or cache[0] != key:          # (heads, N, N) float32, as the stock forward gathers it on every call
        bias = table.detach()[self.relative_position_index.view(-1)].view(n_tok, n_tok, -1).permute(2, 0, 1).float().contiguous()
        cache = self.__dict__["_rgbd_bias_cache"] = (key, bias)
    mask = attention_mask.float().contiguous() if attention_mask is not None else None
    ctx = Fn.window_attention(q.contiguous(), k.contiguous(), v.contiguous(), cache[1], mask, self.num_attention_heads)
    return (ctx,)


def _mha_cross_attention_forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None,
                                 average_attn_weights=True, is_causal=False):
    """nn.MultiheadAttention.forward for the decoder's masked cross-attention (boolean attn_mask, batch_first=False) under bf16
    autocast: the three input projections and out_proj stay F.linear / nn.Linear, the attention core is one kernel.  The attention
    weights are not computed (the decoder layer only forwards them when output_attentions is set -- then the stock path runs)."""
    ok = (not torch.is_grad_enabled() and query.is_cuda and attn_mask is not None and attn_mask.dtype == torch.bool
          and attn_mask.dim() == 3 and key_padding_mask is None and not is_causal and not self.batch_first
          and self._qkv_same_embed_dim and self.in_proj_bias is not None and self.bias_k is None and not self.add_zero_attn
          and self.head_dim == 32 and not (self.training and self.dropout > 0) and query.dim() == 3 and key.shape[0] % 2 == 0
          and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16)
    if not ok:
        return self._rgbd_stock_forward(query, key, value, key_padding_mask=key_padding_mask, need_weights=need_weights,
                                        attn_mask=attn_mask, average_attn_weights=average_attn_weights, is_causal=is_causal)
    e = self.embed_dim
    w, b = self.in_proj_weight, self.in_proj_bias
    q = nn.functional.linear(query, w[:e], b[:e])
    k = nn.functional.linear(key, w[e:2 * e], b[e:2 * e])
    v = nn.functional.linear(value, w[2 * e:], b[2 * e:])
    if q.dtype != torch.bfloat16:
        return self._rgbd_stock_forward(query, key, value, key_padding_mask=key_padding_mask, need_weights=need_weights,
                                        attn_mask=attn_mask, average_attn_weights=average_attn_weights, is_causal=is_causal)
    ctx = Fn.masked_cross_attention(q.contiguous(), k.contiguous(), v.contiguous(), attn_mask.contiguous(), self.num_heads)
    return self.out_proj(ctx), None


def _swin_prenorm_forward(self, x):
    """SwinLayer.layernorm_before / layernorm_after under bf16 autocast: the float32 LayerNorm result is only ever moved (pad,
    roll, window partition) and then cast to bf16 by nn.Linear -- emit that bf16 tensor directly.  (The residual stream is float32
    in the first stage and bf16 afterwards -- SwinPatchMerging ends in an nn.Linear -- so both input dtypes occur.)"""
    c = x.shape[-1]
    if torch.is_grad_enabled() or not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16) or c % 4 or c > 1024 \
            or len(self.normalized_shape) != 1 \
            or self.weight is None or self.bias is None \
            or not (torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16):
        return self._rgbd_stock_forward(x)
    return Fn.layer_norm(x.contiguous(), self.weight.detach(), self.bias.detach(), self.eps, out_dtype=torch.bfloat16)


def _layer_norm_forward(self, x):
    """Any other nn.LayerNorm of the model (post-norms of the decoders, Swin's embedding / patch-merging / stage-output norms):
    same semantics as torch (float32 result for float32 input, and for bf16 input under autocast), one-pass kernel.  ATen's kernel
    needs 335 us for a (201 600, 256) float32 tensor and 935 us for (614 400, 96); this one 87 / 117 us."""
    c = x.shape[-1]
    auto = torch.is_autocast_enabled("cuda")
    if torch.is_grad_enabled() or not x.is_cuda or c % 4 or c > 1024 or len(self.normalized_shape) != 1 or self.weight is None \
            or self.bias is None or not (x.dtype == torch.float32 or (x.dtype == torch.bfloat16 and auto)):
        return self._rgbd_stock_forward(x)
    return Fn.layer_norm(x.contiguous(), self.weight.detach().float(), self.bias.detach().float(), self.eps, out_dtype=torch.float32)


def install_fast_decoder_ops(model: nn.Module, deformable_attention: bool = True, attention_mask: bool = True,
                             window_attention: bool = True, swin_prenorm_bf16: bool = True,
                             masked_cross_attention: bool = True, layer_norm: bool = True) -> nn.Module:
    """Rebind the forwards described in the module docstring on every matching submodule of ``model`` (idempotent).
    ``uninstall_fast_decoder_ops`` restores the stock forwards."""
    from transformers.models.mask2former import modeling_mask2former as m2f
    from transformers.models.swin.modeling_swin import SwinLayer, SwinSelfAttention
    if swin_prenorm_bf16:
        for layer in model.modules():
            if isinstance(layer, SwinLayer):
                for ln in (layer.layernorm_before, layer.layernorm_after):
                    if isinstance(ln, nn.LayerNorm) and not hasattr(ln, "_rgbd_stock_forward"):
                        ln._rgbd_stock_forward = ln.forward
                        ln.forward = types.MethodType(_swin_prenorm_forward, ln)
    if masked_cross_attention:
        for layer in model.modules():
            if isinstance(layer, m2f.Mask2FormerMaskedAttentionDecoderLayer) and isinstance(layer.cross_attn, nn.MultiheadAttention) \
                    and not hasattr(layer.cross_attn, "_rgbd_stock_forward"):
                layer.cross_attn._rgbd_stock_forward = layer.cross_attn.forward
                layer.cross_attn.forward = types.MethodType(_mha_cross_attention_forward, layer.cross_attn)
    for mod in model.modules():
        if hasattr(mod, "_rgbd_stock_forward"):
            continue
        if deformable_attention and isinstance(mod, m2f.Mask2FormerPixelDecoderEncoderMultiscaleDeformableAttention):
            mod._rgbd_stock_forward = mod.forward
            mod.forward = types.MethodType(_msda_attention_forward, mod)
        elif attention_mask and isinstance(mod, m2f.Mask2FormerMaskPredictor):
            mod._rgbd_stock_forward = mod.forward
            mod.forward = types.MethodType(_mask_predictor_forward, mod)
        elif window_attention and isinstance(mod, SwinSelfAttention):
            mod._rgbd_stock_forward = mod.forward
            mod.forward = types.MethodType(_swin_self_attention_forward, mod)
        elif layer_norm and isinstance(mod, nn.LayerNorm):
            mod._rgbd_stock_forward = mod.forward
            mod.forward = types.MethodType(_layer_norm_forward, mod)
    return model


def uninstall_fast_decoder_ops(model: nn.Module) -> nn.Module:
    for mod in model.modules():
        if "_rgbd_stock_forward" in mod.__dict__:
            del mod.__dict__["forward"]
            del mod.__dict__["_rgbd_stock_forward"]
            mod.__dict__.pop("_rgbd_bias_cache", None)
    return model
