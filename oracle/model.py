"""CPU oracle of the WHOLE v0.4.0 model (test infrastructure; also the whole-model leg of bench.py's reference arm).

What the reference runs per frame (mask2former/predictor.py:19-36, :697-703; mask2former/utils/custom_model.py:324-390):
stock Hugging Face Swin encoder -> depth-guidance hot path (here: ``oracle.hotpath.depth_guidance_forward``, the pinned
CPU restatement) -> stock HF pixel decoder -> stock HF transformer decoder + heads -> HF's own
``Mask2FormerImageProcessor.post_process_instance_segmentation``.  Everything except the hot path is the same
``transformers`` code the reference imports; nothing here is on the product path.
"""
from __future__ import annotations

import copy
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import hotpath as O


def cpu_oracle_model(model, guidance_weights: Dict[str, torch.Tensor]):
    """Deep copy of ``model`` (a ``Mask2FormerForUniversalSegmentation`` with this repo's pixel-level module) on the CPU,
    whose pixel-level forward runs the oracle hot path with ``guidance_weights`` (state_dict keys of the reference's
    ``CustomMask2FormerPixelLevelModule`` children: ratio_predictor., dsam0-2., depth_gradient_injection.)."""
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerPixelLevelModuleOutput
    cpu_model = copy.deepcopy(model).cpu().float().eval()
    plm = cpu_model.model.pixel_level_module

    def oracle_forward(pixel_values, output_hidden_states=False):
        feats = plm.encoder(pixel_values[:, 0:3]).feature_maps                                   # CM:330
        fused, _ = O.depth_guidance_forward(guidance_weights, pixel_values, list(feats))         # CM:332-355
        dec = plm.decoder(fused, output_hidden_states=output_hidden_states)                      # CM:383
        return Mask2FormerPixelLevelModuleOutput(encoder_last_hidden_state=fused[-1], encoder_hidden_states=None,
                                                 decoder_last_hidden_state=dec.mask_features,
                                                 decoder_hidden_states=dec.multi_scale_features)
    plm.forward = oracle_forward
    return cpu_model


_PROC = None


def hf_post_process(class_logits: torch.Tensor, mask_logits: torch.Tensor, threshold: float,
                    target_sizes: Optional[Sequence[Tuple[int, int]]], return_binary_maps: bool = False) -> List[Dict]:
    """HF's routine exactly as the reference calls it (model_essential_part.py:86-91 with ``return_binary_maps=True``,
    predictor.py:701-703 without)."""
    global _PROC
    if _PROC is None:
        from transformers.models.mask2former.image_processing_mask2former import Mask2FormerImageProcessor
        _PROC = Mask2FormerImageProcessor()
    return _PROC.post_process_instance_segmentation(
        SimpleNamespace(class_queries_logits=class_logits, masks_queries_logits=mask_logits), threshold=threshold,
        target_sizes=list(target_sizes) if target_sizes is not None else None, return_binary_maps=return_binary_maps)


def predict(cpu_model, pixel_values: torch.Tensor, threshold: float = 0.5,
            target_sizes: Optional[Sequence[Tuple[int, int]]] = None, return_binary_maps: bool = False):
    """Forward + post-processing of one batch on the CPU.  Returns (HF results, class logits, mask logits)."""
    with torch.no_grad():
        out = cpu_model(pixel_values=pixel_values)
    res = hf_post_process(out.class_queries_logits, out.masks_queries_logits, threshold, target_sizes, return_binary_maps)
    return res, out.class_queries_logits, out.masks_queries_logits
