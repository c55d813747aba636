"""``CustomMask2FormerPixelLevelModule`` for ``version == "0.4.0"`` (reference
mask2former/utils/custom_model.py:56-134 ctor, :324-355 + :383-390 forward): the Hugging Face pixel-level
module (stock Swin encoder + stock pixel decoder) with the CUDA depth-guidance hot path in between.  Child
names are the reference's (``encoder, decoder, ratio_predictor, dsam0, dsam1, dsam2,
depth_gradient_injection``) so reference checkpoints load unchanged.  Unlike CM:129-134 the channel list is
taken from ``encoder.channels`` (Swin-B needs [128,256,512,1024])."""
from __future__ import annotations

import torch
from torch import Tensor
from transformers.models.mask2former.modeling_mask2former import (Mask2FormerPixelLevelModule,
                                                                  Mask2FormerPixelLevelModuleOutput)

from .modules import (DepthGradientInjectionResidual, DSAModule, EnhancedDepthImageRatioPredictor,
                      depth_guidance_forward)


class CustomMask2FormerPixelLevelModule(Mask2FormerPixelLevelModule):
    main_input_name = "pixel_values"

    def __init__(self, config, version: str = "0.4.0"):
        super().__init__(config)
        if version != "0.4.0":
            raise NotImplementedError(
                f"rgbd_b200 implements the paper's final variant (version '0.4.0', DGGM + E-DSAM); got {version!r}")
        self.version = version
        c = list(self.encoder.channels)
        self.ratio_predictor = EnhancedDepthImageRatioPredictor(3)
        self.dsam0 = DSAModule(in_channels=c[0], out_channels=c[1], num_depth_regions=3)
        self.dsam1 = DSAModule(in_channels=c[1], out_channels=c[2], num_depth_regions=3)
        self.dsam2 = DSAModule(in_channels=c[2], out_channels=c[3], num_depth_regions=3)
        self.depth_gradient_injection = DepthGradientInjectionResidual(c, 3)

    def forward(self, pixel_values: Tensor, output_hidden_states: bool = False) -> Mask2FormerPixelLevelModuleOutput:
        rgb = pixel_values[:, 0:3, :, :]
        color_feature_map = self.encoder(rgb).feature_maps                                  # CM:330
        backbone_features = depth_guidance_forward(
            self.ratio_predictor, (self.dsam0, self.dsam1, self.dsam2), self.depth_gradient_injection,
            pixel_values.float(), color_feature_map)                                        # CM:332-355
        backbone_features = [f.to(color_feature_map[0].dtype) for f in backbone_features]
        decoder_output = self.decoder(backbone_features, output_hidden_states=output_hidden_states)   # CM:383
        return Mask2FormerPixelLevelModuleOutput(
            encoder_last_hidden_state=backbone_features[-1],
            encoder_hidden_states=tuple(backbone_features) if output_hidden_states else None,
            decoder_last_hidden_state=decoder_output.mask_features,
            decoder_hidden_states=decoder_output.multi_scale_features,
        )


def swin_tiny_mask2former_config(num_labels: int = 80, **overrides):
    """Mask2Former + Swin-T configuration equal to the reference's ``checkpoints/standard/config.json``
    hyper-parameters (depths [2,2,6,2], embed 96, 100 queries), built without reading the reference tree."""
    from transformers import Mask2FormerConfig, SwinConfig
    backbone = SwinConfig(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7,
                          drop_path_rate=0.3, out_features=["stage1", "stage2", "stage3", "stage4"])
    kw = dict(backbone_config=backbone, num_labels=num_labels, num_queries=100, feature_size=256, mask_feature_size=256,
              hidden_dim=256, encoder_layers=6, decoder_layers=10, num_attention_heads=8, dim_feedforward=2048)
    kw.update(overrides)
    return Mask2FormerConfig(**kw)


def build_rgbd_mask2former(config=None, version: str = "0.4.0"):
    """``Mask2FormerForUniversalSegmentation`` whose pixel-level module is the RGB-D one (what
    ``CustomMask2FormerForUniversalSegmentation(config, version)`` builds in the reference, CM:45-53)."""
    from transformers import Mask2FormerForUniversalSegmentation
    config = config or swin_tiny_mask2former_config()
    model = Mask2FormerForUniversalSegmentation(config)
    model.model.pixel_level_module = CustomMask2FormerPixelLevelModule(config, version=version)
    return model
