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

from transformers.backbone_utils import load_backbone

from .modules import (DepthGradientInjectionResidual, DSAModule, EnhancedDepthImageRatioPredictor, RatioPredictor,
                      depth_guidance_forward)


class CustomMask2FormerPixelLevelModule(Mask2FormerPixelLevelModule):
    main_input_name = "pixel_values"

    DGGM_ONLY = ("0.0.3", "0.0.4", "0.0.5", "0.0.6")     # CM:73-75, 156-163: 7-channel input, DGGM on the raw features
    DSAM_ONLY = ("0.1.2",)                                # CM:97-100, 234-256: 6-channel input, DSAM cascade, ratio 0.1
    DEPTH_ENCODER = ("0.1.3", "0.3.0")                    # CM:101-116, 258-322: second backbone on the depth image feeds the
    #                                                       feature-based RatioPredictor; 0.3.0 then applies DGGM to the DSAM result
    SUPPORTED = ("0.0.0", "0.4.0") + DGGM_ONLY + DSAM_ONLY + DEPTH_ENCODER

    def __init__(self, config, version: str = "0.4.0"):
        super().__init__(config)
        if version not in self.SUPPORTED:
            raise NotImplementedError(
                f"rgbd_b200 builds the version branches made of DGGM / E-DSAM ({', '.join(self.SUPPORTED)}); the dual-backbone "
                f"and surface-normal ablations are out of scope (SURVEY section 2 row 7); got {version!r}")
        self.version = version
        c = list(self.encoder.channels)
        if version == "0.4.0":
            self.ratio_predictor = EnhancedDepthImageRatioPredictor(3)
        if version in self.DEPTH_ENCODER:
            self.depth_encoder = load_backbone(config)
            self.ratio_predictor = RatioPredictor(depth_channels_list=c)
        if version == "0.4.0" or version in self.DSAM_ONLY or version in self.DEPTH_ENCODER:
            self.dsam0 = DSAModule(in_channels=c[0], out_channels=c[1], num_depth_regions=3)
            self.dsam1 = DSAModule(in_channels=c[1], out_channels=c[2], num_depth_regions=3)
            self.dsam2 = DSAModule(in_channels=c[2], out_channels=c[3], num_depth_regions=3)
        if version in ("0.4.0", "0.3.0") or version in self.DGGM_ONLY:
            self.depth_gradient_injection = DepthGradientInjectionResidual(c, 3)
        #: SURVEY 8f-2 (inference only, off by default): run the pixel decoder's ``input_projections`` (Conv2d(C_i,256,1) +
        #: GroupNorm(32,256) of the three coarse levels) and its FPN lateral ``adapter_1`` on this library's kernels and hand
        #: the stock decoder the projected maps; everything after the projections stays stock HF code.
        self.fuse_input_projections = False
        self._proj_cache = {}

    def _decode(self, backbone_features, output_hidden_states):
        """``self.decoder(backbone_features)`` (CM:383), optionally with the input projections computed here."""
        dec = self.decoder
        if not (self.fuse_input_projections and not torch.is_grad_enabled() and backbone_features[0].is_cuda):
            return dec(backbone_features, output_hidden_states=output_hidden_states)
        from . import functional as Fn
        n_lv = dec.num_feature_levels
        feats = list(backbone_features)
        projected = {}
        for level, idx in enumerate(range(len(feats) - 1, len(feats) - 1 - n_lv, -1)):       # coarsest first, like HF
            conv, gn = dec.input_projections[level][0], dec.input_projections[level][1]
            projected[idx] = Fn.project_group_norm(feats[idx].float().contiguous(), conv.weight, conv.bias, gn.weight, gn.bias,
                                                   gn.num_groups, gn.eps, self._proj_cache.setdefault(("in", level), {}))
        lateral = {}
        for k, lat in enumerate(dec.lateral_convolutions):                                  # features[:num_fpn_levels][::-1]
            idx = dec.num_fpn_levels - 1 - k
            lateral[idx] = Fn.project_group_norm(feats[idx].float().contiguous(), lat[0].weight, lat[0].bias, lat[1].weight,
                                                 lat[1].bias, lat[1].num_groups, lat[1].eps, self._proj_cache.setdefault(("lat", k), {}))
        dt = backbone_features[0].dtype
        handed = [(projected.get(i, lateral.get(i, f))).to(dt) for i, f in enumerate(feats)]
        # the stock forward applies input_projections[level](x) / lateral_conv(x): make those the identity for this call
        saved_in, saved_lat = dec.input_projections, dec.lateral_convolutions
        try:
            dec.input_projections = torch.nn.ModuleList([torch.nn.Identity() for _ in saved_in])
            dec.lateral_convolutions = [torch.nn.Identity() for _ in saved_lat]
            return dec(handed, output_hidden_states=output_hidden_states)
        finally:
            dec.input_projections, dec.lateral_convolutions = saved_in, saved_lat

    def _dsam_only(self, pixel_values: Tensor, feats, ratio=None):
        """CM:234-256 (ratio 0.1) / CM:258-290 (predicted ratios) batched: cp[k+1] += dsam_k(cp[k], gray(depth), ratio)
        with no detach."""
        from . import functional as Fn
        B = pixel_values.shape[0]
        if ratio is None:
            ratio = torch.full((B,), 0.1, device=pixel_values.device, dtype=torch.float32)   # CM:647 default argument
        ratio = ratio.reshape(-1).float().contiguous()
        cp = [f.float().contiguous() for f in feats]
        dec = Fn.depth_decompose(ratio, [tuple(f.shape[2:]) for f in cp[:3]], depth3=pixel_values[:, 3:6].float())
        for k, dsam in enumerate((self.dsam0, self.dsam1, self.dsam2)):
            cp[k + 1] = dsam.stage_forward(cp[k], dec.pooled[k], dec.bias_variant, residual=cp[k + 1])
        return cp

    def to_grayscale(self, image_data):
        """CM:392-502.  ndarray (H,W,3) / (3,H,W) / (H,W,1) / (1,H,W) / (H,W) -> (H,W) ndarray of the input dtype (host
        arithmetic, as in the reference); tensor (3,H,W) / (H,W,3) / (1,H,W) / (H,W,1) -> (1,H,W), (B,C,H,W) -> (B,1,H,W);
        CUDA tensors go through the device kernel (``(0.299 r + 0.587 g) + 0.114 b`` in float32, bit-exact for float32 and
        integer inputs)."""
        import numpy as np
        from . import functional as Fn
        if isinstance(image_data, np.ndarray):
            if image_data.ndim == 3:
                shape = image_data.shape
                if shape[-1] == 3:
                    gray = 0.299 * image_data[:, :, 0] + 0.587 * image_data[:, :, 1] + 0.114 * image_data[:, :, 2]
                elif shape[0] == 3:
                    gray = 0.299 * image_data[0] + 0.587 * image_data[1] + 0.114 * image_data[2]
                elif shape[-1] == 1 or shape[0] == 1:
                    gray = image_data.squeeze()
                else:
                    raise ValueError("Input NumPy ndarray image should be RGB (H, W, 3) or (C, H, W) with C=3, or grayscale "
                                     "(H, W, 1), (1, H, W) or (H, W).")
            elif image_data.ndim == 2:
                gray = image_data
            else:
                raise ValueError("Input NumPy ndarray image should be 2D (H, W) or 3D (H, W, C) or (C, H, W).")
            return gray.astype(image_data.dtype)
        if not isinstance(image_data, torch.Tensor):
            raise TypeError("Input image_data must be NumPy ndarray or PyTorch Tensor.")
        t = image_data
        if t.ndim == 4:
            channels, batched = t.shape[1], True
        elif t.ndim == 3:
            batched = False
            if t.shape[0] == 3:
                channels = 3
            elif t.shape[-1] == 3:
                channels, t = 3, t.permute(2, 0, 1)
            elif t.shape[0] == 1:
                channels = 1
            elif t.shape[-1] == 1:
                channels, t = 1, t.permute(2, 0, 1)
            else:
                raise ValueError("Input PyTorch Tensor image with ndim=3, but cannot determine channel location.")
        else:
            raise ValueError("Input PyTorch Tensor image should be 3D (C, H, W) or (H, W, 3) or 4D (B, C, H, W).")
        if channels == 1:
            return t
        if channels != 3:
            raise ValueError("Input PyTorch Tensor image should be grayscale or RGB (channels 1 or 3).")
        x = t if batched else t[None]
        # float32 (the dtype of pixel_values) is bit-exact with the reference.  Other dtypes are computed in float32 and cast
        # back (CM:499 ``.to(image_data.dtype)``): identical for integer tensors, whose products torch promotes to float32
        # anyway; float16 / bfloat16 / float64 tensors differ from the reference's per-operation rounding in the last bits.
        gray = Fn.to_grayscale(x.float().contiguous()).to(t.dtype)
        return gray[:, None] if batched else gray

    def forward(self, pixel_values: Tensor, output_hidden_states: bool = False) -> Mask2FormerPixelLevelModuleOutput:
        rgb = pixel_values[:, 0:3, :, :]
        color_feature_map = self.encoder(rgb).feature_maps                                  # CM:330
        if self.version == "0.0.0":
            backbone_features = list(color_feature_map)                                     # CM:145-146
        elif self.version in self.DGGM_ONLY:
            backbone_features = self.depth_gradient_injection(
                [f.float() for f in color_feature_map], pixel_values[:, 3:6].float(), pixel_values[:, 6:7].float())
        elif self.version in self.DSAM_ONLY:
            backbone_features = self._dsam_only(pixel_values, color_feature_map)
        elif self.version in self.DEPTH_ENCODER:
            depth_feature_map = self.depth_encoder(pixel_values[:, 3:6, :, :]).feature_maps              # CM:262, 299
            predicted_ratios = self.ratio_predictor(list(depth_feature_map))                             # CM:266, 304
            backbone_features = self._dsam_only(pixel_values, color_feature_map, predicted_ratios)       # CM:272-290
            if self.version == "0.3.0":                                                                  # CM:322
                backbone_features = self.depth_gradient_injection(backbone_features, pixel_values[:, 6:9].float(),
                                                                  pixel_values[:, 9:10].float())
        else:
            backbone_features = depth_guidance_forward(
                self.ratio_predictor, (self.dsam0, self.dsam1, self.dsam2), self.depth_gradient_injection,
                pixel_values.float(), color_feature_map)                                    # CM:332-355
        backbone_features = [f.to(color_feature_map[0].dtype) for f in backbone_features]
        decoder_output = self._decode(backbone_features, output_hidden_states)                      # CM:383
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
    # The reference builds its pixel-level module inside the model's __init__, i.e. BEFORE Hugging Face's post_init() walks
    # the tree (CM:45-53); swapped in afterwards it would keep torch's default initialisers and, worse, raw
    # ``nn.Parameter(torch.Tensor(...))`` members such as the pixel decoder's ``level_embed`` would stay UNINITIALISED memory.
    # post_init() initialises exactly the modules that are not flagged as initialised yet.
    model.post_init()
    return model
