"""Synthetic NYUv2-shaped RGB-D frames, numpy only.

``map_10channel_case2`` (reference DL:386-425) produces ``pixel_values (10,H,W)``:
[0:3] ImageNet-normalised RGB, [3:6] ImageNet-normalised depth-as-RGB, [6:9] the
normalised Sobel magnitude (3 identical channels), [9] the valid-gradient mask.  There is
no dataset here, so frames are drawn from a frozen ``RandomState`` stream: a floor ramp,
a few boxes at random depths, sensor noise and 5 % invalid (zero) pixels give the
multi-modal depth histogram real indoor scenes have.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np

# image_mean / image_std exactly as the reference's checkpoint stores them (mask2former/checkpoints/standard/
# preprocessor_config.json).  They are float32-rounded ImageNet statistics; std[1] is 0.2239999920129776, ONE ULP BELOW
# float32(0.224) -- using the decimal 0.224 changes 27 % of the normalised green/depth values by one ulp.
IMAGE_MEAN = np.array([0.48500001430511475, 0.4560000002384186, 0.4059999883174896], dtype=np.float32)
IMAGE_STD = np.array([0.2290000021457672, 0.2239999920129776, 0.22499999403953552], dtype=np.float32)


def synth_rgbd_u8(frame_idx: int, height: int = 480, width: int = 640, kind: str = "nyu"
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (rgb uint8 (H,W,3), depth uint8 (H,W)).  ``kind`` selects the stress cases
    of SURVEY.md §8d: nyu | uniform | constant | two_valued | all_invalid."""
    rs = np.random.RandomState(1234 + frame_idx)
    rgb = rs.randint(0, 256, size=(height, width, 3)).astype(np.uint8)
    if kind == "nyu":
        ramp = 40.0 + 170.0 * (np.arange(height, dtype=np.float64)[:, None] / height)
        d = np.broadcast_to(ramp, (height, width)).copy()
        for _ in range(rs.randint(3, 7)):
            bh = rs.randint(height // 8, height // 2)
            bw = rs.randint(width // 8, width // 2)
            y0 = rs.randint(0, height - bh)
            x0 = rs.randint(0, width - bw)
            d[y0:y0 + bh, x0:x0 + bw] = rs.randint(30, 231)
        d += rs.standard_normal(size=d.shape) * 2.0
        d = np.clip(np.rint(d), 1, 255)
        d[rs.uniform(size=d.shape) < 0.05] = 0
        depth = d.astype(np.uint8)
    elif kind == "uniform":
        depth = rs.randint(0, 256, size=(height, width)).astype(np.uint8)
    elif kind == "constant":
        depth = np.full((height, width), 128, dtype=np.uint8)
    elif kind == "two_valued":
        depth = np.where(rs.uniform(size=(height, width)) < 0.5, 60, 200).astype(np.uint8)
    elif kind == "all_invalid":
        depth = np.zeros((height, width), dtype=np.uint8)
    else:
        raise ValueError(f"unknown synthetic kind {kind!r}")
    return rgb, depth


RESCALE_FACTOR = 0.00392156862745098       # preprocessor_config.json "rescale_factor" (1/255 as a double)


def normalise_u8(img_hwc_u8: np.ndarray) -> np.ndarray:
    """The Hugging Face image processor's arithmetic as ``map_10channel_case2`` (DL:405-410) invokes it:
    ``rescale`` = float32(float64(x) * rescale_factor), then ``normalize`` = (x - mean) / std in float32 with
    float32 mean/std (transformers.image_transforms.rescale / normalize).  Returns (3,H,W) float32."""
    x = (img_hwc_u8.astype(np.float64) * RESCALE_FACTOR).astype(np.float32)
    x = (x - IMAGE_MEAN) / IMAGE_STD
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float32)


def assemble_pixel_values(rgb_u8: np.ndarray, depth_u8: np.ndarray,
                          gradient_features: Callable[[np.ndarray], Tuple[np.ndarray, ...]]) -> np.ndarray:
    """(10,H,W) float32 exactly as DL:405-421 lays it out.  ``gradient_features`` maps the float32
    depth image to (normalised magnitude, ..., valid mask) -- the caller supplies it (the GPU
    front-end in the product, the oracle in tests)."""
    rgb = normalise_u8(rgb_u8)
    depth3 = normalise_u8(np.repeat(depth_u8[:, :, None], 3, axis=2))
    out = gradient_features(depth_u8.astype(np.float32))
    norm, vmask = out[0], out[-1]
    return np.concatenate([rgb, depth3, np.stack([norm] * 3, 0), vmask[None]], axis=0).astype(np.float32)
