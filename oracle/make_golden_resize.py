"""Golden vectors for the resize half of the data mapper's front-end -> tests/golden/resize.npz, from the real libraries:
Pillow ``Image.resize(BILINEAR)``, OpenCV ``cv2.resize(INTER_LINEAR)``, and the whole map_10channel_case2 pipeline
(mask2former/utils/dataloader.py:386-425) through HuggingFace's PIL-backend Mask2Former processor loaded from the
reference's preprocessor_config.json (with a smaller ``size`` to keep the fixture small) + the reference's
calculate_gradient_features.  Run in the build container: ``python oracle/make_golden_resize.py``."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import GOLD, REF, import_reference, load_pkg      # noqa: E402

CASES = [  # name, source (H, W), target (h, w)
    ("down", (120, 160), (96, 96)),          # the real use: 480x640 -> 384x384 at a quarter of the size
    ("up", (30, 40), (96, 96)),
    ("odd", (61, 83), (48, 72)),
    ("wide", (40, 200), (64, 50)),           # shrink by 4 in x (9-tap filter), grow in y
    ("same", (32, 32), (32, 32)),
]


def main():
    import cv2
    from PIL import Image
    cm, dp = import_reference()
    synthetic, _ = load_pkg()
    out = {}
    for j, (name, (H, W), (h, w)) in enumerate(CASES):
        rgb, depth = synthetic.synth_rgbd_u8(400 + j, H, W, "nyu" if j % 2 == 0 else "uniform")
        out[f"{name}.rgb"] = rgb
        out[f"{name}.depth"] = depth
        out[f"{name}.rgb_pil"] = np.array(Image.fromarray(rgb).resize((w, h), resample=Image.BILINEAR))
        out[f"{name}.depth_pil"] = np.array(Image.fromarray(depth).resize((w, h), resample=Image.BILINEAR))
        out[f"{name}.depth_cv"] = cv2.resize(depth, (w, h), interpolation=cv2.INTER_LINEAR)
    # the whole mapper on one frame: HF processor (resize + rescale + normalise) on [colour, depth as RGB], cv2 resize +
    # gradient features on the depth
    from transformers import AutoImageProcessor
    proc = AutoImageProcessor.from_pretrained(os.path.join(REF, "mask2former/checkpoints/standard"), backend="pil")
    H, W, h, w = 120, 160, 96, 96
    rgb, depth = synthetic.synth_rgbd_u8(450, H, W, "nyu")
    depth_colorful = np.array(Image.fromarray(depth).convert("RGB"))
    pv = proc(images=[rgb, depth_colorful], size={"height": h, "width": w}, return_tensors="np").pixel_values
    assert pv.shape == (2, 3, h, w), pv.shape
    resized_depth = cv2.resize(depth, (w, h), interpolation=cv2.INTER_LINEAR)      # DL:413 (square target: argument order moot)
    norm, gx, gy, vmask = dp.calculate_gradient_features(resized_depth)
    out["mapper.rgb"] = rgb
    out["mapper.depth"] = depth
    out["mapper.pixel_values"] = np.concatenate([pv.reshape(6, h, w), np.stack([norm] * 3), vmask[None]]).astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "resize.npz"), **out)
    print("resize.npz", os.path.getsize(os.path.join(GOLD, "resize.npz")))


if __name__ == "__main__":
    main()
