"""Golden vectors of the data mapper's front-end (map_10channel_case2, mask2former/utils/dataloader.py:386-425) from the
real thing: HuggingFace's PIL-backend Mask2Former image processor loaded from the reference's own
checkpoints/standard/preprocessor_config.json (rescale + normalize; frames already at model resolution so its resize is
the identity) and the reference's calculate_gradient_features -> tests/golden/frontend.npz.
Run in the build container: ``python oracle/make_golden_frontend.py``."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import GOLD, REF, import_reference, load_pkg      # noqa: E402


def main():
    cm, dp = import_reference()
    synthetic, _ = load_pkg()
    from transformers import AutoImageProcessor
    proc = AutoImageProcessor.from_pretrained(os.path.join(REF, "mask2former/checkpoints/standard"), backend="pil")
    out = {}
    for j, (h, w, kind) in enumerate([(64, 96, "nyu"), (96, 64, "uniform"), (32, 32, "two_valued")]):
        rgb, depth = synthetic.synth_rgbd_u8(300 + j, h, w, kind)
        depth_colorful = np.repeat(depth[:, :, None], 3, axis=2)           # PIL 'L' -> 'RGB' replicates the channel
        pv = proc(images=[rgb, depth_colorful], do_resize=False, return_tensors="np").pixel_values
        assert pv.shape == (2, 3, h, w), pv.shape
        norm, gx, gy, vmask = dp.calculate_gradient_features(depth)        # DL:414 on the (here unresized) depth
        full = np.concatenate([pv.reshape(6, h, w), np.stack([norm, norm, norm]), vmask[None]]).astype(np.float32)
        out[f"f{j}.rgb"] = rgb
        out[f"f{j}.depth"] = depth
        out[f"f{j}.pixel_values"] = full
    out["image_mean"] = np.array(proc.image_mean, dtype=np.float64)
    out["image_std"] = np.array(proc.image_std, dtype=np.float64)
    out["rescale_factor"] = np.float64(proc.rescale_factor)
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **out)
    print("frontend.npz", os.path.getsize(os.path.join(GOLD, "frontend.npz")), type(proc).__name__)


if __name__ == "__main__":
    main()
