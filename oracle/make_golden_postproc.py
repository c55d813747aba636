"""Section 9 of the golden set: outputs of HuggingFace's own ``post_process_instance_segmentation`` (the routine the
reference calls at model_essential_part.py:86-91 / predictor.py:701-703) on seeded inputs -> tests/golden/postproc.npz.
Run in the build container (transformers is installed there): ``python oracle/make_golden_postproc.py``."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def synth_outputs(seed: int, B: int, Q: int, C: int, h: int, w: int):
    """Seeded class / mask logits: smooth blobs so the binarised masks are regions, sharp class scores."""
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(B, Q, C + 1, generator=g) * 3.0
    coarse = torch.randn(B, Q, max(h // 6, 2), max(w // 6, 2), generator=g) * 4.0
    masks = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=False)
    masks = masks + 0.3 * torch.randn(B, Q, h, w, generator=g) - 1.5
    masks[:, 0] = -5.0                     # an all-background query
    return cls.contiguous(), masks.contiguous()


CASES = [  # name, seed, B, Q, C, (h, w), threshold, target size
    ("up480", 11, 2, 20, 8, (30, 40), 0.5, (480, 640)),
    ("down120", 12, 2, 20, 8, (30, 40), 0.3, (120, 160)),
    ("native", 13, 1, 16, 5, (24, 24), 0.5, None),
    ("thr0", 14, 1, 12, 4, (20, 28), 0.0, (100, 132)),
]


def main():
    from transformers.models.mask2former.image_processing_mask2former import Mask2FormerImageProcessor
    proc = Mask2FormerImageProcessor()
    out = {}
    for name, seed, B, Q, C, hw, thr, tgt in CASES:
        cls, masks = synth_outputs(seed, B, Q, C, *hw)
        res = proc.post_process_instance_segmentation(
            SimpleNamespace(class_queries_logits=cls, masks_queries_logits=masks), threshold=thr,
            target_sizes=None if tgt is None else [tgt] * B, return_binary_maps=True)
        for b, r in enumerate(res):
            info = r["segments_info"]
            n = len(info)
            out[f"{name}.{b}.n"] = np.int64(n)
            out[f"{name}.{b}.labels"] = np.array([s["label_id"] for s in info], dtype=np.int64)
            out[f"{name}.{b}.scores"] = np.array([s["score"] for s in info], dtype=np.float64)
            if n:
                m = r["segmentation"].numpy().astype(bool)
                out[f"{name}.{b}.shape"] = np.array(m.shape, dtype=np.int64)
                out[f"{name}.{b}.masks"] = np.packbits(m, axis=None)
        print(name, [len(r["segments_info"]) for r in res])
    np.savez_compressed(os.path.join(GOLD, "postproc.npz"), **out)
    print("postproc.npz", os.path.getsize(os.path.join(GOLD, "postproc.npz")))


if __name__ == "__main__":
    sys.exit(main())
