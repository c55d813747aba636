"""Oracle of the two resizes of the data mapper (oracle/resize.py) against outputs of Pillow, OpenCV and the HF processor
themselves (tests/golden/resize.npz, oracle/make_golden_resize.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rgbd_b200  # noqa: E402,F401
from rgbd_b200 import synthetic  # noqa: E402
from oracle import hotpath as O, resize as R  # noqa: E402
from oracle.make_golden_resize import CASES  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "resize.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_resize_oracle_bit_exact(case):
    name, _, (h, w) = case
    assert np.array_equal(R.pil_bilinear_resize_u8(GOLD[f"{name}.rgb"], (h, w)), GOLD[f"{name}.rgb_pil"])
    assert np.array_equal(R.pil_bilinear_resize_u8(GOLD[f"{name}.depth"], (h, w)), GOLD[f"{name}.depth_pil"])
    assert np.array_equal(R.cv_linear_resize_u8(GOLD[f"{name}.depth"], (h, w)), GOLD[f"{name}.depth_cv"])


def mapper_oracle(rgb, depth, size):
    rgb_r = R.pil_bilinear_resize_u8(rgb, size)
    depth_pil = R.pil_bilinear_resize_u8(depth, size)
    depth_cv = R.cv_linear_resize_u8(depth, size)
    pv = synthetic.assemble_pixel_values(rgb_r, depth_pil, O.gradient_features)
    norm, _, _, vmask = O.gradient_features(depth_cv)
    pv[6:9] = norm
    pv[9] = vmask
    return pv


def test_whole_mapper_matches_hf_processor_and_reference():
    pv = mapper_oracle(GOLD["mapper.rgb"], GOLD["mapper.depth"], (96, 96))
    assert np.array_equal(pv, GOLD["mapper.pixel_values"])
