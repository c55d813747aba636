"""GPU parity of the resize kernels (csrc/resize.cu) and the whole device mapper through the C ABI: bit-exact against the
goldens produced by Pillow / OpenCV / the HF processor and against the oracle at camera resolution."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

from oracle import resize as R                                   # noqa: E402
from oracle.make_golden_resize import CASES                       # noqa: E402
from test_resize_oracle import mapper_oracle                      # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "resize.npz"))


@pytest.fixture(scope="module")
def fn():
    import rgbd_b200  # noqa: F401
    from rgbd_b200 import functional
    functional._lib.load()
    return functional


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_resize_kernels_match_pillow_and_opencv_goldens(fn, case):
    name, _, (h, w) = case
    rgb, depth = GOLD[f"{name}.rgb"], GOLD[f"{name}.depth"]
    # batch of two (second image flipped) to exercise the batch stride
    rgb2 = np.stack([rgb, rgb[::-1].copy()])
    depth2 = np.stack([depth, depth[:, ::-1].copy()])
    got = fn.resize_pil_bilinear(cu(rgb2), (h, w)).cpu().numpy()
    assert np.array_equal(got[0], GOLD[f"{name}.rgb_pil"])
    assert np.array_equal(got[1], R.pil_bilinear_resize_u8(rgb2[1], (h, w)))
    got_d = fn.resize_pil_bilinear(cu(depth2), (h, w)).cpu().numpy()
    assert np.array_equal(got_d[0], GOLD[f"{name}.depth_pil"])
    got_c = fn.resize_cv_linear(cu(depth2), (h, w)).cpu().numpy()
    assert np.array_equal(got_c[0], GOLD[f"{name}.depth_cv"])
    assert np.array_equal(got_c[1], R.cv_linear_resize_u8(depth2[1], (h, w)))


def test_device_mapper_matches_hf_processor_golden(fn):
    pv = fn.map_10channel(cu(GOLD["mapper.rgb"][None]), cu(GOLD["mapper.depth"][None]), (96, 96))
    assert pv.shape == (1, 10, 96, 96)
    assert torch.equal(pv[0].cpu(), torch.from_numpy(GOLD["mapper.pixel_values"]))


def test_device_mapper_at_camera_resolution(fn):
    """480x640 frames -> 384x384 pixel_values (the reference's real sizes), against the oracle."""
    import rgbd_b200  # noqa: F401
    from rgbd_b200 import synthetic
    frames = [synthetic.synth_rgbd_u8(900 + j, 480, 640, "nyu") for j in range(3)]
    rgb = np.stack([f[0] for f in frames])
    depth = np.stack([f[1] for f in frames])
    pv = fn.map_10channel(cu(rgb), cu(depth)).cpu().numpy()
    for j in range(3):
        assert np.array_equal(pv[j], mapper_oracle(rgb[j], depth[j], (384, 384))), j
    # size=None keeps the resolution: identical to the plain packer
    same = fn.map_10channel(cu(rgb[:1]), cu(depth[:1]), None)
    assert torch.equal(same, fn.pack_pixel_values(cu(rgb[:1]), cu(depth[:1])))
