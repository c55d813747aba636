"""Golden set, section 10: the REFERENCE's ``EnhancedDepthImageRatioPredictor`` in ``.train()`` mode (batch-statistics
BatchNorm with running-stat updates + Dropout; mask2former/utils/custom_model.py:1369-1487) -> tests/golden/ratio_train.npz.
Run in the build container: ``python oracle/make_golden_train.py``.

The Dropout keep-masks are seeded numpy draws injected through forward hooks on the two ``nn.Dropout`` modules
(``x * keep / (1 - p)`` is torch's own definition of train-mode dropout), so the fixture does not depend on torch's RNG
stream.  Two consecutive steps are stored: outputs of both and the buffers after each."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import GOLD, import_reference, load_pkg   # noqa: E402

HW, B, SEED_W = (48, 64), 3, 500


def train_inputs(synthetic, step):
    frames = []
    for j in range(B):
        _, d = synthetic.synth_rgbd_u8(70 + 10 * step + j, HW[0], HW[1], "nyu" if j < 2 else "uniform")
        frames.append(synthetic.normalise_u8(np.repeat(d[:, :, None], 3, axis=2)))
    rs = np.random.RandomState(900 + step)
    keep = (torch.from_numpy(rs.rand(B, 128) >= 0.3), torch.from_numpy(rs.rand(B, 64) >= 0.2))
    return torch.from_numpy(np.stack(frames)), keep


def main():
    cm, _ = import_reference()
    synthetic, weights = load_pkg()
    m = cm.EnhancedDepthImageRatioPredictor(3)
    m.load_state_dict(weights.ratio_weights(seed=SEED_W))
    m.train()
    cur = {}
    for idx, j in ((2, 0), (5, 1)):
        drop = m.fc_layers[idx]
        assert isinstance(drop, torch.nn.Dropout)
        drop.register_forward_hook(lambda mod, inp, out, j=j: inp[0] * (cur["keep"][j].float() / (1.0 - mod.p)))
    out = {}
    for step in range(2):
        x, keep = train_inputs(synthetic, step)
        cur["keep"] = keep
        with torch.no_grad():
            r = m(x)
        out[f"step{step}.ratio"] = r.numpy()
        for k, v in m.state_dict().items():
            if "running" in k or "num_batches" in k:
                out[f"step{step}.{k}"] = v.numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "ratio_train.npz"), **out)
    print("ratio_train.npz", os.path.getsize(os.path.join(GOLD, "ratio_train.npz")), out["step0.ratio"].ravel(), out["step1.ratio"].ravel())


if __name__ == "__main__":
    main()
