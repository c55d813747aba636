"""Oracle of the instance post-processing (oracle/postproc.py) against the outputs of HuggingFace's own routine stored in
tests/golden/postproc.npz (oracle/make_golden_postproc.py), and the thin mAP restatement on hand-checkable cases."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postproc as OP                                  # noqa: E402
from oracle.make_golden_postproc import CASES, synth_outputs       # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "postproc.npz"))


def match_segments(got, b, name):
    """Every HF segment (label, rounded score, mask) must be matched one-to-one by a segment of ``got``."""
    n = int(GOLD[f"{name}.{b}.n"])
    assert len(got["labels"]) == n, (name, b, len(got["labels"]), n)
    if n == 0:
        return
    shape = tuple(GOLD[f"{name}.{b}.shape"])
    ref_masks = np.unpackbits(GOLD[f"{name}.{b}.masks"])[:int(np.prod(shape))].reshape(shape).astype(bool)
    used = set()
    gm = np.asarray(got["masks"].cpu()) if torch.is_tensor(got["masks"]) else got["masks"]
    gl = np.asarray(got["labels"].cpu() if torch.is_tensor(got["labels"]) else got["labels"])
    gs = np.asarray(got["scores"].cpu() if torch.is_tensor(got["scores"]) else got["scores"], dtype=np.float64)
    for k in range(n):
        lab, sc = int(GOLD[f"{name}.{b}.labels"][k]), float(GOLD[f"{name}.{b}.scores"][k])
        cands = [j for j in range(n) if j not in used and int(gl[j]) == lab and abs(gs[j] - sc) < 2e-6
                 and np.array_equal(gm[j], ref_masks[k])]
        assert cands, (name, b, k, lab, sc)
        used.add(cands[0])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_huggingface_routine(case):
    name, seed, B, Q, C, hw, thr, tgt = case
    cls, masks = synth_outputs(seed, B, Q, C, *hw)
    res = OP.post_process_instance_segmentation(cls, masks, thr, None if tgt is None else [tgt] * B)
    for b, r in enumerate(res):
        match_segments(r, b, name)
        # defined order: class score descending; segmentation map = id of the last kept segment covering the pixel
        seg = np.full(r["segmentation"].shape, -1, dtype=np.int32)
        for sid in range(len(r["labels"])):
            seg[r["masks"][sid].numpy()] = sid
        assert np.array_equal(seg, r["segmentation"].numpy())


def test_candidate_order_is_descending_with_index_ties():
    cls = torch.zeros(3, 3)                     # all scores tie: flat index order
    s, q, l = OP.select_candidates(cls)
    assert q.tolist() == [0, 0, 1] and l.tolist() == [0, 1, 0]
    cls = torch.tensor([[0.0, 5.0, 0.0], [4.0, 0.0, 0.0]])
    s, q, l = OP.select_candidates(cls)
    assert (q[0], l[0]) == (0, 1) and (q[1], l[1]) == (1, 0) and float(s[0]) > float(s[1])


def test_mask_iou_and_average_precision_hand_cases():
    a = np.zeros((2, 4, 4), bool)
    a[0, :2] = True
    a[1, 2:] = True
    g = np.zeros((2, 4, 4), bool)
    g[0, :2] = True                              # identical to a[0]
    g[1, 1:3] = True                             # half overlap with both
    iou = OP.mask_iou(a, g)
    assert iou[0, 0] == 1.0 and iou[1, 0] == 0.0
    assert abs(iou[0, 1] - 4 / 12) < 1e-12 and abs(iou[1, 1] - 4 / 12) < 1e-12
    assert OP.mask_iou(np.zeros((1, 2, 2), bool), np.zeros((1, 2, 2), bool))[0, 0] == 0.0
    preds = [{"labels": np.array([0, 0]), "scores": np.array([0.9, 0.8])}]
    tgts = [{"labels": np.array([0, 0])}]
    r = OP.average_precision(preds, tgts, [iou])
    # one of two ground truths is found at every threshold, by the top-scoring detection: AP = 0.5 * (51/101 recall pts)
    assert abs(r["map"] - 51 / 101) < 1e-9 and abs(r["map_50"] - 51 / 101) < 1e-9
    perfect = OP.average_precision(preds, tgts, [np.eye(2)])
    assert abs(perfect["map"] - 1.0) < 1e-12
    assert OP.average_precision([{"labels": np.array([1]), "scores": np.array([0.5])}], tgts, [np.zeros((1, 2))])["map"] == 0.0
