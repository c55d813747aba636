"""CPU oracle of the instance post-processing the reference runs on the model outputs -- TEST
INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU leg; never by the product path).

The reference calls a third-party routine it does not vendor:
``image_processor.post_process_instance_segmentation(outputs, threshold=..., target_sizes=..., return_binary_maps=True)``
(mask2former/utils/model_essential_part.py:86-91, mask2former/predictor.py:34-36 and 701-703) of HuggingFace
``transformers`` (``Mask2FormerImageProcessor``; the algorithm below restates transformers 5.5.x,
models/mask2former/image_processing_mask2former.py).  Parity pin: tests/golden/postproc.npz holds the outputs of that
HF routine itself on seeded inputs (oracle/make_golden_postproc.py).

One deliberate, documented difference: HF selects the candidates with ``topk(..., sorted=False)`` whose ORDER is
unspecified (it differs between torch's CPU and CUDA kernels).  The order only decides the list order of the returned
segments / their ids.  The oracle -- and the device kernel -- use a defined order: descending class score, ties broken by
the smaller flattened (query, label) index.  Comparisons against HF are therefore made per (query, label) key.

``mask_iou`` / ``average_precision`` restate thinly what the reference's evaluator gets from torchmetrics'
``MeanAveragePrecision(iou_type="segm")`` (mask2former/utils/model_essential_part.py:111-170; torchmetrics and
pycocotools are not installed here): COCO-style 101-point interpolated AP over IoU thresholds 0.50:0.05:0.95, greedy
matching in descending score order, per class, all areas, at most 100 detections per image.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

GRID = 384      # HF: "Scale back to preprocessed image size - (384, 384) for all models"


def select_candidates(class_logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(Q, C+1) logits -> top-Q of the Q*C (query, label) scores: (scores, query index, label), in the defined order."""
    q, c1 = class_logits.shape
    c = c1 - 1
    scores = F.softmax(class_logits.float(), dim=-1)[:, :-1].reshape(-1)
    flat = torch.arange(q * c)
    # descending score, ascending flat index on ties (lexsort: last key is primary)
    order = np.lexsort((flat.numpy(), -scores.numpy().astype(np.float64)))[:q]
    order = torch.from_numpy(order.astype(np.int64))
    return scores[order], order // c, order % c


def post_process_image(class_logits: torch.Tensor, mask_logits: torch.Tensor, threshold: float = 0.5,
                       target_size: Optional[Tuple[int, int]] = None) -> Dict[str, torch.Tensor]:
    """One image: class_logits (Q, C+1), mask_logits (Q, h, w).  Returns the kept segments in the defined order:
    masks (n, Ht, Wt) bool, labels (n,), scores (n,) float32 (unrounded), query (n,), segmentation (Ht, Wt) int32 with
    -1 background and the id of the LAST kept segment covering a pixel (HF paints in list order)."""
    q = class_logits.shape[0]
    up = F.interpolate(mask_logits.float()[None], size=(GRID, GRID), mode="bilinear", align_corners=False)[0]
    cls_score, query, label = select_candidates(class_logits)
    mp = up[query]
    pm = (mp > 0).float()
    mask_score = (mp.sigmoid().flatten(1) * pm.flatten(1)).sum(1) / (pm.flatten(1).sum(1) + 1e-6)
    pred = cls_score * mask_score
    size = (GRID, GRID) if target_size is None else tuple(int(v) for v in target_size)
    if target_size is not None:
        pm = F.interpolate(pm[None], size=size, mode="nearest")[0]
    keep = [j for j in range(q) if bool(pm[j].any()) and float(pred[j]) >= threshold]
    seg = torch.full(size, -1, dtype=torch.int32)
    for sid, j in enumerate(keep):
        seg[pm[j] == 1] = sid
    idx = torch.tensor(keep, dtype=torch.int64)
    return {"masks": pm[idx].bool() if keep else torch.zeros((0,) + size, dtype=torch.bool),
            "labels": label[idx], "scores": pred[idx], "query": query[idx], "segmentation": seg}


def post_process_instance_segmentation(class_logits: torch.Tensor, mask_logits: torch.Tensor, threshold: float = 0.5,
                                       target_sizes: Optional[Sequence[Tuple[int, int]]] = None) -> List[Dict]:
    return [post_process_image(class_logits[i], mask_logits[i], threshold,
                               None if target_sizes is None else target_sizes[i])
            for i in range(class_logits.shape[0])]


def mask_iou(pred: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """(P,H,W) bool x (G,H,W) bool -> (P,G) float64 IoU (0 where the union is empty)."""
    npix = int(np.prod(pred.shape[1:]))
    p = pred.reshape(pred.shape[0], npix).astype(np.int64)
    g = gt.reshape(gt.shape[0], npix).astype(np.int64)
    inter = p @ g.T
    union = p.sum(1)[:, None] + g.sum(1)[None, :] - inter
    return np.where(union > 0, inter / np.maximum(union, 1), 0.0)


def average_precision(preds: Sequence[Dict], targets: Sequence[Dict], ious: Sequence[np.ndarray],
                      thresholds: Sequence[float] = tuple(0.5 + 0.05 * i for i in range(10)),
                      max_det: int = 100) -> Dict[str, float]:
    """COCO-style segm mAP from per-image IoU matrices.  preds[i]: labels (P,), scores (P,); targets[i]: labels (G,);
    ious[i]: (P, G).  Returns map, map_50, map_75."""
    classes = sorted({int(c) for t in targets for c in np.asarray(t["labels"]).tolist()} |
                     {int(c) for p in preds for c in np.asarray(p["labels"]).tolist()})
    rec_pts = np.linspace(0.0, 1.0, 101)
    ap = np.full((len(thresholds), len(classes)), -1.0)
    for ci, c in enumerate(classes):
        n_gt = sum(int((np.asarray(t["labels"]) == c).sum()) for t in targets)
        if n_gt == 0:
            continue
        for ti, thr in enumerate(thresholds):
            scores, tps = [], []
            for p, t, iou in zip(preds, targets, ious):
                pl, ps = np.asarray(p["labels"]), np.asarray(p["scores"], dtype=np.float64)
                gl = np.asarray(t["labels"])
                pi = np.nonzero(pl == c)[0]
                pi = pi[np.argsort(-ps[pi], kind="mergesort")][:max_det]
                gi = np.nonzero(gl == c)[0]
                taken = np.zeros(len(gi), dtype=bool)
                for k in pi:
                    best, best_j = min(thr, 1 - 1e-10), -1
                    for jj, g in enumerate(gi):
                        if taken[jj] or iou[k, g] < best:
                            continue
                        best, best_j = iou[k, g], jj
                    scores.append(ps[k])
                    tps.append(best_j >= 0)
                    if best_j >= 0:
                        taken[best_j] = True
            if not scores:
                ap[ti, ci] = 0.0
                continue
            order = np.argsort(-np.asarray(scores), kind="mergesort")
            tp = np.asarray(tps, dtype=np.float64)[order]
            ctp, cfp = np.cumsum(tp), np.cumsum(1 - tp)
            rec = ctp / n_gt
            prec = ctp / np.maximum(ctp + cfp, np.finfo(np.float64).eps)
            for k in range(len(prec) - 1, 0, -1):
                prec[k - 1] = max(prec[k - 1], prec[k])
            idx = np.searchsorted(rec, rec_pts, side="left")
            ap[ti, ci] = float(np.mean([prec[i] if i < len(prec) else 0.0 for i in idx]))

    def mean_valid(a):
        v = a[a > -1]
        return float(v.mean()) if v.size else -1.0
    thr = [round(t, 2) for t in thresholds]
    return {"map": mean_valid(ap),
            "map_50": mean_valid(ap[thr.index(0.5)]) if 0.5 in thr else -1.0,
            "map_75": mean_valid(ap[thr.index(0.75)]) if 0.75 in thr else -1.0}
