"""Host-side mirror of the post-processing / evaluation interfaces the reference uses.

* ``post_process_instance_segmentation`` has the signature and return structure of the HuggingFace
  ``Mask2FormerImageProcessor`` method the reference calls (mask2former/utils/model_essential_part.py:86-91,
  mask2former/predictor.py:34-36 and 701-703); the work runs in csrc/postproc.cu with ONE device->host read of the small
  per-segment tables at the end (HF syncs once per query through ``.item()``).  Segments come in a defined order (class
  score descending, flattened (query, label) index on ties); HF's ``topk(sorted=False)`` order is unspecified.
* ``postprocess_prediction_batch`` / ``MaskAP`` mirror ``Evaluator.postprocess_prediction_batch`` and the segm mAP the
  reference gets from torchmetrics (model_essential_part.py:83-170): IoU on the device, the COCO-style accumulation
  (101-point interpolated AP over IoU 0.50:0.05:0.95, per class, <= 100 detections per image) on the host.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import functional as Fn
from ._lib import RgbdB200Error


def _group_by_size(target_sizes: Sequence[Tuple[int, int]]) -> Dict[Tuple[int, int], List[int]]:
    groups: Dict[Tuple[int, int], List[int]] = {}
    for i, s in enumerate(target_sizes):
        groups.setdefault((int(s[0]), int(s[1])), []).append(i)
    return groups


def post_process_instance_segmentation(outputs, threshold: float = 0.5, mask_threshold: float = 0.5,
                                       overlap_mask_area_threshold: float = 0.8,
                                       target_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                                       return_coco_annotation: bool = False,
                                       return_binary_maps: bool = False) -> List[Dict]:
    """Drop-in for ``image_processor.post_process_instance_segmentation``.  ``mask_threshold`` and
    ``overlap_mask_area_threshold`` are accepted and unused, exactly as in HF's instance routine."""
    if return_coco_annotation and return_binary_maps:
        raise ValueError("return_coco_annotation and return_binary_maps can not be both set to True.")
    if return_coco_annotation:
        raise NotImplementedError("run-length encoded output is not built; use the segmentation map or binary maps")
    cls = outputs.class_queries_logits
    msk = outputs.masks_queries_logits
    if not (cls.is_cuda and msk.is_cuda):
        raise RgbdB200Error("post_process_instance_segmentation: model outputs must be CUDA tensors (no CPU path)")
    cls = cls.detach().float().contiguous()
    msk = msk.detach().float().contiguous()
    B = cls.shape[0]
    sizes = [(384, 384)] * B if target_sizes is None else [tuple(int(v) for v in s) for s in target_sizes]
    if len(sizes) != B:
        raise ValueError("Make sure that you pass in as many target sizes as the batch dimension of the logits")
    results: List[Optional[Dict]] = [None] * B
    for size, idxs in _group_by_size(sizes).items():
        sel = torch.as_tensor(idxs, device=cls.device)
        whole = len(idxs) == B
        r = Fn.post_process_instances(cls if whole else cls[sel].contiguous(), msk if whole else msk[sel].contiguous(),
                                      threshold, size, want_segmentation=not return_binary_maps)
        count = r.count.cpu().tolist()                 # the one synchronisation of the batch
        labels, scores = r.labels.cpu(), r.scores.cpu()
        for k, i in enumerate(idxs):
            n = count[k]
            info = [{"id": j, "label_id": int(labels[k, j]), "was_fused": False, "score": round(float(scores[k, j]), 6)}
                    for j in range(n)]
            if return_binary_maps and n:
                seg = r.masks[k, :n].float()
            elif return_binary_maps:
                seg = torch.zeros(size, device=cls.device) - 1
            else:
                seg = r.segmentation[k].float()
            results[i] = {"segmentation": seg, "segments_info": info}
    return results


def postprocess_prediction_batch(prediction_batch, target_sizes, threshold: float = 0.0) -> List[Dict[str, torch.Tensor]]:
    """``Evaluator.postprocess_prediction_batch`` (model_essential_part.py:83-109): dictionaries with "masks" (bool),
    "labels", "scores" per image."""
    from types import SimpleNamespace
    out = post_process_instance_segmentation(
        SimpleNamespace(class_queries_logits=prediction_batch[0], masks_queries_logits=prediction_batch[1]),
        threshold=threshold, target_sizes=target_sizes, return_binary_maps=True)
    res = []
    for image_predictions, target_size in zip(out, target_sizes):
        if image_predictions["segments_info"]:
            res.append({"masks": image_predictions["segmentation"].to(dtype=torch.bool),
                        "labels": torch.tensor([x["label_id"] for x in image_predictions["segments_info"]]),
                        "scores": torch.tensor([x["score"] for x in image_predictions["segments_info"]])})
        else:
            res.append({"masks": torch.zeros([0, *target_size], dtype=torch.bool), "labels": torch.tensor([]),
                        "scores": torch.tensor([])})
    return res


class MaskAP:
    """Accumulates (prediction, target) pairs and computes segm mAP / mAP@50 / mAP@75 like the reference's
    ``MeanAveragePrecision(iou_type="segm")`` use (model_essential_part.py:111-170).  IoU matrices come from the device
    kernel; the matching and precision/recall integration are small host loops."""

    def __init__(self, thresholds: Sequence[float] = tuple(0.5 + 0.05 * i for i in range(10)), max_det: int = 100):
        self.thresholds = [float(t) for t in thresholds]
        self.max_det = max_det
        self.records: List[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = []

    def update(self, preds: Sequence[Dict[str, torch.Tensor]], targets: Sequence[Dict[str, torch.Tensor]]) -> None:
        for p, t in zip(preds, targets):
            pm, gm = p["masks"], t["masks"]
            if pm.shape[0] and gm.shape[0]:
                iou = Fn.mask_iou(pm.cuda().contiguous(), gm.cuda().contiguous()).cpu().numpy().astype(np.float64)
            else:
                iou = np.zeros((pm.shape[0], gm.shape[0]))
            self.records.append((np.asarray(p["labels"].cpu(), dtype=np.int64), np.asarray(p["scores"].cpu(), dtype=np.float64),
                                 np.asarray(t["labels"].cpu(), dtype=np.int64), iou))

    def compute(self) -> Dict[str, float]:
        classes = sorted({int(c) for r in self.records for c in r[0]} | {int(c) for r in self.records for c in r[2]})
        rec_pts = np.linspace(0.0, 1.0, 101)
        ap = np.full((len(self.thresholds), len(classes)), -1.0)
        for ci, c in enumerate(classes):
            n_gt = sum(int((gl == c).sum()) for _, _, gl, _ in self.records)
            if n_gt == 0:
                continue
            for ti, thr in enumerate(self.thresholds):
                score_list, tp_list = [], []
                for pl, ps, gl, iou in self.records:
                    det = np.nonzero(pl == c)[0]
                    det = det[np.argsort(-ps[det], kind="mergesort")][:self.max_det]
                    gts = np.nonzero(gl == c)[0]
                    free = np.ones(len(gts), dtype=bool)
                    for d in det:
                        hit = -1
                        best = min(thr, 1 - 1e-10)
                        for gi, g in enumerate(gts):
                            if free[gi] and iou[d, g] >= best:
                                best, hit = iou[d, g], gi
                        if hit >= 0:
                            free[hit] = False
                        score_list.append(ps[d])
                        tp_list.append(hit >= 0)
                if not score_list:
                    ap[ti, ci] = 0.0
                    continue
                order = np.argsort(-np.asarray(score_list), kind="mergesort")
                tp = np.asarray(tp_list, dtype=np.float64)[order]
                ctp, cfp = np.cumsum(tp), np.cumsum(1.0 - tp)
                recall = ctp / n_gt
                precision = ctp / np.maximum(ctp + cfp, np.finfo(np.float64).eps)
                precision = np.maximum.accumulate(precision[::-1])[::-1]
                at = np.searchsorted(recall, rec_pts, side="left")
                ap[ti, ci] = float(np.where(at < len(precision), precision[np.minimum(at, len(precision) - 1)], 0.0).mean())

        def mean_valid(a: np.ndarray) -> float:
            v = a[a > -1]
            return float(v.mean()) if v.size else -1.0
        thr = [round(t, 2) for t in self.thresholds]
        return {"map": mean_valid(ap), "map_50": mean_valid(ap[thr.index(0.5)]) if 0.5 in thr else -1.0,
                "map_75": mean_valid(ap[thr.index(0.75)]) if 0.75 in thr else -1.0}
