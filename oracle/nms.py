"""NMS-with-payload oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates the default path of /root/reference/ultralytics/utils/ops.py:348-530 (`non_max_suppression_old`: best class only,
no apriori labels, no class filter, class-aware or agnostic, not v10) including the payload the reference threads through it for the OoD
methods -- the per-anchor `extra_item` rows (raw class logits) and `strides` -- and torchvision's `nms` (greedy suppression in
score order, IoU in float32, strict `>` against the threshold) that it calls at :489.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def nms_keep(boxes: np.ndarray, iou_thres: float) -> np.ndarray:
    """torchvision.ops.nms for boxes already in descending score order (its stable sort keeps that order)."""
    boxes = np.asarray(boxes, F32)
    n = len(boxes)
    areas = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    supp = np.zeros(n, bool)
    keep = []
    thr = F32(iou_thres)
    for i in range(n):
        if supp[i]:
            continue
        keep.append(i)
        j = np.arange(i + 1, n)
        w = np.maximum(F32(0), np.minimum(boxes[i, 2], boxes[j, 2]) - np.maximum(boxes[i, 0], boxes[j, 0]))
        h = np.maximum(F32(0), np.minimum(boxes[i, 3], boxes[j, 3]) - np.maximum(boxes[i, 1], boxes[j, 1]))
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[j] - inter)
        supp[j[ovr > thr]] = True
    return np.asarray(keep, np.int64)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300, max_nms=30000, max_wh=7680,
                        extra_item=None, strides=None, agnostic=False):
    """prediction [bs, 4 + nc, A] (cx, cy, w, h, class confidences); extra_item [bs, E, A]; strides [A].
    -> (list of [k, 6] arrays (xyxy, conf, cls), list of [k, E] payload rows, list of [k] strides)."""
    pred = np.asarray(prediction, F32)
    bs, nc = pred.shape[0], pred.shape[1] - 4
    out, extras, strs = [], [], []
    for xi in range(bs):
        x = pred[xi].T                                             # [A, 4 + nc]
        cand = x[:, 4:].max(1) > F32(conf_thres)                   # :412
        x = x[cand]
        e = np.asarray(extra_item[xi], F32).T[cand] if extra_item is not None else None
        s = np.asarray(strides)[cand] if strides is not None else None
        if not len(x):
            out.append(np.zeros((0, 6), F32))
            extras.append(np.zeros((0,), F32))
            strs.append(np.zeros((0,), F32))
            continue
        half = x[:, 2:4] / F32(2)                                  # xywh2xyxy :645-649
        box = np.concatenate([x[:, :2] - half, x[:, :2] + half], 1)
        conf = x[:, 4:].max(1)
        cls = x[:, 4:].argmax(1).astype(F32)
        order = np.argsort(-conf, kind="stable")[:max_nms]         # :478-482 (ties: anchor order)
        det = np.concatenate([box, conf[:, None], cls[:, None]], 1)[order]
        keep = nms_keep(det[:, :4] + det[:, 5:6] * F32(0 if agnostic else max_wh), iou_thres)[:max_det]     # :485-489
        out.append(det[keep])
        extras.append(e[order][keep] if e is not None else None)
        strs.append(s[order][keep] if s is not None else None)
    return out, extras, strs
