"""Prediction / ground-truth matching oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates /root/reference/ood_utils.py:233-292 (`OODMethod.match_predicted_boxes_to_targets`):
  torchvision `box_iou` (float32), the class-equality mask (:251-257), their product = assignment score matrix, scipy's
  `linear_sum_assignment(score, maximize=True)` (:283) and the reference's walk over the assignment (:288-291, quirk Q8: it
  indexes the score matrix with the POSITION in the assignment instead of the assigned row).

`lsap` restates scipy 1.18's solver (scipy/optimize/_lsap.c -> rectangular_lsap/rectangular_lsap.cpp, the shortest
augmenting path algorithm of Crouse 2016): rows are added one at a time, the frontier is scanned in the order of a
`remaining` list initialised in REVERSE column order, ties go to the first candidate in that scan unless a later one is an
unassigned column ("select one which gives us a new sink node").  IoU x mask matrices are mostly zeros, so the result
depends on exactly these rules; tests/test_host_logic.py holds this function to scipy on random and tie-heavy matrices.
"""
from __future__ import annotations

import numpy as np


def box_iou(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """torchvision.ops.box_iou in float32: [P, 4] x [G, 4] xyxy -> [P, G]."""
    a, b = np.asarray(a, np.float32).reshape(-1, 4), np.asarray(b, np.float32).reshape(-1, 4)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (area_a[:, None] + area_b[None, :] - inter)).astype(np.float32)


def lsap(cost: np.ndarray, maximize: bool = False):
    """scipy.optimize.linear_sum_assignment restated: -> (row_ind, col_ind), rows ascending."""
    cost = np.asarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    transpose = nc < nr
    if transpose:
        cost = cost.T.copy()
        nr, nc = nc, nr
    if maximize:
        cost = -cost
    u, v = np.zeros(nr), np.zeros(nc)
    spc = np.empty(nc)
    path = np.full(nc, -1, np.int64)
    col4row = np.full(nr, -1, np.int64)
    row4col = np.full(nc, -1, np.int64)
    for cur in range(nr):
        # augmenting path from row `cur`
        min_val = 0.0
        remaining = [nc - it - 1 for it in range(nc)]
        SR = np.zeros(nr, bool)
        SC = np.zeros(nc, bool)
        spc[:] = np.inf
        sink, i = -1, cur
        while sink == -1:
            index, lowest = -1, np.inf
            SR[i] = True
            for it, j in enumerate(remaining):
                r = min_val + cost[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            remaining[index] = remaining[-1]
            remaining.pop()
        # dual update
        u[cur] += min_val
        for i in range(nr):
            if SR[i] and i != cur:
                u[i] += min_val - spc[col4row[i]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - spc[j]
        # augment
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transpose:
        order = np.argsort(col4row)
        return col4row[order], order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row


def match_predictions(pred_xyxy, pred_cls, gt_xyxy, gt_cls, iou_threshold: float, compat: bool = True):
    """-> (valid_preds list, score matrix, (row_ind, col_ind)) of one image (ood_utils.py:247-291)."""
    score = box_iou(pred_xyxy, gt_xyxy) * (np.asarray(pred_cls)[:, None] == np.asarray(gt_cls)[None, :]).astype(np.float32)
    rows, cols = lsap(score, maximize=True)
    valid = []
    for i, (r, c) in enumerate(zip(rows, cols)):
        rr = i if compat else int(r)                       # Q8: the reference uses the position, not the assigned row
        if score[rr, c] > iou_threshold:
            valid.append(rr)
    return valid, score, (rows, cols)
