"""CPU restatement of the ranking of unknown-object proposals (enhanced unknown localisation).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/ood_utils.py:1031-1084: RoIAlign (1x1 output, spatial_scale 1, aligned=False) of the proposals
on the padded feature map of the selected stride, `activations_transformation` (row normalisation, :2404-2409),
`compute_distance` to the clusters of every class that has some on that stride (:2422-2430), then the fold over the
classes chosen by CUSTOM_HYP.unk.rank.RANK_BOXES_OPERATION.  Pinned by tests/golden/golden_eul_rank.npz, produced with the
reference's own method objects (tests/golden/make_golden.py::golden_eul_rank).
"""
from __future__ import annotations

import numpy as np

from . import distance, roi_align


def distance_matrix(feature_map, proposals, clusters, stride, metric):
    """-> [classes with clusters on `stride`, P] distances to the nearest centroid."""
    fm = np.asarray(feature_map, dtype=np.float32)
    boxes = np.asarray(proposals, dtype=np.float32).reshape(-1, 4)
    rois = np.concatenate([np.zeros((len(boxes), 1), np.float32), boxes], 1)      # (batch_idx, x1, y1, x2, y2)
    feats = roi_align.roi_align_1x1(fm[None], rois, 1.0)
    x = distance.normalize_rows(feats)
    rows = [distance.compute_distance(np.asarray(per_cls[stride], dtype=np.float32), x, metric)
            for per_cls in clusters if len(per_cls[stride]) > 0]
    return np.array(rows)


def fold(d, operation, use_ood_thr_to_remove_props=False):
    """ood_utils.py:1057-1084."""
    d = np.asarray(d)
    if operation == "mean":
        return d.mean(axis=0)
    if operation == "max":
        return d.max(axis=0)
    if operation == "sum":
        return d.sum(axis=0)
    if operation == "min":
        if use_ood_thr_to_remove_props:
            return d.min(axis=0), np.argsort(d, axis=0)[0]
        return d.min(axis=0) * 100
    if operation == "geometric_mean":
        return np.exp(np.log(d).mean(axis=0))
    if operation == "entropy":
        pk = d / d.sum(axis=0)
        pk = pk / pk.sum(axis=0)
        return -(pk * np.log(pk)).sum(axis=0)
    raise NotImplementedError(operation)
