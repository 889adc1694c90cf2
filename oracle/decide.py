"""Decide-on-results oracle (distance + logit methods, fusion).  TEST INFRASTRUCTURE.

Restates the control flow of /root/reference/ood_utils.py:
  * `DistanceMethod.compute_ood_decision_on_results` :2038-2134 and
    `_compute_ood_decision_for_one_result_from_roi_aligned_feature_maps` :2136-2180,
    including the quirks the reference exhibits:
      Q1  the class is looked up with the *in-stride* index (`bbox_idx = idx`, :2152-2154) and the
          per-image output list is stride-major;
      Q4  `if self.thresholds[cls][stride]:` -- `[]`, `0` and `0.0` all mean "no threshold -> OoD";
          a missing cluster gives distance 1000 (:2159-2164) which is still compared.
  * `compute_INDness_scores_on_results` :1498-1581 + `DistanceMethod.compute_indness` :1584-1620 (Q2: -1).
  * `LogitsMethod.compute_ood_decision_on_results` :1195-1208, `compute_INDness_scores_on_results` :1210-1222.
  * `FusionMethod.fuse_ood_decisions` :2906-2940, `TripleFusionMethod.fuse_ood_decisions` :3282-3301.

An image is a plain dict: maps (3 CHW f32 arrays), boxes [M,4], cls [M], strides [M],
logits [M,NC] (optional), img_hw (H, W).
"""
from __future__ import annotations

import numpy as np

from . import distance as D
from . import logits as L
from .roi_align import extract_roi_aligned_features_from_correct_stride

F32 = np.float32


def _has_cluster(clusters, c, s):
    return c < len(clusters) and len(clusters[c][s]) > 0


def distance_decisions(images, clusters, thresholds, metric, compat_q1=True, normalize=True,
                       return_details=False, transform=None):
    """-> List[List[int]] (1 = InD, 0 = OoD), per image in the reference's (stride-major) order.

    compat_q1=False gives the evidently intended behaviour: class of the box being scored,
    decisions returned in box order.
    With return_details also returns per-image lists of (distance, argmin, cls_used, stride, box_idx).
    transform(x [1, C], cls, stride) -> [1, d]: the `activations_transformation` of the SDR methods (:2501-2571, e.g.
    normalise -> trained embedding for ivis :2542-2548); replaces the plain row normalisation (:2404-2409).
    """
    decisions, details = [], []
    for im in images:
        feats = extract_roi_aligned_features_from_correct_stride(
            [m[None] for m in im["maps"]], [im["boxes"]], [im["strides"]], im["img_hw"])[0]   # :2056-2066
        cls_all = np.asarray(im["cls"])
        dec, det = [], []
        for s, (idx_in_stride, fm) in enumerate(feats):                                        # :2147
            if len(idx_in_stride) == 0:
                continue
            for j, v in enumerate(fm):                                                         # :2152
                box = int(idx_in_stride[j])
                c = int(cls_all[j]) if compat_q1 else int(cls_all[box])                        # :2153-2154 (Q1)
                if not _has_cluster(clusters, c, s):
                    dist, arg = 1000, -1                                                       # :2159-2164
                else:
                    x = v.reshape(1, -1)
                    if transform is not None:
                        x = np.asarray(transform(x, c, s), dtype=F32)                          # :2169 -> SDR override
                    else:
                        x = D.normalize_rows(x) if normalize else x.astype(F32)                # :2169 -> :2409
                    pw = D.pairwise(clusters[c][s], x, metric)
                    dist, arg = pw.min(axis=0)[0], int(pw.argmin(axis=0)[0])                   # :2166-2170
                thr = thresholds[c][s] if c < len(thresholds) else []
                d = 1 if (thr and dist < thr) else 0                                           # :2173-2180 (Q4)
                dec.append(d)
                det.append((float(dist), arg, c, s, box))
        if not compat_q1:
            order = np.argsort([t[4] for t in det], kind="stable")
            dec = [dec[i] for i in order]
            det = [det[i] for i in order]
        decisions.append(dec)
        details.append(det)
    return (decisions, details) if return_details else decisions


def distance_indness(images, compat_q2=True, **kw):
    """:1498-1620.  With the reference's defaults every per-stride method returns -1 (Q2)."""
    if not compat_q2:
        raise NotImplementedError("intended formula (:1599-1604) is implemented in the product, not the oracle")
    dec = distance_decisions(images, compat_q1=True, **kw)
    return [[-1 for _ in d] for d in dec]


def logit_decisions(images, method, thresholds, temper=1.0):
    """:1195-1208, box order."""
    out = []
    for im in images:
        if len(im["cls"]) == 0:
            out.append([])
            continue
        sc = L.scores(im["logits"], im["cls"], method, temper)
        out.append([int(v) for v in L.decide(sc, im["cls"], thresholds)])
    return out


def logit_indness(images, method, thresholds, min_score, max_score, temper=1.0, clip=True):
    """:1210-1257, box order."""
    out = []
    for im in images:
        row = []
        if len(im["cls"]):
            sc = L.scores(im["logits"], im["cls"], method, temper)
            for s, c in zip(sc, np.asarray(im["cls"]).astype(int)):
                row.append(L.indness(s, int(c), thresholds, min_score, max_score, clip))
        out.append(row)
    return out


def fuse(d1, d2, strategy):
    """:2906-2940 -- position-wise, whatever order each list is in."""
    out = []
    for a, b in zip(d1, d2):
        assert len(a) == len(b)
        if strategy == "and":
            out.append([max(x, y) for x, y in zip(a, b)])
        elif strategy == "or":
            out.append([min(x, y) for x, y in zip(a, b)])
        elif strategy == "score":
            out.append([1 if (x + y) > 0 else 0 for x, y in zip(a, b)])
        else:
            raise NotImplementedError(strategy)
    return out


def fuse3(d1, d2, d3):
    """:3282-3301 majority vote."""
    return [[1 if (x + y + z) >= 2 else 0 for x, y, z in zip(a, b, c)] for a, b, c in zip(d1, d2, d3)]
