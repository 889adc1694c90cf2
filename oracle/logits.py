"""Logit-method scorers.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates /root/reference/ood_utils.py:1388-1443 (MSP / Energy / ODIN / Sigmoid `compute_scores`),
:1195-1208 (decision), :1224-1257 (`LogitsMethod.compute_indness`) in numpy float32.
The reference computes with torch CPU float32 (`softmax`, `logsumexp`, `sigmoid`); this
restatement uses the max-subtracted formulas torch documents, evaluated in float32.
`MaxLogit` has no counterpart in the reference -- parity unpinned for it;
defined here as `logits.max(axis=1)`.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
METHODS = ("MSP", "Energy", "ODIN", "Sigmoid", "MaxLogit")


def _softmax(z):
    z = z.astype(F32)
    m = z.max(axis=1, keepdims=True)
    e = np.exp((z - m).astype(F32)).astype(F32)
    return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)


def scores(logits, cls, method, temper=1.0):
    """logits [n,NC] f32 (raw, pre-sigmoid), cls [n] int -> [n] f32 score (higher = more InD)."""
    z = np.asarray(logits, dtype=F32).reshape(-1, np.shape(logits)[-1])
    cls = np.asarray(cls).astype(np.int64).reshape(-1)
    r = np.arange(len(z))
    if method == "MSP":                      # ood_utils.py:1394-1397
        return _softmax(z)[r, cls]
    if method == "ODIN":                     # ood_utils.py:1424-1427
        return _softmax((z / F32(temper)).astype(F32))[r, cls]
    if method == "Energy":                   # ood_utils.py:1409-1412
        zt = (z / F32(temper)).astype(F32)
        m = zt.max(axis=1)
        lse = (m + np.log(np.exp((zt - m[:, None]).astype(F32)).sum(axis=1, dtype=F32)).astype(F32)).astype(F32)
        return (F32(temper) * lse).astype(F32)
    if method == "Sigmoid":                  # ood_utils.py:1436-1443 (use_values_before_sigmoid=True)
        s = (F32(1) / (F32(1) + np.exp(-z).astype(F32))).astype(F32)
        return s[r, cls]
    if method == "MaxLogit":
        return z.max(axis=1)
    raise ValueError(method)


def decide(score, cls, thresholds):
    """ood_utils.py:1203-1206: 0 (OoD) if score < thr[cls] else 1 (InD)."""
    thr = np.asarray([float(t) for t in thresholds], dtype=np.float64)
    return np.where(np.asarray(score, dtype=np.float64) < thr[np.asarray(cls, dtype=np.int64)], 0, 1).astype(np.int64)


def indness(score, cls_idx, thr, min_score, max_score, clip=True):
    """ood_utils.py:1224-1257 for one box (python floats, like the reference)."""
    score = float(score)
    t = float(thr[cls_idx])
    if score > t:
        a = 1 / (float(max_score[cls_idx]) - t)
        b = -t / (float(max_score[cls_idx]) - t)
    elif score < t:
        a = -1 / (float(min_score[cls_idx]) - t)
        b = t / (float(min_score[cls_idx]) - t)
    else:
        a = 0
        b = 0
    v = a * score + b
    return max(-1, min(v, 1)) if clip else v
