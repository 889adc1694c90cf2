"""Normalise + pairwise-distance oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates in numpy what the reference calls from scikit-learn 1.9 / scipy 1.18:
  * `DistanceMethod.activations_transformation`  /root/reference/ood_utils.py:2404-2409
      -> sklearn.preprocessing.normalize (site-packages/sklearn/preprocessing/_data.py:2073-2081,
         `row_norms` sklearn/utils/extmath.py:86-95, `_handle_zeros_in_scale` _data.py:101-133)
  * `_PairwiseDistanceClustersPerClassPerStride.compute_distance`  ood_utils.py:2422-2430
      -> sklearn.metrics.pairwise_distances(cluster, activations, metric).min(axis=0)
         'l2'/'euclidean': sklearn/metrics/pairwise.py:376-427,567-638 (float64 upcast, cast to f32, sqrt)
         'l1'/'manhattan': scipy cdist 'cityblock' in float64 (pairwise.py:1108-1109) -> returns float64
         'cosine': 1 - normalize(X) @ normalize(Y).T in float32, clip to [0,2] (pairwise.py:1171-1182)
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
METRIC_ALIASES = {"l1": "l1", "manhattan": "l1", "cityblock": "l1", "l2": "l2", "euclidean": "l2", "cosine": "cosine"}


def normalize_rows(x):
    """sklearn.preprocessing.normalize(x.reshape(n,-1), axis=1) for float32 input."""
    x = np.array(x, dtype=F32, copy=True).reshape(len(x), -1)
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms[norms < 10 * np.finfo(F32).eps] = 1.0
    x /= norms[:, None]
    return x


def pairwise(cluster, acts, metric):
    """pairwise_distances(cluster[K,D], acts[n,D], metric) -> [K,n] (f32, or f64 for l1)."""
    m = METRIC_ALIASES[metric]
    c = np.asarray(cluster, dtype=F32)
    a = np.asarray(acts, dtype=F32)
    if m == "l2":
        c64, a64 = c.astype(np.float64), a.astype(np.float64)
        d = -2.0 * (c64 @ a64.T)
        d += np.einsum("ij,ij->i", c64, c64)[:, None]
        d += np.einsum("ij,ij->i", a64, a64)[None, :]
        d = d.astype(F32)
        np.maximum(d, 0, out=d)
        return np.sqrt(d)
    if m == "l1":
        return np.abs(c.astype(np.float64)[:, None, :] - a.astype(np.float64)[None, :, :]).sum(-1)
    cn, an = normalize_rows(c), normalize_rows(a)
    s = cn @ an.T
    s *= -1
    s += 1
    np.clip(s, 0, 2, out=s)
    return s


def compute_distance(cluster, acts, metric):
    """ood_utils.py:2422-2430: distance of every vector to its nearest centroid."""
    return pairwise(cluster, acts, metric).min(axis=0)


def compute_assignment(cluster, acts, metric):
    """Box-to-cluster assignment = argmin over the same matrix (first minimum)."""
    return pairwise(cluster, acts, metric).argmin(axis=0)
