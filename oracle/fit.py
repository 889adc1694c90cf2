"""Fit-from-activations oracle: clusters, InD scores, thresholds.  TEST INFRASTRUCTURE.

Restates /root/reference/ood_utils.py
  * `DistanceMethod.generate_clusters` :2263-2295
      'one'      -> `generate_one_cluster_per_class_and_stride` :2297-2314 (mean of normalised vectors)
      otherwise  -> `generate_multiple_cluster_per_class_per_stride` :2316-2371: labels from
                    /root/reference/cluster_utils.py:18-186 ('all' :30-33, 'KMeans_<k>' :62-73),
                    centroid = np.mean of the *normalised* members per sorted label (:2359-2366)
  * `compute_scores_from_activations` :1877-1899 -> `compute_scores_clusters_per_class_and_stride` :2000-2024
  * `obtain_min_max_distances` :1901-1915
  * `OODMethod.generate_thresholds` :583-637 (np.percentile(..., method='lower'))
  * `LogitsMethod.compute_scores_from_activations` :1311-1332, `obtain_min_max_distances` :1334-1347
Constants from /root/reference/custom_hyperparams.py: MIN_SAMPLES=3 (:52),
MIN_NUMBER_OF_SAMPLES_FOR_THR=5 (:123).
"""
from __future__ import annotations

import numpy as np

from . import distance as D
from . import logits as L
from .kmeans import kmeans_fit_predict

F32 = np.float32
MIN_SAMPLES = 3
MIN_NUMBER_OF_SAMPLES_FOR_THR = 5


def cluster_labels(x, cluster_method, use_sklearn=False):
    """cluster_utils.py:18-186 for the in-scope methods.  x is the normalised [n,D] matrix."""
    if cluster_method == "one":
        raise ValueError("The 'one' method is not allowed for this function")
    if cluster_method == "all":
        return np.arange(len(x))
    if cluster_method.startswith("KMeans") and cluster_method.split("_")[-1].isdigit():
        k = int(cluster_method.split("_")[-1])
        if k < 2:
            raise ValueError("The number of clusters must be greater than 1")
        k = min(k, len(x))
        if use_sklearn:
            from sklearn.cluster import KMeans
            return KMeans(n_clusters=k, random_state=10).fit_predict(x)
        return kmeans_fit_predict(x, k, random_state=10)[0]
    raise NotImplementedError(f"{cluster_method} is out of scope (SURVEY.md §2 row 4)")


def generate_clusters(ind_tensors, cluster_method, normalize=True, use_sklearn=False):
    """ind_tensors[cls][stride] = ndarray [N,C,1,1] (or empty) -> clusters[cls][stride] = [K,C] f32 or empty."""
    out = [[[] for _ in range(3)] for _ in range(len(ind_tensors))]
    for c, per_stride in enumerate(ind_tensors):
        for s, a in enumerate(per_stride):
            if len(a) > MIN_SAMPLES:
                x = D.normalize_rows(a) if normalize else np.asarray(a, F32).reshape(len(a), -1)
                if cluster_method == "one":
                    out[c][s] = np.mean(x, axis=0)[None, :]
                else:
                    lab = cluster_labels(x, cluster_method, use_sklearn)
                    out[c][s] = np.array([np.mean(x[lab == j], axis=0) for j in sorted(set(lab.tolist()))])
            else:
                out[c][s] = np.empty(0)
    return out


def compute_scores_from_activations(activations, clusters, metric, normalize=True):
    """-> scores[cls][stride] (ndarray, or [] when the class/stride has samples but no cluster,
    or np.empty(0) when it has no samples), min_dist, max_dist."""
    scores = [[[] for _ in range(3)] for _ in range(len(activations))]
    for c, per_stride in enumerate(activations):
        for s, a in enumerate(per_stride):
            if len(a) > 0:
                if len(clusters[c][s]) > 0:
                    x = D.normalize_rows(a) if normalize else np.asarray(a, F32).reshape(len(a), -1)
                    scores[c][s] = D.compute_distance(clusters[c][s], x, metric)
            else:
                scores[c][s] = np.empty(0)
    mn = [[(np.min(v) if len(v) > 0 else 0) for v in row] for row in scores]
    mx = [[(np.max(v) if len(v) > 0 else 0) for v in row] for row in scores]
    return scores, mn, mx


def generate_thresholds(ind_scores, tpr, is_distance_method, per_stride):
    used = 100 * tpr if is_distance_method else (1 - tpr) * 100
    if per_stride:
        thr = [[[] for _ in range(3)] for _ in range(len(ind_scores))]
        for c, row in enumerate(ind_scores):
            for s, v in enumerate(row):
                if len(v) > MIN_NUMBER_OF_SAMPLES_FOR_THR:
                    thr[c][s] = float(np.percentile(v, used, method="lower"))
        return thr
    thr = [0 for _ in range(len(ind_scores))]
    for c, v in enumerate(ind_scores):
        if len(v) > MIN_NUMBER_OF_SAMPLES_FOR_THR:
            thr[c] = float(np.percentile(v, used, method="lower"))
    return thr


def percentile_lower_index(n, q, dtype=np.float32):
    """Index of the order statistic numpy 2.3 returns for `np.percentile(a, q, method='lower')`:
    numpy/lib/_function_base_impl.py:4277 divides q by `a.dtype.type(100)` (so the quantile is
    rounded to the DATA dtype: float32 scores -> float32 quantile, float64 L1 scores -> float64), and
    :141-144 takes floor((n - 1) * quantile) in that same dtype.  q is a python float, as in
    /root/reference/ood_utils.py:593-597 (`100*tpr` or `(1 - tpr)*100`)."""
    dt = np.dtype(dtype).type
    quant = np.true_divide(float(q), dt(100))
    return int(np.floor((int(n) - 1) * quant).astype(np.intp))


def logits_scores_from_activations(activations, method, temper=1.0):
    """LogitsMethod.compute_scores_from_activations :1311-1347 -> scores[cls], min_score, max_score."""
    scores = []
    for c, a in enumerate(activations):
        if len(a) > 0:
            scores.append(L.scores(np.asarray(a, F32), np.full(len(a), c), method, temper))
        else:
            scores.append(np.array([], dtype=F32))
    mn = [(np.min(v) if len(v) > 0 else 0.0) for v in scores]
    mx = [(np.max(v) if len(v) > 0 else 0.0) for v in scores]
    return scores, mn, mx
