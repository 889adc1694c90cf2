"""CPU restatement of the silhouette-searched `KMeans` cluster method.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/cluster_utils.py:
  :75-80    `KMeans`: n_clusters over CUSTOM_HYP.clusters.RANGE_OF_CLUSTERS (2..14), random_state 10
  :203-302  `compute_score_for_all_possible_configurations`: per k fit_predict, the "more than one label, fewer than
            n-1" and ">= MIN_SAMPLES per cluster" checks, silhouette / Calinski-Harabasz score, default score (-1 / 0) on
            any failure
  :316-356  `search_for_best_param`: first maximum of the scores
  :160-177  all scores == -1 -> every sample in cluster 0, else refit with the best k
and the third-party arithmetic it calls (scikit-learn 1.9 `silhouette_samples` = `pairwise_distances_chunked` +
`_silhouette_reduce`, `calinski_harabasz_score`).  Pinned by tests/golden/golden_ksearch.npz (frozen from the reference)
and against sklearn itself in tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import numpy as np

from . import distance, kmeans

RANGE_OF_CLUSTERS = list(range(2, 15))       # custom_hyperparams.py:53
MIN_SAMPLES = 3                              # custom_hyperparams.py:52


def pairwise_full(x, metric):
    """pairwise_distances(x, metric=metric) for float32 rows: [n, n] with an exactly zero diagonal."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if metric == "l1":
        d = np.abs(x[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)).sum(-1)
    elif metric == "l2":
        xx = np.einsum("ij,ij->i", x.astype(np.float64), x.astype(np.float64))
        d2 = xx[:, None] - 2.0 * (x.astype(np.float64) @ x.astype(np.float64).T) + xx[None, :]
        d = np.sqrt(np.maximum(d2, 0)).astype(np.float32)
    elif metric == "cosine":
        u = distance.normalize_rows(x)
        d = np.clip(1.0 - u @ u.T, 0, 2)
    else:
        raise ValueError(metric)
    np.fill_diagonal(d, 0)
    return d


def silhouette_samples(x, labels, metric):
    """sklearn silhouette_samples: a = mean distance to the other members of the own cluster, b = smallest mean distance
    to another cluster, (b - a) / max(a, b); members of one-sample clusters get 0."""
    _, enc = np.unique(labels, return_inverse=True)
    freq = np.bincount(enc)
    d = pairwise_full(x, metric)
    n = len(enc)
    sums = np.zeros((n, len(freq)), dtype=d.dtype)
    for i in range(n):
        sums[i] += np.bincount(enc, weights=d[i], minlength=len(freq))
    idx = (np.arange(n), enc)
    intra = sums[idx].copy()
    sums[idx] = np.inf
    inter = (sums / freq).min(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        intra = intra / (freq - 1).take(enc)
        sil = (inter - intra) / np.maximum(intra, inter)
    return np.nan_to_num(sil)


def silhouette_score(x, labels, metric):
    return float(np.mean(silhouette_samples(x, labels, metric)))


def calinski_harabasz_score(x, labels):
    """sklearn calinski_harabasz_score: between / within dispersion ratio."""
    x = np.asarray(x)
    uniq, enc = np.unique(labels, return_inverse=True)
    n, k = len(x), len(uniq)
    mean = np.mean(x, axis=0)
    extra = intra = 0.0
    for c in range(k):
        xc = x[enc == c]
        mc = np.mean(xc, axis=0)
        extra += len(xc) * np.sum((mc - mean) ** 2)
        intra += np.sum((xc - mc) ** 2)
    return float(1.0 if intra == 0.0 else extra * (n - k) / (intra * (k - 1.0)))


def k_search(x, metric, perf_score_metric="silhouette", labels_for_k=None):
    """-> (scores per k in RANGE_OF_CLUSTERS, final labels).  `labels_for_k(x, k)` overrides the k-means used."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = len(x)
    fit = labels_for_k or (lambda xx, k: kmeans.kmeans_fit_predict(xx, k, random_state=10)[0])
    default = -1 if perf_score_metric == "silhouette" else 0
    scores, cache = [], {}
    for k in RANGE_OF_CLUSTERS:
        score = default
        if k <= n:                                       # KMeans raises for n_samples < n_clusters -> default score
            lab = cache[k] = np.asarray(fit(x, k))
            _, counts = np.unique(lab, return_counts=True)
            if n - 1 > len(counts) > 1 and counts.min() >= MIN_SAMPLES:
                score = silhouette_score(x, lab, metric) if perf_score_metric == "silhouette" else calinski_harabasz_score(x, lab)
        scores.append(score)
    if (np.array(scores) == -1).all():
        return scores, np.zeros(n, dtype=np.int32)
    best = RANGE_OF_CLUSTERS[int(np.argmax(scores))]
    if best > n:
        raise ValueError(f"n_samples={n} should be >= n_clusters={best}.")
    return scores, cache[best]
