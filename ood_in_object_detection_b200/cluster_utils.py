"""Cluster-label selection for one (class, stride) -- the in-scope part of /root/reference/cluster_utils.py:18-186.

`'all'` (:30-33), `KMeans_<k>` (:62-73) and the silhouette / Calinski-Harabasz searched `KMeans` (:75-80, :160-186,
:203-356) are served; k-means (kmeans.py) and the O(n^2 D) part of the silhouette score (ops.silhouette_score, K7) run on
the GPU.  The other clusterers (DBSCAN, HDBSCAN, Agglomerative, Birch, MeanShift, GMM, BGMM) are CPU-library algorithms
outside the hot path and raise NotImplementedError.
"""
from __future__ import annotations

import os
from logging import Logger
from typing import List, Optional, Tuple

import numpy as np

from .constants import is_valid_cluster_method, kmeans_k
from .custom_hyperparams import CUSTOM_HYP


def search_number_of_clusters(x, metric: str, perf_score_metric: str, logger: Logger, random_state: int = 10):
    """`KMeans` of cluster_utils.py:75-80: for every k in CUSTOM_HYP.clusters.RANGE_OF_CLUSTERS fit k-means (random_state
    10) and score the labels (:203-302), take the FIRST best k (:350 `np.argmax`) and its labels (:177 refits the same
    deterministic estimator: the labels of the search are reused).  A k that fails the reference's checks -- more
    clusters than samples (sklearn raises), a single label, n - 1 labels or more, a cluster with fewer than MIN_SAMPLES
    members -- scores the default (-1 silhouette / 0 Calinski-Harabasz); if every silhouette is -1 all samples go to
    cluster 0 (:160-166).  x: float32 [n, D] device tensor.  -> (labels int32 device tensor, scores, ks)."""
    import torch
    from . import kmeans, ops
    if perf_score_metric not in ("silhouette", "calinski_harabasz"):
        raise ValueError(f"Invalid performance score metric: {perf_score_metric}")
    n = int(x.shape[0])
    default = -1 if perf_score_metric == "silhouette" else 0
    ks = list(CUSTOM_HYP.clusters.RANGE_OF_CLUSTERS)
    assert len(ks) > 1, "Parameter n_clusters must have more than one value to evaluate"
    scores: List[float] = []
    found = {}
    pairs = None                                              # the pair distances of x, shared by every candidate k
    # Every candidate k is an independent KMeans(n_clusters=k, random_state) on the same rows: they are fitted together as
    # the segments of ONE segmented fit over copies of x (the small fits are launch-bound one by one), when the copies fit
    # the budget (OODB200_KSEARCH_BATCH_GB, default 4 GB).
    fitted = {}
    valid = [k for k in ks if k <= n]
    batch_bytes = 4 * n * int(x.shape[1]) * len(valid)
    if len(valid) > 1 and batch_bytes <= float(os.environ.get("OODB200_KSEARCH_BATCH_GB", "4")) * (1 << 30):
        seeding = "device" if n >= kmeans.DEVICE_SEEDING_MIN_ROWS else "host"      # what a fit of x alone would use
        res = kmeans.kmeans_fit_predict_single(x.repeat(len(valid), 1), [n] * len(valid), max(valid), random_state=random_state,
                                               seg_k=valid, seeding=seeding)
        fitted = {k: res.labels[i * n:(i + 1) * n] for i, k in enumerate(valid)}
    for k in ks:
        score = default
        if k > n:
            logger.error(f"Error with parameters {{'n_clusters': {k}, 'random_state': {random_state}}}: "
                         f"n_samples={n} should be >= n_clusters={k}.")
        else:
            labels = fitted[k] if k in fitted else kmeans.kmeans_fit_predict_single(x, [n], k, random_state=random_state).labels
            counts = torch.bincount(labels.long(), minlength=k).cpu().numpy()
            present = counts[counts > 0]
            if n - 1 > len(present) > 1:
                if present.min() < CUSTOM_HYP.clusters.MIN_SAMPLES:
                    bad = int(np.flatnonzero((counts > 0) & (counts < CUSTOM_HYP.clusters.MIN_SAMPLES))[0])
                    logger.error(f"Error with parameters {{'n_clusters': {k}, 'random_state': {random_state}}}: "
                                 f"Cluster {bad} has less than {CUSTOM_HYP.clusters.MIN_SAMPLES} samples.")
                elif perf_score_metric == "silhouette":
                    if pairs is None:
                        pairs = ops.PairDistances(x, metric)
                    score = ops.silhouette_score(x, labels, metric, pairs=pairs)
                    logger.debug(f"Silhouette score: {score}")
                else:
                    score = ops.calinski_harabasz_score(x, labels)
                    logger.debug(f"Calinski-Harabasz score: {score}")
            else:
                logger.debug("Clustering resulted in a single cluster, skipping.")
            found[k] = labels
        scores.append(score)
    if (np.array(scores) == -1).all():
        logger.warning("All configurations resulted in a single cluster. Assigning all samples to the same cluster.")
        return torch.zeros(n, dtype=torch.int32, device=x.device), scores, ks
    best = ks[int(np.argmax(scores))]
    logger.info(f"Best parameters: {{'n_clusters': {best}, 'random_state': {random_state}}}")
    if best not in found:
        raise ValueError(f"n_samples={n} should be >= n_clusters={best}.")
    return found[best], scores, ks


def find_optimal_number_of_clusters_one_class_one_stride_and_return_labels(
        feature_maps: np.ndarray, cluster_method: str, metric: str, perf_score_metric: str,
        string_for_visualization: str, logger: Logger, visualize: Optional[bool] = False) -> np.ndarray:
    assert is_valid_cluster_method(cluster_method), f"Invalid clustering method: {cluster_method}"
    if cluster_method == 'one':
        raise ValueError("The 'one' method is not allowed for this function")
    if cluster_method == 'all':
        return np.arange(len(feature_maps))
    k = kmeans_k(cluster_method)
    if k is not None or cluster_method == 'KMeans':
        import torch
        from . import kmeans, ops
        x = ops.h2d(np.ascontiguousarray(feature_maps, dtype=np.float32), ops.default_device())
        if k is None:
            return search_number_of_clusters(x, metric, perf_score_metric, logger)[0].cpu().numpy()
        if k < 2:
            raise ValueError("The number of clusters must be greater than 1")
        res = kmeans.kmeans_fit_predict_single(x, [len(feature_maps)], min(k, len(feature_maps)), random_state=10)
        return res.labels.cpu().numpy()
    raise NotImplementedError(f"cluster_method '{cluster_method}' is a CPU-library clusterer outside the GPU hot path")
