"""Cluster-label selection for one (class, stride) -- the in-scope part of /root/reference/cluster_utils.py:18-186.

`'all'` (:30-33) and `KMeans_<k>` (:62-73) are served; k-means runs on the GPU (kmeans.py).  The other clusterers
(DBSCAN, HDBSCAN, Agglomerative, Birch, MeanShift, GMM, silhouette-searched `KMeans`) are CPU-library algorithms outside
the hot path and raise NotImplementedError.
"""
from __future__ import annotations

from logging import Logger
from typing import Optional

import numpy as np

from .constants import is_valid_cluster_method, kmeans_k


def find_optimal_number_of_clusters_one_class_one_stride_and_return_labels(
        feature_maps: np.ndarray, cluster_method: str, metric: str, perf_score_metric: str,
        string_for_visualization: str, logger: Logger, visualize: Optional[bool] = False) -> np.ndarray:
    assert is_valid_cluster_method(cluster_method), f"Invalid clustering method: {cluster_method}"
    if cluster_method == 'one':
        raise ValueError("The 'one' method is not allowed for this function")
    if cluster_method == 'all':
        return np.arange(len(feature_maps))
    k = kmeans_k(cluster_method)
    if k is not None:
        if k < 2:
            raise ValueError("The number of clusters must be greater than 1")
        import torch
        from . import kmeans, ops
        x = torch.as_tensor(np.ascontiguousarray(feature_maps, dtype=np.float32)).to(ops.default_device())
        res = kmeans.kmeans_fit_predict_single(x, [len(feature_maps)], min(k, len(feature_maps)), random_state=10)
        return res.labels.cpu().numpy()
    raise NotImplementedError(f"cluster_method '{cluster_method}' is a CPU-library clusterer outside the GPU hot path")
