"""String enums the OoD method classes validate against (contract: /root/reference/constants.py:12-47).
`KMeans_<k>` is accepted for any k >= 2 (the reference whitelists 3/5/10 but its parser takes any integer,
cluster_utils.py:63-73; BASELINE configs need KMeans_16 / KMeans_64)."""
import re

STRIDES_RATIO = [8, 16, 32]
UNKNOWN_CLASS_INDEX = 80

LOGITS_METHODS = ['NoMethod', 'MSP', 'Energy', 'ODIN', 'Sigmoid', 'MaxLogit']
DISTANCE_METHODS = ['L1_cl_stride', 'L2_cl_stride', 'Cosine_cl_stride', 'Umap', 'CosineIvis', 'L1Ivis', 'L2Ivis']
OOD_METHOD_CHOICES = LOGITS_METHODS + DISTANCE_METHODS

FTMAPS_RELATED_OPTIONS = ['roi_aligned_ftmaps', 'all_ftmaps', 'ftmaps_and_strides', 'ftmaps_and_strides_exact_pos']
LOGITS_RELATED_OPTIONS = ['logits']
INTERNAL_ACTIVATIONS_EXTRACTION_OPTIONS = FTMAPS_RELATED_OPTIONS + LOGITS_RELATED_OPTIONS + ['none']

AVAILABLE_CLUSTERING_METHODS = ['one', 'all', 'DBSCAN', 'KMeans', 'KMeans_3', 'KMeans_5', 'KMeans_10', 'HDBSCAN',
                                'AgglomerativeClustering', 'OPTICS', 'Birch', 'MeanShift', 'SpectralClustering', 'GMM', 'BGMM']
GPU_CLUSTERING_METHODS = ['one', 'all']          # + KMeans_<k>; the other clusterers are CPU libraries (out of scope)
AVAILABLE_CLUSTER_OPTIMIZATION_METRICS = ['silhouette', 'calinski_harabasz']

TARGETS_RELATED_OPTIONS = ['all_targets_one_stride', 'all_targets_all_strides']
PREDICTIONS_RELATED_OPTIONS = ['valid_preds_one_stride', 'valid_preds_all_strides', 'all_preds_all_strides']
IND_INFO_CREATION_OPTIONS = TARGETS_RELATED_OPTIONS + PREDICTIONS_RELATED_OPTIONS

_KMEANS_K = re.compile(r"^KMeans_(\d+)$")


def kmeans_k(cluster_method: str):
    """`KMeans_<k>` -> k, else None."""
    m = _KMEANS_K.match(cluster_method)
    return int(m.group(1)) if m else None


def is_valid_cluster_method(cluster_method: str) -> bool:
    return cluster_method in AVAILABLE_CLUSTERING_METHODS or kmeans_k(cluster_method) is not None
