"""sklearn.cluster.KMeans restated in numpy.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference's k-means fit is one library call,
`KMeans(n_clusters=k, random_state=10).fit_predict(X)`  (/root/reference/cluster_utils.py:62-73),
made per (class, stride) from /root/reference/ood_utils.py:2345.  scikit-learn is a third-party
dependency that is not vendored under /root/reference (requirements.txt:3 `>=1.3.0`; this image
has 1.9.0), so its published algorithm is restated here, single-threaded (n_threads=1: chunks of
256 rows processed in order, which is a defined summation order -- SURVEY.md §7):

  KMeans.fit                 sklearn/cluster/_kmeans.py:1436-1563  (mean-centre, 1 init, Lloyd)
  _tolerance                 _kmeans.py:285-293
  _kmeans_plusplus           _kmeans.py:180-278   (numpy RandomState stream: choice, then uniform(size=2+int(ln k)))
  _kmeans_single_lloyd       _kmeans.py:630-758
  lloyd_iter_chunked_dense   sklearn/cluster/_k_means_lloyd.pyx:23-218 (argmin over ||c||^2 - 2 x.c, first minimum)
  _relocate_empty_clusters_dense / _average_centers / _center_shift   _k_means_common.pyx:167-311

Pinned against the library itself in tests/test_oracle_vs_golden.py (identical labels on
clustered data) and against golden vectors produced by the reference's own call site.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
CHUNK = 256


def _sq_dists_upcast(Y, X):
    """_euclidean_distances(Y, X, squared=True) for float32 inputs: float64 expansion, cast to f32, clamp."""
    Y64, X64 = Y.astype(np.float64), X.astype(np.float64)
    d = -2.0 * (Y64 @ X64.T)
    d += np.einsum("ij,ij->i", Y64, Y64)[:, None]
    d += np.einsum("ij,ij->i", X64, X64)[None, :]
    d = d.astype(F32)
    np.maximum(d, 0, out=d)
    return d


def kmeans_plusplus(X, k, random_state):
    """_kmeans_plusplus with unit sample weights.  X is already mean-centred float32."""
    n = X.shape[0]
    w = np.ones(n, dtype=X.dtype)
    centers = np.empty((k, X.shape[1]), dtype=X.dtype)
    n_local_trials = 2 + int(np.log(k))
    center_id = random_state.choice(n, p=w / w.sum())
    indices = np.full(k, -1, dtype=int)
    centers[0] = X[center_id]
    indices[0] = center_id
    closest = _sq_dists_upcast(centers[0, None], X)
    current_pot = closest @ w
    for c in range(1, k):
        rand_vals = random_state.uniform(size=n_local_trials) * current_pot
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)
        np.clip(cand, None, closest.size - 1, out=cand)
        d = _sq_dists_upcast(X[cand], X)
        np.minimum(closest, d, out=d)
        pots = d @ w.reshape(-1, 1)
        best = int(np.argmin(pots))
        current_pot = pots[best]
        closest = d[best]
        centers[c] = X[cand[best]]
        indices[c] = cand[best]
    return centers, indices


def _assign(X, centers):
    csn = np.einsum("ij,ij->i", centers, centers).astype(X.dtype)
    labels = np.empty(X.shape[0], dtype=np.int32)
    for s in range(0, X.shape[0], CHUNK):
        pd = csn[None, :] + F32(-2.0) * (X[s:s + CHUNK] @ centers.T)       # gemm(alpha=-2, beta=1)
        labels[s:s + CHUNK] = pd.argmin(axis=1)                           # first minimum
    return labels


def lloyd_iter(X, centers_old, update_centers=True):
    k, dim = centers_old.shape
    labels = _assign(X, centers_old)
    if not update_centers:
        return labels, None, None
    centers_new = np.zeros_like(centers_old)
    weight = np.zeros(k, dtype=X.dtype)
    for j in range(k):
        m = labels == j
        cnt = int(m.sum())
        if cnt:
            centers_new[j] = np.cumsum(X[m], axis=0, dtype=X.dtype)[-1]     # sequential f32 sum in row order
            weight[j] = cnt
    empty = np.where(weight == 0)[0]
    if empty.size:                                                         # _relocate_empty_clusters_dense
        dist = ((X - centers_old[labels]) ** 2).sum(axis=1)
        if np.max(dist) != 0:
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            for idx, new_id in enumerate(empty):
                fi = far[idx]
                old_id = labels[fi]
                centers_new[old_id] -= X[fi]
                centers_new[new_id] = X[fi]
                weight[new_id] = 1
                weight[old_id] -= 1
    amax = int(np.argmax(weight))
    for j in range(k):                                                     # _average_centers
        if weight[j] > 0:
            centers_new[j] *= X.dtype.type(1.0 / weight[j])
        else:
            centers_new[j] = centers_new[amax]
    shift = np.sqrt(((centers_new - centers_old) ** 2).sum(axis=1)).astype(X.dtype)   # _center_shift
    return labels, centers_new, shift


def kmeans_fit_predict(X, n_clusters, random_state=10, max_iter=300, tol=1e-4, init_centers=None):
    """-> labels[int32], centers (in original coordinates), n_iter, strict_convergence."""
    X = np.array(X, dtype=F32, order="C", copy=True)
    tol_abs = 0 if tol == 0 else np.mean(np.var(X, axis=0)) * tol
    mean = X.mean(axis=0)
    X -= mean
    if init_centers is None:
        rs = np.random.RandomState(random_state) if not isinstance(random_state, np.random.RandomState) else random_state
        centers, _ = kmeans_plusplus(X, n_clusters, rs)
    else:
        centers = np.array(init_centers, dtype=F32) - mean
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    n_iter = 0
    for i in range(max_iter):
        labels, centers_new, shift = lloyd_iter(X, centers)
        centers = centers_new
        n_iter = i + 1
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels, _, _ = lloyd_iter(X, centers, update_centers=False)
    return labels, centers + mean, n_iter, strict
