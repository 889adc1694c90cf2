"""GPU parity tests of the fit stage: K2 on vectors, K5 exact percentile, K4 k-means (labels bit-exact vs sklearn
on clustered data, and vs the golden labels produced by the reference's own call site)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from ood_in_object_detection_b200 import ops
    return ops.default_device()


def test_vec_score_matches_oracle(dev):
    from oracle import distance as D
    from ood_in_object_detection_b200 import ops
    rng = np.random.default_rng(0)
    dim, sizes, ks = 72, [50, 0, 300, 7], [3, 2, 10, 1]
    x = np.abs(rng.standard_normal((sum(sizes), dim))).astype(np.float32)
    cents = [(np.abs(rng.standard_normal((k, dim))) / np.sqrt(dim)).astype(np.float32) for k in ks]
    seg_off = np.concatenate([[0], np.cumsum(sizes)])
    cent = np.concatenate(cents)
    unit = np.concatenate([ops._unit_rows(c) for c in cents])
    crow = np.concatenate([[0], np.cumsum(ks)])[:-1]
    d, a = ops.vec_score(torch.from_numpy(x).to(dev), seg_off.tolist(), torch.from_numpy(cent).to(dev),
                         torch.from_numpy(unit).to(dev), crow.tolist(), ks, 0b111, True)
    d, a = d.cpu().numpy(), a.cpu().numpy()
    for m, metric in enumerate(("l1", "l2", "cosine")):
        for g in range(len(sizes)):
            if sizes[g] == 0:
                continue
            xs = D.normalize_rows(x[seg_off[g]:seg_off[g + 1]])
            pw = D.pairwise(cents[g], xs, metric)
            np.testing.assert_allclose(d[m, seg_off[g]:seg_off[g + 1]], pw.min(0), rtol=1e-5, atol=5e-7 if m == 2 else 0)
            assert (a[m, seg_off[g]:seg_off[g + 1]] != pw.argmin(0)).sum() <= 1


def test_exact_percentile_lower(dev, golden):
    from ood_in_object_detection_b200 import select
    g = golden("golden_thresholds.npz")
    rng = np.random.default_rng(1)
    segs = [g[f"v_{n}"] for n in (6, 11, 21, 101, 1001, 4097)]
    segs += [rng.standard_normal(5000).astype(np.float32) * 3, np.zeros(0, np.float32), np.full(17, -2.5, np.float32),
             rng.uniform(0, 1e-30, 300).astype(np.float32)]
    off = np.concatenate([[0], np.cumsum([len(s) for s in segs])])
    scores = torch.from_numpy(np.concatenate(segs)).to(dev)
    for q in (95.0, (1 - 0.95) * 100, 100 * 0.9, (1 - 0.9) * 100, 0.0, 100.0):
        ranks = [select.lower_index(len(s), q) if len(s) else None for s in segs]
        vals, mn, mx = select.segment_select(scores, off.tolist(), ranks)
        for i, s in enumerate(segs):
            if len(s) == 0:
                assert vals[i] is None and mn[i] is None
                continue
            assert vals[i] == float(np.percentile(s, q, method="lower")), (i, q)
            assert mn[i] == float(s.min()) and mx[i] == float(s.max())
    # golden: the reference's generate_thresholds on float32 scores
    for n in (6, 11, 21, 101, 1001, 4097):
        for tpr in (0.9, 0.95, 0.99, 0.8):
            v = g[f"v_{n}"]
            t = torch.from_numpy(v).to(dev)
            r, _, _ = select.segment_select(t, [0, n], [select.lower_index(n, 100 * tpr)])
            assert r[0] == g[f"dist_{n}_{tpr}"][0]
            r, _, _ = select.segment_select(t, [0, n], [select.lower_index(n, (1 - tpr) * 100)])
            assert r[0] == g[f"logit_{n}_{tpr}"][0]


def test_kmeans_labels_match_sklearn_and_reference(dev, golden):
    """Bit-exact labels on clustered data (strict convergence), several ragged segments in one launch."""
    from sklearn.cluster import KMeans
    from threadpoolctl import threadpool_limits
    from ood_in_object_detection_b200 import kmeans, synth
    g = golden("golden_kmeans.npz")
    # (1) the reference's own call site, KMeans_10 / KMeans_5 / n < k
    for tag in "abc":
        x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
        res = kmeans.kmeans_fit_predict_single(torch.from_numpy(x).to(dev), [len(x)], min(k, len(x)))
        assert np.array_equal(res.labels.cpu().numpy(), g[f"{tag}_labels"]), tag
    # (2) many segments at once, different sizes, vs sklearn per segment
    segs = [synth.blob_vectors(s, n, 48, kk, 7.0)[0] for s, n, kk in ((1, 3000, 8), (2, 700, 8), (3, 5, 3), (4, 20000, 8))]
    sizes = [len(s) for s in segs]
    res = kmeans.kmeans_fit_predict_single(torch.from_numpy(np.concatenate(segs)).to(dev), sizes, 8)
    lab = res.labels.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(sizes)])
    for i, s in enumerate(segs):
        with threadpool_limits(1):
            km = KMeans(n_clusters=min(8, len(s)), random_state=10).fit(s)
        assert np.array_equal(lab[off[i]:off[i + 1]], km.labels_), i
        assert res.n_iter[i] == km.n_iter_
        np.testing.assert_allclose(res.centers[i, :min(8, len(s))].cpu().numpy(), km.cluster_centers_, rtol=1e-4, atol=1e-5)


def test_kmeans_c3_shape_fast_path(dev):
    """D = 576, K = 16 (the TMA-fed register-tiled step kernel, 5 x 128 columns, ragged last chunk) vs sklearn."""
    from sklearn.cluster import KMeans
    from threadpoolctl import threadpool_limits
    from ood_in_object_detection_b200 import kmeans, synth
    segs = [synth.blob_vectors(40 + i, n, 576, 16, 8.0)[0] for i, n in enumerate((2500, 1037))]
    sizes = [len(s) for s in segs]
    x = torch.from_numpy(np.concatenate(segs)).to(dev)
    res = kmeans.kmeans_fit_predict_single(x, sizes, 16)
    lab = res.labels.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(sizes)])
    for i, s in enumerate(segs):
        with threadpool_limits(1):
            km = KMeans(n_clusters=16, random_state=10).fit(s)
        assert np.array_equal(lab[off[i]:off[i + 1]], km.labels_), i
        np.testing.assert_allclose(res.centers[i].cpu().numpy(), km.cluster_centers_, rtol=1e-4, atol=1e-5)
    means, counts = kmeans.member_means(x, sizes, res.labels, 16)         # update == 2 mode of the same kernel
    for i, s in enumerate(segs):
        for j in range(16):
            m = lab[off[i]:off[i + 1]] == j
            assert counts[i, j].item() == m.sum()
            np.testing.assert_allclose(means[i, j].cpu().numpy(), s[m].mean(0), rtol=1e-5, atol=1e-6)


def test_kmeans_realistic_overlap_agreement(dev):
    """Unstructured data: labels are compared through agreement / inertia (sklearn itself is not reproducible across
    thread counts there), not bit for bit."""
    from sklearn.cluster import KMeans
    from threadpoolctl import threadpool_limits
    from ood_in_object_detection_b200 import kmeans
    rng = np.random.default_rng(3)
    x = np.abs(rng.standard_normal((6000, 64))).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    res = kmeans.kmeans_fit_predict_single(torch.from_numpy(x).to(dev), [len(x)], 8)
    with threadpool_limits(1):
        km = KMeans(n_clusters=8, random_state=10).fit(x)
    lab = res.labels.cpu().numpy()
    agree = (lab == km.labels_).mean()
    c = res.centers[0].cpu().numpy()
    inertia = ((x - c[lab]) ** 2).sum()
    assert agree > 0.9 and inertia <= km.inertia_ * 1.01
