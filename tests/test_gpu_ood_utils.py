"""The reference's class surface (ood_in_object_detection_b200.ood_utils) against golden vectors frozen from the
reference's own classes (tests/golden/make_golden.py).  Reads like the reference's usage in ood_evaluation.py:
construct -> extract activations -> generate_clusters -> compute_scores_from_activations -> generate_thresholds ->
compute_ood_decision_on_results."""
from __future__ import annotations

import logging

import numpy as np
import pytest
import torch

from tests.helpers import scoring_case, split, train_case, unpack_nested

pytestmark = pytest.mark.gpu
RTOL = 1e-5
LOG = logging.getLogger("test")
LOG.setLevel(logging.ERROR)

DIST_KW = dict(agg_method="mean", cluster_method="one", cluster_optimization_metric="silhouette",
               ind_info_creation_option="valid_preds_one_stride", which_internal_activations="ftmaps_and_strides",
               iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)
LOGIT_KW = dict(per_class=True, per_stride=False, iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15,
                min_conf_threshold_test=0.15, use_values_before_sigmoid=True)
COMMON = dict(iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)


@pytest.fixture(scope="module")
def ou():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from ood_in_object_detection_b200 import ood_utils
    return ood_utils


def _results(maps, boxes, cls, strides, img, device="cuda"):
    """Our Results stand-ins for a batch; maps = 3 arrays [B, C, H, W] (kept as per-image CHW device tensors)."""
    from ood_in_object_detection_b200.results import Results, batch_shape
    tm = [torch.from_numpy(np.ascontiguousarray(m)).to(device) for m in maps]
    out = []
    for i in range(len(boxes)):
        b6 = np.concatenate([boxes[i], np.full((len(boxes[i]), 1), 0.5, np.float32), cls[i][:, None]], 1).astype(np.float32)
        out.append(Results(orig_img=batch_shape(len(boxes), img, img), boxes=torch.from_numpy(b6).to(device),
                           extra_item=([t[i] for t in tm], torch.from_numpy(strides[i]).to(device))))
    return out


def _classes(ou):
    return (("l1", ou.L1DistanceOneClusterPerStride), ("l2", ou.L2DistanceOneClusterPerStride),
            ("cos", ou.CosineDistanceOneClusterPerStride))


@pytest.mark.parametrize("name", ["golden_c1_one.npz", "golden_small_kmeans5.npz"])
def test_distance_methods_fit_and_decide(ou, golden, name):
    g = golden(name)
    nc, img = int(g["nc"]), int(g["img"])
    cluster_method = str(g["cluster_method"])
    tmaps, tboxes, tcls, tstrides = train_case(g)
    train = _results(tmaps, tboxes, tcls, tstrides, img)
    for r, b in zip(train, tboxes):
        r.valid_preds = list(range(len(b)))            # the golden fit takes every prediction as valid
    images, maps = scoring_case(g)
    test = _results(maps, [im["boxes"] for im in images], [im["cls"] for im in images], [im["strides"] for im in images], img)
    n_exempt = 0
    for tag, cls in _classes(ou):
        m = cls(**dict(DIST_KW, cluster_method=cluster_method))
        acts = m._empty_activation_lists(nc)
        m.extract_internal_activations(train, acts, None)
        m.format_internal_activations(acts)
        # ---- fit: clusters, InD scores, thresholds
        clusters = m.generate_clusters(acts, LOG)
        ref_clusters = unpack_nested(g, f"{tag}_clusters", nc)
        for c in range(nc):
            for s in range(3):
                assert clusters[c][s].shape == ref_clusters[c][s].shape, (tag, c, s)
                if ref_clusters[c][s].size:
                    np.testing.assert_allclose(clusters[c][s], ref_clusters[c][s], rtol=RTOL, atol=1e-7)
        m.clusters = ref_clusters                       # pin the state so the next stages compare like for like
        scores = m.compute_scores_from_activations(acts, LOG)
        for c in range(nc):
            for s in range(3):
                ref = g[f"{tag}_fitscores_{c}_{s}"]
                assert len(scores[c][s]) == len(ref), (tag, c, s)
                if len(ref):
                    np.testing.assert_allclose(scores[c][s], ref, rtol=RTOL, atol=1e-6)   # a lone member is its own centroid: 0 vs 1 ulp * D
                    np.testing.assert_allclose(m.min_dist[c][s], float(g[f"{tag}_mindist_{c}_{s}"]), rtol=RTOL, atol=1e-6)
                    np.testing.assert_allclose(m.max_dist[c][s], float(g[f"{tag}_maxdist_{c}_{s}"]), rtol=RTOL, atol=1e-6)
        thr = m.generate_thresholds(scores, 0.95, LOG)
        ref_thr = unpack_nested(g, f"{tag}_thr", nc, as_threshold=True)
        for c in range(nc):
            for s in range(3):
                if ref_thr[c][s] == []:
                    assert thr[c][s] == [], (tag, c, s)
                else:
                    assert isinstance(thr[c][s], float)
                    np.testing.assert_allclose(thr[c][s], ref_thr[c][s], rtol=RTOL, atol=1e-6)
        # ---- decide, with the reference's fitted state
        m.thresholds = ref_thr
        dec = m.compute_ood_decision_on_results(test, LOG)
        assert [len(d) for d in dec] == [int(v) for v in g["n_boxes"]]
        flat = np.array([v for d in dec for v in d], np.int8)
        ref_d = g[f"{tag}_dist"]
        thr_flat = np.array([ref_thr[c][s] if ref_thr[c][s] != [] else np.nan
                             for c, s in zip(g[f"{tag}_cls_used"], g[f"{tag}_stride_of"])], np.float64)
        near = np.abs(ref_d - thr_flat) <= RTOL * np.abs(thr_flat)
        n_exempt += int(near.sum())
        assert np.array_equal(flat[~near], g[f"{tag}_decisions"][~near]), tag
        r = m.score_results(test)
        np.testing.assert_allclose(r["dist"], ref_d, rtol=RTOL, atol=5e-7 if tag == "cos" else 0)
        assert np.array_equal(r["cls_used"], g[f"{tag}_cls_used"]) and np.array_equal(r["stride"], g[f"{tag}_stride_of"])
        if tag == "cos":                                # Q2: the reference's INDness is -1 for every box
            ind = m.compute_INDness_scores_on_results(test, LOG)
            assert np.array_equal(np.array([v for d in ind for v in d], np.float64), g["cos_indness"])
            m.min_dist = unpack_nested(g, "cos_mindist", nc)
            m.max_dist = unpack_nested(g, "cos_maxdist", nc)
            m.reference_compat = False                  # intended semantics: in [-1, 1], sign agrees with the decision
            ind2 = np.array([v for d in m.compute_INDness_scores_on_results(test, LOG) for v in d])
            dec2 = np.array([v for d in m.compute_ood_decision_on_results(test, LOG) for v in d])
            assert ind2.min() >= -1 and ind2.max() <= 1
            assert np.all((ind2 > 0) <= (dec2 == 1)) and np.all(ind2[dec2 == 0] <= 0)
    assert n_exempt <= 2


def test_pre_pooled_and_box_order_modes(ou, golden):
    """'roi_aligned_ftmaps' (vectors pooled by the detector side) gives the decisions of 'ftmaps_and_strides';
    reference_compat=False returns them in box order with the box's own class."""
    g = golden("golden_small_kmeans5.npz")
    nc, img = int(g["nc"]), int(g["img"])
    images, maps = scoring_case(g)
    boxes, cls, strides = [im["boxes"] for im in images], [im["cls"] for im in images], [im["strides"] for im in images]
    test = _results(maps, boxes, cls, strides, img)
    m = ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, cluster_method="KMeans_5"))
    m.clusters = unpack_nested(g, "l2_clusters", nc)
    m.thresholds = unpack_nested(g, "l2_thr", nc, as_threshold=True)
    fused = m.score_results(test)
    tm = [torch.from_numpy(x).cuda() for x in maps]
    pooled = ou.extract_roi_aligned_features_from_correct_stride(
        tm, [torch.from_numpy(b) for b in boxes], [torch.from_numpy(s) for s in strides], (img, img), "cuda")
    from ood_in_object_detection_b200.results import Results, batch_shape
    pre = [Results(orig_img=batch_shape(len(boxes), img, img), boxes=r.boxes, extra_item=pooled[i]) for i, r in enumerate(test)]
    m2 = ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, cluster_method="KMeans_5", which_internal_activations="roi_aligned_ftmaps"))
    m2.clusters, m2.thresholds = m.clusters, m.thresholds
    r2 = m2.score_results(pre)
    np.testing.assert_allclose(r2["dist"], fused["dist"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(r2["decision"], fused["decision"]) and np.array_equal(r2["cls_used"], fused["cls_used"])
    assert m2.compute_ood_decision_on_results(pre, LOG) == m.compute_ood_decision_on_results(test, LOG)
    # box-order semantics: same multiset of (box -> distance) when the class lookup happens to agree, and always the
    # box's own class
    m.reference_compat = m2.reference_compat = False
    a, b = m.score_results(test), m2.score_results(pre)
    assert np.array_equal(a["cls_used"], np.concatenate(cls).astype(np.int64))
    assert np.array_equal(a["stride"], np.concatenate(strides).astype(np.int64))
    np.testing.assert_allclose(b["dist"], a["dist"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(a["decision"], b["decision"])


def test_fused_multi_method_helper(ou, golden):
    """compute_ood_decisions_fused (maps uploaded once, metrics in one launch) == each method on its own; host inputs."""
    g = golden("golden_small_kmeans5.npz")
    nc, img = int(g["nc"]), int(g["img"])
    images, maps = scoring_case(g)
    boxes, cls, strides = [im["boxes"] for im in images], [im["cls"] for im in images], [im["strides"] for im in images]
    test = _results(maps, boxes, cls, strides, img, device="cpu")         # per-image views of one host tensor: one H2D copy
    shared = unpack_nested(g, "l2_clusters", nc)
    methods = []
    for tag, klass in _classes(ou):
        m = klass(**dict(DIST_KW, cluster_method="KMeans_5"))
        m.clusters = shared
        m.thresholds = unpack_nested(g, f"{tag}_thr", nc, as_threshold=True)
        methods.append(m)
    rng = np.random.default_rng(0)
    from ood_in_object_detection_b200.results import Results, batch_shape
    lres = [Results(orig_img=batch_shape(len(boxes), img, img), boxes=r.boxes,
                    extra_item=torch.from_numpy(rng.normal(size=(len(b), nc)).astype(np.float32))) for r, b in zip(test, boxes)]
    msp = ou.MSP(**LOGIT_KW)
    msp.thresholds = [0.2] * nc
    out = ou.compute_ood_decisions_fused(methods + [msp], test, LOG, logits_results=lres)
    assert set(out) == {"L1DistancePerStride", "L2DistancePerStride", "CosineDistancePerStride", "MSP"}
    for m in methods:
        assert out[m.name] == m.compute_ood_decision_on_results(test, LOG), m.name
    assert out["MSP"] == msp.compute_ood_decision_on_results(lres, LOG)
    # several logit methods: one upload + one launch with the method mask == each method on its own
    en, ml = ou.Energy(temper=2, **LOGIT_KW), ou.MaxLogit(**LOGIT_KW)
    en.thresholds = list(np.linspace(3.0, 9.0, nc))
    ml.thresholds = list(np.linspace(1.2, 2.6, nc))
    out3 = ou.compute_ood_decisions_fused([methods[0], methods[2], msp, en, ml], test, LOG, logits_results=lres)
    assert out3[methods[0].name] == out[methods[0].name] and out3[methods[2].name] == out[methods[2].name]
    for m in (msp, en, ml):
        alone = m.compute_ood_decision_on_results(lres, LOG)
        assert out3[m.name] == alone, m.name
        assert 0 < sum(map(sum, alone)) < sum(map(len, alone)), m.name       # thresholds that actually split the boxes
    out4 = ou.compute_ood_decisions_fused([methods[0], methods[2]], test, LOG)        # second call: cached tables
    assert out4[methods[0].name] == out[methods[0].name] and out4[methods[2].name] == out[methods[2].name]
    methods[2].thresholds = [[t * 0.5 if t else t for t in per] for per in methods[2].thresholds]   # new thresholds: no stale hit
    out5 = ou.compute_ood_decisions_fused([methods[0], methods[2]], test, LOG)
    assert out5[methods[2].name] == methods[2].compute_ood_decision_on_results(test, LOG)
    own = ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, cluster_method="KMeans_5"))   # different clusters object: own pass
    own.clusters = [list(row) for row in shared]
    own.thresholds = methods[1].thresholds
    out2 = ou.compute_ood_decisions_fused([methods[0], own], test, LOG)
    assert out2["L2DistancePerStride"] == out["L2DistancePerStride"] and out2["L1DistancePerStride"] == out["L1DistancePerStride"]


def test_extractor_helper_matches_reference(ou, golden):
    g = golden("golden_roi_edges.npz")
    img = int(g["img"])
    n = g["n_boxes"]
    boxes, strides = split(g["boxes"], n), split(g["strides"], n)
    maps = [torch.from_numpy(g[f"map{s}"]).cuda() for s in range(3)]
    for all_strides in (False, True):
        out = ou.extract_roi_aligned_features_from_correct_stride(
            maps, [torch.from_numpy(b) for b in boxes], [torch.from_numpy(s) for s in strides], (img, img), "cuda",
            extract_all_strides=all_strides)
        for i in range(3):
            for s in range(3):
                idx, feats = out[i][s]
                ref_idx, ref = g[f"all{int(all_strides)}_idx_{i}_{s}"], g[f"all{int(all_strides)}_feat_{i}_{s}"]
                assert idx.dtype == torch.int16 and np.array_equal(idx.cpu().numpy(), ref_idx)
                if len(ref_idx):
                    assert tuple(feats.shape) == (len(ref_idx), ref.shape[1], 1, 1)
                    got = feats.reshape(len(ref_idx), -1).cpu().numpy()
                    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-30
                    assert np.all(np.abs(got - ref) <= RTOL * scale)


def test_quirks_through_the_class(ou, golden):
    g = golden("golden_quirks.npz")
    img = int(g["img"])
    maps = [g[f"map{s}"] for s in range(3)]
    res = _results(maps, [g["boxes"]], [g["cls"]], [g["strides"]], img)
    clusters = unpack_nested(g, "clusters", 3)
    m = ou.L2DistanceOneClusterPerStride(**DIST_KW)
    m.clusters = clusters
    cases = {"q1": [[1e9] * 3, [1e-9] * 3, [1e-9] * 3], "q4_zero": [[0.0] * 3, [1e9] * 3, [1e9] * 3],
             "q4_empty": [[[]] * 3, [1e9] * 3, [1e9] * 3]}
    for k, thr in cases.items():
        m.thresholds = thr
        assert m.compute_ood_decision_on_results(res, LOG) == [g[f"{k}_decisions"].tolist()], k
    m.thresholds = [[1e9] * 3] * 3
    m.clusters = [clusters[0], [np.empty(0), clusters[1][1], clusters[1][2]], clusters[2]]
    assert m.compute_ood_decision_on_results(res, LOG) == [g["missing_cluster_decisions"].tolist()]
    for c in range(3):                                  # in-place edit of the thresholds is honoured
        m.thresholds[c] = [999.0] * 3
    assert m.compute_ood_decision_on_results(res, LOG) == [g["missing_cluster_thr999_decisions"].tolist()]
    assert m.compute_ood_decision_on_results([], LOG) == []
    with pytest.raises(AssertionError):
        ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, which_internal_activations="logits"))
    with pytest.raises(AssertionError):
        ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, cluster_method="nope"))


def test_logits_methods(ou, golden):
    g = golden("golden_logits.npz")
    nc = int(g["nc"])
    from ood_in_object_detection_b200.results import Results, batch_shape
    n = g["n_boxes"]
    cls, logits = split(g["cls"], n), split(g["logits"], n)
    keep = split(g["fitted_mask"], n)

    def mk(mask=None):
        out = []
        for i in range(len(n)):
            sel = np.ones(len(cls[i]), bool) if mask is None else mask[i]
            b6 = np.zeros((int(sel.sum()), 6), np.float32)
            b6[:, 5] = cls[i][sel]
            out.append(Results(orig_img=batch_shape(len(n), 640, 640), boxes=torch.from_numpy(b6),
                               extra_item=torch.from_numpy(logits[i][sel])))
        return out
    results, results_fitted = mk(), mk(keep)
    acts = [torch.from_numpy(g["train_logits"][g["train_cls"] == c]) if (g["train_cls"] == c).any() else torch.tensor([])
            for c in range(nc)]
    for tag, m in (("MSP", ou.MSP(**LOGIT_KW)), ("Energy", ou.Energy(temper=1, **LOGIT_KW)),
                   ("ODIN", ou.ODIN(temper=1000, **LOGIT_KW)), ("Sigmoid", ou.Sigmoid(**LOGIT_KW))):
        scores = m.compute_scores_from_activations(acts, LOG)
        np.testing.assert_allclose(np.concatenate(scores), g[f"{tag}_fitscores"], rtol=RTOL)
        np.testing.assert_allclose(m.min_score, g[f"{tag}_min"], rtol=RTOL)
        np.testing.assert_allclose(m.max_score, g[f"{tag}_max"], rtol=RTOL)
        thr = m.generate_thresholds(scores, 0.95, LOG)
        assert thr[7] == 0                                   # class without samples keeps 0
        np.testing.assert_allclose(thr, g[f"{tag}_thr"], rtol=RTOL)
        m.thresholds = g[f"{tag}_thr"].tolist()
        m.min_score, m.max_score = g[f"{tag}_min"].tolist(), g[f"{tag}_max"].tolist()
        sc = np.array([m.compute_scores(torch.from_numpy(z), int(c))[0] for z, c in zip(g["logits"][:40], g["cls"][:40])])
        np.testing.assert_allclose(sc, g[f"{tag}_scores"][:40], rtol=RTOL)
        all_sc = m.compute_scores(torch.from_numpy(g["logits"]), g["cls"])
        np.testing.assert_allclose(all_sc, g[f"{tag}_scores"], rtol=RTOL)
        dec = m.compute_ood_decision_on_results(results, LOG)
        assert [len(d) for d in dec] == [int(v) for v in n]
        flat = np.array([v for d in dec for v in d], np.int8)
        thr_box = np.asarray(m.thresholds)[g["cls"].astype(int)]
        near = np.abs(g[f"{tag}_scores"] - thr_box) <= RTOL * np.abs(thr_box)
        assert near.sum() <= 2 and np.array_equal(flat[~near], g[f"{tag}_decisions"][~near]), tag
        ind = m.compute_INDness_scores_on_results(results_fitted, LOG)
        tol = 2e-2 if tag == "ODIN" else 1e-4           # ODIN's InD score range is ~1e-4 wide: INDness amplifies f32 rounding
        np.testing.assert_allclose(np.array([v for d in ind for v in d]), g[f"{tag}_indness"], atol=tol)
    no = ou.NoMethod(**LOGIT_KW)
    assert no.compute_ood_decision_on_results(results, LOG) == [[1] * int(v) for v in n]
    bad = ou.Sigmoid(**LOGIT_KW)                         # the reference asserts cls == argmax (ood_utils.py:1442)
    z = g["logits"][:4].copy()
    z[0, (int(g["cls"][0]) + 1) % nc] = 50.0
    with pytest.raises(AssertionError):
        bad.compute_scores(torch.from_numpy(z), g["cls"][:4])
    ml = ou.MaxLogit(**LOGIT_KW)
    np.testing.assert_array_equal(ml.compute_scores(torch.from_numpy(g["logits"]), g["cls"]), g["logits"].max(1))


def test_sigmoid_on_post_sigmoid_values_fused_and_per_method(ou, golden):
    """`use_values_before_sigmoid=False` (ood_utils.py:1438-1439): the hooked values already went through the detector's
    sigmoid, the Sigmoid score is the value itself.  The fused multi-method pass must decide exactly like the per-method
    call for both flag values."""
    g = golden("golden_logits.npz")
    from ood_in_object_detection_b200.results import Results, batch_shape
    n = g["n_boxes"]
    cls = split(g["cls"], n)
    for before in (True, False):
        vals = g["logits"] if before else torch.sigmoid(torch.from_numpy(g["logits"])).numpy()
        rows = split(vals, n)
        results = []
        for i in range(len(n)):
            b6 = np.zeros((len(cls[i]), 6), np.float32)
            b6[:, 5] = cls[i]
            results.append(Results(orig_img=batch_shape(len(n), 640, 640), boxes=torch.from_numpy(b6),
                                   extra_item=torch.from_numpy(np.ascontiguousarray(rows[i]))))
        kw = dict(LOGIT_KW, use_values_before_sigmoid=before)
        sg, msp = ou.Sigmoid(**kw), ou.MSP(**kw)
        sc = sg.compute_scores(torch.from_numpy(vals), g["cls"])
        want = torch.sigmoid(torch.from_numpy(g["logits"])).numpy()[np.arange(len(g["cls"])), g["cls"].astype(int)]
        if before:
            np.testing.assert_allclose(sc, want, rtol=RTOL)
        else:
            np.testing.assert_array_equal(sc, want)          # the reference returns the stored value (:1443)
        sg.thresholds = [float(np.median(sc))] * int(g["nc"])
        msp.thresholds = [0.5] * int(g["nc"])
        fused = ou.compute_ood_decisions_fused([sg, msp], results, LOG)
        assert list(fused) == ["MSP", "MSP#1"]                # Sigmoid's name is 'MSP' in the reference too
        assert fused["MSP"] == sg.compute_ood_decision_on_results(results, LOG)
        assert fused["MSP#1"] == msp.compute_ood_decision_on_results(results, LOG)
        flat = np.array([v for d in fused["MSP"] for v in d])
        assert 0 < flat.sum() < len(flat)


def test_fusion_rules(ou, golden):
    g = golden("golden_fusion.npz")
    n = g["n"]
    lists = lambda a, cast: [[cast(v) for v in p] for p in split(a, n)]
    m1, m2, m3 = ou.MSP(**LOGIT_KW), ou.MSP(**LOGIT_KW), ou.MSP(**LOGIT_KW)
    for strat in ("and", "or", "score"):
        f = ou.FusionMethod(m1, m2, strat, fusion_method_name="fusion-MSP-MSP", cluster_method="one", **COMMON)
        a, b = (lists(g["s1"], float), lists(g["s2"], float)) if strat == "score" else (lists(g["d1"], int), lists(g["d2"], int))
        out = f.fuse_ood_decisions(a, b)
        assert [len(v) for v in out] == n.tolist()
        assert [v for p in out for v in p] == g[f"fuse_{strat}"].tolist(), strat
    t = ou.TripleFusionMethod(m1, m2, m3, cluster_method="one", **COMMON)
    out = t.fuse_ood_decisions(lists(g["d1"], int), lists(g["d2"], int), lists(g["d3"], int))
    assert [v for p in out for v in p] == g["fuse_majority"].tolist()
    f = ou.FusionMethod(m1, m2, "and", fusion_method_name="x", cluster_method="one", **COMMON)
    f.thresholds = ([1.0], [2.0])
    assert m1.thresholds == [1.0] and m2.thresholds == [2.0] and f.thresholds == ([1.0], [2.0])
    with pytest.raises(ValueError):
        f.thresholds = ([1.0],)
    with pytest.raises(ValueError):
        f.clusters = []
    assert f.cluster_method == 'None' and not f.is_distance_method


def test_generate_thresholds_matches_numpy_lower(ou, golden):
    g = golden("golden_thresholds.npz")
    dm = ou.L2DistanceOneClusterPerStride(**DIST_KW)
    lm = ou.MSP(**LOGIT_KW)
    for n in (5, 6, 11, 21, 101, 1001, 4097):
        v = g[f"v_{n}"]
        for tpr in (0.9, 0.95, 0.99, 0.8):
            t = dm.generate_thresholds([[v, v.astype(np.float64), np.empty(0)]], tpr, LOG)[0]
            got = np.array([x if x != [] else np.nan for x in t], np.float64)
            assert np.array_equal(got, g[f"dist_{n}_{tpr}"], equal_nan=True), (n, tpr)
            assert np.array_equal(np.array(lm.generate_thresholds([v], tpr, LOG), np.float64), g[f"logit_{n}_{tpr}"]), (n, tpr)


def test_kmeans_fit_through_the_class(ou, golden):
    g = golden("golden_kmeans.npz")
    acts = unpack_nested(g, "fit_acts", 3)
    m = ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, cluster_method="KMeans_5"))
    m.clusters = m.generate_clusters(acts, LOG)
    ref = unpack_nested(g, "fit_clusters", 3)
    for c in range(3):
        for s in range(3):
            assert m.clusters[c][s].shape == ref[c][s].shape, (c, s)
    # (class 0, stride 0): 5 blobs, k = 5 -> well separated, centroids to float32 rounding
    np.testing.assert_allclose(m.clusters[0][0], ref[0][0], rtol=RTOL, atol=1e-7)
    # (class 1, stride 2): 4 blobs, k = 5 -> one blob is split in two; which of its points go where depends on the
    # summation order of the x.c dot products (BLAS in the reference), so only the 3 unsplit centroids are comparable
    # to rounding; the split pair must still partition the same blob with the same inertia (k-means parity)
    got, want = m.clusters[1][2], ref[1][2]
    d = np.abs(got[:, None, :] - want[None, :, :]).max(-1)
    assert (d.min(1) < 1e-5).sum() == 3 and np.array_equal(d.argmin(1), np.arange(5))
    x = m.activations_transformation(acts[1][2])
    inertia = lambda cen: float(((x[:, None, :] - cen[None]) ** 2).sum(-1).min(1).sum())
    assert abs(inertia(got) - inertia(want)) <= 2e-3 * inertia(want)
    m.clusters = ref
    scores = m.compute_scores_from_activations(acts, LOG)
    thr = m.generate_thresholds(scores, 0.95, LOG)
    ref_thr = unpack_nested(g, "fit_thr", 3, as_threshold=True)
    for c in range(3):
        for s in range(3):
            if ref_thr[c][s] == []:
                assert thr[c][s] == []
            else:
                np.testing.assert_allclose(thr[c][s], ref_thr[c][s], rtol=RTOL, atol=1e-6)
    # labels through the cluster_utils mirror (cluster_utils.py:62-73)
    from ood_in_object_detection_b200 import cluster_utils
    for tag in "abc":
        lab = cluster_utils.find_optimal_number_of_clusters_one_class_one_stride_and_return_labels(
            g[f"{tag}_x"], f"KMeans_{int(g[f'{tag}_k'])}", "l2", "silhouette", "", LOG)
        assert np.array_equal(lab, g[f"{tag}_labels"]), tag


def test_unknown_proposal_ranking(ou, golden):
    """DistanceMethod.rank_unknown_proposals == ood_utils.py:1031-1084 run with the reference's objects: every fold of the
    [classes, proposals] distance matrix within 1e-5, the closest class exact; channels-last map gives the same ranks."""
    from ood_in_object_detection_b200.custom_hyperparams import CUSTOM_HYP
    g = golden("golden_eul_rank.npz")
    nc, s = int(g["nc"]), int(g["stride"])
    clusters = unpack_nested(g, "clusters", nc)
    fm = torch.from_numpy(g["fm"])
    for tag, klass in _classes(ou):
        m = klass(**DIST_KW)
        m.clusters = clusters
        for op in ("mean", "max", "sum", "min", "geometric_mean", "entropy"):
            got = m.rank_unknown_proposals(fm, g["props"], s, operation=op)
            np.testing.assert_allclose(got, g[f"{tag}_{op}"], rtol=1e-5, atol=1e-6 if tag == "cos" else 0, err_msg=f"{tag} {op}")
        assert CUSTOM_HYP.unk.rank.RANK_BOXES_OPERATION == "entropy"
        np.testing.assert_allclose(m.rank_unknown_proposals(fm.cuda(), torch.from_numpy(g["props"]), s), g[f"{tag}_entropy"],
                                   rtol=1e-5, atol=1e-6)
        CUSTOM_HYP.unk.rank.USE_OOD_THR_TO_REMOVE_PROPS = True
        try:
            mn, closest = m.rank_unknown_proposals(fm, g["props"], s, operation="min")
        finally:
            CUSTOM_HYP.unk.rank.USE_OOD_THR_TO_REMOVE_PROPS = False
        assert np.array_equal(closest, g[f"{tag}_closest"])
        np.testing.assert_allclose(mn, g[f"{tag}_minthr"], rtol=1e-5, atol=1e-6 if tag == "cos" else 0)
        cl = fm[None].contiguous(memory_format=torch.channels_last)[0]          # same values, [H, W, C] memory
        np.testing.assert_allclose(m.rank_unknown_proposals(cl, g["props"], s, operation="sum"), g[f"{tag}_sum"],
                                   rtol=1e-5, atol=1e-6)
    assert len(m.rank_unknown_proposals(fm, np.zeros((0, 4), np.float32), s)) == 0


def test_c4_shaped_methods_through_the_classes(ou, golden):
    """BASELINE config 4 (YOLOv8l maps 256/512/512, K = 10) against golden_c4.npz, frozen from the reference's classes:
    vanilla Cosine, the SDR path of IvisMethodCosine (pool -> normalise -> supplied 32-d reducer -> un-normalised
    scoring, decided on Results), MSP, fusion-MSP-Cosine_cl_stride with and / or / score, EUL ranking of 3 proposals per
    image against all classes."""
    from ood_in_object_detection_b200.results import Results, batch_shape
    from ood_in_object_detection_b200 import synth
    from tests.helpers import Projection
    g = golden("golden_c4.npz")
    nc, img, B, k = int(g["nc"]), int(g["img"]), int(g["batch"]), int(g["k"])
    ch = tuple(int(c) for c in g["channels"])
    maps = synth.feature_maps(int(g["seed"]) + 2, B, ch, tuple(img // s for s in synth.STRIDES))
    n = g["n_boxes"]
    boxes, cls, strides, logit = split(g["boxes"], n), split(g["cls"], n), split(g["strides"], n), split(g["logits"], n)
    results = _results(maps, boxes, cls, strides, img)
    results_l = []
    for i in range(B):
        b6 = np.concatenate([boxes[i], np.full((len(boxes[i]), 1), 0.5, np.float32), cls[i][:, None]], 1).astype(np.float32)
        results_l.append(Results(orig_img=batch_shape(B, img, img), boxes=torch.from_numpy(b6), extra_item=torch.from_numpy(logit[i])))
    kw = dict(DIST_KW, cluster_method=f"KMeans_{k}")
    cos, ivis = ou.CosineDistanceOneClusterPerStride(**kw), ou.IvisMethodCosine(**kw)
    assert ivis.name == "IvisCosineDistancePerStride" and ivis.metric == "cosine"
    ivis.set_reducers([Projection(900 + s, ch[s], 32) for s in range(3)])          # one reducer per stride, like the reference
    flat = lambda d: np.array([v for im in d for v in im], np.int8)
    dec = {}
    for tag, m in (("cos", cos), ("ivis", ivis)):
        m.clusters = unpack_nested(g, f"{tag}_clusters", nc)
        m.thresholds = unpack_nested(g, f"{tag}_thr", nc, as_threshold=True)
        r = m.score_results(results)
        np.testing.assert_allclose(r["dist"], g[f"{tag}_dist"], rtol=RTOL, atol=5e-7)
        assert np.array_equal(r["cls_used"], g[f"{tag}_cls_used"]) and np.array_equal(r["stride"], g[f"{tag}_stride_of"])
        d = m.compute_ood_decision_on_results(results, LOG)
        assert [len(v) for v in d] == [int(v) for v in n]
        thr_box = np.array([m.thresholds[c][s] if m.thresholds[c][s] != [] else np.inf
                            for c, s in zip(g[f"{tag}_cls_used"], g[f"{tag}_stride_of"])])
        near = np.abs(g[f"{tag}_dist"] - thr_box) <= RTOL * thr_box
        assert near.sum() <= 2 and np.array_equal(flat(d)[~near], g[f"{tag}_decisions"][~near]), tag
        dec[tag] = (d, near)
    msp = ou.MSP(**LOGIT_KW)
    msp.thresholds, msp.min_score, msp.max_score = g["msp_thr"].tolist(), g["msp_min"].tolist(), g["msp_max"].tolist()
    d_msp = msp.compute_ood_decision_on_results(results_l, LOG)
    assert np.array_equal(flat(d_msp), g["msp_decisions"])
    near = dec["cos"][1]
    for strat in ("and", "or", "score"):
        thr_pair = (msp.thresholds, cos.thresholds)
        f = ou.FusionMethod(msp, cos, strat, fusion_method_name="fusion-MSP-Cosine_cl_stride", cluster_method=f"KMeans_{k}", **COMMON)
        f.thresholds = thr_pair
        out = f.compute_ood_decision_on_results(results_l, LOG, results2=results)
        assert np.array_equal(flat(out)[~near], g[f"fusion_{strat}"][~near]), strat
    ind = msp.compute_INDness_scores_on_results(results_l, LOG)
    np.testing.assert_allclose(np.array([v for im in ind for v in im]), g["msp_indness"], atol=1e-4)
    # EUL: 3 proposals per image on the stride-1 map against every class that has clusters there
    for i in range(B):
        fm = torch.from_numpy(maps[1][i]).cuda()
        for op, key in (("entropy", "eul_entropy"), ("min", "eul_min")):
            ranks = cos.rank_unknown_proposals(fm, g["eul_props"][i], 1, operation=op)
            np.testing.assert_allclose(ranks, g[key][i], rtol=2e-4 if op == "entropy" else RTOL, atol=1e-6)
    # an SDR method fits in the embedded space too: clusters of the reference's fit have 32 columns
    assert ivis.clusters[0][0].shape[1] == 32
    with pytest.raises(RuntimeError):
        ou.IvisMethodL2(**kw).activations_transformation(np.zeros((2, ch[0]), np.float32), cls_idx=0, stride_idx=0)


def test_fit_with_segments_of_4096_rows_and_more(ou, golden):
    """Segments of 4096 .. 6000 rows through the class surface (generate_clusters KMeans_5 -> compute_scores_from_activations ->
    generate_thresholds) against golden_bigfit.npz from the reference: FP32 fit scores (the arithmetic of the decision
    path) to 1e-5, and the opt-in tensor-core scorer (vec_score_one(tensor_core=True)) to its 1e-3 tier."""
    from ood_in_object_detection_b200 import ops
    from tests.test_oracle_vs_golden import bigfit_activations
    g = golden("golden_bigfit.npz")
    acts = bigfit_activations(g)
    for tag, cls in _classes(ou):
        m = cls(**dict(DIST_KW, cluster_method="KMeans_5"))
        clusters = m.generate_clusters(acts, LOG)
        for c in range(3):
            for s in range(3):
                ref = g[f"{tag}_clusters_{c}_{s}"]
                assert np.asarray(clusters[c][s]).shape == ref.shape, (tag, c, s)
                if ref.size:
                    np.testing.assert_allclose(clusters[c][s], ref, rtol=RTOL, atol=1e-7)
        m.clusters = unpack_nested(g, f"{tag}_clusters", 3)
        scores = m.compute_scores_from_activations(acts, LOG)
        thr = m.generate_thresholds(scores, 0.95, LOG)
        for c in range(3):
            for s in range(3):
                ref = g[f"{tag}_scores_{c}_{s}"]
                # cosine distances of these tight blobs are ~3e-3 = 1 - (a float32 dot product near 1): both sides carry the
                # float32 rounding of that difference (~1e-7 absolute each), hence the absolute floor for cosine
                atol = 1.2e-6 if tag == "cos" else 2e-7
                if ref.size:
                    np.testing.assert_allclose(np.asarray(scores[c][s], np.float64), ref, rtol=RTOL, atol=atol)
                gt = g[f"{tag}_thr_{c}_{s}"]
                assert (thr[c][s] == [] and gt.ndim == 1) or thr[c][s] == pytest.approx(float(gt), rel=RTOL, abs=atol)
        if tag in ("l2", "cos"):                            # the tcgen05 scorer on the same segments
            for c, s in ((0, 0), (1, 1), (1, 0)):
                x = ops.normalize_rows(torch.from_numpy(acts[c][s].reshape(len(acts[c][s]), -1)).cuda())
                cent = torch.from_numpy(np.ascontiguousarray(m.clusters[c][s])).cuda()
                unit = torch.from_numpy(ops._unit_rows(m.clusters[c][s])).cuda()
                slot = ops.METRIC_SLOT["l2" if tag == "l2" else "cosine"]
                d, _ = ops.vec_score_one(x, [0, x.shape[0]], cent, unit, [0], [cent.shape[0]], slot, normalize=False, tensor_core=True)
                np.testing.assert_allclose(d[slot].cpu().numpy(), g[f"{tag}_scores_{c}_{s}"], rtol=1e-3, atol=3e-6)
