"""The CPU oracle against the golden vectors frozen from the real reference (CPU-only tests)."""
import numpy as np
import pytest

from oracle import decide, distance, fit, kmeans, logits, roi_align
from tests.helpers import scoring_case, split, train_case, unpack_nested

F32 = np.float32


def test_roi_align_matches_torchvision():
    """numpy restatement == torchvision's C++ op, bit for bit, incl. Q5 border boxes."""
    import torch
    import torchvision
    rng = np.random.default_rng(0)
    for (C, H, W, img) in [(16, 80, 80, 640), (8, 20, 20, 640), (8, 25, 23, 736)]:
        fm = rng.standard_normal((3, C, H, W)).astype(F32)
        K = 60
        ctr = rng.uniform(0, img, (K, 2))
        sz = np.exp(rng.uniform(np.log(4), np.log(img * 0.9), (K, 2)))
        rois = np.stack([rng.integers(0, 3, K), np.clip(ctr[:, 0] - sz[:, 0] / 2, 0, img), np.clip(ctr[:, 1] - sz[:, 1] / 2, 0, img),
                         np.clip(ctr[:, 0] + sz[:, 0] / 2, 0, img), np.clip(ctr[:, 1] + sz[:, 1] / 2, 0, img)], 1).astype(F32)
        rois[0, 1:] = [img - 10, img - 10, img, img]
        rois[1, 1:] = [0, 0, img, img]
        rois[2, 1:] = [5, 5, 5, 5]
        rois[3, 1:] = [-20, -30, 50, 60]
        rois[4, 1:] = [img - 5, img - 5, img + 40, img + 40]
        ref = torchvision.ops.roi_align(torch.from_numpy(fm), torch.from_numpy(rois), (1, 1), spatial_scale=W / img,
                                        aligned=False).numpy()[:, :, 0, 0]
        assert np.array_equal(roi_align.roi_align_1x1(fm, rois, W / img), ref)


def test_extractor_edges(golden):
    g = golden("golden_roi_edges.npz")
    n = g["n_boxes"]
    maps = [g[f"map{s}"] for s in range(3)]
    boxes, strides = split(g["boxes"], n), split(g["strides"], n)
    for all_strides in (False, True):
        out = roi_align.extract_roi_aligned_features_from_correct_stride(maps, boxes, strides, (int(g["img"]),) * 2,
                                                                         extract_all_strides=all_strides)
        for i in range(3):
            for s in range(3):
                assert np.array_equal(out[i][s][0], g[f"all{int(all_strides)}_idx_{i}_{s}"])
                exp = g[f"all{int(all_strides)}_feat_{i}_{s}"]
                got = np.asarray(out[i][s][1], F32).reshape(exp.shape)
                assert np.array_equal(got, exp)


def test_separable_weights_equal_sample_sum(golden):
    """The factorised form the CUDA kernel uses is the same function (to float32 rounding)."""
    g = golden("golden_roi_edges.npz")
    fm = g["map0"]
    H = W = fm.shape[-1]
    sc = W / int(g["img"])
    for b in g["boxes"]:
        ref = roi_align.roi_align_1x1(fm, np.concatenate([[0], b])[None], sc)[0]
        sw, sh, rw, rh, gw, gh = roi_align.roi_params(b, sc)
        y0, wy = roi_align.axis_weights(sh, rh, gh, H)
        x0, wx = roi_align.axis_weights(sw, rw, gw, W)
        if len(wy) == 0 or len(wx) == 0:
            assert not ref.any()
            continue
        win = fm[0, :, y0:y0 + len(wy), x0:x0 + len(wx)].astype(np.float64)
        sep = np.einsum("cyx,y,x->c", win, wy.astype(np.float64), wx.astype(np.float64)) / max(gh * gw, 1)
        np.testing.assert_allclose(sep, ref, rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["golden_c1_one.npz", "golden_small_kmeans5.npz"])
def test_decisions_and_distances(golden, name):
    g = golden(name)
    images, _ = scoring_case(g)
    nc = int(g["nc"])
    for tag, metric in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine")):
        clusters = unpack_nested(g, f"{tag}_clusters", nc)
        thr = unpack_nested(g, f"{tag}_thr", nc, as_threshold=True)
        dec, det = decide.distance_decisions(images, clusters, thr, metric, return_details=True)
        assert np.array_equal(np.concatenate([np.asarray(d, np.int8) for d in dec]), g[f"{tag}_decisions"])
        d = np.array([t[0] for im in det for t in im])
        np.testing.assert_allclose(d, g[f"{tag}_dist"], rtol=1e-6, atol=1e-7)
        assert np.array_equal(np.array([t[2] for im in det for t in im]), g[f"{tag}_cls_used"])
        assert np.array_equal(np.array([t[4] for im in det for t in im]), g[f"{tag}_box_of"])
    assert np.all(g["cos_indness"] == -1)       # Q2


def test_fit_clusters_scores_thresholds(golden):
    g = golden("golden_c1_one.npz")
    nc = int(g["nc"])
    maps, boxes, cls, strides = train_case(g)
    img = int(g["img"])
    acts = [[[] for _ in range(3)] for _ in range(nc)]
    for i in range(len(boxes)):
        feats = roi_align.extract_roi_aligned_features_from_correct_stride([m[i:i + 1] for m in maps], [boxes[i]], [strides[i]], (img, img))[0]
        for s, (idx, fm) in enumerate(feats):
            for j, b in enumerate(idx):
                acts[int(cls[i][int(b)])][s].append(fm[j])
    acts = [[np.stack(v, 0) if len(v) else np.empty(0) for v in row] for row in acts]
    for tag, metric in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine")):
        clusters = fit.generate_clusters(acts, "one")
        scores, mn, mx = fit.compute_scores_from_activations(acts, clusters, metric)
        thr = fit.generate_thresholds(scores, 0.95, True, True)
        for c in range(nc):
            for s in range(3):
                np.testing.assert_allclose(clusters[c][s], g[f"{tag}_clusters_{c}_{s}"], rtol=1e-6, atol=1e-7)
                np.testing.assert_allclose(np.asarray(scores[c][s], np.float64), g[f"{tag}_fitscores_{c}_{s}"], rtol=2e-6, atol=2e-7)
                gt = g[f"{tag}_thr_{c}_{s}"]
                if gt.ndim == 0:
                    assert thr[c][s] == pytest.approx(float(gt), rel=2e-6, abs=2e-7)
                else:
                    assert thr[c][s] == []


def test_quirks(golden):
    g = golden("golden_quirks.npz")
    img = int(g["img"])
    im = dict(maps=[g[f"map{s}"][0] for s in range(3)], boxes=g["boxes"], cls=g["cls"], strides=g["strides"], img_hw=(img, img))
    clusters = unpack_nested(g, "clusters", 3)
    cases = {"q1": [[1e9] * 3, [1e-9] * 3, [1e-9] * 3], "q4_zero": [[0.0] * 3, [1e9] * 3, [1e9] * 3],
             "q4_empty": [[[]] * 3, [1e9] * 3, [1e9] * 3]}
    for k, thr in cases.items():
        assert decide.distance_decisions([im], clusters, thr, "l2")[0] == g[f"{k}_decisions"].tolist()
    assert decide.distance_decisions([im], clusters, cases["q1"], "l2", compat_q1=False)[0] == [0, 1, 0]
    miss = [clusters[0], [np.empty(0), clusters[1][1], clusters[1][2]], clusters[2]]
    assert decide.distance_decisions([im], miss, [[1e9] * 3] * 3, "l2")[0] == g["missing_cluster_decisions"].tolist()
    assert decide.distance_decisions([im], miss, [[999.0] * 3] * 3, "l2")[0] == g["missing_cluster_thr999_decisions"].tolist()


def test_logits(golden):
    g = golden("golden_logits.npz")
    n = g["n_boxes"]
    images = [dict(cls=c, logits=z) for c, z in zip(split(g["cls"], n), split(g["logits"], n))]
    nc = int(g["nc"])
    acts = [g["train_logits"][g["train_cls"] == c] for c in range(nc)]
    for tag, temper in (("MSP", 1.0), ("Energy", 1.0), ("ODIN", 1000.0), ("Sigmoid", 1.0)):
        sc = logits.scores(g["logits"], g["cls"], tag, temper)
        np.testing.assert_allclose(sc, g[f"{tag}_scores"], rtol=2e-6, atol=1e-7)
        fs, mn, mx = fit.logits_scores_from_activations(acts, tag, temper)
        np.testing.assert_allclose(np.concatenate(fs), g[f"{tag}_fitscores"], rtol=2e-6, atol=1e-7)
        thr = fit.generate_thresholds(fs, 0.95, False, False)
        np.testing.assert_allclose(thr, g[f"{tag}_thr"], rtol=2e-6, atol=1e-7)
        # decide with the reference's thresholds so that rounding of the oracle's own fit cannot move a decision
        dec = decide.logit_decisions(images, tag, g[f"{tag}_thr"], temper)
        assert np.array_equal(np.concatenate([np.asarray(d, np.int8) for d in dec]), g[f"{tag}_decisions"])
        fitted = [dict(cls=im["cls"][im["cls"] != 7], logits=im["logits"][im["cls"] != 7]) for im in images]
        ind = decide.logit_indness(fitted, tag, g[f"{tag}_thr"], g[f"{tag}_min"], g[f"{tag}_max"], temper)
        tol = 2e-2 if tag == "ODIN" else 1e-4     # ODIN's score range is ~1e-4 wide: INDness amplifies f32 rounding
        np.testing.assert_allclose(np.concatenate([np.asarray(v) for v in ind]), g[f"{tag}_indness"], atol=tol)


def test_fusion(golden):
    g = golden("golden_fusion.npz")
    n = g["n"]
    d1, d2, d3 = (list(map(list, split(g[k], n))) for k in ("d1", "d2", "d3"))
    s1, s2 = (list(map(list, split(g[k], n))) for k in ("s1", "s2"))
    cat = lambda x: np.concatenate([np.asarray(v, np.int8) for v in x])
    assert np.array_equal(cat(decide.fuse(d1, d2, "and")), g["fuse_and"])
    assert np.array_equal(cat(decide.fuse(d1, d2, "or")), g["fuse_or"])
    assert np.array_equal(cat(decide.fuse(s1, s2, "score")), g["fuse_score"])
    assert np.array_equal(cat(decide.fuse3(d1, d2, d3)), g["fuse_majority"])


def test_thresholds_percentile_lower(golden):
    g = golden("golden_thresholds.npz")
    for n in (5, 6, 11, 21, 101, 1001, 4097):
        v = g[f"v_{n}"]
        for tpr in (0.9, 0.95, 0.99, 0.8):
            t = fit.generate_thresholds([[v, v.astype(np.float64), np.empty(0)]], tpr, True, True)[0]
            exp = g[f"dist_{n}_{tpr}"]
            for a, b in zip(t, exp):
                assert (a == [] and np.isnan(b)) or a == b
            assert np.array_equal(np.asarray(fit.generate_thresholds([v], tpr, False, False), np.float64), g[f"logit_{n}_{tpr}"])
            if n > 5:   # explicit order-statistic form used by the GPU radix select
                k = fit.percentile_lower_index(n, 100 * tpr)
                assert float(np.sort(v)[k]) == exp[0]
                k = fit.percentile_lower_index(n, (1 - tpr) * 100)
                assert float(np.sort(v)[k]) == g[f"logit_{n}_{tpr}"][0]


def test_kmeans_labels(golden):
    """numpy restatement of sklearn KMeans == labels from the reference's call site."""
    g = golden("golden_kmeans.npz")
    for tag in "abc":
        x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
        lab = kmeans.kmeans_fit_predict(x, min(k, len(x)), random_state=10)[0]
        assert np.array_equal(lab, g[f"{tag}_labels"]), tag
    acts = unpack_nested(g, "fit_acts", 3)
    clusters = fit.generate_clusters(acts, "KMeans_5")
    scores, _, _ = fit.compute_scores_from_activations(acts, clusters, "l2")
    thr = fit.generate_thresholds(scores, 0.95, True, True)
    for c in range(3):
        for s in range(3):
            np.testing.assert_allclose(clusters[c][s], g[f"fit_clusters_{c}_{s}"], rtol=1e-5, atol=1e-6)
            gt = g[f"fit_thr_{c}_{s}"]
            assert (thr[c][s] == [] and gt.ndim == 1) or thr[c][s] == pytest.approx(float(gt), rel=1e-5)


def test_cpu_port_matches_golden(golden):
    """The timing port (same library calls as the reference) returns the reference's decisions."""
    import torch
    from oracle import cpu_path
    g = golden("golden_small_kmeans5.npz")
    images, _ = scoring_case(g)
    nc = int(g["nc"])
    timg = [dict(maps=[torch.from_numpy(m) for m in im["maps"]], boxes=torch.from_numpy(im["boxes"]),
                 cls=torch.from_numpy(im["cls"]), strides=torch.from_numpy(im["strides"]), img_hw=im["img_hw"]) for im in images]
    for tag, metric in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine")):
        dec = cpu_path.distance_decisions(timg, unpack_nested(g, f"{tag}_clusters", nc),
                                          unpack_nested(g, f"{tag}_thr", nc, as_threshold=True), metric)
        assert np.array_equal(np.concatenate([np.asarray(d, np.int8) for d in dec]), g[f"{tag}_decisions"])
    gl = golden("golden_logits.npz")
    n = gl["n_boxes"]
    limg = [dict(cls=torch.from_numpy(c), logits=torch.from_numpy(z)) for c, z in zip(split(gl["cls"], n), split(gl["logits"], n))]
    for tag, temper in (("MSP", 1.0), ("Energy", 1.0), ("ODIN", 1000.0), ("Sigmoid", 1.0)):
        dec = cpu_path.logit_decisions(limg, tag, gl[f"{tag}_thr"].tolist(), temper)
        assert np.array_equal(np.concatenate([np.asarray(d, np.int8) for d in dec]), gl[f"{tag}_decisions"])


def test_silhouette_and_k_search(golden):
    """oracle/silhouette.py against sklearn (the dependency it restates) and against the per-k scores and final labels the
    reference's own k-search produced (cluster_utils.py:203-356; tests/golden/make_golden.py::golden_ksearch)."""
    from sklearn.metrics import calinski_harabasz_score, silhouette_score
    from oracle import silhouette as S
    g = golden("golden_ksearch.npz")
    for tag in "abcde":
        x, metric, perf = g[f"{tag}_x"], str(g[f"{tag}_metric"]), str(g[f"{tag}_perf"])
        ref_scores, ref_labels = g[f"{tag}_scores"], g[f"{tag}_labels"]
        assert g[f"{tag}_ks"].tolist() == S.RANGE_OF_CLUSTERS
        if len(set(ref_labels.tolist())) > 1:
            assert abs(S.silhouette_score(x, ref_labels, metric) - silhouette_score(x, ref_labels, metric=metric)) < 1e-6
            assert np.isclose(S.calinski_harabasz_score(x, ref_labels), calinski_harabasz_score(x, ref_labels), rtol=1e-6)
        scores, labels = S.k_search(x, metric, perf)
        np.testing.assert_allclose(scores, ref_scores, rtol=2e-6, atol=2e-6)
        assert np.array_equal(labels, ref_labels), tag


def test_eul_proposal_ranking(golden):
    """oracle/eul.py against the distances and folds produced with the reference's own method objects
    (ood_utils.py:1031-1084; tests/golden/make_golden.py::golden_eul_rank)."""
    from oracle import eul
    from tests.helpers import unpack_nested
    g = golden("golden_eul_rank.npz")
    clusters = unpack_nested(g, "clusters", int(g["nc"]))
    for tag, metric in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine")):
        d = eul.distance_matrix(g["fm"], g["props"], clusters, int(g["stride"]), metric)
        np.testing.assert_allclose(d, g[f"{tag}_matrix"], rtol=1e-5, atol=5e-7 if metric == "cosine" else 0)
        for op in ("mean", "max", "sum", "min", "geometric_mean", "entropy"):
            np.testing.assert_allclose(eul.fold(g[f"{tag}_matrix"], op), g[f"{tag}_{op}"], rtol=1e-6)
        mn, closest = eul.fold(g[f"{tag}_matrix"], "min", True)
        assert np.array_equal(closest, g[f"{tag}_closest"]) and np.allclose(mn, g[f"{tag}_minthr"])


def _c4_images(g):
    from ood_in_object_detection_b200 import synth
    seed, img, B = int(g["seed"]), int(g["img"]), int(g["batch"])
    ch = tuple(int(c) for c in g["channels"])
    maps = synth.feature_maps(seed + 2, B, ch, tuple(img // s for s in synth.STRIDES))
    n = g["n_boxes"]
    boxes, cls, strides, logit = split(g["boxes"], n), split(g["cls"], n), split(g["strides"], n), split(g["logits"], n)
    return [dict(maps=[m[i] for m in maps], boxes=boxes[i], cls=cls[i], strides=strides[i], logits=logit[i], img_hw=(img, img))
            for i in range(B)], maps


def test_c4_shaped_methods(golden):
    """BASELINE config 4 shapes (YOLOv8l maps, K = 10): vanilla Cosine, the SDR path (normalise -> 32-d embedding ->
    un-normalised scoring), MSP, fusion and / or / score, EUL proposal ranking -- oracle vs the reference's outputs."""
    from oracle import eul
    from tests.helpers import Projection
    g = golden("golden_c4.npz")
    nc = int(g["nc"])
    images, maps = _c4_images(g)
    ch = [int(c) for c in g["channels"]]
    proj = [Projection(900 + s, ch[s], 32) for s in range(3)]
    embed = lambda x, c, s: proj[s].transform(distance.normalize_rows(x))
    dec = {}
    for tag, tf in (("cos", None), ("ivis", embed)):
        clusters = unpack_nested(g, f"{tag}_clusters", nc)
        thr = unpack_nested(g, f"{tag}_thr", nc, as_threshold=True)
        d, det = decide.distance_decisions(images, clusters, thr, "cosine", return_details=True, transform=tf)
        dist = np.array([t[0] for im in det for t in im])
        np.testing.assert_allclose(dist, g[f"{tag}_dist"], rtol=1e-5, atol=2e-7)
        flat = np.concatenate([np.asarray(v, np.int8) for v in d])
        thr_box = np.array([thr[c][s] if thr[c][s] != [] else np.inf for c, s in zip(g[f"{tag}_cls_used"], g[f"{tag}_stride_of"])])
        near = np.abs(g[f"{tag}_dist"] - thr_box) <= 1e-5 * thr_box
        assert near.sum() <= 2 and np.array_equal(flat[~near], g[f"{tag}_decisions"][~near]), tag
        assert np.array_equal(np.array([t[2] for im in det for t in im]), g[f"{tag}_cls_used"])
        dec[tag] = d
        assert clusters[0][0].shape[1] == (32 if tag == "ivis" else ch[0])
    d_msp = decide.logit_decisions(images, "MSP", g["msp_thr"], 1.0)
    assert np.array_equal(np.concatenate([np.asarray(v, np.int8) for v in d_msp]), g["msp_decisions"])
    cat = lambda x: np.concatenate([np.asarray(v, np.int8) for v in x])
    d_cos = [list(v) for v in split(g["cos_decisions"], g["n_boxes"])]
    assert np.array_equal(cat(decide.fuse(d_msp, d_cos, "and")), g["fusion_and"])
    assert np.array_equal(cat(decide.fuse(d_msp, d_cos, "or")), g["fusion_or"])
    ind_msp = decide.logit_indness(images, "MSP", g["msp_thr"], g["msp_min"], g["msp_max"], 1.0)
    np.testing.assert_allclose(np.concatenate([np.asarray(v) for v in ind_msp]), g["msp_indness"], atol=1e-4)
    ind_cos = decide.distance_indness(images, clusters=unpack_nested(g, "cos_clusters", nc),
                                      thresholds=unpack_nested(g, "cos_thr", nc, as_threshold=True), metric="cosine")
    assert np.all(g["cos_indness"] == -1) and all(v == -1 for row in ind_cos for v in row)       # Q2
    assert np.array_equal(cat(decide.fuse(ind_msp, ind_cos, "score")), g["fusion_score"])
    clusters = unpack_nested(g, "cos_clusters", nc)
    for i in range(len(images)):
        d = eul.distance_matrix(maps[1][i], g["eul_props"][i], clusters, 1, "cosine")
        np.testing.assert_allclose(d, g["eul_matrix"][i], rtol=1e-5, atol=5e-7)
        np.testing.assert_allclose(eul.fold(g["eul_matrix"][i], "entropy"), g["eul_entropy"][i], rtol=1e-6)
        np.testing.assert_allclose(eul.fold(g["eul_matrix"][i], "min"), g["eul_min"][i], rtol=1e-6)


def bigfit_activations(g):
    from ood_in_object_detection_b200 import synth
    acts = [[np.empty(0) for _ in range(3)] for _ in range(3)]
    for c, s, sd, n, d in g["spec"]:
        acts[int(c)][int(s)] = np.abs(synth.blob_vectors(int(sd), int(n), int(d), 5, 7.0, unit_norm=False)[0])[:, :, None, None]
    return acts


def test_fit_with_segments_of_4096_rows_and_more(golden):
    """Fit on segments of 4096 .. 6000 rows (KMeans_5 -> member means -> scores -> thresholds) against the reference's
    outputs: the sizes at which the CUDA fit scorer runs its large-input path."""
    g = golden("golden_bigfit.npz")
    acts = bigfit_activations(g)
    clusters = fit.generate_clusters(acts, "KMeans_5")
    for tag, metric in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine")):
        scores, mn, mx = fit.compute_scores_from_activations(acts, clusters, metric)
        thr = fit.generate_thresholds(scores, 0.95, True, True)
        for c in range(3):
            for s in range(3):
                ref_cl = g[f"{tag}_clusters_{c}_{s}"]
                assert np.asarray(clusters[c][s]).shape == ref_cl.shape, (tag, c, s)
                if ref_cl.size:
                    np.testing.assert_allclose(clusters[c][s], ref_cl, rtol=1e-5, atol=1e-7)
                    np.testing.assert_allclose(np.asarray(scores[c][s], np.float64), g[f"{tag}_scores_{c}_{s}"], rtol=2e-5, atol=2e-7)
                gt = g[f"{tag}_thr_{c}_{s}"]
                assert (thr[c][s] == [] and gt.ndim == 1) or thr[c][s] == pytest.approx(float(gt), rel=2e-5)


def test_nms_with_payload(golden):
    """oracle/nms.py against `non_max_suppression_old` of the reference's ultralytics fork (boxes, confidences, classes, the
    gathered logits and strides), bit for bit, at two threshold settings."""
    from oracle import nms
    from tests.helpers import nms_inputs
    g = golden("golden_nms.npz")
    pred, logits, strides = nms_inputs(int(g["seed"]))
    for tag in ("a", "b", "agn"):
        conf, iou, max_det = g[f"{tag}_cfg"]
        out, extra, st = nms.non_max_suppression(pred, conf, iou, int(max_det), extra_item=logits, strides=strides, agnostic=tag == "agn")
        assert [len(o) for o in out] == g[f"{tag}_n"].tolist()
        assert np.array_equal(np.concatenate(out), g[f"{tag}_det"])
        assert np.array_equal(np.concatenate([e.reshape(len(o), -1) for e, o in zip(extra, out)]), g[f"{tag}_extra"])
        assert np.array_equal(np.concatenate(st), g[f"{tag}_strides"])


def test_postprocess_restated(golden):
    """The OoD branch of `DetectionPredictor.postprocess` restated with oracle/nms.py (NMS + payload, box clipping to the
    input, raw-logit heads pass through a sigmoid first) against the frozen outputs of the reference's method."""
    from oracle import nms
    from tests.helpers import postprocess_inputs
    g = golden("golden_postprocess.npz")
    pred, logits, _ = postprocess_inputs()
    stride_of = np.concatenate([np.full((320 // s) ** 2, i, np.float32) for i, s in enumerate((8, 16, 32))])
    for tag, conf, extra_item, strides in (("fs", 0.25, None, stride_of), ("fs_hi", 0.97, None, stride_of),
                                           ("pos", 0.25, None, np.arange(len(stride_of), dtype=np.float32)),
                                           ("lg_raw", 0.25, np.concatenate([pred[:, :4], logits], 1), None),
                                           ("lg", 0.25, pred, None)):
        head = pred
        if tag == "lg_raw":                                      # predict.py:199-209: the head emits raw logits, NMS sees their sigmoid
            import torch
            head = np.concatenate([pred[:, :4], torch.from_numpy(logits).sigmoid().numpy()], 1)
        res = nms.non_max_suppression(head, conf, 0.45, 300, extra_item=extra_item, strides=strides)
        out, payload = res[0], (res[2] if strides is not None else res[1])
        assert [len(o) for o in out] == g[f"{tag}_n"].tolist()
        boxes = np.concatenate([o.reshape(-1, 6) for o in out]).copy()
        boxes[:, [0, 2]] = boxes[:, [0, 2]].clip(0, 320)
        boxes[:, [1, 3]] = boxes[:, [1, 3]].clip(0, 320)
        assert np.array_equal(boxes, g[f"{tag}_boxes"])
        if strides is not None:
            assert np.array_equal(np.concatenate([np.asarray(p).reshape(-1) for p in payload]).astype(np.float64), g[f"{tag}_extra"])
        else:
            assert np.array_equal(np.concatenate([np.asarray(p).reshape(len(o), -1)[:, 4:] for p, o in zip(payload, out)]), g[f"{tag}_extra"])

