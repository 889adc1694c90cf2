"""Freeze golden vectors by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Every output array below comes out of a reference function
(imported through oracle/ref_shim.py); inputs are either stored next to the outputs or
regenerated bit-identically from the stored seed by ood_in_object_detection_b200.synth.

The reference has no tests or fixtures for this path, so these files are the
pin for both the oracle (tests/test_oracle_vs_golden.py) and the CUDA path (tests/test_gpu_*.py).
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from ood_in_object_detection_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LOG = logging.getLogger("golden")
LOG.setLevel(logging.ERROR)
F32 = np.float32


def _images(ref, maps, det, img):
    """Reference `Results` for a batch; maps = list of 3 [B,C,H,W] arrays."""
    B = len(det["boxes"])
    tm = [torch.from_numpy(m) for m in maps]
    boxes6 = [torch.from_numpy(np.concatenate([det["boxes"][i], det["conf"][i][:, None], det["cls"][i][:, None]], 1))
              for i in range(B)]
    return ref_shim.make_results(ref, [[t[i] for t in tm] for i in range(B)], boxes6,
                                 strides=[torch.from_numpy(s) for s in det["strides"]],
                                 batch_hw=(img, img), n_batch=B)


def _collect_activations(ref, maps, det, img, nc):
    """Per (class, stride) pooled vectors of every box, grouped the way the reference's fit path
    groups them (`ood_utils.py:1764-1776`, correct bbox_idx, all predictions taken as valid)."""
    acts = [[[] for _ in range(3)] for _ in range(nc)]
    tm = [torch.from_numpy(m) for m in maps]
    for i in range(len(det["boxes"])):
        feats = ref.extract_roi_aligned_features_from_correct_stride(
            ftmaps=[t[i][None] for t in tm], boxes=[torch.from_numpy(det["boxes"][i])],
            strides=[torch.from_numpy(det["strides"][i])], img_shape=(img, img), device="cpu")[0]
        for s, (idx, fm) in enumerate(feats):
            for j, b in enumerate(idx):
                acts[int(det["cls"][i][int(b)])][s].append(fm[j].numpy())
    for c in range(nc):
        for s in range(3):
            acts[c][s] = np.stack(acts[c][s], 0) if len(acts[c][s]) else np.empty(0)
    return acts


def _ref_distances_q1(ref, method, results):
    """Per-box distance / class / stride in the reference's own (Q1, stride-major) order, obtained by
    calling the reference's extractor, `activations_transformation` and `compute_distance`."""
    dist, cls_used, stride_of, box_of = [], [], [], []
    for res in results:
        ftmaps, strides = res.extra_item
        feats = ref.extract_roi_aligned_features_from_correct_stride(
            ftmaps=[ft[None] for ft in ftmaps], boxes=[res.boxes.xyxy], strides=[strides],
            img_shape=res.orig_img.shape[1:3], device="cpu")[0]
        cls_all = res.boxes.cls.cpu()
        d_i, c_i, s_i, b_i = [], [], [], []
        for s, (idx, fm) in enumerate(feats):
            for j, v in enumerate(fm):
                c = int(cls_all[j])
                if len(method.clusters[c][s]) == 0:
                    d = 1000.0
                else:
                    x = method.activations_transformation(v.unsqueeze(0).numpy(), cls_idx=c, stride_idx=s)
                    d = float(method.compute_distance(method.clusters[c][s], x)[0])
                d_i.append(d), c_i.append(c), s_i.append(s), b_i.append(int(idx[j]))
        dist.append(np.array(d_i, np.float64)), cls_used.append(np.array(c_i, np.int32))
        stride_of.append(np.array(s_i, np.int32)), box_of.append(np.array(b_i, np.int32))
    return dist, cls_used, stride_of, box_of


def _pack_nested(prefix, nested, store):
    """clusters/thresholds/scores [cls][stride] -> flat npz entries."""
    for c, row in enumerate(nested):
        for s, v in enumerate(row):
            if isinstance(v, list):
                v = np.zeros(0, F32) if len(v) == 0 else np.asarray(v)
            store[f"{prefix}_{c}_{s}"] = np.asarray(v)


def _cat(lst, dtype):
    return np.concatenate([np.asarray(a, dtype).reshape(-1) for a in lst]) if len(lst) else np.zeros(0, dtype)


def golden_scoring(ref, name, wl, cluster_method, seed, lam_train, channels=None):
    """Fit (clusters, scores, thresholds) on a synthetic train batch and decide on a test batch, with the
    reference's L1 / L2 / Cosine methods."""
    ou = ref.ood_utils
    ch = channels or wl.channels
    hw = tuple(wl.img // s for s in synth.STRIDES)
    train_maps = synth.feature_maps(seed, wl.batch, ch, hw)
    train_det = synth.detections(seed + 1, wl.batch, wl.img, wl.nc, lam_train)
    test_maps = synth.feature_maps(seed + 2, wl.batch, ch, hw)
    test_det = synth.detections(seed + 3, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
    if wl.batch > 2:                       # ragged edge cases: an image with no boxes, one with a single box
        for k in test_det:
            test_det[k][1] = test_det[k][1][:0]
            test_det[k][2] = test_det[k][2][:1]
    acts = _collect_activations(ref, train_maps, train_det, wl.img, wl.nc)
    results = _images(ref, test_maps, test_det, wl.img)
    store = dict(seed=seed, img=wl.img, batch=wl.batch, nc=wl.nc, channels=np.array(ch), lam_train=lam_train,
                 cluster_method=cluster_method,
                 n_boxes=np.array([len(b) for b in test_det["boxes"]]),
                 boxes=_cat(test_det["boxes"], F32).reshape(-1, 4), cls=_cat(test_det["cls"], F32),
                 strides=_cat(test_det["strides"], F32), conf=_cat(test_det["conf"], F32),
                 train_n_boxes=np.array([len(b) for b in train_det["boxes"]]),
                 train_boxes=_cat(train_det["boxes"], F32).reshape(-1, 4), train_cls=_cat(train_det["cls"], F32),
                 train_strides=_cat(train_det["strides"], F32))
    kw = dict(ref_shim.DIST_KW, cluster_method=cluster_method)
    for tag, cls in (("l1", ou.L1DistanceOneClusterPerStride), ("l2", ou.L2DistanceOneClusterPerStride),
                     ("cos", ou.CosineDistanceOneClusterPerStride)):
        m = cls(**kw)
        m.clusters = m.generate_clusters(acts, LOG)
        scores = m.compute_scores_from_activations(acts, LOG)
        m.thresholds = m.generate_thresholds(scores, 0.95, LOG)
        dec = m.compute_ood_decision_on_results(results, LOG)
        dist, cls_used, stride_of, box_of = _ref_distances_q1(ref, m, results)
        _pack_nested(f"{tag}_clusters", m.clusters, store)
        _pack_nested(f"{tag}_thr", m.thresholds, store)
        _pack_nested(f"{tag}_fitscores", scores, store)
        _pack_nested(f"{tag}_mindist", m.min_dist, store)
        _pack_nested(f"{tag}_maxdist", m.max_dist, store)
        store[f"{tag}_decisions"] = _cat(dec, np.int8)
        store[f"{tag}_dist"] = _cat(dist, np.float64)
        store[f"{tag}_cls_used"] = _cat(cls_used, np.int32)
        store[f"{tag}_stride_of"] = _cat(stride_of, np.int32)
        store[f"{tag}_box_of"] = _cat(box_of, np.int32)
        indness = m.compute_INDness_scores_on_results(results, LOG) if tag == "cos" else None
        if indness is not None:
            store["cos_indness"] = _cat(indness, np.float64)
    # pooled vectors of the test boxes straight from the reference extractor (box order, per stride)
    tm = [torch.from_numpy(m) for m in test_maps]
    pooled = []
    for i in range(wl.batch):
        feats = ref.extract_roi_aligned_features_from_correct_stride(
            ftmaps=[t[i][None] for t in tm], boxes=[torch.from_numpy(test_det["boxes"][i])],
            strides=[torch.from_numpy(test_det["strides"][i])], img_shape=(wl.img, wl.img), device="cpu")[0]
        vec = [None] * len(test_det["boxes"][i])
        for s, (idx, fm) in enumerate(feats):
            for j, b in enumerate(idx):
                vec[int(b)] = fm[j].numpy().reshape(-1)
        pooled += vec
    for s in range(3):
        sel = [v for v, st in zip(pooled, _cat(test_det["strides"], F32)) if int(st) == s]
        store[f"pooled_s{s}"] = np.stack(sel, 0) if sel else np.zeros((0, ch[s]), F32)
    np.savez_compressed(os.path.join(OUT, name), **store)
    print(name, "boxes", int(store["n_boxes"].sum()), "InD frac l2", float(store["l2_decisions"].mean()))


def golden_logits(ref, seed=77, batch=8, nc=20):
    ou = ref.ood_utils
    rng = np.random.default_rng(seed)
    train_cls = rng.integers(0, nc, size=3000).astype(F32)
    train_cls[train_cls == 7] = 8                       # class 7 gets no samples -> threshold stays 0
    train_logits = synth.logits_for(rng, train_cls, nc)
    acts = [torch.from_numpy(train_logits[train_cls == c]) if (train_cls == c).any() else torch.tensor([])
            for c in range(nc)]
    det = synth.detections(seed + 1, batch, 640, nc, 40)
    det["boxes"][3], det["cls"][3], det["conf"][3], det["logits"][3] = (det[k][3][:0] for k in ("boxes", "cls", "conf", "logits"))
    boxes6 = [torch.from_numpy(np.concatenate([det["boxes"][i], det["conf"][i][:, None], det["cls"][i][:, None]], 1))
              for i in range(batch)]
    results = ref_shim.make_results(ref, None, boxes6, logits=[torch.from_numpy(z) for z in det["logits"]], n_batch=batch)
    keep = [c != 7 for c in det["cls"]]
    results_fitted = ref_shim.make_results(ref, None, [b[torch.from_numpy(k)] for b, k in zip(boxes6, keep)],
                                           logits=[torch.from_numpy(z[k]) for z, k in zip(det["logits"], keep)], n_batch=batch)
    store = dict(seed=seed, nc=nc, n_boxes=np.array([len(b) for b in det["boxes"]]), cls=_cat(det["cls"], F32),
                 logits=np.concatenate(det["logits"], 0), train_logits=train_logits, train_cls=train_cls,
                 fitted_mask=_cat(keep, bool))
    for tag, m in (("MSP", ou.MSP(**ref_shim.LOGIT_KW)), ("Energy", ou.Energy(temper=1, **ref_shim.LOGIT_KW)),
                   ("ODIN", ou.ODIN(temper=1000, **ref_shim.LOGIT_KW)), ("Sigmoid", ou.Sigmoid(**ref_shim.LOGIT_KW))):
        scores = m.compute_scores_from_activations(acts, LOG)
        m.thresholds = m.generate_thresholds(scores, 0.95, LOG)
        store[f"{tag}_thr"] = np.asarray(m.thresholds, np.float64)
        store[f"{tag}_min"] = np.asarray(m.min_score, np.float64)
        store[f"{tag}_max"] = np.asarray(m.max_score, np.float64)
        store[f"{tag}_fitscores"] = np.concatenate([np.asarray(s, F32) for s in scores])
        store[f"{tag}_decisions"] = _cat(m.compute_ood_decision_on_results(results, LOG), np.int8)
        # the reference divides by (max_score - thr) = 0 for a class without InD samples (ZeroDivisionError),
        # so INDness is frozen only for boxes whose class was fitted
        store[f"{tag}_indness"] = _cat(m.compute_INDness_scores_on_results(results_fitted, LOG), np.float64)
        sc = [m.compute_scores(torch.from_numpy(z), int(c))[0] for z, c in zip(store["logits"], store["cls"])]
        store[f"{tag}_scores"] = np.asarray(sc, F32)
    np.savez_compressed(os.path.join(OUT, "golden_logits.npz"), **store)
    print("golden_logits boxes", len(store["cls"]), "MSP InD frac", float(store["MSP_decisions"].mean()))


def golden_fusion(ref):
    ou = ref.ood_utils
    rng = np.random.default_rng(5)
    d1 = [list(rng.integers(0, 2, size=n)) for n in (7, 0, 12)]
    d2 = [list(rng.integers(0, 2, size=n)) for n in (7, 0, 12)]
    d3 = [list(rng.integers(0, 2, size=n)) for n in (7, 0, 12)]
    s1 = [list(rng.uniform(-1, 1, size=n)) for n in (7, 0, 12)]
    s2 = [[-1.0] * n for n in (7, 0, 12)]
    m1, m2 = ou.MSP(**ref_shim.LOGIT_KW), ou.MSP(**ref_shim.LOGIT_KW)
    common = dict(iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)
    store = dict(d1=_cat(d1, np.int8), d2=_cat(d2, np.int8), d3=_cat(d3, np.int8), s1=_cat(s1, np.float64),
                 s2=_cat(s2, np.float64), n=np.array([7, 0, 12]))
    for strat in ("and", "or", "score"):
        f = ou.FusionMethod(m1, m2, strat, fusion_method_name="fusion-MSP-MSP", cluster_method="one", **common)
        a, b = (s1, s2) if strat == "score" else (d1, d2)
        store[f"fuse_{strat}"] = _cat(f.fuse_ood_decisions(a, b), np.int8)
    t = ou.TripleFusionMethod(m1, m2, ou.MSP(**ref_shim.LOGIT_KW), cluster_method="one", **common)
    store["fuse_majority"] = _cat(t.fuse_ood_decisions(d1, d2, d3), np.int8)
    np.savez_compressed(os.path.join(OUT, "golden_fusion.npz"), **store)
    print("golden_fusion ok")


def golden_roi_edges(ref):
    """The extractor on hand-picked boxes: thin border boxes (Q5), full image, degenerate, ragged batch."""
    rng = np.random.default_rng(11)
    img = 160
    ch, hw = (8, 12, 16), (20, 10, 5)
    maps = [synth.feature_map(rng, (3, c, h, h)) for c, h in zip(ch, hw)]
    boxes = [np.array([[150, 150, 160, 160], [0, 0, 160, 160], [5, 5, 5, 5], [159, 0, 160, 3], [10.3, 20.7, 90.1, 33.3],
                       [0, 0, 7.9, 7.9], [152.5, 3, 160, 158], [40, 40, 41, 120]], F32),
             np.zeros((0, 4), F32),
             np.array([[30, 30, 130, 130], [156, 156, 160, 160], [0, 100, 160, 104]], F32)]
    strides = [np.array([2, 0, 1, 2, 0, 0, 1, 2], F32), np.zeros(0, F32), np.array([2, 2, 0], F32)]
    store = dict(img=img, n_boxes=np.array([len(b) for b in boxes]), boxes=np.concatenate(boxes), strides=np.concatenate(strides))
    for s, m in enumerate(maps):
        store[f"map{s}"] = m
    for all_strides in (False, True):
        out = ref.extract_roi_aligned_features_from_correct_stride(
            ftmaps=[torch.from_numpy(m) for m in maps], boxes=[torch.from_numpy(b) for b in boxes],
            strides=[torch.from_numpy(s) for s in strides], img_shape=(img, img), device="cpu",
            extract_all_strides=all_strides)
        for i in range(3):
            for s in range(3):
                idx, fm = out[i][s]
                store[f"all{int(all_strides)}_idx_{i}_{s}"] = np.asarray(idx, np.int16)
                fm = np.asarray(fm, F32)
                store[f"all{int(all_strides)}_feat_{i}_{s}"] = fm.reshape(len(idx), -1) if len(idx) else np.zeros((0, ch[s]), F32)
    np.savez_compressed(os.path.join(OUT, "golden_roi_edges.npz"), **store)
    print("golden_roi_edges ok")


def golden_quirks(ref):
    """Q1 probe from SURVEY.md §8 (3 boxes, (cls,stride) = (1,2),(0,0),(2,2)) and Q4 (falsy thresholds,
    missing cluster), through the reference's own decision function."""
    ou = ref.ood_utils
    rng = np.random.default_rng(3)
    img, ch, hw = 160, (8, 12, 16), (20, 10, 5)
    maps = [synth.feature_map(rng, (1, c, h, h)) for c, h in zip(ch, hw)]
    box = np.array([[10, 10, 120, 130], [30, 40, 50, 60], [5, 60, 150, 150]], F32)
    cls = np.array([1, 0, 2], F32)
    strides = np.array([2, 0, 2], F32)
    det = dict(boxes=[box], cls=[cls], strides=[strides], conf=[np.full(3, 0.5, F32)])
    results = _images(ref, maps, det, img)
    clusters = [[rng.uniform(0, 0.3, size=(2, c)).astype(F32) for c in ch] for _ in range(3)]
    m = ou.L2DistanceOneClusterPerStride(**ref_shim.DIST_KW)
    m.clusters = clusters
    store = dict(img=img, boxes=box, cls=cls, strides=strides)
    for s, mp in enumerate(maps):
        store[f"map{s}"] = mp
    _pack_nested("clusters", clusters, store)
    cases = {
        "q1": [[1e9] * 3, [1e-9] * 3, [1e-9] * 3],                 # class 0 always InD, classes 1,2 always OoD
        "q4_zero": [[0.0] * 3, [1e9] * 3, [1e9] * 3],              # thr 0.0 is falsy -> OoD
        "q4_empty": [[[]] * 3, [1e9] * 3, [1e9] * 3],              # thr [] -> OoD
    }
    for k, thr in cases.items():
        m.thresholds = thr
        store[f"{k}_decisions"] = _cat(m.compute_ood_decision_on_results(results, LOG), np.int8)
    m.thresholds = [[1e9] * 3] * 3
    m.clusters = [clusters[0], [np.empty(0), clusters[1][1], clusters[1][2]], clusters[2]]   # (class 1, stride 0) missing
    store["missing_cluster_decisions"] = _cat(m.compute_ood_decision_on_results(results, LOG), np.int8)
    m.thresholds = [[999.0] * 3] * 3                                                          # 1000 !< 999 -> OoD
    store["missing_cluster_thr999_decisions"] = _cat(m.compute_ood_decision_on_results(results, LOG), np.int8)
    np.savez_compressed(os.path.join(OUT, "golden_quirks.npz"), **store)
    print("golden_quirks", {k: store[k].tolist() for k in store if k.endswith("decisions")})


def golden_kmeans(ref):
    """Labels from the reference's own call site (cluster_utils.py:62-73) on clustered vectors, plus the
    centroids/thresholds the reference's fit derives from them."""
    ou = ref.ood_utils
    store = {}
    for tag, (n, dim, k, sep) in {"a": (1500, 24, 10, 6.0), "b": (900, 16, 5, 8.0), "c": (7, 8, 10, 6.0)}.items():
        x, _ = synth.blob_vectors(100 + len(tag) + n, n, dim, min(k, 6), sep)
        lab = ref.cluster_utils.find_optimal_number_of_clusters_one_class_one_stride_and_return_labels(
            x, f"KMeans_{k}", "l2", "silhouette", "", LOG)
        store[f"{tag}_x"], store[f"{tag}_labels"], store[f"{tag}_k"] = x, np.asarray(lab, np.int32), k
    # full fit through the reference's DistanceMethod with KMeans_5 on [cls][stride] activations
    acts = [[np.empty(0) for _ in range(3)] for _ in range(3)]
    acts[0][0] = synth.blob_vectors(7, 600, 16, 5, 7.0, unit_norm=False)[0][:, :, None, None]
    acts[1][2] = synth.blob_vectors(8, 400, 24, 4, 7.0, unit_norm=False)[0][:, :, None, None]
    acts[2][1] = synth.blob_vectors(9, 3, 12, 2, 7.0, unit_norm=False)[0][:, :, None, None]      # <= MIN_SAMPLES -> no cluster
    m = ou.L2DistanceOneClusterPerStride(**dict(ref_shim.DIST_KW, cluster_method="KMeans_5"))
    m.clusters = m.generate_clusters(acts, LOG)
    scores = m.compute_scores_from_activations(acts, LOG)
    thr = m.generate_thresholds(scores, 0.95, LOG)
    _pack_nested("fit_acts", acts, store)
    _pack_nested("fit_clusters", m.clusters, store)
    _pack_nested("fit_scores", scores, store)
    _pack_nested("fit_thr", thr, store)
    np.savez_compressed(os.path.join(OUT, "golden_kmeans.npz"), **store)
    print("golden_kmeans ok", [len(set(store[f"{t}_labels"].tolist())) for t in "abc"])


def golden_thresholds(ref):
    """np.percentile(method='lower') through the reference's generate_thresholds, incl. the
    (1-0.9)*100 = 9.999999999999998 index case."""
    ou = ref.ood_utils
    rng = np.random.default_rng(21)
    store = {}
    dm = ou.L2DistanceOneClusterPerStride(**ref_shim.DIST_KW)
    lm = ou.MSP(**ref_shim.LOGIT_KW)
    for n in (5, 6, 11, 21, 101, 1001, 4097):
        v = rng.uniform(0, 1, size=n).astype(F32)
        v[: n // 3] = v[0]                                   # ties
        store[f"v_{n}"] = v
        for tpr in (0.9, 0.95, 0.99, 0.8):
            t = dm.generate_thresholds([[v, v.astype(np.float64), np.empty(0)]], tpr, LOG)[0]
            store[f"dist_{n}_{tpr}"] = np.array([x if x != [] else np.nan for x in t], np.float64)
            store[f"logit_{n}_{tpr}"] = np.array(lm.generate_thresholds([v], tpr, LOG), np.float64)
    np.savez_compressed(os.path.join(OUT, "golden_thresholds.npz"), **store)
    print("golden_thresholds ok")


def golden_matching(ref):
    """`match_predicted_boxes_to_targets` (ood_utils.py:233-292) and `create_targets_dict` (:201-231) on jittered
    copies of ground-truth boxes: more predictions than targets, fewer, equal, none."""
    ou = ref.ood_utils
    rng = np.random.default_rng(31)
    store = {}
    shapes = [(9, 4), (3, 7), (5, 5), (0, 3), (4, 0), (12, 6)]
    store["shapes"] = np.array(shapes)
    for k, (P, G) in enumerate(shapes):
        c = rng.uniform(40, 600, size=(G, 2))
        wh = rng.uniform(20, 120, size=(G, 2))
        gt = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(F32)
        gcls = rng.integers(0, 3, size=G).astype(F32)
        if G:
            src = rng.integers(0, G, size=P)
            pred = gt[src] + rng.normal(0, 6, size=(P, 4)).astype(F32)
            pcls = np.where(rng.uniform(size=P) < 0.8, gcls[src], (gcls[src] + 1) % 3).astype(F32)
        else:
            pred = rng.uniform(0, 600, size=(P, 4)).astype(F32)
            pred[:, 2:] += pred[:, :2]
            pcls = rng.integers(0, 3, size=P).astype(F32)
        b6 = torch.from_numpy(np.concatenate([pred, np.full((P, 1), 0.5, F32), pcls[:, None]], 1).astype(F32))
        res = ref_shim.make_results(ref, None, [b6], logits=[torch.zeros(P, 3)], n_batch=1)
        targets = dict(bboxes=[torch.from_numpy(gt)], cls=[torch.from_numpy(gcls)])
        for thr in (0.5, 0.3):
            ou.OODMethod.match_predicted_boxes_to_targets(res, targets, thr)
            store[f"valid_{k}_{thr}"] = np.asarray(res[0].valid_preds, np.int64)
        store[f"pred_{k}"], store[f"pcls_{k}"], store[f"gt_{k}"], store[f"gcls_{k}"] = pred, pcls, gt, gcls
    data = dict(im_file=["a", "b", "c"], batch_idx=torch.tensor([0., 2., 0., 2., 2.]),
                bboxes=torch.from_numpy(rng.uniform(0.1, 0.6, size=(5, 4)).astype(F32)),
                cls=torch.from_numpy(rng.integers(0, 5, size=(5, 1)).astype(F32)),
                resized_shape=[(640, 480), (320, 320), (512, 640)])
    t = ou.OODMethod.create_targets_dict(data)
    store["td_batch_idx"], store["td_bboxes"], store["td_cls"] = data["batch_idx"].numpy(), data["bboxes"].numpy(), data["cls"].numpy()
    store["td_shapes"] = np.array(data["resized_shape"])
    for i in range(3):
        store[f"td_out_bboxes_{i}"] = np.asarray(t["bboxes"][i], np.float64)
        store[f"td_out_cls_{i}"] = np.asarray(t["cls"][i], F32)
    np.savez_compressed(os.path.join(OUT, "golden_matching.npz"), **store)
    print("golden_matching", {k: store[k].tolist() for k in store if k.startswith("valid_") and k.endswith("0.5")})


def golden_ksearch(ref):
    """Silhouette / Calinski-Harabasz search of the number of k-means clusters through the reference's own functions
    (cluster_utils.py:75-80 params, :203-302 scores per k, :160-186 best k + refit): per-k scores and final labels."""
    from sklearn.cluster import KMeans
    cu = ref.cluster_utils
    store = {}
    cases = {"a": (600, 24, 5, 7.0, "l2", "silhouette"), "b": (420, 16, 3, 8.0, "cosine", "silhouette"),
             "c": (500, 12, 4, 7.0, "l1", "silhouette"), "d": (450, 16, 6, 8.0, "l2", "calinski_harabasz"),
             "e": (9, 8, 2, 8.0, "l2", "silhouette")}        # e: fewer samples than most k -> default scores
    for tag, (n, dim, k_true, sep, metric, perf) in cases.items():
        x, _ = synth.blob_vectors(300 + n, n, dim, k_true, sep)
        params = {"n_clusters": list(range(2, 15)), "random_state": [10]}
        scores, configs = cu.compute_score_for_all_possible_configurations(x, KMeans, params, "n_clusters", perf, metric, LOG)
        lab = cu.find_optimal_number_of_clusters_one_class_one_stride_and_return_labels(x, "KMeans", metric, perf, "", LOG)
        store[f"{tag}_x"] = x
        store[f"{tag}_scores"] = np.asarray(scores, np.float64)
        store[f"{tag}_ks"] = np.asarray([c["n_clusters"] for c in configs], np.int32)
        store[f"{tag}_labels"] = np.asarray(lab, np.int32)
        store[f"{tag}_metric"], store[f"{tag}_perf"] = np.array(metric), np.array(perf)
        print("ksearch", tag, "best k", int(store[f"{tag}_ks"][int(np.argmax(scores))]), np.round(scores, 4).tolist())
    np.savez_compressed(os.path.join(OUT, "golden_ksearch.npz"), **store)


def golden_eul_rank(ref):
    """Ranking of unknown proposals: the statements of ood_utils.py:1031-1084 executed with the reference's own objects
    (torchvision roi_align on the padded map with spatial_scale 1, `activations_transformation`, `compute_distance` of the
    reference's method classes, the numpy / scipy fold).  The block sits inside a 500-line method that needs the whole
    EUL pipeline (saliency maps, skimage region proposals), so it cannot be called in isolation."""
    import torchvision.ops as t_ops
    from scipy.stats import entropy, gmean
    ou = ref.ood_utils
    rng = np.random.default_rng(77)
    store = {}
    C, H, W, nc, P = 24, 22, 26, 6, 9
    fm = np.abs(rng.standard_normal((C, H, W))).astype(F32)
    x1, y1 = rng.uniform(-1, W - 3, P), rng.uniform(-1, H - 3, P)
    props = np.stack([x1, y1, x1 + rng.uniform(0.3, 9, P), y1 + rng.uniform(0.3, 9, P)], 1).astype(F32)
    clusters = [[np.empty(0)] * 3 for _ in range(nc)]
    for c in range(nc):
        if c != 2:                                           # class 2 has no clusters on this stride
            k = int(rng.integers(1, 5))
            v = np.abs(rng.standard_normal((k, C))).astype(F32)
            clusters[c][1] = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(F32)
    store["fm"], store["props"], store["stride"], store["nc"] = fm, props, np.int32(1), np.int32(nc)
    _pack_nested("clusters", clusters, store)
    feats = t_ops.roi_align(input=torch.from_numpy(fm).unsqueeze(0), boxes=[torch.from_numpy(props).float()],
                            output_size=(1, 1), spatial_scale=1.0, aligned=False)
    for tag, klass in (("l1", ou.L1DistanceOneClusterPerStride), ("l2", ou.L2DistanceOneClusterPerStride),
                       ("cos", ou.CosineDistanceOneClusterPerStride)):
        m = klass(**ref_shim.DIST_KW)
        m.clusters = clusters
        dpp = []
        for idx_cls, cluster in enumerate(m.clusters):
            if len(cluster[1]) > 0:
                dpp.append(m.compute_distance(cluster[1], m.activations_transformation(feats, cls_idx=idx_cls, stride_idx=1)))
        dpp = np.array(dpp)
        store[f"{tag}_matrix"] = dpp.astype(np.float64)
        store[f"{tag}_mean"], store[f"{tag}_max"], store[f"{tag}_sum"] = dpp.mean(axis=0), dpp.max(axis=0), dpp.sum(axis=0)
        store[f"{tag}_min"] = dpp.min(axis=0) * 100
        store[f"{tag}_minthr"], store[f"{tag}_closest"] = dpp.min(axis=0), np.argsort(dpp, axis=0)[0]
        store[f"{tag}_geometric_mean"] = gmean(dpp, axis=0)
        store[f"{tag}_entropy"] = entropy(dpp / dpp.sum(axis=0), axis=0)
    np.savez_compressed(os.path.join(OUT, "golden_eul_rank.npz"), **store)
    print("golden_eul_rank", store["l2_matrix"].shape, np.round(store["l2_entropy"], 4).tolist())


def golden_c4(ref):
    """BASELINE config 4 shapes (YOLOv8l maps 256/512/512 @ 80/40/20, K = 10 per (class, stride)) through the reference's own
    classes: the vanilla Cosine method with `KMeans_10`, an SDR method (IvisMethodCosine: normalise -> 32-d embedding ->
    un-normalised scoring, ood_utils.py:2501-2571; the trained ivis model is replaced by the fixed projection
    tests/helpers.py::Projection, the rest is the reference's code), MSP, `fusion-MSP-Cosine_cl_stride` with and / or /
    score (fuse_ood_decisions over the two methods' own outputs, :2942-2975), and the EUL ranking of 3 unknown proposals
    per image against all classes (:1031-1084 statements, as in golden_eul_rank)."""
    import torchvision.ops as t_ops
    from scipy.stats import entropy
    from tests.helpers import Projection
    ou = ref.ood_utils
    nc, img, k = 8, 640, 10
    wl = synth.Workload("C4-shaped YOLOv8l K=10", "l", img, 4, nc, k, 50)
    ch, hw, seed = wl.channels, wl.map_hw, 4400
    train_maps = synth.feature_maps(seed, 3, ch, hw)
    train_det = synth.detections(seed + 1, 3, img, nc, 300, fixed=True)
    test_maps = synth.feature_maps(seed + 2, wl.batch, ch, hw)
    test_det = synth.detections(seed + 3, wl.batch, img, nc, wl.lam)
    acts = _collect_activations(ref, train_maps, train_det, img, nc)
    results = _images(ref, test_maps, test_det, img)
    store = dict(seed=seed, img=img, batch=wl.batch, nc=nc, k=k, channels=np.array(ch),
                 n_boxes=np.array([len(b) for b in test_det["boxes"]]),
                 boxes=_cat(test_det["boxes"], F32).reshape(-1, 4), cls=_cat(test_det["cls"], F32),
                 strides=_cat(test_det["strides"], F32), conf=_cat(test_det["conf"], F32),
                 logits=np.concatenate(test_det["logits"], 0),
                 train_cls=_cat(train_det["cls"], F32), train_logits=np.concatenate(train_det["logits"], 0))
    kw = dict(ref_shim.DIST_KW, cluster_method=f"KMeans_{k}")
    cos = ou.CosineDistanceOneClusterPerStride(**kw)
    ivis = ou.IvisMethodCosine(**kw)
    ivis.ivis = [Projection(900 + s, ch[s], 32) for s in range(3)]
    ivis.is_dimensionality_reduction_trained = True
    for tag, m in (("cos", cos), ("ivis", ivis)):
        m.clusters = m.generate_clusters(acts, LOG)
        scores = m.compute_scores_from_activations(acts, LOG)
        m.thresholds = m.generate_thresholds(scores, 0.95, LOG)
        dec = m.compute_ood_decision_on_results(results, LOG)
        dist, cls_used, stride_of, box_of = _ref_distances_q1(ref, m, results)
        _pack_nested(f"{tag}_clusters", m.clusters, store)
        _pack_nested(f"{tag}_thr", m.thresholds, store)
        store[f"{tag}_decisions"] = _cat(dec, np.int8)
        store[f"{tag}_dist"] = _cat(dist, np.float64)
        store[f"{tag}_cls_used"] = _cat(cls_used, np.int32)
        store[f"{tag}_stride_of"] = _cat(stride_of, np.int32)
    # MSP on the same detections
    boxes6 = [torch.from_numpy(np.concatenate([test_det["boxes"][i], test_det["conf"][i][:, None], test_det["cls"][i][:, None]], 1))
              for i in range(wl.batch)]
    results_l = ref_shim.make_results(ref, None, boxes6, logits=[torch.from_numpy(z) for z in test_det["logits"]],
                                      batch_hw=(img, img), n_batch=wl.batch)
    msp = ou.MSP(**ref_shim.LOGIT_KW)
    tl, tc = store["train_logits"], store["train_cls"]
    lacts = [torch.from_numpy(tl[tc == c]) if (tc == c).any() else torch.tensor([]) for c in range(nc)]
    lscores = msp.compute_scores_from_activations(lacts, LOG)
    msp.thresholds = msp.generate_thresholds(lscores, 0.95, LOG)
    store["msp_thr"] = np.asarray(msp.thresholds, np.float64)
    store["msp_min"], store["msp_max"] = np.asarray(msp.min_score, np.float64), np.asarray(msp.max_score, np.float64)
    d_msp = msp.compute_ood_decision_on_results(results_l, LOG)
    d_cos = cos.compute_ood_decision_on_results(results, LOG)
    store["msp_decisions"] = _cat(d_msp, np.int8)
    common = dict(iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)
    for strat in ("and", "or", "score"):
        thr_pair = (msp.thresholds, cos.thresholds)
        f = ou.FusionMethod(msp, cos, strat, fusion_method_name="fusion-MSP-Cosine_cl_stride", cluster_method=f"KMeans_{k}", **common)
        f.thresholds = thr_pair                              # the constructor resets both methods' thresholds (OODMethod.__init__, :89)
        if strat == "score":                                 # :2958-2975: INDness of both methods, summed, > 0 = InD
            a = msp.compute_INDness_scores_on_results(results_l, LOG)
            b = cos.compute_INDness_scores_on_results(results, LOG)
            store["msp_indness"], store["cos_indness"] = _cat(a, np.float64), _cat(b, np.float64)
        else:
            a, b = d_msp, d_cos
        store[f"fusion_{strat}"] = _cat(f.fuse_ood_decisions(a, b), np.int8)
    # EUL: 3 unknown proposals per image on the stride-1 map, ranked against every class (cosine clusters above)
    rng = np.random.default_rng(seed + 9)
    W = hw[1]
    x1, y1 = rng.uniform(-1, W - 4, (wl.batch, 3)), rng.uniform(-1, W - 4, (wl.batch, 3))
    props = np.stack([x1, y1, x1 + rng.uniform(0.5, 12, (wl.batch, 3)), y1 + rng.uniform(0.5, 12, (wl.batch, 3))], -1).astype(F32)
    store["eul_props"] = props
    mats = []
    for i in range(wl.batch):
        feats = t_ops.roi_align(input=torch.from_numpy(test_maps[1][i]).unsqueeze(0), boxes=[torch.from_numpy(props[i]).float()],
                                output_size=(1, 1), spatial_scale=1.0, aligned=False)
        dpp = np.array([cos.compute_distance(cl[1], cos.activations_transformation(feats, cls_idx=c, stride_idx=1))
                        for c, cl in enumerate(cos.clusters) if len(cl[1]) > 0])
        mats.append(dpp.astype(np.float64))
    store["eul_matrix"] = np.stack(mats)                      # [image, classes with clusters, proposal]
    store["eul_entropy"] = np.stack([entropy(m / m.sum(axis=0), axis=0) for m in mats])
    store["eul_min"] = np.stack([m.min(axis=0) * 100 for m in mats])
    np.savez_compressed(os.path.join(OUT, "golden_c4.npz"), **store)
    print("golden_c4 boxes", int(store["n_boxes"].sum()), "InD frac cos / ivis / msp",
          float(store["cos_decisions"].mean()), float(store["ivis_decisions"].mean()), float(store["msp_decisions"].mean()),
          "fusion and/or/score", [float(store[f"fusion_{s}"].mean()) for s in ("and", "or", "score")])


def golden_bigfit(ref):
    """Segments with >= 4096 rows through the reference's fit (generate_clusters with KMeans_5 -> member means,
    compute_scores_from_activations, generate_thresholds) for L1 / L2 / Cosine: pins the fit-time scores at the sizes where
    K2 used to switch to the tensor-core kernel.  The activations are regenerated from the seeds (synth.blob_vectors)."""
    ou = ref.ood_utils
    spec = {(0, 0): (71, 6000, 128), (1, 1): (72, 4500, 256), (1, 0): (73, 4096, 128), (2, 2): (74, 40, 160)}
    acts = [[np.empty(0) for _ in range(3)] for _ in range(3)]
    store = dict(spec=np.array([[c, s, sd, n, d] for (c, s), (sd, n, d) in spec.items()]))
    for (c, s), (sd, n, d) in spec.items():
        acts[c][s] = np.abs(synth.blob_vectors(sd, n, d, 5, 7.0, unit_norm=False)[0])[:, :, None, None]   # 5 separated blobs for KMeans_5
    for tag, cls in (("l1", ou.L1DistanceOneClusterPerStride), ("l2", ou.L2DistanceOneClusterPerStride),
                     ("cos", ou.CosineDistanceOneClusterPerStride)):
        m = cls(**dict(ref_shim.DIST_KW, cluster_method="KMeans_5"))
        m.clusters = m.generate_clusters(acts, LOG)
        scores = m.compute_scores_from_activations(acts, LOG)
        thr = m.generate_thresholds(scores, 0.95, LOG)
        _pack_nested(f"{tag}_clusters", m.clusters, store)
        _pack_nested(f"{tag}_scores", [[np.asarray(v, F32) for v in row] for row in scores], store)
        _pack_nested(f"{tag}_thr", thr, store)
        _pack_nested(f"{tag}_mindist", m.min_dist, store)
        _pack_nested(f"{tag}_maxdist", m.max_dist, store)
    np.savez_compressed(os.path.join(OUT, "golden_bigfit.npz"), **store)
    print("golden_bigfit ok", {k: float(np.asarray(store[k])) for k in ("l2_thr_0_0", "cos_thr_1_1")})


def golden_nms(ref):
    """`non_max_suppression_old` (ultralytics/utils/ops.py:348-530) with the OoD payload (extra_item = raw logits, strides),
    default path, at two confidence / IoU settings."""
    from ultralytics.utils import ops as uops
    from tests.helpers import nms_inputs
    pred, logits, strides = nms_inputs()
    store = dict(seed=91)
    for tag, (conf, iou, max_det) in {"a": (0.25, 0.45, 300), "b": (0.15, 0.7, 50), "agn": (0.25, 0.45, 300)}.items():
        out, extra, st = uops.non_max_suppression_old(torch.from_numpy(pred), conf, iou, max_det=max_det, max_time_img=100.0,
                                                     agnostic=tag == "agn",
                                                     extra_item=torch.from_numpy(logits), strides=torch.from_numpy(strides))
        # (max_time_img: the reference abandons the remaining images after 0.5 s + 0.05 s per image on a slow CPU)
        store[f"{tag}_n"] = np.array([len(o) for o in out])
        store[f"{tag}_det"] = np.concatenate([o.numpy() for o in out]).astype(F32)
        store[f"{tag}_extra"] = np.concatenate([e.numpy().reshape(len(o), -1) for e, o in zip(extra, out)]).astype(F32)
        store[f"{tag}_strides"] = np.concatenate([s.numpy().reshape(-1) for s in st]).astype(F32)
        store[f"{tag}_cfg"] = np.array([conf, iou, max_det], np.float64)
    np.savez_compressed(os.path.join(OUT, "golden_nms.npz"), **store)
    print("golden_nms kept per image", {t: store[f"{t}_n"].tolist() for t in ("a", "b", "agn")})


def golden_postprocess(ref):
    """The OoD branch of `DetectionPredictor.postprocess` (ultralytics/models/yolo/detect/predict.py:117-363) driven with a
    stand-in predictor: every extraction mode, raw-logit and sigmoid heads, a confidence that empties an image."""
    from ultralytics.models.yolo.detect.predict import DetectionPredictor
    from tests.helpers import fake_predictor, postprocess_inputs
    pred, logits, maps = postprocess_inputs()
    p = torch.from_numpy(pred)
    raw = torch.cat([p[:, :4], torch.from_numpy(logits)], 1)
    tmaps = [torch.from_numpy(m) for m in maps]
    img = torch.zeros((4, 3, 320, 320))
    store = dict(seed=93)
    cases = {"fs": ("ftmaps_and_strides", False, 0.25), "fs_hi": ("ftmaps_and_strides", False, 0.97),
             "pos": ("ftmaps_and_strides_exact_pos", False, 0.25), "all": ("all_ftmaps", False, 0.25),
             "roi": ("roi_aligned_ftmaps", False, 0.25), "lg": ("logits", False, 0.25), "lg_raw": ("logits", True, 0.25),
             "lg_hi": ("logits", True, 0.97)}
    for tag, (mode, before, conf) in cases.items():
        head = raw.clone() if before else p.clone()
        res = DetectionPredictor.postprocess(fake_predictor(mode, before, conf), ((head,), None if mode == "logits" else tmaps), img, img)
        n = np.array([len(r.boxes) for r in res])
        store[f"{tag}_n"] = n
        store[f"{tag}_boxes"] = np.concatenate([r.boxes.data.numpy().reshape(-1, 6) for r in res]).astype(F32)
        store[f"{tag}_shape"] = np.array(res[0].orig_img.shape)
        if mode == "logits":
            store[f"{tag}_extra"] = np.concatenate([r.extra_item.numpy().reshape(len(r.boxes), 20) for r in res]).astype(F32)
        elif mode in ("ftmaps_and_strides", "ftmaps_and_strides_exact_pos"):
            store[f"{tag}_extra"] = np.concatenate([r.extra_item[1].numpy().reshape(-1) for r in res]).astype(np.float64)
            assert all(torch.equal(r.extra_item[0][s], tmaps[s][i]) for i, r in enumerate(res) for s in range(3))
        elif mode == "roi_aligned_ftmaps":
            for s in range(3):
                store[f"{tag}_idx{s}"] = np.concatenate([r.extra_item[s][0].numpy().reshape(-1).astype(np.int64) for r in res])
                store[f"{tag}_cnt{s}"] = np.array([len(r.extra_item[s][0]) for r in res])
                store[f"{tag}_feat{s}"] = np.concatenate([r.extra_item[s][1].numpy().reshape(len(r.extra_item[s][0]), -1) for r in res]).astype(F32)
        else:
            assert all(torch.equal(r.extra_item[s], tmaps[s][i]) for i, r in enumerate(res) for s in range(3))
    np.savez_compressed(os.path.join(OUT, "golden_postprocess.npz"), **store)
    print("golden_postprocess kept per image", {t: store[f"{t}_n"].tolist() for t in cases})


def main():
    global OUT
    if "--out" in sys.argv:                                  # regenerate into another directory (tests/test_live_reference.py)
        OUT = sys.argv[sys.argv.index("--out") + 1]
        os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    only = {"--only-c4": golden_c4, "--only-bigfit": golden_bigfit, "--only-eul": golden_eul_rank, "--only-matching": golden_matching,
            "--only-ksearch": golden_ksearch, "--only-quirks": golden_quirks, "--only-fusion": golden_fusion,
            "--only-thresholds": golden_thresholds, "--only-nms": golden_nms, "--only-postprocess": golden_postprocess}
    picked = [fn for flag, fn in only.items() if flag in sys.argv]
    if picked:
        for fn in picked:
            fn(ref)
        return
    golden_c4(ref)
    golden_bigfit(ref)
    golden_nms(ref)
    golden_postprocess(ref)
    golden_ksearch(ref)
    golden_eul_rank(ref)
    golden_matching(ref)
    golden_roi_edges(ref)
    golden_quirks(ref)
    golden_fusion(ref)
    golden_logits(ref)
    golden_thresholds(ref)
    golden_kmeans(ref)
    golden_scoring(ref, "golden_c1_one.npz", synth.CONFIGS["C1"], "one", seed=1235, lam_train=60)
    small = synth.Workload("small KMeans_5", "n", 320, 6, 4, 5, 40, channels=(16, 24, 32))
    golden_scoring(ref, "golden_small_kmeans5.npz", small, "KMeans_5", seed=4321, lam_train=120)


if __name__ == "__main__":
    main()
