"""Drop-in check at the real boundary: the reference's UNMODIFIED evaluation driver runs on this repo's
classes.  `ood_evaluation.py` does `from ood_utils import ...`; with `sys.modules['ood_utils']` bound to
ood_in_object_detection_b200.ood_utils its own `select_ood_detection_method` (ood_evaluation.py:179-288) builds the methods and
its own `execute_pipeline_for_in_distribution_configuration` (:398-594) drives the fit -- activation collection over a
(fake) detector and loader, clusters, scores, thresholds, files on disk -- on the GPU.  The fitted state and the decisions
of a test batch are then compared with the CPU oracle.

The reference's modules come from oracle/ref_shim.py: /root/reference in the build container, the byte-compiled build
oracle/_ref on the GPU box (skipped when neither is there).
"""
from __future__ import annotations

import json
import logging
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from ood_in_object_detection_b200 import synth

pytestmark = pytest.mark.gpu
NC, IMG, BATCH = 4, 320, 4
CHANNELS = (16, 24, 32)


class _Logger(logging.Logger):
    def flush(self):                                    # the reference's own logger class has one (log.py)
        pass


def _batches(seed, n_batches, lam):
    """Per batch: maps (3 x [B, C, H, W]) and detections; the loader hands out uint8 images whose first pixel is the
    batch number so that the fake detector knows which detections to return."""
    hw = tuple(IMG // s for s in synth.STRIDES)
    out = []
    for i in range(n_batches):
        out.append(dict(maps=synth.feature_maps(seed + 10 * i, BATCH, CHANNELS, hw),
                        det=synth.detections(seed + 10 * i + 1, BATCH, IMG, NC, lam)))
    return out


class _Loader:
    """Yields ultralytics-style batch dicts (create_targets_dict reads im_file / batch_idx / bboxes / cls / resized_shape):
    the ground truth of every image is its own set of predicted boxes, so every prediction is a valid one."""

    def __init__(self, batches):
        self.batches = batches
        self.batch_size = BATCH

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        for i, b in enumerate(self.batches):
            det = b["det"]
            img = torch.zeros((BATCH, 3, IMG, IMG), dtype=torch.uint8)
            img[:, 0, 0, 0] = i
            bidx = np.concatenate([np.full(len(det["boxes"][k]), k, np.float32) for k in range(BATCH)])
            xyxy = np.concatenate(det["boxes"]) / IMG
            cxcywh = np.stack([(xyxy[:, 0] + xyxy[:, 2]) / 2, (xyxy[:, 1] + xyxy[:, 3]) / 2,
                               xyxy[:, 2] - xyxy[:, 0], xyxy[:, 3] - xyxy[:, 1]], 1).astype(np.float32)
            yield dict(img=img, im_file=[f"b{i}_{k}.jpg" for k in range(BATCH)], batch_idx=torch.from_numpy(bidx),
                       bboxes=torch.from_numpy(cxcywh), cls=torch.from_numpy(np.concatenate(det["cls"])[:, None]),
                       resized_shape=[(IMG, IMG)] * BATCH)


class _Detector:
    """Stands in for the patched ultralytics YOLO: `predict` returns this repo's Results with the extra item the configured
    method asks for (feature maps + strides, or raw logits)."""

    def __init__(self, batch_sets):
        self.sets = batch_sets                          # {first-pixel id: batches}
        self.names = {i: str(i) for i in range(NC)}
        self.ckpt = {"train_args": {"name": "fake_detector"}}
        self.mode = "ftmaps"
        self.which = None

    def predict(self, imgs, save=False, verbose=False, conf=0.0, device=None):
        from ood_in_object_detection_b200.results import Results, batch_shape
        i = int(round(float(imgs[0, 0, 0, 0]) * 255))
        b = self.which[i]
        det, maps = b["det"], [torch.from_numpy(m).cuda() for m in b["maps"]]
        out = []
        for k in range(BATCH):
            b6 = np.concatenate([det["boxes"][k], det["conf"][k][:, None], det["cls"][k][:, None]], 1).astype(np.float32)
            extra = ([m[k] for m in maps], torch.from_numpy(det["strides"][k]).cuda()) if self.mode == "ftmaps" \
                else torch.from_numpy(det["logits"][k])
            out.append(Results(orig_img=batch_shape(BATCH, IMG, IMG), boxes=torch.from_numpy(b6).cuda(), extra_item=extra))
        return out


@pytest.fixture(scope="module")
def reference_driver():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("neither /root/reference nor the compiled build oracle/_ref is present")
    ref_shim.load()                                     # stubs for matplotlib / tap / ..., sys.path, custom_hyperparams patch
    import ood_in_object_detection_b200.ood_utils as ours
    saved = sys.modules.get("ood_utils")
    sys.modules["ood_utils"] = ours                     # what a user of the reference does: point the import at this package
    sys.modules.pop("ood_evaluation", None)
    try:
        import ood_evaluation
        assert ood_evaluation.L2DistanceOneClusterPerStride is ours.L2DistanceOneClusterPerStride
        yield ood_evaluation, ours
    finally:
        sys.modules.pop("ood_evaluation", None)
        if saved is not None:
            sys.modules["ood_utils"] = saved


def _args(**kw):
    base = dict(ood_method="L2_cl_stride", cluster_method="KMeans_5", cluster_optimization_metric="silhouette",
                ind_info_creation_option="valid_preds_one_stride", which_internal_activations="ftmaps_and_strides",
                enhanced_unk_localization=False, conf_thr_train=0.15, conf_thr_test=0.15, use_values_before_sigmoid=True,
                temperature_energy=1, temperature_odin=1000, fusion_strategy="and", which_split="train", tpr_thr=0.95,
                load_thresholds=False, load_ind_activations=False, load_clusters=False)
    base.update(kw)
    return SimpleNamespace(**base)


def test_reference_driver_fits_and_decides_with_our_classes(reference_driver, tmp_path):
    ood_evaluation, ours = reference_driver
    from oracle import decide, fit, roi_align
    ood_evaluation.STORAGE_PATH = tmp_path
    log = _Logger("pipeline")
    log.setLevel(logging.ERROR)
    train, test = _batches(500, 3, 110), _batches(900, 1, 30)
    det = _Detector(None)
    det.which = train
    # ---- every method string of the reference's CLI constructs through its own selector
    for name in ("NoMethod", "MSP", "Energy", "ODIN", "Sigmoid", "L1_cl_stride", "L2_cl_stride", "Cosine_cl_stride", "Umap", "L1Ivis",
                 "L2Ivis", "CosineIvis", "fusion-MSP-Cosine_cl_stride", "fusion-MSP-Energy-L2_cl_stride"):
        m = ood_evaluation.select_ood_detection_method(_args(ood_method=name))
        assert isinstance(m, ours.OODMethod), name
    # ---- a distance method through the reference's fit pipeline
    args = _args()
    method = ood_evaluation.select_ood_detection_method(args)
    assert type(method) is ours.L2DistanceOneClusterPerStride and method.cluster_method == "KMeans_5"
    ood_evaluation.execute_pipeline_for_in_distribution_configuration(method, det, "cuda:0", _Loader(train), _Loader(train), log, args)
    files = sorted(os.listdir(tmp_path))
    assert any("_activations" in f and f.endswith(".pt") for f in files) and any("_clusters_KMeans_5_" in f for f in files) \
        and any("_thresholds_KMeans_5" in f and f.endswith(".json") for f in files), files
    thr_file = [f for f in files if f.endswith(".json")][0]
    assert json.load(open(os.path.join(tmp_path, thr_file))) == method.thresholds          # json-able, as the reference stores it
    # the oracle on the same activations: same clusters (well-separated or not, k-means labels follow sklearn), thresholds
    acts = [[[] for _ in range(3)] for _ in range(NC)]
    for b in train:
        for k in range(BATCH):
            feats = roi_align.extract_roi_aligned_features_from_correct_stride([m[k:k + 1] for m in b["maps"]], [b["det"]["boxes"][k]],
                                                                               [b["det"]["strides"][k]], (IMG, IMG))[0]
            for s, (idx, fm) in enumerate(feats):
                for j, bi in enumerate(idx):
                    acts[int(b["det"]["cls"][k][int(bi)])][s].append(fm[j])
    acts = [[np.stack(v, 0) if len(v) else np.empty(0) for v in row] for row in acts]
    n_cells = 0
    for c in range(NC):
        for s in range(3):
            assert len(method.clusters[c][s]) == (min(5, len(acts[c][s])) if len(acts[c][s]) > 3 else 0), (c, s)
            n_cells += len(method.clusters[c][s]) > 0
    assert n_cells >= 10
    scores, _, _ = fit.compute_scores_from_activations(acts, method.clusters, "l2")
    thr = fit.generate_thresholds(scores, 0.95, True, True)
    for c in range(NC):
        for s in range(3):
            assert (thr[c][s] == [] and method.thresholds[c][s] == []) or method.thresholds[c][s] == pytest.approx(thr[c][s], rel=1e-5), (c, s)
    # ---- decisions of a test batch: class surface (CUDA) vs oracle with the fitted state
    det.which = test
    results = det.predict(torch.zeros((BATCH, 3, IMG, IMG)))
    dec = method.compute_ood_decision_on_results(results, log)
    images = [dict(maps=[m[k] for m in test[0]["maps"]], boxes=test[0]["det"]["boxes"][k], cls=test[0]["det"]["cls"][k],
                   strides=test[0]["det"]["strides"][k], img_hw=(IMG, IMG)) for k in range(BATCH)]
    ref_dec, detail = decide.distance_decisions(images, method.clusters, method.thresholds, "l2", return_details=True)
    near = [abs(t[0] - (method.thresholds[t[2]][t[3]] or np.inf)) <= 1e-5 * abs(t[0]) for im in detail for t in im]
    a, b = np.array([v for d in dec for v in d]), np.array([v for d in ref_dec for v in d])
    assert [len(d) for d in dec] == [len(d) for d in ref_dec] and sum(near) <= 2
    assert np.array_equal(a[~np.array(near)], b[~np.array(near)])
    assert 0 < a.sum() < len(a)
    # ---- a logits method through the same driver
    det.mode, det.which = "logits", train
    args = _args(ood_method="MSP")
    msp = ood_evaluation.select_ood_detection_method(args)
    ood_evaluation.execute_pipeline_for_in_distribution_configuration(msp, det, "cuda:0", _Loader(train), _Loader(train), log, args)
    assert len(msp.thresholds) == NC and all(0 < t < 1 for t in msp.thresholds)
