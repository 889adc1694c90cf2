"""Host-side logic of the class surface that needs no GPU: prediction <-> ground-truth matching and the target
dictionary (golden vectors from the reference, tests/golden/make_golden.py::golden_matching), constructor
validation, fitted-state plumbing of the fusion classes."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from ood_in_object_detection_b200 import ood_utils as ou
from ood_in_object_detection_b200.results import Results, batch_shape

DIST_KW = dict(agg_method="mean", cluster_method="one", cluster_optimization_metric="silhouette",
               ind_info_creation_option="valid_preds_one_stride", which_internal_activations="ftmaps_and_strides",
               iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)
LOGIT_KW = dict(per_class=True, per_stride=False, iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15,
                min_conf_threshold_test=0.15, use_values_before_sigmoid=True)
COMMON = dict(iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)


def test_matching_oracle_matches_reference(golden):
    """oracle/matching.py (box_iou, scipy's assignment restated, the reference's walk with quirk Q8) against the valid_preds
    the reference produced (golden_matching.npz), and its assignment solver against scipy on random and tie-heavy matrices.
    The CUDA kernel (csrc/matching.cu) is held to both in tests/test_gpu_matching.py."""
    from scipy.optimize import linear_sum_assignment
    from oracle import matching
    g = golden("golden_matching.npz")
    for k, (P, G) in enumerate(g["shapes"]):
        for thr in (0.5, 0.3):
            valid, score, _ = matching.match_predictions(g[f"pred_{k}"], g[f"pcls_{k}"], g[f"gt_{k}"], g[f"gcls_{k}"], thr)
            assert valid == g[f"valid_{k}_{thr}"].tolist(), (k, thr)
        valid, score, _ = matching.match_predictions(g[f"pred_{k}"], g[f"pcls_{k}"], g[f"gt_{k}"], g[f"gcls_{k}"], 0.5, compat=False)
        assert all(float(score[i].max()) > 0.5 for i in valid)              # intended semantics: a real same-class overlap
    rng = np.random.default_rng(0)
    for trial in range(800):
        P, G = int(rng.integers(0, 14)), int(rng.integers(0, 10))
        kind = trial % 4
        c = rng.uniform(0, 1, (P, G))
        if kind == 1:
            c = c * (rng.uniform(size=(P, G)) < 0.25)                     # mostly zeros, like IoU x mask
        elif kind == 2:
            c = rng.integers(0, 3, (P, G)).astype(float)                  # small integers: heavy ties
        elif kind == 3:
            c = np.round(c, 1) * (rng.uniform(size=(P, G)) < 0.5)
        c = c.astype(np.float32)
        for mx in (True, False):
            a, b = linear_sum_assignment(c, maximize=mx), matching.lsap(c, mx)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (trial, mx)


def test_targets_dict_matches_reference(golden):
    g = golden("golden_matching.npz")
    data = dict(im_file=["a", "b", "c"], batch_idx=torch.from_numpy(g["td_batch_idx"]), bboxes=torch.from_numpy(g["td_bboxes"]),
                cls=torch.from_numpy(g["td_cls"]), resized_shape=[tuple(int(v) for v in r) for r in g["td_shapes"]])
    t = ou.OODMethod.create_targets_dict(data)
    for i in range(3):
        np.testing.assert_allclose(np.asarray(t["bboxes"][i], np.float64), g[f"td_out_bboxes_{i}"], rtol=1e-6, atol=1e-4)
        assert np.array_equal(np.asarray(t["cls"][i]), g[f"td_out_cls_{i}"])


def test_constructor_contracts():
    m = ou.L2DistanceOneClusterPerStride(**DIST_KW)
    assert (m.name, m.metric, m.is_distance_method, m.per_class, m.per_stride) == ("L2DistancePerStride", "l2", True, True, True)
    assert ou.L1DistanceOneClusterPerStride(**DIST_KW).name == "L1DistancePerStride"
    assert ou.CosineDistanceOneClusterPerStride(**DIST_KW).metric == "cosine"
    assert ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, cluster_method="KMeans_16")).cluster_method == "KMeans_16"
    for bad in (dict(which_internal_activations="logits"), dict(cluster_method="nope"), dict(agg_method="max"),
                dict(ind_info_creation_option="x"), dict(cluster_optimization_metric="x")):
        with pytest.raises(AssertionError):
            ou.L2DistanceOneClusterPerStride(**dict(DIST_KW, **bad))
    a, e, o, s = ou.MSP(**LOGIT_KW), ou.Energy(temper=1, **LOGIT_KW), ou.ODIN(temper=1000, **LOGIT_KW), ou.Sigmoid(**LOGIT_KW)
    assert (a.name, e.name, o.name, s.name) == ("MSP", "Energy", "ODIN", "MSP")          # Sigmoid's name is 'MSP' upstream too
    assert a.which_internal_activations == "logits" and a.cluster_method == "None" and not a.is_distance_method
    assert ou.NoMethod(**LOGIT_KW).compute_scores(torch.zeros(4, 20), 3).tolist() == [1.0] * 4
    with pytest.raises(NotImplementedError):
        a.compute_distance(None, None)
    assert a.compute_indness.__doc__ and m.thresholds is None and m.clusters is None


def test_fusion_state_plumbing():
    a, m = ou.MSP(**LOGIT_KW), ou.L2DistanceOneClusterPerStride(**DIST_KW)
    f = ou.FusionMethod(a, m, "score", fusion_method_name="fusion-MSP-L2_cl_stride", cluster_method="one", **COMMON)
    assert f.is_distance_method and f.cluster_method == "one" and f.which_internal_activations == "none"
    f.clusters = [[np.ones((1, 4), np.float32)] * 3]
    assert m.clusters is f.clusters
    f.thresholds = ([0.1], [[1.0, 2.0, 3.0]])
    assert a.thresholds == [0.1] and m.thresholds == [[1.0, 2.0, 3.0]]
    f.thresholds = None
    assert a.thresholds is None and m.thresholds is None
    t = ou.TripleFusionMethod(a, ou.Energy(temper=1, **LOGIT_KW), m, cluster_method="one", **COMMON)
    assert t.name == "fusion-MSP-Energy_L2DistancePerStride" and t.fusion_strategy == "majority_voting"
    t.clusters = [[np.zeros((1, 4), np.float32)] * 3]
    assert m.clusters is t.clusters
    with pytest.raises(ValueError):
        t.thresholds = ([], [])

    class _M:                                           # configure_extra_output_of_the_model is attribute plumbing
        class model:
            model = [type("Head", (), {})()]
        ckpt_path = "yolov8n.pt"
    ou.configure_extra_output_of_the_model(_M, a)
    assert _M.model.which_layers_to_extract == "logits" and _M.model.model[-1].output_values_before_sigmoid is True
    ou.configure_extra_output_of_the_model(_M, m)
    assert _M.model.which_layers_to_extract == "convolutional_layers" and _M.model.extraction_mode == "ftmaps_and_strides"
    assert _M.model.model[-1].output_values_before_sigmoid is False


def test_first_centre_draw_is_randomstate_choice():
    """kmeans._sklearn_first_center restates RandomState.choice(n, p=uniform float32 weights) without its O(n) checks:
    same index and same stream position for every n (sklearn _kmeans.py:234)."""
    from ood_in_object_detection_b200 import kmeans
    for n in (1, 2, 3, 7, 100, 4097, 200000, 1234567):
        for seed in (10, 0, 123):
            a, b = np.random.RandomState(seed), np.random.RandomState(seed)
            w = np.ones(n, dtype=np.float32)
            ref = int(a.choice(n, p=w / w.sum()))
            assert kmeans._sklearn_first_center(b, n) == ref, (n, seed)
            assert a.uniform() == b.uniform()


def test_parallel_evaluation_of_the_sequential_float32_cumsum():
    """oracle/scan_model.py (the arithmetic of csrc/seed.cu::seed_scan_kernel in numpy: integer increments per binade, tie
    parity through a scan of parity maps, crossing blocks redone with the float chain) == np.cumsum(float32) at every
    128-value boundary, bit for bit, on data where ties, binade jumps, zeros and denormals occur."""
    from oracle import scan_model
    rng = np.random.default_rng(23)
    n = 30001
    jumps = rng.random(n).astype(np.float32) ** 6
    jumps[[5, 1000, 1001, 20000]] = [3e4, 7e6, 1e-30, 4e9]
    cases = {"integers": rng.integers(0, 1024, n).astype(np.float32) * np.float32(40.0),                     # ties once ulp(sum) >= 80
             "odd": np.where(np.arange(n) % 7 == 0, 1025.0, 513.0).astype(np.float32) * np.float32(64.0),     # EVERY add a tie later on
             "tails": np.exp(rng.standard_normal(n) * 4).astype(np.float32), "jumps": jumps,
             "zeros": np.concatenate([np.zeros(5000, np.float32), rng.random(n - 5000).astype(np.float32) * 1e-3]),
             "denormals": (rng.random(n) * 1e-41).astype(np.float32), "plain": (rng.random(n) ** 4).astype(np.float32),
             "short": np.float32([0.1, 0.2, 0.3]), "empty_tail": np.full(256, 0.25, np.float32)}
    ties = 0
    for name, v in cases.items():
        cum = np.cumsum(v)
        ends = np.concatenate([cum[127::128], cum[-1:]]) if len(v) % 128 else cum[127::128]
        got = scan_model.boundary_sums(v)
        assert np.array_equal(got.view(np.uint32), ends.view(np.uint32)), name
        assert np.array_equal(scan_model.boundary_sums(v, blocks_per_pass=7).view(np.uint32), ends.view(np.uint32)), name
        s = cum[len(cum) // 2]
        if 1e-30 < s < 1e30:                                                     # ties against the unit of the sum half-way through
            with np.errstate(over="ignore", invalid="ignore"):
                x = v * np.float32(2.0 ** (23 - (np.frexp(s)[1] - 1)))
                ties += int((x - np.floor(x) == 0.5).sum())
    assert ties > 1000                                                                                     # the tie rule is exercised
