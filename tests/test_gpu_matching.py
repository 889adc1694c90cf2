"""csrc/matching.cu (IoU x class mask, scipy-compatible assignment, valid predictions; one launch per batch) against the
golden vectors of the reference's `match_predicted_boxes_to_targets` (ood_utils.py:233-292) and against the oracle
(oracle/matching.py, itself held to scipy) on batches with ties, more predictions than targets, fewer, and empty images."""
from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ou():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from ood_in_object_detection_b200 import ood_utils
    return ood_utils


def test_matching_through_the_class_matches_reference(ou, golden):
    from ood_in_object_detection_b200.results import Results, batch_shape
    g = golden("golden_matching.npz")
    shapes = [tuple(int(v) for v in s) for s in g["shapes"]]
    for thr in (0.5, 0.3):                              # every golden case as ONE batch
        res, bb, cc = [], [], []
        for k, (P, G) in enumerate(shapes):
            b6 = np.concatenate([g[f"pred_{k}"], np.full((P, 1), 0.5, np.float32), g[f"pcls_{k}"][:, None]], 1).astype(np.float32)
            res.append(Results(orig_img=batch_shape(len(shapes), 640, 640), boxes=torch.from_numpy(b6).reshape(P, 6).cuda()))
            bb.append(torch.from_numpy(g[f"gt_{k}"]))
            cc.append(torch.from_numpy(g[f"gcls_{k}"]))
        ou.OODMethod.match_predicted_boxes_to_targets(res, dict(bboxes=bb, cls=cc), thr)
        for k, (P, G) in enumerate(shapes):
            assert res[k].valid_preds == g[f"valid_{k}_{thr}"].tolist(), (k, thr)
            assert tuple(res[k].assignment_score_matrix.shape) == (P, G) and len(res[k].assignment[0]) == min(P, G)
    ou.OODMethod.match_predicted_boxes_to_targets(res, dict(bboxes=bb, cls=cc), 0.5, compat=False)
    for k, (P, G) in enumerate(shapes):
        m = res[k].assignment_score_matrix
        assert all(float(m[i].max()) > 0.5 for i in res[k].valid_preds)
        if P <= G:
            assert res[k].valid_preds == g[f"valid_{k}_0.5"].tolist()


def test_matching_kernel_matches_oracle_on_ragged_batches(ou):
    from scipy.optimize import linear_sum_assignment
    from ood_in_object_detection_b200 import ops
    from oracle import matching
    rng = np.random.default_rng(3)
    for rep in range(4):
        preds, pcls, gts, gcls = [], [], [], []
        for i in range(48):
            G = int(rng.integers(0, 40)) if i % 7 else 0
            P = int(rng.integers(0, 120)) if i % 11 else 0
            if rep == 3 and i == 5:
                P, G = 300, 90                          # max_det predictions against a crowded image
            c = rng.uniform(40, 600, size=(max(G, 1), 2))
            wh = rng.uniform(20, 160, size=(max(G, 1), 2))
            gt = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)[:G]
            gc = rng.integers(0, 3, size=G).astype(np.float32)
            if G:
                src = rng.integers(0, G, size=P)
                jitter = rng.normal(0, 8, size=(P, 4)).astype(np.float32) * (rng.uniform(size=(P, 1)) < 0.7)   # 30 % exact copies: ties
                pr = gt[src] + jitter
                pc = np.where(rng.uniform(size=P) < 0.8, gc[src], (gc[src] + 1) % 3).astype(np.float32)
            else:
                pr = rng.uniform(0, 600, size=(P, 4)).astype(np.float32)
                pr[:, 2:] += pr[:, :2]
                pc = rng.integers(0, 3, size=P).astype(np.float32)
            preds.append(pr), pcls.append(pc), gts.append(gt), gcls.append(gc)
        for compat in (True, False):
            out = ops.match_boxes(preds, pcls, gts, gcls, 0.5, compat=compat)
            for i, (valid, score, (rows, cols)) in enumerate(out):
                ref_valid, ref_score, (rr, rc) = matching.match_predictions(preds[i], pcls[i], gts[i], gcls[i], 0.5, compat=compat)
                assert np.array_equal(score, ref_score), i                  # float32 IoU, bit for bit
                assert np.array_equal(rows, rr) and np.array_equal(cols, rc), (rep, i)
                assert valid == ref_valid, (rep, i)
                if score.size:
                    a = linear_sum_assignment(score, maximize=True)
                    assert np.array_equal(a[0], rows) and np.array_equal(a[1], cols)
