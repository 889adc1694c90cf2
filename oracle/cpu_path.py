"""Loop-for-loop CPU port of the reference's scoring / fit path, for TIMING.  TEST INFRASTRUCTURE.

The reference is Python and cannot travel to the GPU box (/root/reference does not exist
there), so `bench.py`'s `cpu_baseline` / `--impl reference` legs time this port instead
(`cpu_baseline.kind = "port"`).  It performs the same work the reference performs, in the same
loop structure, with the SAME third-party calls the reference makes (torchvision
`roi_align`, sklearn `normalize` / `pairwise_distances` / `KMeans`, numpy `percentile`,
torch `softmax` / `logsumexp`), on torch CPU tensors:

  per image -> per stride -> per box, one `.cpu().numpy()`, one `normalize`, one
  `pairwise_distances` per box        /root/reference/ood_utils.py:2038-2180
  the per-stride extractor             /root/reference/ultralytics/models/yolo/detect/predict.py:13-90
  logit methods, per box               /root/reference/ood_utils.py:1195-1208, 1388-1443
  fit: clusters, scores, thresholds    /root/reference/ood_utils.py:2263-2371, 1877-1915, 583-637;
                                       /root/reference/cluster_utils.py:62-73

It is checked against the golden vectors produced by the real reference in
tests/test_oracle_vs_golden.py, so its results (not only its cost) are the reference's.
"""
from __future__ import annotations

import numpy as np
import torch

F32 = np.float32


def extract_roi_aligned_features_from_correct_stride(ftmaps, boxes, strides, img_shape, device="cpu",
                                                     extract_all_strides=False):
    """predict.py:13-90 with torch + torchvision, same tensor ops in the same order."""
    from torchvision.ops import roi_align
    batch_size = len(boxes)
    strides_cat = torch.cat(strides)
    img_indices = torch.cat([torch.full((len(b),), i, dtype=torch.long, device=device) for i, b in enumerate(boxes)])
    boxes_cat = torch.cat(boxes)
    rois = torch.cat([img_indices.unsqueeze(1), boxes_cat], dim=1)
    out = [[[[] for _ in range(2)] for _ in range(len(ftmaps))] for _ in range(batch_size)]
    for stride_idx, ftmap in enumerate(ftmaps):
        if extract_all_strides:
            relevant = rois
            stride_mask = torch.full((relevant.shape[0],), True, dtype=torch.bool, device=device)
        else:
            stride_mask = strides_cat == stride_idx
            relevant = rois[stride_mask]
        if relevant.shape[0] == 0:
            feats = torch.empty(0, device=device)
        else:
            feats = roi_align(input=ftmap, boxes=relevant, output_size=(1, 1),
                              spatial_scale=ftmap.shape[-1] / img_shape[1], aligned=False)
        for idx_img in range(batch_size):
            img_mask = rois[:, 0] == idx_img
            img_mask_cur = img_mask[stride_mask]
            stride_mask_cur = stride_mask[img_mask]
            idx_in_img = torch.arange(sum(img_mask), device=device, dtype=torch.int16)   # python sum(), as in :84
            if not extract_all_strides:
                idx_in_img = idx_in_img[stride_mask_cur]
            out[idx_img][stride_idx][0] = idx_in_img
            out[idx_img][stride_idx][1] = feats[img_mask_cur, ...]
    return out


def distance_decisions(images, clusters, thresholds, metric):
    """ood_utils.py:2038-2180.  images: dicts with torch tensors maps (3 x CHW), boxes, cls, strides."""
    from sklearn.metrics import pairwise_distances
    from sklearn.preprocessing import normalize
    ood_decision = []
    for im in images:
        dec = []
        feats = extract_roi_aligned_features_from_correct_stride(
            ftmaps=[ft[None, ...] for ft in im["maps"]], boxes=[im["boxes"]], strides=[im["strides"]],
            img_shape=im["img_hw"], device=im["boxes"].device)[0]
        cls_all = im["cls"].cpu()
        for stride_idx, (bbox_idx_in_one_stride, ftmaps) in enumerate(feats):
            if len(bbox_idx_in_one_stride) > 0:
                for idx, ftmap in enumerate(ftmaps):
                    cls_idx = int(cls_all[idx])                                         # Q1
                    ftmap = ftmap.cpu().unsqueeze(0).numpy()
                    if len(clusters[cls_idx][stride_idx]) == 0:
                        distance = 1000
                    else:
                        x = normalize(ftmap.reshape(ftmap.shape[0], -1), axis=1)
                        distance = pairwise_distances(clusters[cls_idx][stride_idx], x, metric=metric).min(axis=0)[0]
                    thr = thresholds[cls_idx][stride_idx]
                    if thr:
                        dec.append(1 if distance < thr else 0)
                    else:
                        dec.append(0)
        ood_decision.append(dec)
    return ood_decision


def logit_score_torch(logits, cls_idx, method, temper=1.0):
    """ood_utils.py:1394-1443 on a torch CPU tensor."""
    if len(logits.shape) == 1:
        logits = logits.unsqueeze(0)
    if method == "MSP":
        return torch.nn.functional.softmax(logits, dim=1)[:, cls_idx].numpy()
    if method == "Energy":
        return temper * torch.logsumexp(logits / temper, dim=1).numpy()
    if method == "ODIN":
        return torch.nn.functional.softmax(logits / temper, dim=1)[:, cls_idx].numpy()
    if method == "Sigmoid":
        s = torch.sigmoid(logits).numpy()
        assert (cls_idx == s.argmax(axis=1)).all(), "The max logit is not the one of the predicted class"
        return s[:, cls_idx]
    raise ValueError(method)


def logit_decisions(images, method, thresholds, temper=1.0):
    """ood_utils.py:1195-1208 -- per box `.cpu()` + tiny torch ops."""
    out = []
    for im in images:
        dec = []
        for idx_bbox in range(len(im["cls"])):
            cls_idx = int(im["cls"][idx_bbox].cpu())
            logits = im["logits"][idx_bbox].cpu()
            score = logit_score_torch(logits, cls_idx, method, temper)[0]
            dec.append(0 if score < thresholds[cls_idx] else 1)
        out.append(dec)
    return out


def fit_distance(activations, cluster_method, metric, tpr=0.95):
    """generate_clusters + compute_scores_from_activations + generate_thresholds with sklearn, as the
    reference calls them.  activations[cls][stride] = ndarray [N,C,1,1] or empty.  -> (clusters, scores, thr, n_iter)."""
    from sklearn.cluster import KMeans
    from sklearn.metrics import pairwise_distances
    from sklearn.preprocessing import normalize
    ncls = len(activations)
    clusters = [[[] for _ in range(3)] for _ in range(ncls)]
    scores = [[[] for _ in range(3)] for _ in range(ncls)]
    thr = [[[] for _ in range(3)] for _ in range(ncls)]
    n_iter = []
    for c in range(ncls):
        for s in range(3):
            a = activations[c][s]
            if len(a) > 3:
                x = normalize(a.reshape(a.shape[0], -1), axis=1)
                if cluster_method == "one":
                    clusters[c][s] = np.mean(x, axis=0)[None, :]
                else:
                    k = min(int(cluster_method.split("_")[-1]), len(x))
                    km = KMeans(n_clusters=k, random_state=10)
                    lab = km.fit_predict(x)
                    n_iter.append(km.n_iter_)
                    clusters[c][s] = np.array([np.mean(x[lab == j], axis=0) for j in sorted(set(lab))])
            else:
                clusters[c][s] = np.empty(0)
    for c in range(ncls):
        for s in range(3):
            a = activations[c][s]
            if len(a) > 0:
                if len(clusters[c][s]) > 0:
                    x = normalize(a.reshape(a.shape[0], -1), axis=1)
                    scores[c][s] = pairwise_distances(clusters[c][s], x, metric=metric).min(axis=0)
            else:
                scores[c][s] = np.empty(0)
            if len(scores[c][s]) > 5:
                thr[c][s] = float(np.percentile(scores[c][s], 100 * tpr, method="lower"))
    return clusters, scores, thr, n_iter
