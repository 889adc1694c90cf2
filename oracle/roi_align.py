"""RoIAlign 1x1 adaptive-grid oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, in numpy float32 with the same operation order:
  * torchvision 0.26 `torch.ops.torchvision.roi_align` forward for
    `output_size=(1,1), sampling_ratio=-1, aligned=False` -- the call made at
    /root/reference/ultralytics/models/yolo/detect/predict.py:64-70.  The C++/CUDA
    kernel source is not in this image (only `_C.so`); the written spec is
    `site-packages/torchvision/ops/roi_align.py:35-200` (`_roi_align`,
    `_bilinear_interpolate`) plus the published kernel behaviour that samples
    with y<-1, y>H, x<-1 or x>W contribute 0 (verified against the op
    in tests/test_oracle_vs_golden.py::test_roi_align_matches_torchvision).
  * `extract_roi_aligned_features_from_correct_stride`
    (/root/reference/ultralytics/models/yolo/detect/predict.py:13-90).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def roi_params(box_xyxy, spatial_scale):
    """predict.py:64-70 -> roi_align: start/size/grid of one ROI, all in float32."""
    sc = F32(spatial_scale)
    x1, y1, x2, y2 = (F32(v) for v in box_xyxy)
    sw, sh, ew, eh = x1 * sc, y1 * sc, x2 * sc, y2 * sc          # roi_align.py:136-139 (offset = 0)
    rw = max(F32(ew - sw), F32(1.0))                              # roi_align.py:144-145 (aligned=False)
    rh = max(F32(eh - sh), F32(1.0))
    gh = int(np.ceil(rh))                                         # roi_align.py:152-153, pooled size 1
    gw = int(np.ceil(rw))
    return sw, sh, rw, rh, gw, gh


def _axis_samples(start, size, grid, extent):
    """Sample coordinates along one axis and their bilinear split.

    Returns (low, high, l, h, valid) for the `grid` samples; kernel formula
    `start + (i + .5f) * size / grid` evaluated left to right in float32.
    """
    i = np.arange(grid, dtype=F32)
    c = (start + ((i + F32(0.5)) * size) / F32(grid)).astype(F32)
    valid = ~((c < F32(-1.0)) | (c > F32(extent)))
    c = np.where(c <= 0, F32(0), c).astype(F32)                   # roi_align.py:46-47 clamp(min=0)
    low = c.astype(np.int32)
    edge = low >= extent - 1                                      # roi_align.py:50-56
    high = np.where(edge, extent - 1, low + 1).astype(np.int32)
    low = np.where(edge, extent - 1, low).astype(np.int32)
    c = np.where(edge, low.astype(F32), c).astype(F32)
    l = (c - low.astype(F32)).astype(F32)
    h = (F32(1.0) - l).astype(F32)
    return low, high, l, h, valid


def roi_align_1x1(fmap, rois, spatial_scale):
    """fmap [N,C,H,W] f32, rois [K,5] (batch_idx,x1,y1,x2,y2) f32 -> [K,C] f32.

    Accumulates the samples sequentially (iy outer, ix inner) in float32 like the kernel.
    """
    fmap = np.ascontiguousarray(fmap, dtype=F32)
    rois = np.asarray(rois, dtype=F32).reshape(-1, 5)
    _, C, H, W = fmap.shape
    out = np.zeros((rois.shape[0], C), dtype=F32)
    for k, roi in enumerate(rois):
        b = int(roi[0])
        sw, sh, rw, rh, gw, gh = roi_params(roi[1:], spatial_scale)
        yl, yh, ly, hy, vy = _axis_samples(sh, rh, gh, H)
        xl, xh, lx, hx, vx = _axis_samples(sw, rw, gw, W)
        img = fmap[b]                                             # [C,H,W]
        w1 = (hy[:, None] * hx[None, :]).astype(F32)
        w2 = (hy[:, None] * lx[None, :]).astype(F32)
        w3 = (ly[:, None] * hx[None, :]).astype(F32)
        w4 = (ly[:, None] * lx[None, :]).astype(F32)
        v1 = img[:, yl[:, None], xl[None, :]]
        v2 = img[:, yl[:, None], xh[None, :]]
        v3 = img[:, yh[:, None], xl[None, :]]
        v4 = img[:, yh[:, None], xh[None, :]]
        val = ((w1 * v1 + w2 * v2) + w3 * v3) + w4 * v4           # [C,gh,gw] float32
        val = np.where(vy[None, :, None] & vx[None, None, :], val, F32(0)).astype(F32)
        acc = np.cumsum(val.reshape(C, -1), axis=1, dtype=F32)[:, -1] if gh * gw > 0 else np.zeros(C, F32)
        out[k] = acc / F32(max(gh * gw, 1))
    return out


def axis_weights(start, size, grid, extent):
    """Separable form used by the CUDA kernel: per-row (or per-column) summed bilinear weights.

    sum_{iy,ix} bilinear(y_iy, x_ix) == sum_r sum_c wy[r] * wx[c] * v[r, c]  because the
    bilinear weights and the out-of-range mask both factor over the two axes.
    Returns (first_index, weights[float32]).  Same math, different summation order, so it
    agrees with `roi_align_1x1` to float32 rounding (not bit-exactly).
    """
    low, high, l, h, valid = _axis_samples(start, size, grid, extent)
    w = np.zeros(extent, dtype=np.float64)
    np.add.at(w, low[valid], h[valid].astype(np.float64))
    np.add.at(w, high[valid], l[valid].astype(np.float64))
    nz = np.nonzero(w)[0]
    if nz.size == 0:
        return 0, np.zeros(0, F32)
    return int(nz[0]), w[nz[0]:nz[-1] + 1].astype(F32)


def extract_roi_aligned_features_from_correct_stride(ftmaps, boxes, strides, img_shape, extract_all_strides=False):
    """predict.py:13-90 on numpy inputs.

    ftmaps: list of S arrays [N,C_s,H_s,W_s]; boxes: list of N arrays [M_i,4];
    strides: list of N arrays [M_i] in {0,1,2}; img_shape (H, W).
    Returns out[img][stride] = [idx_in_img (int16), feats [m,C_s,1,1]] like the reference.
    """
    n_img = len(boxes)
    strides_cat = np.concatenate([np.asarray(s, dtype=F32) for s in strides]) if n_img else np.zeros(0, F32)
    img_idx = np.concatenate([np.full(len(b), i, dtype=np.int64) for i, b in enumerate(boxes)])
    boxes_cat = np.concatenate([np.asarray(b, dtype=F32).reshape(-1, 4) for b in boxes])
    rois = np.concatenate([img_idx[:, None].astype(F32), boxes_cat], axis=1)
    out = [[[[], []] for _ in ftmaps] for _ in range(n_img)]
    for s, fm in enumerate(ftmaps):
        mask = np.ones(len(rois), bool) if extract_all_strides else (strides_cat == s)   # predict.py:52-58
        rel = rois[mask]
        if rel.shape[0] == 0:
            feats = np.zeros((0,), F32)
        else:
            feats = roi_align_1x1(fm, rel, fm.shape[-1] / img_shape[1])[:, :, None, None]  # predict.py:64-70
        for i in range(n_img):                                                           # predict.py:78-88
            in_img = img_idx == i
            idx = np.arange(int(in_img.sum()), dtype=np.int16)
            if not extract_all_strides:
                idx = idx[mask[in_img]]
            out[i][s][0] = idx
            out[i][s][1] = feats[in_img[mask]] if rel.shape[0] else feats
    return out
