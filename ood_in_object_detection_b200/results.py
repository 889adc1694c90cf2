"""Minimal stand-ins for the detector-side types that cross the hot-path boundary.

The OoD method classes only read `res.boxes.xyxy / .cls / .conf`, `res.extra_item`, `res.orig_img.shape` and
(fit side) `res.valid_preds` from `ultralytics.engine.results.Results`
(/root/reference/ultralytics/engine/results.py:237-281, Boxes :1114-1148).  The real ultralytics objects work
unchanged (duck typing); these two classes exist so that the hot path can be driven and tested without the
detector package.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch


class Boxes:
    """`data` = [M, 6] (x1, y1, x2, y2, conf, cls), as `ultralytics.engine.results.Boxes` stores it."""

    def __init__(self, data: torch.Tensor, orig_shape=None):
        if not isinstance(data, torch.Tensor):
            data = torch.as_tensor(np.asarray(data))
        if data.ndim == 1:
            data = data[None, :]
        assert data.shape[-1] == 6, f"expected [M, 6] boxes (xyxy, conf, cls), got {tuple(data.shape)}"
        self.data = data
        self.orig_shape = orig_shape

    @property
    def xyxy(self) -> torch.Tensor:
        return self.data[:, :4]

    @property
    def conf(self) -> torch.Tensor:
        return self.data[:, 4]

    @property
    def cls(self) -> torch.Tensor:
        return self.data[:, 5]

    def __len__(self) -> int:
        return int(self.data.shape[0])


class Results:
    """One image's detections plus the extra item the detector hooks out.

    orig_img   the WHOLE uint8 batch [B, H, W, 3] when the source is a tensor (predict.py:342-354), so that
               `orig_img.shape[1:3] == (H, W)` of the network input (ood_utils.py:2061); an object with `.shape`
               is enough
    extra_item `(ftmaps: list of 3 [C_s, H_s, W_s] tensors, strides [M])` for 'ftmaps_and_strides',
               a per-stride list of `(idx_in_img, feats [m_s, C_s, 1, 1])` for 'roi_aligned_ftmaps',
               or the raw class logits [M, NC] for the logits methods
    """

    def __init__(self, orig_img, path: str = "", names: Optional[Dict[int, str]] = None, boxes=None, extra_item: Any = None):
        self.orig_img = orig_img
        self.orig_shape = tuple(orig_img.shape[:2]) if hasattr(orig_img, "shape") else None
        self.path = path
        self.names = names or {}
        self.boxes = boxes if isinstance(boxes, Boxes) or boxes is None else Boxes(boxes, self.orig_shape)
        self.extra_item = extra_item
        self.valid_preds = []

    def __len__(self) -> int:
        return 0 if self.boxes is None else len(self.boxes)


class _Shape:
    """Shape-only placeholder for `orig_img` (avoids allocating a uint8 batch just to carry H and W)."""

    def __init__(self, *shape):
        self.shape = tuple(shape)


def batch_shape(n_batch: int, h: int, w: int) -> _Shape:
    return _Shape(n_batch, h, w, 3)
