"""Shared helpers for the parity tests: rebuild inputs of the golden cases."""
from __future__ import annotations

import numpy as np

from ood_in_object_detection_b200 import synth

F32 = np.float32


def unpack_nested(g, prefix, ncls, as_threshold=False):
    """npz -> [cls][stride] nested lists (inverse of make_golden._pack_nested)."""
    out = []
    for c in range(ncls):
        row = []
        for s in range(3):
            v = g[f"{prefix}_{c}_{s}"]
            if as_threshold:
                row.append(float(v) if v.ndim == 0 else [])
            else:
                row.append(v)
        out.append(row)
    return out


def split(a, counts):
    idx = np.cumsum(counts)[:-1]
    return np.split(a, idx)


def scoring_case(g):
    """Rebuild the test images of a golden_scoring file: maps from the seed, detections from the file."""
    seed, img, B = int(g["seed"]), int(g["img"]), int(g["batch"])
    ch = tuple(int(c) for c in g["channels"])
    hw = tuple(img // s for s in synth.STRIDES)
    maps = synth.feature_maps(seed + 2, B, ch, hw)
    n = g["n_boxes"]
    boxes, cls, strides = split(g["boxes"], n), split(g["cls"], n), split(g["strides"], n)
    images = [dict(maps=[m[i] for m in maps], boxes=boxes[i], cls=cls[i], strides=strides[i], img_hw=(img, img))
              for i in range(B)]
    return images, maps


def train_case(g):
    seed, img, B = int(g["seed"]), int(g["img"]), int(g["batch"])
    ch = tuple(int(c) for c in g["channels"])
    hw = tuple(img // s for s in synth.STRIDES)
    maps = synth.feature_maps(seed, B, ch, hw)
    n = g["train_n_boxes"]
    return maps, split(g["train_boxes"], n), split(g["train_cls"], n), split(g["train_strides"], n)


class Projection:
    """Fixed stand-in for a TRAINED dimensionality reducer (ivis / umap model) of the SDR methods: tanh of a seeded random
    projection to `d_out` dimensions.  Used on both sides of the C4 golden (the reference's IvisMethodCosine and ours)."""

    def __init__(self, seed: int, d_in: int, d_out: int = 32):
        rng = np.random.default_rng(seed)
        self.w = (rng.standard_normal((d_in, d_out)) * (4.0 / np.sqrt(d_in))).astype(np.float64)

    def transform(self, a) -> np.ndarray:
        a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
        return np.tanh(a @ self.w).astype(np.float32)


def nms_inputs(seed=91, bs=4, nc=20, img=640):
    """Detector-head-shaped inputs of golden_nms.npz / golden_postprocess.npz (the generator lives in the package: synth.head_output)."""
    from ood_in_object_detection_b200 import synth
    return synth.head_output(seed, bs, nc, img)


def postprocess_inputs():
    """Head output + hooked maps of golden_postprocess.npz (make_golden.py and tests/test_gpu_postprocess.py build the same)."""
    from ood_in_object_detection_b200 import synth
    pred, logits, _ = nms_inputs(seed=93, bs=4, nc=20, img=320)
    maps = synth.feature_maps(7, 4, (16, 24, 32), (40, 20, 10))
    return pred, logits, maps


def fake_predictor(mode, before_sigmoid, conf, device="cpu"):
    """The attributes `DetectionPredictor.postprocess` reads (predict.py:117-363), nothing else."""
    from types import SimpleNamespace as NS
    return NS(args=NS(conf=conf, iou=0.45, agnostic_nms=False, max_det=300, classes=None, model="yolov8s.pt", task="detect"),
              model=NS(model=NS(extraction_mode=mode, model=[NS(output_values_before_sigmoid=before_sigmoid)]),
                       names={i: str(i) for i in range(20)}),
              device=__import__('torch').device(device), batch=[[f"im{i}.jpg" for i in range(4)]])
