"""Synthetic workloads of the shapes BASELINE.json names.

Feature maps are generated with integer draws and exactly-rounded IEEE operations only
(no transcendental functions), so the same seed gives bit-identical maps on every machine;
that lets the golden fixtures under tests/golden/ store only the seed for the maps.
Value range mimics the post-SiLU activations the reference reports
(`inspect_activations.ipynb` cell 17: min ~ -0.28, max 5.6-9.6).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

F32 = np.float32

# hooked-map channels at strides 8/16/32 (DetectionModel('yolov8{n,s,m,l,x}.yaml'))
YOLOV8_CHANNELS = {"n": (64, 128, 256), "s": (128, 256, 512), "m": (192, 384, 576), "l": (256, 512, 512),
                   "x": (320, 640, 640)}
STRIDES = (8, 16, 32)


@dataclass
class Workload:
    name: str
    model: str
    img: int
    batch: int
    nc: int
    k: int
    lam: float            # mean boxes per image (Poisson), or exact count when fixed_boxes
    fixed_boxes: bool = False
    channels: tuple = field(default=None)

    def __post_init__(self):
        if self.channels is None:
            self.channels = YOLOV8_CHANNELS[self.model]

    @property
    def map_hw(self):
        return tuple(self.img // s for s in STRIDES)


# BASELINE.json configs[0..4]
CONFIGS = {
    "C1": Workload("C1 YOLOv8n 640 B=8 L2 K=1", "n", 640, 8, 20, 1, 30),
    "C2": Workload("C2 YOLOv8s 640 B=64 L1/cos+logits K=10", "s", 640, 64, 20, 10, 50),
    "C4": Workload("C4 YOLOv8l 640 B=32/GPU K=10", "l", 640, 32, 20, 10, 50),
    "C5": Workload("C5 YOLOv8x 1280 B=16 300 boxes K=64", "x", 1280, 16, 20, 64, 300, fixed_boxes=True),
}


def feature_map(rng: np.random.Generator, shape) -> np.ndarray:
    """Exact, platform-independent float32 map: 7*u^4 - 0.27 with u = k/65536."""
    u = rng.integers(0, 65536, size=shape, dtype=np.uint16).astype(np.float64) / 65536.0
    u *= u
    u *= u
    return (7.0 * u - 0.27).astype(F32)


def feature_maps(seed: int, batch: int, channels, map_hw) -> List[np.ndarray]:
    rng = np.random.default_rng(seed)
    return [feature_map(rng, (batch, c, hw, hw)) for c, hw in zip(channels, map_hw)]


def boxes_for_image(rng: np.random.Generator, img: int, nc: int, m: int):
    """Boxes per SURVEY.md §8d: uniform centre, log-uniform side in [16, 0.6*img]*(img/640),
    aspect in U[0.5,2], clipped; FPN-like stride id by size; uniform class."""
    scale = img / 640.0
    ctr = rng.uniform(0, img, size=(m, 2))
    side = np.exp(rng.uniform(np.log(16 * scale), np.log(0.6 * img), size=m))
    asp = rng.uniform(0.5, 2.0, size=m)
    w, h = side * np.sqrt(asp), side / np.sqrt(asp)
    x1 = np.clip(ctr[:, 0] - w / 2, 0, img)
    x2 = np.clip(ctr[:, 0] + w / 2, 0, img)
    y1 = np.clip(ctr[:, 1] - h / 2, 0, img)
    y2 = np.clip(ctr[:, 1] + h / 2, 0, img)
    box = np.stack([x1, y1, x2, y2], axis=1).astype(F32)
    mx = np.maximum(box[:, 2] - box[:, 0], box[:, 3] - box[:, 1])
    stride = np.where(mx < 64 * scale, 0, np.where(mx < 192 * scale, 1, 2)).astype(F32)
    cls = rng.integers(0, nc, size=m).astype(F32)
    conf = rng.uniform(0.15, 1.0, size=m).astype(F32)
    return box, cls, stride, conf


def logits_for(rng: np.random.Generator, cls: np.ndarray, nc: int) -> np.ndarray:
    """2*randn - 4, +U[6,9] on the predicted class, lifted above the runner-up when needed so that the
    reference's `Sigmoid` assert (`ood_utils.py:1442`, cls == argmax) holds."""
    z = 2.0 * rng.standard_normal((len(cls), nc)) - 4.0
    r, c = np.arange(len(cls)), cls.astype(int)
    own = z[r, c] + rng.uniform(6, 9, size=len(cls))
    z[r, c] = -np.inf
    z[r, c] = np.maximum(own, z.max(axis=1) + 0.25) if len(cls) else own   # predicted class is the arg-max
    return z.astype(F32)


def detections(seed: int, batch: int, img: int, nc: int, lam: float, fixed: bool = False, max_det: int = 300):
    """-> per-image lists: boxes [M,4], cls [M], strides [M], conf [M], logits [M,nc] (all float32)."""
    rng = np.random.default_rng(seed)
    out = {"boxes": [], "cls": [], "strides": [], "conf": [], "logits": []}
    for _ in range(batch):
        m = int(lam) if fixed else int(min(max_det, rng.poisson(lam)))
        b, c, s, cf = boxes_for_image(rng, img, nc, m)
        out["boxes"].append(b)
        out["cls"].append(c)
        out["strides"].append(s)
        out["conf"].append(cf)
        out["logits"].append(logits_for(rng, c, nc))
    return out


def head_output(seed=91, bs=4, nc=20, img=640):
    """Detector-head-shaped tensors: pred [bs, 4 + nc, A] (cx, cy, w, h, class confidences), raw logits [bs, nc, A], the stride of
    every anchor [A]; a few objects per image with clusters of overlapping anchors around them (what NMS is for).  The inputs of
    golden_nms.npz / golden_postprocess.npz and of the bench's from-the-head measurement."""
    rng = np.random.default_rng(seed)
    strides = np.concatenate([np.full((img // s) ** 2, s, F32) for s in (8, 16, 32)])
    A = len(strides)
    pred = np.zeros((bs, 4 + nc, A), F32)
    logits = (2.0 * rng.standard_normal((bs, nc, A)) - 7.0).astype(F32)
    for b in range(bs):
        n_obj = [6, 0, 14, 30][b % 4]
        centres = rng.uniform(40, img - 40, (max(n_obj, 1), 2))
        sizes = rng.uniform(30, 260, (max(n_obj, 1), 2))
        obj = rng.integers(0, max(n_obj, 1), A)
        pred[b, 0:2] = (centres[obj] + rng.normal(0, 6, (A, 2))).T
        pred[b, 2:4] = (sizes[obj] * rng.uniform(0.85, 1.15, (A, 2))).T
        if n_obj:
            hot = rng.uniform(size=A) < 0.04
            cls_obj = rng.integers(0, nc, n_obj)
            logits[b, cls_obj[obj[hot]], np.nonzero(hot)[0]] += rng.uniform(6, 11, int(hot.sum())).astype(F32)
        pred[b, 4:] = 1.0 / (1.0 + np.exp(-logits[b].astype(np.float64)))
    return pred, logits, strides


def blob_vectors(seed: int, n: int, dim: int, k: int, sep: float = 6.0, unit_norm: bool = True):
    """Fit vectors for the k-means configs: mixture of k Gaussians; `sep` = centre spread in
    units of the within-cluster sigma*sqrt(dim) (large = well separated -> strict convergence)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((k, dim)) * (sep / np.sqrt(dim)) + 1.0
    lab = rng.integers(0, k, size=n)
    x = centres[lab] + rng.standard_normal((n, dim)) / np.sqrt(dim)
    if unit_norm:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(F32), lab
