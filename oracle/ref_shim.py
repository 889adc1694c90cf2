"""Import shim for the UNMODIFIED reference at /root/reference.  Test infrastructure.

Only usable where /root/reference exists (the build container).  It never
travels to the GPU box: nothing in `-m gpu` tests, `smoke()` or `bench.py`
calls `load()`.  Used by `tests/golden/make_golden.py` (to freeze golden
vectors) and by the live-reference half of `tests/test_oracle_vs_golden.py`.

Why a shim is needed:
  * `ood_utils.py:26-37` pulls in matplotlib / hdbscan / skimage / tap / ...
    which are not installed -> stub them with MagicMock packages;
  * `custom_hyperparams.py:30-33,126-138` uses dataclass instances as field
    defaults, a ValueError on Python >= 3.11 -> wrap `dataclasses.dataclass`
    with `unsafe_hash=True` while that one module is imported.
Nothing under /root/reference is modified or copied.
"""
from __future__ import annotations

import dataclasses
import importlib.abc
import importlib.machinery
import os
import sys
import warnings
from types import SimpleNamespace
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"
_MISSING = ("matplotlib", "hdbscan", "skimage", "seaborn", "tap", "natsort", "umap", "ivis",
            "openpyxl", "adjustText")
_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ood_utils.py"))


class _Stub(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        m.__name__ = spec.name
        return m

    def exec_module(self, module):
        pass


def load() -> SimpleNamespace:
    """Import the reference modules; returns a namespace with the pieces the tests use."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present on this machine")
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp")
    warnings.filterwarnings("ignore")
    sys.meta_path.insert(0, _Stub())
    # our own package has modules called ood_utils / cluster_utils too, but they live inside
    # the package namespace, so the top-level names resolve to the reference here.
    sys.path.insert(0, REFERENCE_ROOT)
    _dc = dataclasses.dataclass

    def _patched(cls=None, **kw):
        kw = {"unsafe_hash": True, **kw}
        if cls is not None:
            return _dc(cls, **kw)
        return lambda c: _dc(c, **kw)

    dataclasses.dataclass = _patched
    try:
        import custom_hyperparams  # noqa: F401
    finally:
        dataclasses.dataclass = _dc
    import cluster_utils
    import ood_utils
    from ultralytics.engine.results import Results
    from ultralytics.models.yolo.detect.predict import extract_roi_aligned_features_from_correct_stride

    _loaded = SimpleNamespace(
        ood_utils=ood_utils,
        cluster_utils=cluster_utils,
        custom_hyperparams=custom_hyperparams,
        Results=Results,
        extract_roi_aligned_features_from_correct_stride=extract_roi_aligned_features_from_correct_stride,
    )
    return _loaded


DIST_KW = dict(agg_method="mean", cluster_method="one", cluster_optimization_metric="silhouette",
               ind_info_creation_option="valid_preds_one_stride", which_internal_activations="ftmaps_and_strides",
               iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)
LOGIT_KW = dict(per_class=True, per_stride=False, iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15,
                min_conf_threshold_test=0.15, use_values_before_sigmoid=True)


def make_results(ref, maps_per_image, boxes6, strides=None, logits=None, batch_hw=(640, 640), n_batch=1):
    """Build the reference's `Results` objects for synthetic inputs.

    maps_per_image: list (per image) of 3 CHW float32 torch tensors, or None for logits methods.
    boxes6: list of [M,6] tensors (xyxy, conf, cls).  strides: list of [M] float tensors.
    `orig_img` is the whole uint8 batch [B,H,W,3] as in `predict.py:342-354`, so that
    `orig_img.shape[1:3] == (H, W)` (`ood_utils.py:2061`).
    """
    import numpy as np
    orig = np.zeros((n_batch, batch_hw[0], batch_hw[1], 3), np.uint8)
    out = []
    for i, b in enumerate(boxes6):
        if logits is not None:
            extra = logits[i]
        else:
            extra = (list(maps_per_image[i]), strides[i])
        out.append(ref.Results(orig_img=orig, path="synthetic", names={k: str(k) for k in range(80)},
                               boxes=b, extra_item=extra))
    return out
