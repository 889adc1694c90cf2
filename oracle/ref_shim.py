"""Import shim for the UNMODIFIED reference.  Test / measurement infrastructure.

Source of the modules: /root/reference where it exists (the build container), else oracle/_ref, the byte-compiled build
of the same files that `oracle/make_ref.py` produces in the build container and that travels to the GPU box with the
snapshot (git-ignored, like a compiled .so).  Used by `tests/golden/make_golden.py` (to freeze golden vectors),
`tests/test_live_reference.py`, `tests/test_gpu_pipeline.py` (the reference's evaluation driver over this repo's
classes) and by the reference arm / `cpu_baseline` leg of `bench.py` (the reference's own per-box code timed on the
host cores).  Nothing under ood_in_object_detection_b200/ imports it.

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

_HERE = os.path.dirname(os.path.abspath(__file__))
_COMPILED = os.path.join(_HERE, "_ref")


def _root() -> str:
    forced = os.environ.get("OODB200_REF_ROOT")                  # tests: force the compiled build
    if forced:
        return forced
    return "/root/reference" if os.path.isfile("/root/reference/ood_utils.py") else _COMPILED


REFERENCE_ROOT = _root()
_MISSING = ("matplotlib", "hdbscan", "skimage", "seaborn", "tap", "natsort", "umap", "ivis",
            "openpyxl", "adjustText")
_loaded = None


def available() -> bool:
    return any(os.path.isfile(os.path.join(REFERENCE_ROOT, "ood_utils" + ext)) for ext in (".py", ".pyb"))


def kind() -> str:
    """"source" (/root/reference) or "compiled" (oracle/_ref)."""
    return "source" if os.path.isfile(os.path.join(REFERENCE_ROOT, "ood_utils.py")) else "compiled"


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


class _Compiled(importlib.abc.MetaPathFinder):
    """Finder for the byte-compiled build (oracle/make_ref.py): module a.b -> <root>/a/b.pyb or <root>/a/b/__init__.pyb,
    loaded by importlib's SourcelessFileLoader (`__file__` stays a real path: ultralytics reads its yaml defaults next to it)."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, name, path, target=None):
        import importlib.util
        base = os.path.join(self.root, *name.split("."))
        for cand, pkg in ((base + ".pyb", False), (os.path.join(base, "__init__.pyb"), True)):
            if os.path.isfile(cand):
                loader = importlib.machinery.SourcelessFileLoader(name, cand)
                return importlib.util.spec_from_file_location(name, cand, loader=loader,
                                                              submodule_search_locations=[base] if pkg else None)
        return None


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
    if kind() == "compiled":
        sys.meta_path.insert(0, _Compiled(REFERENCE_ROOT))
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
    assert os.path.dirname(os.path.abspath(ood_utils.__file__)) == os.path.abspath(REFERENCE_ROOT), \
        f"`ood_utils` resolved to {ood_utils.__file__}, not to the reference under {REFERENCE_ROOT}"
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
