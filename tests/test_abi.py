"""CPU-only: the C-ABI library builds, loads and exports every symbol include/oodb200.h declares."""
import ctypes
import os

from ood_in_object_detection_b200 import _lib, build


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _lib.declared_symbols()
    assert "oodb200_fmap_score_f32" in declared and "oodb200_logit_score_f32" in declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/oodb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table out of sync with the header"


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.oodb200_abi_version() == 2
    # argument validation happens on the host, before any launch: no GPU needed
    rc = lib.oodb200_fuse_u8(None, None, None, -1, 0, None, None)
    assert rc == -1 and b"negative" in lib.oodb200_last_error()
    rc = lib.oodb200_logit_score_f32(None, None, 4, 20, 0, 1.0, 1000.0, None, None, None, 1, None, None, None, None, None)
    assert rc == -1 and b"method_mask" in lib.oodb200_last_error()


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (oracle/__init__.py)."""
    root = os.path.dirname(build.HERE)
    for dirpath, _, files in os.walk(build.HERE):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    assert os.path.isdir(os.path.join(root, "oracle"))


def test_product_path_fails_loudly_without_a_cuda_device():
    """No CPU fallback anywhere on the product path: with host tensors (or no CUDA device at all) the entry points raise."""
    import numpy as np
    import pytest
    import torch
    from ood_in_object_detection_b200 import nms, ops
    from ood_in_object_detection_b200.postprocess import postprocess
    from tests.helpers import fake_predictor
    pred = torch.zeros((2, 4 + 3, 100))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nms.non_max_suppression(pred, 0.25, 0.45)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        postprocess(fake_predictor("all_ftmaps", False, 0.25), ((pred,), [torch.zeros((2, 8, 4, 4))] * 3),
                    torch.zeros((2, 3, 32, 32)), torch.zeros((2, 3, 32, 32)))
    with pytest.raises(TypeError):
        ops.PairDistances(torch.zeros((4, 8)), "l2")                      # host rows
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ops.default_device()
        with pytest.raises(RuntimeError):
            ops.match_boxes([np.zeros((1, 4), np.float32)], [np.zeros(1, np.int32)], [np.zeros((1, 4), np.float32)],
                            [np.zeros(1, np.int32)], 0.5)
