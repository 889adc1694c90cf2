"""ctypes binding of liboodb200.so (the C ABI declared in include/oodb200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or a
call returns an error code, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "oodb200.h")

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_L = C.c_int64

# name -> argtypes (restype is int unless listed in _RESTYPE); must mirror include/oodb200.h
SIGNATURES = {
    "oodb200_abi_version": [],
    "oodb200_last_error": [],
    "oodb200_roi_pool_f32": [_P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _L, _P],
    "oodb200_fmap_workspace_bytes": [_I, _I, _P],
    "oodb200_fmap_score_f32": [_P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P,
                               _P, _I, _P, _P, _P, _L, _P],
    "oodb200_q1_plan_i32": [_P, _P, _P, _I, _P, _P, _P],
    "oodb200_roi_pool_nhwc_f32": [_P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _L, _P],
    "oodb200_fmap_score_nhwc_f32": [_P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P,
                                    _P, _I, _P, _P, _P, _L, _P],
    "oodb200_match_boxes_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _F, _I, _P, _P, _P, _P, _P, _P],
    "oodb200_nms_workspace_bytes": [_I, _I],
    "oodb200_nms_payload_f32": [_P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _I, _I, _P, _P, _P, _P, _P, _P, _L, _P],
    "oodb200_logit_score_f32": [_P, _P, _I, _I, _I, _F, _F, _P, _P, _P, _I, _P, _P, _P, _P, _P],
    "oodb200_fuse_u8": [_P, _P, _P, _I, _I, _P, _P],
    "oodb200_fuse_score_f32": [_P, _P, _I, _P, _P],
    "oodb200_vec_score_f32": [_P, _L, _I, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "oodb200_dist_indness_f32": [_P, _P, _L, _P, _P, _P, _I, _P, _P],
    "oodb200_normalize_rows_f32": [_P, _L, _I, _L, _P, _L, _P],
    "oodb200_radix_hist_u32": [_P, _P, _I, _L, _P, _I, _I, _P, _P, _P],
    "oodb200_kmeans_smem_bytes": [_I, _I],
    "oodb200_kmeans_step_f32": [_P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _P],
    "oodb200_kmeans_reduce_f32": [_P, _P, _I, _L, _P, _P],
    "oodb200_kmeans_reduce_step_f32": [_P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P],
    "oodb200_kmeans_update_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "oodb200_kmeans_update_peers_f32": [_P, _I, _L, _L, _P, _I, C.c_uint32, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "oodb200_kmeans_converge_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
    "oodb200_sqdist_cand_f32": [_P, _I, _P, _I, _L, _P, _I, _P, _P, _P, _P],
    "oodb200_vec_score_tc_workspace_bytes": [_I, _L, _I],
    "oodb200_vec_score_tc_f32": [_P, _L, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "oodb200_segment_scratch_doubles": [_I, _I],
    "oodb200_segment_colsum_f64": [_P, _I, _P, _I, _P, _P, _P],
    "oodb200_segment_center_f32": [_P, _I, _P, _I, _P, _P, _P, _P, _P],
    "oodb200_kmeans_tc_workspace_bytes": [_I, _I, _I],
    "oodb200_kmeans_step_tc_f32": [_P, _L, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _P, _P],
    "oodb200_seed_grid": [_L],
    "oodb200_seed_sqdist_f32": [_P, _I, _P, _I, _L, _P, _I, _P, _P, _P, _P, _P],
    "oodb200_seed_scan_f32": [_P, _P, _P, _I, _I, _P, _P, _P, _P, _I, _P, _L, _P, _P],
    "oodb200_seed_gather_f32": [_P, _I, _P, _I, _I, _P, _P, _P, _P],
    "oodb200_seed_pick_f32": [_P, _P, _P, _I, _I, _P, _L, _P, _P, _I, _P, _P, _P, _L, _P, _P],
    "oodb200_pair_cluster_sums_f32": [_P, _I, _I, _L, _P, _I, _I, _P, _P],
    "oodb200_pair_dist_matrix_f32": [_P, _I, _I, _L, _I, _P, _L, _P],
    "oodb200_matrix_cluster_sums_f32": [_P, _I, _L, _P, _P, _I, _P, _P],
}
_RESTYPE = {"oodb200_last_error": C.c_char_p, "oodb200_fmap_workspace_bytes": C.c_int64,
            "oodb200_kmeans_smem_bytes": C.c_int64, "oodb200_kmeans_tc_workspace_bytes": C.c_int64,
            "oodb200_segment_scratch_doubles": C.c_int64, "oodb200_vec_score_tc_workspace_bytes": C.c_int64, "oodb200_nms_workspace_bytes": C.c_int64}

_lib = None


def declared_symbols() -> list:
    """Function names declared in include/oodb200.h."""
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(oodb200_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load (building first if needed) the shared library; raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("OODB200_LIB") or _build.build()     # OODB200_LIB: a pre-built variant (tuning runs)
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    if lib.oodb200_abi_version() != 2:
        raise RuntimeError("liboodb200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().oodb200_last_error()
        raise RuntimeError(f"{what} failed ({code}): {msg.decode() if msg else ''}")
