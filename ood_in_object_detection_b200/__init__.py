"""B200-native OoD scoring for YOLO detections: the hot path of
aitor-martinez-seras/OoD_in_Object_Detection rebuilt as sm_100a CUDA kernels behind the
reference's `ood_utils.py` class surface.  See DESIGN.md / INTEGRATION.md.

The CUDA library (liboodb200.so) is loaded lazily by `_lib.load()`; importing the package
does not need a GPU, calling any scoring function does (there is no CPU fallback).
"""
__version__ = "0.1.0"
