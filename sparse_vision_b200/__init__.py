"""sparse_vision_b200 — B200-native (sm_100a) implementation of sparse-vision's SAE training-and-attribution hot path.

Python modules mirror the reference's layout (models/, losses/, utils.py, model_pipeline.py, compute_ie.py) and call
hand-written CUDA through the C ABI in include/svb.h (libsvb.so, bound with ctypes in _lib.py).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
