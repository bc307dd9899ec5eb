"""depthhead_b200 — B200-native (sm_100a) Hough-forest head-pose prediction path of depthhead.

The product is ``libdepthhead_cuda.so`` (hand-written CUDA behind the C ABI in
``include/depthhead_cuda.h``); this package holds its build script, a ctypes binding, the
host-side mirror of the reference API and the synthetic-input generators.
"""
from .api import Context, HoughPrediction, IntrinsicMatrix, PredictionResult, default_context  # noqa: F401
from .capi import DhError  # noqa: F401

__all__ = ["Context", "HoughPrediction", "IntrinsicMatrix", "PredictionResult", "default_context", "DhError"]
