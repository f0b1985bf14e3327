"""strolle_b200 — B200-native implementation of Strolle's per-pixel GI hot path.

BVH traversal + ray/triangle intersection, ReSTIR DI / GI temporal + spatial resampling and SVGF
(temporal accumulation + à-trous) as hand-written sm_100a CUDA kernels behind a C ABI
(include/strolle_b200.h).  `strolle_b200.Engine` is a thin ctypes mirror of `strolle::Engine`.
There is no CPU fallback: without the built library or without a CUDA device, construction fails.
"""
from .engine import Engine, MultiEngine, StrolleError, lib_path, load_library, PASS_NAMES  # noqa: F401
from . import scenes  # noqa: F401
