"""strolle_b200 — B200-native implementation of Strolle's per-pixel GI hot path.

BVH traversal + ray/triangle intersection, ReSTIR DI / GI temporal + spatial resampling and SVGF
(temporal accumulation + à-trous) as hand-written sm_100a CUDA kernels behind a C ABI
(include/strolle_b200.h).  `strolle_b200.Engine` is a thin ctypes mirror of `strolle::Engine`.
There is no CPU fallback: without the built library or without a CUDA device, construction fails.
"""
import os as _os

# The strip transport lets a stream spin (k_strip_wait) until a flag is raised — for the copy-engine halo pushes, by a kernel on another
# stream of the same device.  Streams that share a hardware work queue serialise behind each other, and CUDA maps streams onto only 8
# queues by default: give every stream its own (takes effect if CUDA is not initialised yet; a host that initialises CUDA first sets it itself).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .engine import Engine, MultiEngine, StrolleError, lib_path, load_library, PASS_NAMES  # noqa: F401
from . import scenes  # noqa: F401
