"""perception_b200 — B200-native drop-in for the cuboid_detection / object_detection point-cloud hot path.

Only what the path needs lives here: ``csrc/`` (sm_100a CUDA kernels + the C ABI of libcuboid_cuda.so,
declared in ``include/cuboid_cuda.h``), a ctypes host mirror of that ABI (``api``), the template/PCD
helpers (``pcd``) and the seeded synthetic D435 frame generator (``synth``) that stands in for the bag
files the reference does not ship. There is no CPU fallback: ``api.load()`` raises if the CUDA library
has not been built, and ``cuboid_create`` fails without a CUDA device.
"""
from . import build as build  # noqa: F401

__all__ = ["api", "pcd", "synth", "build", "params"]
