"""path_tracer_ai_b200 — B200-native wavefront path tracer behind the reference's renderer interface.

Everything that computes runs in ``libb2pt.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/b2pt.h``); this package is the thin Python face used by tests and ``bench.py``:

* :mod:`._capi`     ctypes binding of the C ABI (:class:`Engine`)
* :mod:`.renderer`  ``B200Renderer`` / ``Scene`` / ``Camera`` mirrors of the reference's C++ classes
* :mod:`.scenes`    procedural scenes for the BASELINE configurations
* :mod:`.build`     in-tree nvcc build

There is no CPU fallback: importing works without a GPU (so CPU-only tests can check the ABI), but creating
an :class:`Engine` without an sm_100 device raises.
"""
from ._capi import (B2ptError, Engine, FLAG_COUNT_FETCHES, FLAG_EXACT_ONLY, FLAG_LANE_KERNELS, FLAG_NO_LEARN_ORDER, FLAG_NO_SORT, FLAG_POOL_EXTEND, LIB_PATH, REFERENCE_LIGHTS,  # noqa: F401
                    camera_from_cam13, load_library, make_camera, reference_order)
from .renderer import B200Renderer, Camera, Scene, Settings  # noqa: F401

__all__ = ["Engine", "B2ptError", "B200Renderer", "Scene", "Camera", "Settings", "reference_order", "load_library",
           "make_camera", "camera_from_cam13", "FLAG_COUNT_FETCHES", "FLAG_EXACT_ONLY", "FLAG_LANE_KERNELS", "FLAG_NO_LEARN_ORDER", "FLAG_NO_SORT", "FLAG_POOL_EXTEND", "REFERENCE_LIGHTS", "LIB_PATH"]
