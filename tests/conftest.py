import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built():
    """Builds libb2pt.so / the port oracle if they are stale (nvcc cross-compiles without a GPU)."""
    from path_tracer_ai_b200 import build as b
    import oracle
    b.build_lib()
    b.build_cli()
    oracle.build("port")
    if os.path.isdir("/root/reference/include"):
        oracle.build("ref")
    return True


@pytest.fixture(scope="session")
def engine(built):
    import path_tracer_ai_b200 as pt
    eng = pt.Engine()
    yield eng
    eng.close()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def prebuild_from_scene(sc):
    """Loader's pre-build triangle list (what an oracle constructor wants) from a product Scene."""
    inv = np.empty_like(sc.order)
    inv[sc.order] = np.arange(len(sc.order), dtype=np.int32)
    return sc.pos[inv], sc.nrm[inv], sc.mat[inv]


def cam13_of(cam):
    return np.concatenate([cam.getPosition(), cam.getForward(), cam.getRight(), cam.getUp(), [cam.getFOV()]]).astype(np.float32)


def reference_tonemap(fb, gamma):
    """Renderer::saveImage's pixel maths (reference src/renderer.cpp:8-17) with the function the reference calls:
    glm::pow(float) is std::pow(float, float), i.e. libm's powf (numpy's float32 power is a different implementation
    and differs from it in the last bit for ~18 % of inputs).  clamp -> powf(c, 1/gamma) -> (unsigned char)(c * 255)."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    v = np.ascontiguousarray(fb, np.float32)
    inv = float(np.float32(1.0) / np.float32(gamma))
    flat = np.clip(v.reshape(-1), 0, 1)
    out = np.empty(flat.shape[0], np.uint8)
    for i, x in enumerate(flat):
        c = np.float32(libm.powf(float(x), inv)) * np.float32(255.0)
        out[i] = int(c) if np.isfinite(c) else 0
    return out.reshape(v.shape)
