"""ctypes front-ends for the two CPU oracles — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``path_tracer_ai_b200``) never does.

* :class:`PortOracle`  — ``oracle/libpt_oracle.so`` built from ``oracle/pt_oracle.cpp`` (the in-repo CPU
  restatement; can be rebuilt anywhere g++ exists).
* :class:`RefOracle`   — ``oracle/_ref/libref_oracle.so``: the reference's own unmodified headers compiled
  against the glm shim (needs ``/root/reference`` to BUILD; the built ``.so`` travels with the snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libpt_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")

# Reference defaults: src/main.cpp:46-51
DEFAULT_CAM = dict(pos=(0.0, 2.0, 5.0), target=(0.0, 1.8, 0.0), up=(0.0, 1.0, 0.0), fov=45.0)


def build(which: str = "all", quiet: bool = True) -> None:
    """Runs oracle/Makefile (``port``, ``ref`` or ``all``).  ``ref`` is a no-op without /root/reference."""
    subprocess.run(["make", "-C", _HERE, which], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _opt(arr, dtype):
    if arr is None:
        return None
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pack_materials(materials) -> np.ndarray:
    """materials: iterable of (type, (r,g,b), roughness, metallic, ior) -> (n, 8) float32."""
    out = np.zeros((len(materials), 8), dtype=np.float32)
    for i, (ty, alb, rough, metal, ior) in enumerate(materials):
        out[i] = (ty, alb[0], alb[1], alb[2], rough, metal, ior, 0.0)
    return out


class PortOracle:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(PORT_SO):
                build("port")
            L = C.CDLL(PORT_SO)
            L.pto_create.restype = C.c_void_p
            L.pto_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            L.pto_free.argtypes = [C.c_void_p]
            L.pto_ntri.argtypes = [C.c_void_p]
            L.pto_nnodes.argtypes = [C.c_void_p]
            L.pto_get_order.argtypes = [C.c_void_p, i32p]
            L.pto_get_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.pto_get_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.pto_trace_closest.argtypes = [C.c_void_p, f32p, f32p, C.c_void_p, C.c_int64, i32p, C.c_void_p, C.c_void_p, C.c_int]
            L.pto_trace_any.argtypes = [C.c_void_p, f32p, f32p, C.c_void_p, C.c_int64, u8p, C.c_int]
            L.pto_philox.argtypes = [u32p, u32p, u32p]
            L.pto_camera_from_lookat.argtypes = [f32p, f32p, f32p, C.c_float, f32p]
            L.pto_camera_rays.argtypes = [f32p, f32p, C.c_int64, f32p]
            L.pto_render.restype = C.c_double
            L.pto_render.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                     C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_void_p, C.c_int]
            L.pto_omp_max_threads.restype = C.c_int
            cls._lib = L
        return cls._lib

    def __init__(self, pos, nrm=None, mat=None, materials8=None, lights7=None):
        L = self.lib()
        pos = _f32(pos).reshape(-1, 9)
        self.ntri = pos.shape[0]
        nrm = _opt(nrm, np.float32)
        mat = _opt(mat, np.int32)
        m8 = np.zeros((0, 8), np.float32) if materials8 is None else _f32(materials8).reshape(-1, 8)
        if lights7 is None:
            l7, nl = None, -1
        else:
            l7 = _f32(lights7).reshape(-1, 7)
            nl = l7.shape[0]
        self.h = L.pto_create(_ptr(pos), _ptr(nrm), _ptr(mat), self.ntri, _ptr(m8), m8.shape[0], _ptr(l7), nl)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().pto_free(self.h)
            self.h = None

    def order(self):
        out = np.empty(self.ntri, np.int32)
        self.lib().pto_get_order(self.h, out)
        return out

    def triangles(self):
        pos = np.empty((self.ntri, 9), np.float32)
        nrm = np.empty((self.ntri, 9), np.float32)
        mat = np.empty(self.ntri, np.int32)
        self.lib().pto_get_triangles(self.h, _ptr(pos), _ptr(nrm), _ptr(mat))
        return pos, nrm, mat

    def nodes(self):
        n = self.lib().pto_nnodes(self.h)
        boxes = np.empty((n, 6), np.float32)
        ranges = np.empty((n, 4), np.int32)
        self.lib().pto_get_nodes(self.h, _ptr(boxes), _ptr(ranges))
        return boxes, ranges

    def trace_closest(self, o, d, tmax=None, nthreads=0):
        o = _f32(o).reshape(-1, 3)
        d = _f32(d).reshape(-1, 3)
        n = o.shape[0]
        tm = _opt(tmax, np.float32)
        tri = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        uv = np.empty((n, 2), np.float32)
        self.lib().pto_trace_closest(self.h, o, d, _ptr(tm), n, tri, _ptr(t), _ptr(uv), nthreads)
        return tri, t, uv

    def trace_any(self, o, d, tmax=None, nthreads=0):
        o = _f32(o).reshape(-1, 3)
        d = _f32(d).reshape(-1, 3)
        n = o.shape[0]
        tm = _opt(tmax, np.float32)
        occ = np.empty(n, np.uint8)
        self.lib().pto_trace_any(self.h, o, d, _ptr(tm), n, occ, nthreads)
        return occ

    @classmethod
    def philox(cls, ctr4, key2):
        out = np.empty(4, np.uint32)
        cls.lib().pto_philox(np.asarray(ctr4, np.uint32), np.asarray(key2, np.uint32), out)
        return out

    @classmethod
    def camera(cls, pos=DEFAULT_CAM["pos"], target=DEFAULT_CAM["target"], up=DEFAULT_CAM["up"], fov=DEFAULT_CAM["fov"]):
        cam = np.empty(13, np.float32)
        cls.lib().pto_camera_from_lookat(_f32(pos), _f32(target), _f32(up), fov, cam)
        return cam

    @classmethod
    def camera_rays(cls, cam13, uv):
        uv = _f32(uv).reshape(-1, 2)
        out = np.empty((uv.shape[0], 6), np.float32)
        cls.lib().pto_camera_rays(_f32(cam13), uv, uv.shape[0], out)
        return out

    def render(self, cam13, width, height, spp, bounces, seed=1234, window=None, nthreads=0):
        """Returns (fb[H,W,3] float32, seconds, (extend_rays, shadow_rays))."""
        fb = np.zeros((height, width, 3), np.float32)
        x0, y0, x1, y1 = window if window is not None else (0, 0, width, height)
        stats = np.zeros(2, np.int64)
        secs = self.lib().pto_render(self.h, _f32(cam13), width, height, spp, bounces, seed, x0, y0, x1, y1,
                                     fb.reshape(-1), _ptr(stats), nthreads)
        return fb, secs, (int(stats[0]), int(stats[1]))

    @classmethod
    def max_threads(cls):
        return cls.lib().pto_omp_max_threads()


def ref_available() -> bool:
    return os.path.exists(REF_SO)


class RefOracle:
    """The reference's own code (unmodified headers) behind oracle/ref_harness.cpp."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                build("ref")
            if not os.path.exists(REF_SO):
                raise RuntimeError("oracle/_ref/libref_oracle.so is missing and /root/reference is not present to build it")
            L = C.CDLL(REF_SO)
            L.ref_scene_from_arrays.restype = C.c_void_p
            L.ref_scene_from_arrays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            L.ref_scene_from_obj.restype = C.c_void_p
            L.ref_scene_from_obj.argtypes = [C.c_char_p]
            L.ref_scene_free.argtypes = [C.c_void_p]
            L.ref_scene_ntri.argtypes = [C.c_void_p]
            L.ref_scene_nmat.argtypes = [C.c_void_p]
            L.ref_scene_get_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_scene_get_materials.argtypes = [C.c_void_p, f32p]
            L.ref_scene_get_order.argtypes = [C.c_void_p, i32p]
            L.ref_scene_id_bvh_matches.argtypes = [C.c_void_p]
            L.ref_trace_closest.argtypes = [C.c_void_p, f32p, f32p, C.c_void_p, C.c_int64, i32p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            L.ref_render.restype = C.c_double
            L.ref_render.argtypes = [C.c_void_p, f32p, f32p, f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int]
            L.ref_render_window.restype = C.c_double
            L.ref_render_window.argtypes = [C.c_void_p, f32p, f32p, f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_int]
            L.ref_camera_ray.argtypes = [f32p, f32p, f32p, C.c_float, f32p, C.c_int64, f32p]
            L.ref_camera_basis.argtypes = [f32p, f32p, f32p, C.c_float, f32p]
            L.ref_tree_stats.argtypes = [C.c_void_p, i64p]
            L.ref_omp_max_threads.restype = C.c_int
            cls._lib = L
        return cls._lib

    def __init__(self, pos=None, nrm=None, mat=None, materials8=None, obj_path=None):
        L = self.lib()
        if obj_path is not None:
            self.h = L.ref_scene_from_obj(os.fsencode(obj_path))
            if not self.h:
                raise RuntimeError(f"reference loader failed on {obj_path}")
        else:
            pos = _f32(pos).reshape(-1, 9)
            nrm = _opt(nrm, np.float32)
            mat = _opt(mat, np.int32)
            m8 = np.zeros((0, 8), np.float32) if materials8 is None else _f32(materials8).reshape(-1, 8)
            self.h = L.ref_scene_from_arrays(_ptr(pos), _ptr(nrm), _ptr(mat), pos.shape[0], _ptr(m8), m8.shape[0])
        self.ntri = L.ref_scene_ntri(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().ref_scene_free(self.h)
            self.h = None

    def order(self):
        out = np.empty(self.ntri, np.int32)
        self.lib().ref_scene_get_order(self.h, out)
        return out

    def triangles(self):
        pos = np.empty((self.ntri, 9), np.float32)
        nrm = np.empty((self.ntri, 9), np.float32)
        mat = np.empty(self.ntri, np.int32)
        self.lib().ref_scene_get_triangles(self.h, _ptr(pos), _ptr(nrm), _ptr(mat))
        return pos, nrm, mat

    def materials(self):
        n = self.lib().ref_scene_nmat(self.h)
        out = np.empty((n, 8), np.float32)
        self.lib().ref_scene_get_materials(self.h, out)
        return out

    def id_bvh_matches(self):
        return bool(self.lib().ref_scene_id_bvh_matches(self.h))

    def tree_stats(self):
        out = np.zeros(4, np.int64)
        self.lib().ref_tree_stats(self.h, out)
        return dict(nodes=int(out[0]), leaves=int(out[1]), depth=int(out[2]), flat=int(out[3]))

    def trace_closest(self, o, d, tmax=None, nthreads=0, want_geom=False):
        o = _f32(o).reshape(-1, 3)
        d = _f32(d).reshape(-1, 3)
        n = o.shape[0]
        tm = _opt(tmax, np.float32)
        tri = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        pos = np.empty((n, 3), np.float32) if want_geom else None
        nrm = np.empty((n, 3), np.float32) if want_geom else None
        self.lib().ref_trace_closest(self.h, o, d, _ptr(tm), n, tri, _ptr(t), _ptr(pos), _ptr(nrm), 0, nthreads)
        return (tri, t, pos, nrm) if want_geom else (tri, t)

    def render(self, width, height, spp, bounces, cam=DEFAULT_CAM, nthreads=0):
        fb = np.zeros((height, width, 3), np.float32)
        secs = self.lib().ref_render(self.h, _f32(cam["pos"]), _f32(cam["target"]), _f32(cam["up"]), cam["fov"],
                                     width, height, spp, bounces, _ptr(fb), nthreads)
        return fb, secs

    def render_window(self, width, height, window, spp, bounces, cam=DEFAULT_CAM, nthreads=0):
        """The reference's per-sample code (camera ray + tracePath) on the pixels of window = (x0, y0, x1, y1) of a
        width x height frame, `spp` samples per pixel split over independent Renderer objects — one per thread, so the
        member RNG that Renderer::render races under OpenMP is private to each.  Returns (fb[y1-y0, x1-x0, 3], seconds)."""
        x0, y0, x1, y1 = window
        fb = np.zeros((y1 - y0, x1 - x0, 3), np.float32)
        secs = self.lib().ref_render_window(self.h, _f32(cam["pos"]), _f32(cam["target"]), _f32(cam["up"]), cam["fov"],
                                            width, height, x0, y0, x1, y1, spp, bounces, _ptr(fb), nthreads)
        return fb, secs

    @classmethod
    def camera_rays(cls, uv, cam=DEFAULT_CAM):
        uv = _f32(uv).reshape(-1, 2)
        out = np.empty((uv.shape[0], 6), np.float32)
        cls.lib().ref_camera_ray(_f32(cam["pos"]), _f32(cam["target"]), _f32(cam["up"]), cam["fov"], uv, uv.shape[0], out)
        return out

    @classmethod
    def camera_basis(cls, cam=DEFAULT_CAM):
        out = np.empty(12, np.float32)
        cls.lib().ref_camera_basis(_f32(cam["pos"]), _f32(cam["target"]), _f32(cam["up"]), cam["fov"], out)
        return out

    @classmethod
    def max_threads(cls):
        return cls.lib().ref_omp_max_threads()
