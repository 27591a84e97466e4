"""ctypes binding of include/b2pt.h — every call goes through the C ABI of libb2pt.so.

There is deliberately no fallback of any kind: if the library is missing or no sm_100 GPU is present
the constructor raises (the reference's CPU fallback, src/main.cpp:98-113, is removed by design).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2PT_LIB") or os.path.join(_HERE, "libb2pt.so")   # B2PT_LIB: an experiment build of the same library

DIFFUSE, SPECULAR, DIELECTRIC = 0, 1, 2
FLAG_COUNT_FETCHES = 1
FLAG_EXACT_ONLY = 2
FLAG_LANE_KERNELS = 4
FLAG_NO_LEARN_ORDER = 8
FLAG_POOL_EXTEND = 16
FLAG_NO_SORT = 32


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("albedo", C.c_float * 3), ("roughness", C.c_float),
                ("metallic", C.c_float), ("ior", C.c_float), ("_pad", C.c_float)]


class Light(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3), ("intensity", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("forward", C.c_float * 3), ("right", C.c_float * 3),
                ("up", C.c_float * 3), ("fov", C.c_float)]


class Settings(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("max_bounces", C.c_int32), ("gamma", C.c_float)]


class Partition(C.Structure):
    _fields_ = [("tile_rank", C.c_int32), ("tile_world", C.c_int32), ("tile_size", C.c_int32),
                ("sample_begin", C.c_int32), ("sample_count", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_int32), ("max_paths_in_flight", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("extend_rays", C.c_int64), ("shadow_rays", C.c_int64), ("samples", C.c_int64),
                ("fallback_rays", C.c_int64), ("node_fetches", C.c_int64), ("tri_fetches", C.c_int64),
                ("kernel_launches", C.c_int64), ("gpu_seconds", C.c_double), ("trace_seconds", C.c_double),
                ("build_seconds", C.c_double), ("extend_seconds", C.c_double), ("shadow_seconds", C.c_double),
                ("extend_launches", C.c_int64), ("shadow_launches", C.c_int64), ("order_seconds", C.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


EXPORTS = [
    "b2pt_create", "b2pt_destroy", "b2pt_last_error", "b2pt_reference_order", "b2pt_upload_scene",
    "b2pt_trace_closest", "b2pt_trace_any", "b2pt_trace_closest_device", "b2pt_trace_any_device",
    "b2pt_render", "b2pt_render_device", "b2pt_tonemap", "b2pt_tonemap_last", "b2pt_tonemap_thresholds",
    "b2pt_progressive_begin", "b2pt_progressive_pass", "b2pt_get_stats", "b2pt_get_accel_info",
    "b2pt_stream", "b2pt_version",
    "b2pt_multi_create", "b2pt_multi_destroy", "b2pt_multi_last_error", "b2pt_multi_device_count", "b2pt_multi_ctx",
    "b2pt_multi_upload_scene", "b2pt_multi_render", "b2pt_multi_tonemap_last", "b2pt_multi_get_stats",
]

_lib = None


class B2ptError(RuntimeError):
    pass


def load_library():
    """Loads libb2pt.so (no GPU needed for loading) and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m path_tracer_ai_b200.build` "
                          "(there is no CPU or pure-Python fallback for the GPU path)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.b2pt_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.b2pt_destroy.argtypes = [vp]
    L.b2pt_destroy.restype = None
    L.b2pt_last_error.argtypes = [vp]
    L.b2pt_last_error.restype = C.c_char_p
    L.b2pt_reference_order.argtypes = [vp, i64, vp]
    L.b2pt_upload_scene.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, i32]
    L.b2pt_trace_closest.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    L.b2pt_trace_any.argtypes = [vp, vp, vp, vp, i64, vp]
    L.b2pt_trace_closest_device.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    L.b2pt_trace_any_device.argtypes = [vp, vp, vp, vp, i64, vp]
    L.b2pt_render.argtypes = [vp, C.POINTER(Camera), C.POINTER(Settings), C.c_uint64, C.POINTER(Partition), vp]
    L.b2pt_render_device.argtypes = [vp, C.POINTER(Camera), C.POINTER(Settings), C.c_uint64, C.POINTER(Partition), vp]
    L.b2pt_tonemap.argtypes = [vp, vp, i32, i32, C.c_float, i32, vp]
    L.b2pt_tonemap_last.argtypes = [vp, C.c_float, i32, vp]
    L.b2pt_tonemap_thresholds.argtypes = [C.c_float, vp]
    L.b2pt_progressive_begin.argtypes = [vp, C.POINTER(Camera), C.POINTER(Settings), C.c_uint64]
    L.b2pt_progressive_pass.argtypes = [vp, i32, vp, C.POINTER(i32)]
    L.b2pt_multi_create.argtypes = [vp, i32, i32, i64, C.POINTER(vp)]
    L.b2pt_multi_destroy.argtypes = [vp]
    L.b2pt_multi_destroy.restype = None
    L.b2pt_multi_last_error.argtypes = [vp]
    L.b2pt_multi_last_error.restype = C.c_char_p
    L.b2pt_multi_device_count.argtypes = [vp]
    L.b2pt_multi_ctx.argtypes = [vp, i32]
    L.b2pt_multi_ctx.restype = vp
    L.b2pt_multi_upload_scene.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, i32]
    L.b2pt_multi_render.argtypes = [vp, C.POINTER(Camera), C.POINTER(Settings), C.c_uint64, vp]
    L.b2pt_multi_tonemap_last.argtypes = [vp, C.c_float, i32, vp]
    L.b2pt_multi_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b2pt_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b2pt_get_accel_info.argtypes = [vp, vp]
    L.b2pt_stream.argtypes = [vp]
    L.b2pt_stream.restype = vp
    L.b2pt_version.restype = C.c_char_p
    for name in EXPORTS:
        getattr(L, name)   # AttributeError if a declared symbol is not exported
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def reference_order(pos) -> np.ndarray:
    """Host: the permutation the reference's BVH::build applies (include/bvh.hpp:27-72)."""
    L = load_library()
    pos = _f32(pos).reshape(-1, 9)
    order = np.empty(pos.shape[0], np.int32)
    rc = L.b2pt_reference_order(_p(pos), pos.shape[0], _p(order))
    if rc != 0:
        raise B2ptError(f"b2pt_reference_order failed ({rc})")
    return order


def make_camera(position, forward, right, up, fov) -> Camera:
    cam = Camera()
    cam.position[:] = [float(x) for x in position]
    cam.forward[:] = [float(x) for x in forward]
    cam.right[:] = [float(x) for x in right]
    cam.up[:] = [float(x) for x in up]
    cam.fov = float(fov)
    return cam


def camera_from_cam13(cam13) -> Camera:
    c = np.asarray(cam13, np.float32)
    return make_camera(c[0:3], c[3:6], c[6:9], c[9:12], c[12])


REFERENCE_LIGHTS = [   # include/scene.hpp:55-80
    ((2.0, 3.5, 2.0), (1.0, 0.95, 0.8), 9.0),
    ((-1.5, 2.0, 1.5), (0.8, 0.9, 1.0), 2.0),
    ((0.0, 2.0, -2.0), (1.0, 1.0, 1.0), 1.0),
    ((0.0, 0.1, 0.0), (0.9, 0.9, 1.0), 2.0),
]


def check_scene_arrays(pos, nrm, mat, materials8, lights):
    """Shapes the C ABI takes on trust: pos/nrm (ntri, 9) float32, mat (ntri,) int32, materials8 (nmat, 8), <= 16 lights.
    Returns the converted arrays; raises ValueError instead of letting the library read past a short buffer."""
    pos = _f32(pos)
    if pos.size % 9:
        raise ValueError(f"pos has {pos.size} floats: not a whole number of triangles (9 floats each)")
    pos = pos.reshape(-1, 9)
    ntri = pos.shape[0]
    if nrm is not None:
        nrm = _f32(nrm)
        if nrm.size != pos.size:
            raise ValueError(f"nrm has {nrm.size} floats, pos has {pos.size}: one normal per vertex is required")
        nrm = nrm.reshape(-1, 9)
    if mat is not None:
        mat = np.ascontiguousarray(mat, np.int32).reshape(-1)
        if mat.shape[0] != ntri:
            raise ValueError(f"mat has {mat.shape[0]} entries for {ntri} triangles")
    m8 = np.zeros((0, 8), np.float32) if materials8 is None else _f32(materials8)
    if m8.size % 8:
        raise ValueError("materials8 must be (nmat, 8): type, r, g, b, roughness, metallic, ior, 0")
    m8 = m8.reshape(-1, 8)
    lights = list(lights or [])
    if len(lights) > 16:
        raise ValueError(f"{len(lights)} lights: the engine takes at most 16")
    for l in lights:
        if len(l) != 3 or len(l[0]) != 3 or len(l[1]) != 3:
            raise ValueError("a light is ((x, y, z), (r, g, b), intensity)")
    return pos, nrm, mat, m8, lights


def check_ray_arrays(o, d, tmax):
    """o, d: (n, 3) float32; tmax: (n,) or None."""
    o, d = _f32(o), _f32(d)
    if o.size % 3 or d.size != o.size:
        raise ValueError(f"ray origins ({o.size} floats) and directions ({d.size} floats) must both be (n, 3)")
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    tm = None
    if tmax is not None:
        tm = _f32(tmax).reshape(-1)
        if tm.shape[0] != o.shape[0]:
            raise ValueError(f"tmax has {tm.shape[0]} entries for {o.shape[0]} rays")
    return o, d, tm


class Engine:
    """One b2pt context (one GPU)."""

    def __init__(self, device: int = 0, flags: int = 0, max_paths: int = 0):
        self._L = load_library()
        self._h = C.c_void_p()
        cfg = Config(device, flags, max_paths)
        rc = self._L.b2pt_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise B2ptError(f"b2pt_create failed ({rc}): {self._L.b2pt_last_error(None).decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.b2pt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise B2ptError(f"{what} failed ({rc}): {self._L.b2pt_last_error(self._h).decode()}")

    # ---- scene
    def upload_scene(self, pos, nrm=None, mat=None, materials8=None, lights=REFERENCE_LIGHTS):
        """pos/nrm: (ntri, 9) in the reference's post-build order; materials8: (nmat, 8) rows of
        (type, r, g, b, roughness, metallic, ior, 0); lights: [(pos3, color3, intensity)]."""
        pos, nrm, mat, m8, lights = check_scene_arrays(pos, nrm, mat, materials8, lights)
        ntri = pos.shape[0]
        mats = (Material * max(len(m8), 1))()
        for i, row in enumerate(m8):
            mats[i].type = int(row[0])
            mats[i].albedo[:] = [float(row[1]), float(row[2]), float(row[3])]
            mats[i].roughness, mats[i].metallic, mats[i].ior = float(row[4]), float(row[5]), float(row[6])
        ls = (Light * max(len(lights), 1))()
        for i, (p, c, inten) in enumerate(lights):
            ls[i].position[:] = [float(x) for x in p]
            ls[i].color[:] = [float(x) for x in c]
            ls[i].intensity = float(inten)
        rc = self._L.b2pt_upload_scene(self._h, _p(pos), _p(nrm), _p(mat), ntri, C.cast(mats, C.c_void_p), len(m8),
                                       C.cast(ls, C.c_void_p), len(lights))
        self._check(rc, "b2pt_upload_scene")
        self.ntri = ntri

    # ---- queries (host buffers)
    def trace_closest(self, o, d, tmax=None, want_uv=True):
        o, d, tm = check_ray_arrays(o, d, tmax)
        n = o.shape[0]
        tri = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        uv = np.empty((n, 2), np.float32) if want_uv else None
        self._check(self._L.b2pt_trace_closest(self._h, _p(o), _p(d), _p(tm), n, _p(tri), _p(t), _p(uv)), "b2pt_trace_closest")
        return tri, t, uv

    def trace_any(self, o, d, tmax=None):
        o, d, tm = check_ray_arrays(o, d, tmax)
        n = o.shape[0]
        occ = np.empty(n, np.uint8)
        self._check(self._L.b2pt_trace_any(self._h, _p(o), _p(d), _p(tm), n, _p(occ)), "b2pt_trace_any")
        return occ

    # ---- queries (device pointers, e.g. torch tensors' data_ptr())
    def trace_closest_device(self, d_o, d_d, d_tmax, n, d_tri, d_t, d_uv=None):
        self._check(self._L.b2pt_trace_closest_device(self._h, d_o, d_d, d_tmax, n, d_tri, d_t, d_uv), "b2pt_trace_closest_device")

    def trace_any_device(self, d_o, d_d, d_tmax, n, d_occ):
        self._check(self._L.b2pt_trace_any_device(self._h, d_o, d_d, d_tmax, n, d_occ), "b2pt_trace_any_device")

    # ---- render
    @staticmethod
    def _settings(width, height, spp, bounces, gamma=2.2):
        return Settings(width, height, spp, bounces, gamma)

    @staticmethod
    def _partition(part):
        if part is None:
            return None
        return Partition(part.get("tile_rank", 0), part.get("tile_world", 0), part.get("tile_size", 0),
                         part.get("sample_begin", 0), part.get("sample_count", 0))

    def render(self, cam: Camera, width, height, spp, bounces, seed=1234, part=None, out=None):
        """Returns fb[H, W, 3] float32, row 0 = bottom of the view (reference frameBuffer order).  `out`: a
        C-contiguous float32 array of that shape to render into (saves the page faults of a fresh 25 MB array)."""
        if out is not None and (out.shape != (height, width, 3) or out.dtype != np.float32 or not out.flags.c_contiguous):
            out = None
        fb = np.empty((height, width, 3), np.float32) if out is None else out
        st = self._settings(width, height, spp, bounces)
        pt = self._partition(part)
        rc = self._L.b2pt_render(self._h, C.byref(cam), C.byref(st), seed, None if pt is None else C.byref(pt), _p(fb))
        self._check(rc, "b2pt_render")
        return fb

    def render_device(self, cam: Camera, width, height, spp, bounces, d_rgb, seed=1234, part=None):
        st = self._settings(width, height, spp, bounces)
        pt = self._partition(part)
        rc = self._L.b2pt_render_device(self._h, C.byref(cam), C.byref(st), seed, None if pt is None else C.byref(pt), d_rgb)
        self._check(rc, "b2pt_render_device")

    def tonemap(self, d_rgb, width, height, gamma=2.2, flip=False):
        """Device frame -> (H, W, 3) bytes, byte-exact Renderer::saveImage pixel maths (src/renderer.cpp:8-17)."""
        out = np.empty((height, width, 3), np.uint8)
        self._check(self._L.b2pt_tonemap(self._h, d_rgb, width, height, gamma, 1 if flip else 0, _p(out)), "b2pt_tonemap")
        return out

    def tonemap_last(self, width, height, gamma=2.2, flip=False):
        """The frame of the last render() / progressive pass, still resident on the device."""
        out = np.empty((height, width, 3), np.uint8)
        self._check(self._L.b2pt_tonemap_last(self._h, gamma, 1 if flip else 0, _p(out)), "b2pt_tonemap_last")
        return out

    def progressive_begin(self, cam: Camera, width, height, spp, bounces, seed=1234):
        st = self._settings(width, height, spp, bounces)
        self._check(self._L.b2pt_progressive_begin(self._h, C.byref(cam), C.byref(st), seed), "b2pt_progressive_begin")
        self._prog_shape = (height, width, 3)

    def progressive_pass(self, sample_count, want_frame=True):
        """Traces the next `sample_count` samples per pixel; returns (samples_done, running-mean frame or None)."""
        fb = np.empty(self._prog_shape, np.float32) if want_frame else None
        done = C.c_int32(0)
        self._check(self._L.b2pt_progressive_pass(self._h, sample_count, _p(fb), C.byref(done)), "b2pt_progressive_pass")
        return int(done.value), fb

    # ---- introspection
    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.b2pt_get_stats(self._h, C.byref(s)), "b2pt_get_stats")
        return s.as_dict()

    def accel_info(self) -> dict:
        out = np.zeros(8, np.int64)
        self._check(self._L.b2pt_get_accel_info(self._h, _p(out)), "b2pt_get_accel_info")
        return dict(wide_nodes=int(out[0]), wide_node_bytes=int(out[1]), ref_leaves=int(out[2]), ref_nodes=int(out[3]),
                    tri_bytes=int(out[4]), ploc_iterations=int(out[5]), wide_levels=int(out[6]), hoisted_leaves=int(out[7]))

    @property
    def stream(self) -> int:
        return int(self._L.b2pt_stream(self._h) or 0)


def tonemap_thresholds(gamma: float) -> np.ndarray:
    """thr[k] = smallest float the reference's tonemap maps to a byte >= k (host only)."""
    out = np.empty(256, np.float32)
    if load_library().b2pt_tonemap_thresholds(gamma, _p(out)) != 0:
        raise B2ptError("b2pt_tonemap_thresholds failed")
    return out


class MultiEngine:
    """Several GPUs behind one renderer object, one process (include/b2pt.h, b2pt_multi_*)."""

    def __init__(self, devices=None, flags: int = 0, max_paths: int = 0):
        self._L = load_library()
        self._h = C.c_void_p()
        devs = None if devices is None else np.ascontiguousarray(devices, np.int32)
        rc = self._L.b2pt_multi_create(_p(devs), 0 if devs is None else len(devs), flags, max_paths, C.byref(self._h))
        if rc != 0:
            raise B2ptError(f"b2pt_multi_create failed ({rc}): {self._L.b2pt_multi_last_error(None).decode()}")
        self.ndev = int(self._L.b2pt_multi_device_count(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.b2pt_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise B2ptError(f"{what} failed ({rc}): {self._L.b2pt_multi_last_error(self._h).decode()}")

    def upload_scene(self, pos, nrm=None, mat=None, materials8=None, lights=REFERENCE_LIGHTS):
        pos, nrm, mat, m8, lights = check_scene_arrays(pos, nrm, mat, materials8, lights)
        mats = (Material * max(len(m8), 1))()
        for i, row in enumerate(m8):
            mats[i].type = int(row[0])
            mats[i].albedo[:] = [float(row[1]), float(row[2]), float(row[3])]
            mats[i].roughness, mats[i].metallic, mats[i].ior = float(row[4]), float(row[5]), float(row[6])
        ls = (Light * max(len(lights), 1))()
        for i, (p, c, inten) in enumerate(lights):
            ls[i].position[:] = [float(x) for x in p]
            ls[i].color[:] = [float(x) for x in c]
            ls[i].intensity = float(inten)
        self._check(self._L.b2pt_multi_upload_scene(self._h, _p(pos), _p(nrm), _p(mat), pos.shape[0], C.cast(mats, C.c_void_p), len(m8),
                                                    C.cast(ls, C.c_void_p), len(lights)), "b2pt_multi_upload_scene")

    def render(self, cam: Camera, width, height, spp, bounces, seed=1234, out=None):
        if out is not None and (out.shape != (height, width, 3) or out.dtype != np.float32 or not out.flags.c_contiguous):
            out = None
        fb = np.empty((height, width, 3), np.float32) if out is None else out
        st = Settings(width, height, spp, bounces, 2.2)
        self._check(self._L.b2pt_multi_render(self._h, C.byref(cam), C.byref(st), seed, _p(fb)), "b2pt_multi_render")
        return fb

    def tonemap_last(self, width, height, gamma=2.2, flip=False):
        out = np.empty((height, width, 3), np.uint8)
        self._check(self._L.b2pt_multi_tonemap_last(self._h, gamma, 1 if flip else 0, _p(out)), "b2pt_multi_tonemap_last")
        return out

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.b2pt_multi_get_stats(self._h, C.byref(s)), "b2pt_multi_get_stats")
        return s.as_dict()
