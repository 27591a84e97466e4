"""Python mirrors of the reference's host classes, bound to the C ABI (include/b2pt.h, include/b2pt_host.h).

``B200Renderer`` has the lifecycle of the reference's ``OptixRenderer`` (include/gpu/optix_renderer.hpp:11-42):
``Settings`` -> ``initialize()`` -> ``uploadScene(scene)`` -> ``render(camera)`` -> ``saveImage(path)``, with the
same error behaviour (``render`` before ``initialize`` raises, every engine failure raises ``B2ptError`` carrying
the C layer's message).  ``Scene`` / ``Camera`` wrap the C++ host implementations (host/scene.hpp, host/camera.hpp)
so that Python and the command line share one loader.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import B2ptError, Engine, Light, Material


@dataclass
class Settings:
    """OptixRenderer::Settings (include/gpu/optix_renderer.hpp:14-24), same defaults."""
    width: int = 800
    height: int = 450
    samplesPerPixel: int = 10
    maxBounces: int = 3
    gamma: float = 2.2


def _host_lib():
    L = _capi.load_library()
    if not getattr(L, "_host_ready", False):
        vp = C.c_void_p
        L.b2pt_scene_load_obj.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.b2pt_scene_free.argtypes = [vp]
        L.b2pt_scene_free.restype = None
        L.b2pt_scene_num_triangles.argtypes = [vp]
        L.b2pt_scene_num_triangles.restype = C.c_int64
        L.b2pt_scene_num_materials.argtypes = [vp]
        L.b2pt_scene_num_lights.argtypes = [vp]
        L.b2pt_scene_get_triangles.argtypes = [vp, vp, vp, vp, vp]
        L.b2pt_scene_get_materials.argtypes = [vp, vp]
        L.b2pt_scene_get_lights.argtypes = [vp, vp]
        L.b2pt_camera_look_at.argtypes = [vp, vp, vp, C.c_float, C.POINTER(_capi.Camera)]
        L.b2pt_write_png.argtypes = [C.c_char_p, C.c_int32, C.c_int32, vp]
        L.b2pt_obj_parser_selfcheck.argtypes = [C.c_char_p, C.c_int32, C.c_int64]
        L.b2pt_obj_parser_selfcheck.restype = C.c_int
        L._host_ready = True
    return L


HOST_EXPORTS = ["b2pt_scene_load_obj", "b2pt_scene_free", "b2pt_scene_num_triangles", "b2pt_scene_num_materials",
                "b2pt_scene_num_lights", "b2pt_scene_get_triangles", "b2pt_scene_get_materials", "b2pt_scene_get_lights",
                "b2pt_camera_look_at", "b2pt_write_png", "b2pt_obj_parser_selfcheck"]


class Camera:
    """Camera(position, target, up, fov) — include/camera.hpp:9-16; defaults are src/main.cpp:46-51."""

    def __init__(self, position=(0.0, 2.0, 5.0), target=(0.0, 1.8, 0.0), up=(0.0, 1.0, 0.0), fov=45.0):
        L = _host_lib()
        self.c = _capi.Camera()
        p, t, u = (np.asarray(v, np.float32) for v in (position, target, up))
        rc = L.b2pt_camera_look_at(p.ctypes.data, t.ctypes.data, u.ctypes.data, float(fov), C.byref(self.c))
        if rc != 0:
            raise B2ptError("b2pt_camera_look_at failed")

    def getPosition(self): return np.array(self.c.position[:], np.float32)
    def getForward(self): return np.array(self.c.forward[:], np.float32)
    def getRight(self): return np.array(self.c.right[:], np.float32)
    def getUp(self): return np.array(self.c.up[:], np.float32)
    def getFOV(self): return float(self.c.fov)


class Scene:
    """Scene (include/scene.hpp:39-115): four fixed lights, loadFromObj, getTriangles/Materials/Lights.
    Triangles are held in the reference's post-BVH-build order."""

    def __init__(self):
        self.pos = np.zeros((0, 9), np.float32)
        self.nrm = np.zeros((0, 9), np.float32)
        self.mat = np.zeros(0, np.int32)
        self.order = np.zeros(0, np.int32)
        self.materials8 = np.zeros((0, 8), np.float32)
        self.lights = list(_capi.REFERENCE_LIGHTS)

    def loadFromObj(self, path: str) -> bool:
        L = _host_lib()
        h = C.c_void_p()
        if L.b2pt_scene_load_obj(os.fsencode(path), C.byref(h)) != 0:
            return False
        try:
            n = L.b2pt_scene_num_triangles(h)
            self.pos = np.empty((n, 9), np.float32)
            self.nrm = np.empty((n, 9), np.float32)
            self.mat = np.empty(n, np.int32)
            self.order = np.empty(n, np.int32)
            L.b2pt_scene_get_triangles(h, self.pos.ctypes.data, self.nrm.ctypes.data, self.mat.ctypes.data, self.order.ctypes.data)
            nm = L.b2pt_scene_num_materials(h)
            mats = (Material * max(nm, 1))()
            L.b2pt_scene_get_materials(h, C.cast(mats, C.c_void_p))
            self.materials8 = np.array([[m.type, m.albedo[0], m.albedo[1], m.albedo[2], m.roughness, m.metallic, m.ior, 0.0]
                                        for m in mats[:nm]], np.float32).reshape(nm, 8)
            nl = L.b2pt_scene_num_lights(h)
            ls = (Light * max(nl, 1))()
            L.b2pt_scene_get_lights(h, C.cast(ls, C.c_void_p))
            self.lights = [(tuple(l.position[:]), tuple(l.color[:]), float(l.intensity)) for l in ls[:nl]]
        finally:
            L.b2pt_scene_free(h)
        return True

    def setContents(self, pos, nrm, mat, materials8):
        """Triangles in arbitrary order -> applies the reference BVH::build ordering (bvh.hpp:27-72)."""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 9)
        order = _capi.reference_order(pos)
        self.order = order
        self.pos = pos[order]
        self.nrm = (np.zeros_like(pos) if nrm is None else np.ascontiguousarray(nrm, np.float32).reshape(-1, 9))[order]
        self.mat = (np.zeros(len(pos), np.int32) if mat is None else np.ascontiguousarray(mat, np.int32))[order]
        self.materials8 = np.ascontiguousarray(materials8, np.float32).reshape(-1, 8)

    def getTriangles(self): return self.pos, self.nrm, self.mat
    def getMaterials(self): return self.materials8
    def getLights(self): return self.lights


class B200Renderer:
    """Drop-in for OptixRenderer: initialize / uploadScene / render / saveImage."""

    def __init__(self, settings: Settings | None = None, device: int = 0, seed: int = 1234, flags: int = 0, max_paths: int = 0):
        self.settings = settings or Settings()
        self.device, self.seed, self.flags, self.max_paths = device, seed, flags, max_paths
        self.engine: Engine | None = None
        self.frameBuffer: np.ndarray | None = None

    def initialize(self):
        self.engine = Engine(self.device, self.flags, self.max_paths)

    def _require(self, who):
        if self.engine is None:
            raise B2ptError(f"B200Renderer.{who} called before initialize()")   # optix_renderer.cu:421-423

    def uploadScene(self, scene: Scene):
        self._require("uploadScene")
        self.engine.upload_scene(scene.pos, scene.nrm, scene.mat, scene.materials8, scene.lights)

    def render(self, camera: Camera, part=None):
        self._require("render")
        s = self.settings
        # like the reference's frameBuffer member, the array is reused by the next render() of this renderer
        self.frameBuffer = self.engine.render(camera.c, s.width, s.height, s.samplesPerPixel, s.maxBounces, self.seed, part,
                                              out=self.frameBuffer)
        return self.frameBuffer

    def renderProgressive(self, camera: Camera, samples_per_pass: int):
        """Progressive, resumable accumulation: yields (samples_done, estimate) after every pass of
        `samples_per_pass` samples per pixel.  Pass k renders the sample range [k*n, (k+1)*n) of the frame through
        `b2pt_partition.sample_begin/sample_count`, so the union of the passes is exactly the sample set of the
        one-shot frame (same Philox streams); only the order of the float additions differs.  The estimate after a
        pass is the running mean; the last one equals render() up to that rounding."""
        self._require("renderProgressive")
        s = self.settings
        total = s.samplesPerPixel
        accum = np.zeros((s.height, s.width, 3), np.float64)
        done = 0
        while done < total:
            n = min(samples_per_pass, total - done)
            part = self.engine.render(camera.c, s.width, s.height, total, s.maxBounces, self.seed,
                                      dict(sample_begin=done, sample_count=n))
            accum += part.astype(np.float64) * total     # a pass returns (sum of its samples) / total
            done += n
            self.frameBuffer = (accum / done).astype(np.float32)
            yield done, self.frameBuffer

    def tonemapped(self) -> np.ndarray:
        """Renderer::saveImage's pixel maths (src/renderer.cpp:8-17) on the host: clamp, pow(1/gamma), truncate."""
        if self.frameBuffer is None:
            raise B2ptError("saveImage: nothing rendered yet")
        c = np.clip(self.frameBuffer, 0.0, 1.0).astype(np.float32)
        c = np.power(c, np.float32(1.0 / self.settings.gamma), dtype=np.float32)
        return (c * np.float32(255.0)).astype(np.uint8)

    def saveImage(self, filename: str):
        px = np.ascontiguousarray(self.tonemapped())
        L = _host_lib()
        if L.b2pt_write_png(os.fsencode(filename), self.settings.width, self.settings.height, px.ctypes.data) != 0:
            raise B2ptError(f"saveImage: cannot write {filename}")

    def stats(self):
        self._require("stats")
        return self.engine.stats()
