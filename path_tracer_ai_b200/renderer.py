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
        L.b2pt_scene_load_obj_cached.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
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
        L.b2pt_write_pfm.argtypes = [C.c_char_p, C.c_int32, C.c_int32, vp]
        L.b2pt_obj_parser_selfcheck.argtypes = [C.c_char_p, C.c_int32, C.c_int64]
        L.b2pt_obj_parser_selfcheck.restype = C.c_int
        L._host_ready = True
    return L


HOST_EXPORTS = ["b2pt_scene_load_obj", "b2pt_scene_load_obj_cached", "b2pt_scene_free", "b2pt_scene_num_triangles", "b2pt_scene_num_materials",
                "b2pt_scene_num_lights", "b2pt_scene_get_triangles", "b2pt_scene_get_materials", "b2pt_scene_get_lights",
                "b2pt_camera_look_at", "b2pt_write_png", "b2pt_write_pfm", "b2pt_obj_parser_selfcheck"]


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

    def loadFromObj(self, path: str, cache: bool | str = False) -> bool:
        """cache: True / a file name = go through the binary scene cache (b2pt_scene_load_obj_cached)."""
        L = _host_lib()
        h = C.c_void_p()
        if cache:
            rc = L.b2pt_scene_load_obj_cached(os.fsencode(path), None if cache is True else os.fsencode(cache), C.byref(h))
        else:
            rc = L.b2pt_scene_load_obj(os.fsencode(path), C.byref(h))
        if rc != 0:
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
        pos, nrm, mat, m8, _ = _capi.check_scene_arrays(pos, nrm, mat, materials8, self.lights)   # ValueError on mismatched lengths
        order = _capi.reference_order(pos)
        self.order = order
        self.pos = pos[order]
        self.nrm = (np.zeros_like(pos) if nrm is None else nrm)[order]
        self.mat = (np.zeros(len(pos), np.int32) if mat is None else mat)[order]
        self.materials8 = m8

    def getTriangles(self): return self.pos, self.nrm, self.mat
    def getMaterials(self): return self.materials8
    def getLights(self): return self.lights


class B200Renderer:
    """Drop-in for OptixRenderer: initialize / uploadScene / render / saveImage.  `devices=[...]` puts several GPUs of
    one box behind the same object (b2pt_multi_*: scene replicated, interleaved pixel runs, one gather per frame)."""

    def __init__(self, settings: Settings | None = None, device: int = 0, seed: int = 1234, flags: int = 0, max_paths: int = 0,
                 devices=None):
        self.settings = settings or Settings()
        self.device, self.seed, self.flags, self.max_paths = device, seed, flags, max_paths
        self.devices = None if devices is None else list(devices)
        self.engine: Engine | None = None
        self.multi: _capi.MultiEngine | None = None
        self.frameBuffer: np.ndarray | None = None

    def initialize(self):
        if self.devices is not None and len(self.devices) > 1:
            self.multi = _capi.MultiEngine(self.devices, self.flags, self.max_paths)
        else:
            self.engine = Engine(self.device if self.devices is None else self.devices[0], self.flags, self.max_paths)

    def _require(self, who):
        if self.engine is None and self.multi is None:
            raise B2ptError(f"B200Renderer.{who} called before initialize()")   # optix_renderer.cu:421-423

    def uploadScene(self, scene: Scene):
        self._require("uploadScene")
        (self.multi or self.engine).upload_scene(scene.pos, scene.nrm, scene.mat, scene.materials8, scene.lights)

    def render(self, camera: Camera, part=None):
        self._require("render")
        s = self.settings
        # like the reference's frameBuffer member, the array is reused by the next render() of this renderer
        if self.multi is not None:
            if part is not None:
                raise B2ptError("render: a multi-device renderer partitions the frame itself")
            self.frameBuffer = self.multi.render(camera.c, s.width, s.height, s.samplesPerPixel, s.maxBounces, self.seed, out=self.frameBuffer)
        else:
            self.frameBuffer = self.engine.render(camera.c, s.width, s.height, s.samplesPerPixel, s.maxBounces, self.seed, part,
                                                  out=self.frameBuffer)
        return self.frameBuffer

    def renderProgressive(self, camera: Camera, samples_per_pass: int):
        """Progressive, resumable accumulation through b2pt_progressive_begin / _pass: yields
        (samples_done, estimate) after every pass of `samples_per_pass` samples per pixel.  The per-pixel sums stay on
        the device and samples are added in sample order, so the LAST estimate is bit-identical to render(); the
        magenta "no valid sample" colour (renderer.hpp:78) is decided on the last pass only."""
        self._require("renderProgressive")
        if self.engine is None:
            raise B2ptError("renderProgressive: one device only")
        s = self.settings
        self.engine.progressive_begin(camera.c, s.width, s.height, s.samplesPerPixel, s.maxBounces, self.seed)
        done = 0
        while done < s.samplesPerPixel:
            done, fb = self.engine.progressive_pass(samples_per_pass)
            self.frameBuffer = fb
            yield done, fb

    def tonemapped(self, flip: bool = False) -> np.ndarray:
        """Renderer::saveImage's pixel maths (src/renderer.cpp:8-17) — clamp, pow(1/gamma), truncate — on the GPU,
        byte-exact, from the frame the last render left on the device."""
        if self.frameBuffer is None:
            raise B2ptError("saveImage: nothing rendered yet")
        s = self.settings
        return (self.multi or self.engine).tonemap_last(s.width, s.height, s.gamma, flip)

    def saveImage(self, filename: str, flip: bool = False):
        """PNG (rows in framebuffer order like the reference, upright with flip=True), or the linear float frame when
        the name ends in .pfm."""
        L = _host_lib()
        s = self.settings
        if filename.lower().endswith(".pfm"):
            if self.frameBuffer is None:
                raise B2ptError("saveImage: nothing rendered yet")
            fb = np.ascontiguousarray(self.frameBuffer, np.float32)
            if L.b2pt_write_pfm(os.fsencode(filename), s.width, s.height, fb.ctypes.data) != 0:
                raise B2ptError(f"saveImage: cannot write {filename}")
            return
        px = np.ascontiguousarray(self.tonemapped(flip))
        if L.b2pt_write_png(os.fsencode(filename), s.width, s.height, px.ctypes.data) != 0:
            raise B2ptError(f"saveImage: cannot write {filename}")

    def stats(self):
        self._require("stats")
        return (self.multi or self.engine).stats()
