import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import path_tracer_ai_b200 as pt
from path_tracer_ai_b200 import scenes
ms = scenes.mesh_scene(1_000_000, seed=1234)
sc = pt.Scene(); sc.setContents(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
r = pt.B200Renderer(pt.Settings(width=1920, height=1080, samplesPerPixel=16, maxBounces=8), max_paths=32 << 20)
r.initialize()
cam = pt.Camera()
for i in range(4):
    t0 = time.perf_counter(); r.uploadScene(sc); t1 = time.perf_counter(); fb = r.render(cam); t2 = time.perf_counter()
    st = r.stats()
    print(f"upload {1e3*(t1-t0):.1f} ms  render {1e3*(t2-t1):.1f} ms (gpu {1e3*st['gpu_seconds']:.1f} ms) build {1e3*st['build_seconds']:.2f} ms")
