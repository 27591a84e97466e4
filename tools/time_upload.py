"""Wall-clock split of the reference-facing call sequence (uploadScene + render with host buffers):
    python tools/time_upload.py [cornell|mesh] [W H SPP B]"""
import sys, time, tempfile
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import path_tracer_ai_b200 as pt
from path_tracer_ai_b200 import scenes
kind = sys.argv[1] if len(sys.argv) > 1 else "mesh"
W, H, SPP, B = (int(x) for x in sys.argv[2:6]) if len(sys.argv) > 5 else (1920, 1080, 16, 8)
sc = pt.Scene()
if kind == "cornell":
    with tempfile.TemporaryDirectory() as tmp:
        assert sc.loadFromObj(scenes.write_cornell_obj(tmp))
else:
    ms = scenes.mesh_scene(1_000_000, seed=1234)
    sc.setContents(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
r = pt.B200Renderer(pt.Settings(width=W, height=H, samplesPerPixel=SPP, maxBounces=B), max_paths=32 << 20)
r.initialize()
cam = pt.Camera()
for i in range(5):
    t0 = time.perf_counter(); r.uploadScene(sc); t1 = time.perf_counter(); fb = r.render(cam); t2 = time.perf_counter()
    st = r.stats()
    print(f"upload {1e3*(t1-t0):.2f} ms  render {1e3*(t2-t1):.2f} ms (gpu {1e3*st['gpu_seconds']:.2f} ms) build {1e3*st['build_seconds']:.2f} ms")
