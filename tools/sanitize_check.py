"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): a sorted-bounce render of a 6000-triangle mesh
scene in two batch sizes, a progressive pass, closest-hit and any-hit queries.

    python tools/sanitize_check.py        (or under compute-sanitizer --tool memcheck / racecheck where the pool allows it)
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import path_tracer_ai_b200 as pt
from path_tracer_ai_b200 import scenes

ms = scenes.mesh_scene(6000, seed=11)
order = pt.reference_order(ms["pos"])
cam = pt.Camera()
frames = []
for max_paths in (0, 3000):
    eng = pt.Engine(max_paths=max_paths)
    eng.upload_scene(ms["pos"][order], ms["nrm"][order], ms["mat"][order], ms["materials8"])
    assert eng.accel_info()["wide_nodes"] > 64          # the sorted (unfused) path
    frames.append(eng.render(cam.c, 64, 36, 3, 5, seed=3))
    frames.append(eng.render(cam.c, 64, 36, 3, 5, seed=3))
    rng = np.random.default_rng(1)
    o = (rng.random((20000, 3)) * 6 - 3).astype(np.float32)
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    eng.trace_closest(o, d)
    eng.trace_any(o, d, np.full(20000, 2.0, np.float32))
    eng.close()
assert all(np.array_equal(f.view(np.uint32), frames[0].view(np.uint32)) for f in frames)
print("sanitize_check ok", float(frames[0].mean()))
