"""Render-throughput experiments: python tools/bench_render.py [--scene cornell|mesh] [--tris N] [-w W -h H -s SPP -b B] [--reps R]"""
import argparse, json, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import path_tracer_ai_b200 as pt
from path_tracer_ai_b200 import scenes
ap = argparse.ArgumentParser(add_help=False)
ap.add_argument("--scene", default="cornell"); ap.add_argument("--tris", type=int, default=1_000_000)
ap.add_argument("-w", type=int, default=1920); ap.add_argument("-h", type=int, default=1080)
ap.add_argument("-s", type=int, default=16); ap.add_argument("-b", type=int, default=5); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--max-paths", type=int, default=8 << 20); ap.add_argument("--tag", default=""); ap.add_argument("--flags", type=int, default=0); ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
sc = pt.Scene()
if a.scene == "cornell":
    with tempfile.TemporaryDirectory() as tmp:
        assert sc.loadFromObj(scenes.write_cornell_obj(tmp))
else:
    ms = scenes.mesh_scene(a.tris, seed=1234)
    sc.setContents(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
eng = pt.Engine(max_paths=a.max_paths, flags=a.flags)
eng.upload_scene(sc.pos, sc.nrm, sc.mat, sc.materials8, sc.lights)
d_rgb = torch.empty(a.w * a.h * 3, dtype=torch.float32, device="cuda:0")
torch.cuda.synchronize()
best = None
for r in range(a.reps + 1):
    eng.render_device(pt.Camera().c, a.w, a.h, a.s, a.b, d_rgb.data_ptr(), seed=a.seed)
    st = eng.stats()
    if r and (best is None or st["gpu_seconds"] < best["gpu_seconds"]):
        best = st
rays = best["extend_rays"] + best["shadow_rays"]
print(json.dumps(dict(tag=a.tag, scene=a.scene, tris=len(sc.pos), msamples_s=round(best["samples"] / best["gpu_seconds"] * 1e-6, 1),
                      mrays_s=round(rays / best["gpu_seconds"] * 1e-6, 1), ms=round(best["gpu_seconds"] * 1e3, 1),
                      extend_ms=round(best["extend_seconds"] * 1e3, 1), order_ms=round(best["order_seconds"] * 1e3, 1), shadow_ms=round(best["shadow_seconds"] * 1e3, 1),
                      extend_mrays_s=round(best["extend_rays"] / best["extend_seconds"] * 1e-6, 1),
                      shadow_mrays_s=round(best["shadow_rays"] / max(best["shadow_seconds"], 1e-9) * 1e-6, 1),
                      rays_per_sample=round(rays / best["samples"], 2), fallback=best["fallback_rays"], mean=float(d_rgb.mean()))))
