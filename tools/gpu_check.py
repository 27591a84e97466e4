"""First-contact GPU check: trace + render parity against the port oracle, and a first timing."""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import path_tracer_ai_b200 as pt
from path_tracer_ai_b200 import scenes
from oracle import PortOracle

out = {}
eng = pt.Engine()
# ---- soup trace parity
pos = scenes.random_soup(20000, 1)
order = pt.reference_order(pos)
P = PortOracle(pos)
assert np.array_equal(order, P.order())
ppos, _, _ = P.triangles()
eng.upload_scene(ppos)
rng = np.random.default_rng(5)
o = (rng.random((400000, 3)) * 2.4 - 1.2).astype(np.float32)
d = rng.normal(size=(400000, 3)).astype(np.float32)
tri, t, uv = eng.trace_closest(o, d)
st = eng.stats()
rt, rtt, ruv = P.trace_closest(o, d)
out["soup_ids_equal"] = bool(np.array_equal(tri, rt)); out["soup_t_equal"] = bool(np.array_equal(t.view(np.uint32), rtt.view(np.uint32)))
out["soup_uv_equal"] = bool(np.array_equal(uv.view(np.uint32), ruv.view(np.uint32)))
out["soup_mismatch"] = int((tri != rt).sum()); out["soup_stats"] = st
occ = eng.trace_any(o, d, np.full(len(o), 0.7, np.float32))
rocc = P.trace_any(o, d, np.full(len(o), 0.7, np.float32))
out["soup_any_equal"] = bool(np.array_equal(occ, rocc))
print(json.dumps(out, indent=1)); sys.stdout.flush()

# ---- mesh trace parity + speed
ms = scenes.mesh_scene(200000)
P2 = PortOracle(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
p2, n2, m2 = P2.triangles()
eng.upload_scene(p2, n2, m2, ms["materials8"])
print("build", eng.stats()["build_seconds"], eng.accel_info())
o, d = scenes.random_rays(2_000_000, ms["lo"], ms["hi"], 7)
tri, t, uv = eng.trace_closest(o, d)
st = eng.stats()
t0 = time.time(); rt, rtt, ruv = P2.trace_closest(o[:300000], d[:300000]); cpu_s = time.time() - t0
n = 300000
out["mesh_ids_equal"] = bool(np.array_equal(tri[:n], rt)); out["mesh_t_equal"] = bool(np.array_equal(t[:n].view(np.uint32), rtt.view(np.uint32)))
out["mesh_mismatch"] = int((tri[:n] != rt).sum()); out["mesh_hit_frac"] = float((tri >= 0).mean())
out["mesh_stats"] = st; out["mesh_mrays_s"] = 2.0 / st["trace_seconds"]; out["mesh_cpu_mrays_s"] = n / cpu_s * 1e-6
print(json.dumps(out, indent=1)); sys.stdout.flush()

# ---- cornell render parity
dd = tempfile.mkdtemp()
objp = scenes.write_cornell_obj(dd)
sc = pt.Scene(); assert sc.loadFromObj(objp)
# oracle gets the loader's PRE-build list: undo the order
inv = np.empty_like(sc.order); inv[sc.order] = np.arange(len(sc.order), dtype=np.int32)
P3 = PortOracle(sc.pos[inv], sc.nrm[inv], sc.mat[inv], sc.materials8)
assert np.array_equal(P3.order(), sc.order)
eng.upload_scene(sc.pos, sc.nrm, sc.mat, sc.materials8, sc.lights)
cam = pt.Camera()
W, H, SPP, B = 160, 90, 8, 5
fb = eng.render(cam.c, W, H, SPP, B, seed=99)
st = eng.stats()
cam13 = np.concatenate([cam.getPosition(), cam.getForward(), cam.getRight(), cam.getUp(), [cam.getFOV()]]).astype(np.float32)
ofb, secs, rays = P3.render(cam13, W, H, SPP, B, seed=99)
out["render_equal"] = bool(np.array_equal(fb.view(np.uint32), ofb.view(np.uint32)))
out["render_maxabs"] = float(np.abs(fb - ofb).max()); out["render_mean"] = [float(fb.mean()), float(ofb.mean())]
out["render_ndiff_px"] = int((np.abs(fb - ofb).max(-1) > 0).sum())
out["render_stats"] = st; out["oracle_rays"] = rays
print(json.dumps(out, indent=1)); sys.stdout.flush()
# ---- cornell speed
for (w, h, spp) in [(800, 450, 10), (1920, 1080, 16)]:
    fb = eng.render(cam.c, w, h, spp, 5, seed=1)
    st = eng.stats()
    out[f"cornell_{w}x{h}x{spp}"] = dict(msamples_s=st["samples"] / st["gpu_seconds"] * 1e-6, mrays_s=(st["extend_rays"] + st["shadow_rays"]) / st["gpu_seconds"] * 1e-6,
                                        trace_frac=st["trace_seconds"] / st["gpu_seconds"], **st)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
