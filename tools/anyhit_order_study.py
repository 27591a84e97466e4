"""Drives tools/anyhit_order_sim.cpp: builds realistic shadow rays for a scene on the CPU (camera rays and one
diffuse bounce traced with the port oracle, shadow rays towards the four reference lights from every vertex whose
light is above the horizon, exactly the set k_hitinfo queues) and prints the work each visiting order costs.

    python tools/anyhit_order_study.py cornell|mesh [ntri] [width height]

Test/experiment infrastructure only (uses oracle/); no GPU needed."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import path_tracer_ai_b200 as pt
from path_tracer_ai_b200 import scenes
from oracle import PortOracle

kind = sys.argv[1] if len(sys.argv) > 1 else "cornell"
ntri = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (240, 135)
sc = pt.Scene()
if kind == "cornell":
    with tempfile.TemporaryDirectory() as tmp:
        assert sc.loadFromObj(scenes.write_cornell_obj(tmp))
else:
    ms = scenes.mesh_scene(ntri, seed=1234)
    sc.setContents(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
pos, nrm = sc.pos, sc.nrm                      # reference post-build order
inv = np.empty_like(sc.order); inv[sc.order] = np.arange(len(sc.order), dtype=np.int32)
P = PortOracle(pos[inv], nrm[inv], sc.mat[inv], sc.materials8)
assert np.array_equal(P.order(), sc.order)
cam = P.camera()
rng = np.random.default_rng(7)
uv = np.stack(np.meshgrid((np.arange(W) + 0.5) / (W - 1), (np.arange(H) + 0.5) / (H - 1)), -1).reshape(-1, 2).astype(np.float32)
rays = P.camera_rays(cam, uv)
o, d = rays[:, :3].copy(), rays[:, 3:].copy()
lights = np.float32([[2, 3.5, 2], [-1.5, 2, 1.5], [0, 2, -2], [0, 0.1, 0]])
shadow = []
for depth in range(2):
    tri, t, uvb = P.trace_closest(o, d)
    hit = tri >= 0
    o, d, tri, t, uvb = o[hit], d[hit], tri[hit], t[hit], uvb[hit]
    X = o + d * t[:, None]
    n9 = nrm[tri].reshape(-1, 3, 3)
    w = 1 - uvb[:, 0] - uvb[:, 1]
    n = w[:, None] * n9[:, 0] + uvb[:, 0:1] * n9[:, 1] + uvb[:, 1:2] * n9[:, 2]
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    so = X + n * 0.001
    for L in lights:
        ld = L[None] - X
        dist = np.linalg.norm(ld, axis=1)
        ld = ld / dist[:, None]
        keep = (np.einsum("ij,ij->i", n, ld) > 0) & (dist > 1e-4)
        shadow.append(np.concatenate([so[keep], ld[keep], (dist[keep] - 0.001)[:, None]], 1))
    # one uniform-hemisphere bounce
    v = rng.normal(size=X.shape); v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[np.einsum("ij,ij->i", v, n) < 0] *= -1
    o, d = so.astype(np.float32), v.astype(np.float32)
R = np.concatenate(shadow).astype(np.float32)
R = R[rng.permutation(len(R))]
with tempfile.TemporaryDirectory() as tmp:
    exe = os.path.join(tmp, "sim")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "anyhit_order_sim.cpp")], check=True)
    with open(os.path.join(tmp, "t.bin"), "wb") as f:
        f.write(np.int32(len(pos)).tobytes()); f.write(np.ascontiguousarray(pos, np.float32).tobytes())
    with open(os.path.join(tmp, "r.bin"), "wb") as f:
        f.write(np.int32(len(R)).tobytes()); f.write(R.tobytes())
    print(f"{kind}: {len(pos)} triangles, {len(R)} shadow rays from {W}x{H} camera paths (2 vertices each)")
    print(subprocess.run([exe, os.path.join(tmp, "t.bin"), os.path.join(tmp, "r.bin")], capture_output=True, text=True).stdout)
