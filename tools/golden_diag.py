"""Where does a converged GPU frame differ from a reference golden?  python tools/golden_diag.py tests/golden/render_mesh.npz [spp_log2]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import path_tracer_ai_b200 as pt
g = np.load(sys.argv[1])
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 21
order = pt.reference_order(g["pos"])
eng = pt.Engine()
eng.upload_scene(g["pos"][order], g["nrm"][order], g["mat"][order], g["materials8"])
ref = g["fb_ref"]; H, W, _ = ref.shape
a = eng.render(pt.Camera().c, W, H, 1 << lg, int(g["bounces"]), seed=77)
b = eng.render(pt.Camera().c, W, H, 1 << lg, int(g["bounces"]), seed=78)
def rel(x, y): return np.sqrt(((x - y) ** 2).mean((0, 1))) / y.mean((0, 1))
print("gpu vs gpu ", rel(a, b)); print("gpu vs ref ", rel(a, ref)); print("gpu2 vs ref", rel(b, ref))
print("lum ratio", a.mean() / ref.mean(), b.mean() / ref.mean())
d = np.abs(a - ref).sum(2)
idx = np.argsort(d.ravel())[::-1][:12]
for i in idx:
    y, x = divmod(int(i), W)
    print(f"pixel ({x:2d},{y:2d}) ref {ref[y, x]} gpu {a[y, x]} gpu2 {b[y, x]}")
d2 = ((a - ref) ** 2).sum(2)
print("share of squared error in the 12 worst pixels:", d2.ravel()[idx].sum() / d2.sum())
k = 2
A = a.reshape(H // k, k, W // k, k, 3).mean((1, 3)); R = ref.reshape(H // k, k, W // k, k, 3).mean((1, 3))
print("2x2 blocks  ", rel(A, R))
