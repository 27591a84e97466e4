"""Closest-hit / any-hit microbench (BASELINE configs[3]): random rays vs the 1M (or 10M) triangle mesh scene,
rays resident in HBM, timed with CUDA events on the engine's stream (b2pt_stats.trace_seconds).

    python tools/bench_trace.py [--tris 1000000] [--rays 16000000] [--reps 5] [--check 200000] [--any]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import path_tracer_ai_b200 as pt
from path_tracer_ai_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=1_000_000)
ap.add_argument("--rays", type=int, default=16_000_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", type=int, default=0, help="oracle-check this many rays (stratified)")
ap.add_argument("--any", action="store_true")
ap.add_argument("--soup", action="store_true", help="random triangle soup instead of the mesh scene")
ap.add_argument("--coherent", action="store_true", help="camera-like primary rays instead of random rays")
ap.add_argument("--out", default="")
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--sort", default="", help="reorder the batch before tracing: octant | morton (origin cell + direction octant)")
args = ap.parse_args()

t0 = time.time()
if args.soup:
    pos = scenes.random_soup(args.tris, 1234, size=2.0 / np.cbrt(args.tris))
    ms = dict(pos=pos, nrm=None, mat=None, materials8=None, lo=pos.reshape(-1, 3).min(0), hi=pos.reshape(-1, 3).max(0))
else:
    ms = scenes.mesh_scene(args.tris, seed=1234, room=False)
t1 = time.time()
order = pt.reference_order(ms["pos"])
t2 = time.time()
pos = ms["pos"][order]
eng = pt.Engine(flags=args.flags)
eng.upload_scene(pos, None if ms["nrm"] is None else ms["nrm"][order], None if ms["mat"] is None else ms["mat"][order], ms["materials8"])
build_s = eng.stats()["build_seconds"]
t3 = time.time()
n = args.rays
if args.coherent:
    rng = np.random.default_rng(1)
    side = int(np.sqrt(n)); n = side * side
    u, v = np.meshgrid(np.linspace(-1, 1, side, dtype=np.float32), np.linspace(-0.6, 0.6, side, dtype=np.float32))
    o = np.tile(np.float32([0, 2.0, 5.0]), (n, 1))
    d = np.stack([u.ravel() * 0.5, v.ravel() * 0.5 - 0.04, -np.ones(n, np.float32)], 1).astype(np.float32)
else:
    o, d = scenes.random_rays(n, ms["lo"], ms["hi"], 99)
if args.sort:
    octant = ((d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2))
    key = octant
    if args.sort == "morton":
        lo_, hi_ = o.min(0), o.max(0)
        q = np.clip(((o - lo_) / (hi_ - lo_) * 32).astype(np.int64), 0, 31)   # 32^3 origin cells
        cell = np.zeros(n, np.int64)
        for b in range(5):
            for a in range(3):
                cell |= ((q[:, a] >> b) & 1) << (3 * b + a)
        key = (cell << 3) | octant
    perm = np.argsort(key, kind="stable")
    o, d = np.ascontiguousarray(o[perm]), np.ascontiguousarray(d[perm])
dev = torch.device("cuda:0")
to, td = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
tri = torch.empty(n, dtype=torch.int32, device=dev); tt = torch.empty(n, dtype=torch.float32, device=dev)
occ = torch.empty(n, dtype=torch.uint8, device=dev)
tmax = torch.full((n,), 1.5, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
times = []
for r in range(args.reps + 2):
    if args.any:
        eng.trace_any_device(to.data_ptr(), td.data_ptr(), tmax.data_ptr(), n, occ.data_ptr())
    else:
        eng.trace_closest_device(to.data_ptr(), td.data_ptr(), None, n, tri.data_ptr(), tt.data_ptr(), None)
    st = eng.stats()
    if r >= 2:
        times.append(st["trace_seconds"])
best, med = min(times), float(np.median(times))
res = dict(tris=int(len(pos)), rays=n, mode="any" if args.any else "closest", mrays_s_median=n / med * 1e-6, mrays_s_best=n / best * 1e-6,
           ms_median=med * 1e3, fallback_rays=st["fallback_rays"], gen_s=t1 - t0, host_order_s=t2 - t1, build_ms=build_s * 1e3,
           accel=eng.accel_info())
if args.any:
    res["occluded_frac"] = float(occ.float().mean())
else:
    res["hit_frac"] = float((tri >= 0).float().mean())
# instrumented counting build of the same kernel on the same BVH and a 1/8 sample of the batch
ce = pt.Engine(flags=pt.FLAG_COUNT_FETCHES | args.flags)
ce.upload_scene(pos)
m = max(n // 8, 1)
if args.any:
    ce.trace_any_device(to.data_ptr(), td.data_ptr(), tmax.data_ptr(), m, occ.data_ptr())
else:
    ce.trace_closest_device(to.data_ptr(), td.data_ptr(), None, m, tri.data_ptr(), tt.data_ptr(), None)
cs = ce.stats()
res["nodes_per_ray"] = cs["node_fetches"] / m; res["tris_per_ray"] = cs["tri_fetches"] / m
info = ce.accel_info()
bpr = 32 + (4 if args.any else 16) + res["nodes_per_ray"] * info["wide_node_bytes"] + res["tris_per_ray"] * info["tri_bytes"]
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
res["bytes_per_ray"] = bpr; res["achieved_gbs"] = n * bpr / med * 1e-9; res["roofline_frac"] = res["achieved_gbs"] / peak
ce.close()
if args.check and not args.any:
    from oracle import PortOracle
    # re-run so that tri/tt hold the full-batch results again
    eng.trace_closest_device(to.data_ptr(), td.data_ptr(), None, n, tri.data_ptr(), tt.data_ptr(), None)
    P = PortOracle(ms["pos"])
    assert np.array_equal(P.order(), order)
    sel = np.arange(0, n, max(n // args.check, 1))[: args.check]
    c0 = time.time(); rt, rtt, _ = P.trace_closest(o[sel], d[sel]); c1 = time.time()
    g_tri = tri.cpu().numpy()[sel]; g_t = tt.cpu().numpy()[sel]
    res["oracle_checked"] = int(len(sel)); res["oracle_id_mismatch"] = int((g_tri != rt).sum())
    res["oracle_t_mismatch"] = int((g_t.view(np.uint32) != rtt.view(np.uint32)).sum())
    res["cpu_mrays_s"] = len(sel) / (c1 - c0) * 1e-6; res["cpu_threads"] = PortOracle.max_threads()
print(json.dumps(res))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
