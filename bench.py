#!/usr/bin/env python
"""bench.py — headline benchmark of the per-pixel Monte-Carlo hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1|c2|c3|c5]
                    [--scaling strong|weak] [--spp S] [--max-paths P] [--no-cpu-baseline] [--no-c4] [--c4-rays R]

Workload (default): BASELINE.json configs[2] — the 1M-triangle tessellated mesh scene (mixed materials, plus the 8 room
triangles Scene::loadFromObj always adds), 1920x1080, 256 spp, 8 bounces: the configuration the north-star's
">= 1.5 Grays/s per B200 on a 1M-triangle scene" is quoted on.  One "step" = one full frame.  metric = Msamples/s
(camera paths per second, whole job); Mrays/s is reported beside it.  c1 / c2 / c5 are configs[0] / [1] / [4].
configs[3] — the closest-hit microbench, 100M random rays against the 1M-triangle mesh, hit ids checked bit-for-bit
against the reference CPU BVH on a stratified subset — runs in the same process and is reported under `c4`.

* value   — frames rendered with the scene resident in HBM and the frame left on the device; timed between
            barrier + synchronize pairs (max over ranks), an L2 flush between iterations.
* e2e     — the reference-facing call sequence with HOST buffers: B200Renderer.uploadScene + render, i.e. the
            region src/main.cpp:87-92 times (scene H2D from pinned host arrays and framebuffer D2H inside the timed region).
* roofline— the dominant traversal kernel: algorithmic bytes (SURVEY 8d accounting, fetch counts from the
            instrumented build of the same kernels) over its CUDA-event time on the engine's stream, against the
            measured HBM peak.  `traffic` (DRAM bytes per launch), `frac_hbm_measured` and `issue` come from the
            committed ncu capture (profiles/r02_traffic.json) and are only attached when that capture was taken from
            the same kernel sources (source hash); `bound` says what actually limits the kernel.
* cpu_baseline / --impl reference — the reference's own Renderer::render (oracle/_ref, unmodified headers) on all
            host threads, on a bounded sample of the same frame.

N>1: one process per GPU (torchrun), scene replicated.  STRONG scaling (default): the fixed frame is split into
interleaved runs of 1024 pixels (run k -> rank k mod N), every pixel is computed wholly by one rank with Philox keyed by
(pixel, sample), and the per-rank buffers are combined by ONE NCCL sum-reduce per frame inside the timed region: the
combined frame is bit-identical for every N (`frame_hash`).  `--scaling weak` gives every rank its own range of
SPP samples of a frame with SPP*N samples per pixel instead.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, spp, bounces, triangles, description)
    "c1": (800, 450, 10, 5, 0, "BASELINE configs[0]: Cornell box 800x450 10 spp 5 bounces"),
    "c2": (1920, 1080, 100, 5, 0, "BASELINE configs[1]: Cornell box 1920x1080 100 spp 5 bounces"),
    "c3": (1920, 1080, 256, 8, 1_000_000, "BASELINE configs[2]: 1M-triangle mesh scene 1920x1080 256 spp 8 bounces"),
    "c5": (3840, 2160, 1024, 8, 10_000_000, "BASELINE configs[4]: 10M-triangle dielectric-heavy scene 3840x2160 1024 spp 8 bounces"),
}
DEFAULT_MAX_PATHS = 64 << 20
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_traffic.json")
KERNEL_SOURCES = ["exact.cuh", "ctx.cuh", "traverse.cuh", "traverse_rtc.cuh", "traverse_thread.cuh", "traverse_pool.cuh", "render.cu", "trace.cu", "build.cu"]


def kernel_source_hash() -> str:
    """Hash of the kernel sources — comments and white space removed, so that only a change of CODE changes it: an ncu
    capture is only quoted beside numbers measured from the same code."""
    import re
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        p = os.path.join(ROOT, "path_tracer_ai_b200", "csrc", f)
        if os.path.exists(p):
            src = open(p, "r", encoding="utf-8", errors="replace").read()
            src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)     # block comments
            src = re.sub(r"//[^\n]*", " ", src)                   # line comments (no string literal in these files holds //)
            h.update(" ".join(src.split()).encode())
    return h.hexdigest()[:16]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_dict(args, world, ntri):
    """The same dict in both arms (the driver compares them)."""
    W, H, SPP, B, _, desc = WORKLOADS[args.workload]
    if args.spp > 0:
        SPP = args.spp
        desc += f" (spp overridden to {SPP})"
    spp_total = SPP * world if args.scaling == "weak" else SPP
    if world <= 1:
        par = "single GPU"
    elif args.scaling == "weak":
        par = f"sample ranges x{world}, scene replicated, 1 NCCL reduce/frame"
    else:
        par = f"interleaved 1024-pixel runs x{world}, scene replicated, 1 NCCL reduce/frame"
    return {"workload": desc, "width": W, "height": H, "spp": spp_total, "bounces": B, "triangles": int(ntri),
            "parallelism": par, "max_paths_in_flight": args.max_paths,
            "l2": "256 MB buffer written between timed iterations (L2 flush); per-batch path state (GBs) also exceeds L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- scenes -----------------------------------------------------------------------------------------
def prebuild_arrays(workload: str, tmpdir: str):
    """The workload's scene as the reference sees it BEFORE BVH::build: ('obj', path) or ('arrays', dict).  Pure
    numpy — does not load libb2pt.so (the reference arm must not map the product library)."""
    from path_tracer_ai_b200 import scenes
    if workload in ("c1", "c2"):
        return "obj", scenes.write_cornell_obj(tmpdir, seed=1234)
    if workload == "c3":
        return "arrays", scenes.mesh_scene(1_000_000, seed=1234)
    return "arrays", scenes.mesh_scene(10_000_000, seed=1234, dielectric_fraction=0.6)


def make_scene(kind, payload):
    import path_tracer_ai_b200 as pt
    sc = pt.Scene()
    if kind == "obj":
        assert sc.loadFromObj(payload)
    else:
        sc.setContents(payload["pos"], payload["nrm"], payload["mat"], payload["materials8"])
    return sc


def make_cpu_oracle(kind, payload):
    """The reference's own CPU path on the workload's scene (oracle/_ref when built, else the in-repo port)."""
    import oracle
    if oracle.ref_available():
        R = oracle.RefOracle(obj_path=payload) if kind == "obj" else \
            oracle.RefOracle(payload["pos"], payload["nrm"], payload["mat"], payload["materials8"])
        return R, "reference", oracle.RefOracle.max_threads()
    if kind == "obj":
        import path_tracer_ai_b200 as pt   # the port has no OBJ loader of its own
        sc = pt.Scene(); assert sc.loadFromObj(payload)
        inv = np.empty_like(sc.order); inv[sc.order] = np.arange(len(sc.order), dtype=np.int32)
        P = oracle.PortOracle(sc.pos[inv], sc.nrm[inv], sc.mat[inv], sc.materials8)
    else:
        P = oracle.PortOracle(payload["pos"], payload["nrm"], payload["mat"], payload["materials8"])
    return P, "port", oracle.PortOracle.max_threads()


def cpu_render(orc, kind, W, H, spp, bounces):
    """One full-resolution CPU frame of `spp` samples per pixel.  Returns (seconds, samples)."""
    import oracle
    if kind == "reference":
        _, secs = orc.render(W, H, spp, bounces)
    else:
        _, secs, _ = orc.render(oracle.PortOracle.camera(), W, H, spp, bounces, seed=1)
    return secs, W * H * spp


def cpu_sample_text(W, H, spp, SPP, B, cores, kind):
    who = "reference Renderer::render (oracle/_ref, unmodified headers)" if kind == "reference" else "oracle port of Renderer::render"
    return f"{W}x{H}, {spp} of {SPP} spp, {B} bounces per step; {who} on {cores} threads"


def reference_arm(args, rank, world, emit):
    """--impl reference: the reference's own CPU path, rank 0 only; libb2pt.so is never loaded here."""
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is ONE process that should use every host thread
    # it is allowed to (set before the oracle library, and with it the OpenMP runtime, is loaded).
    if world > 1 or "TORCHELASTIC_RUN_ID" in os.environ:
        try:
            ncpu = len(os.sched_getaffinity(0))
        except AttributeError:
            ncpu = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(ncpu)
    W, H, SPP, B, _, desc = WORKLOADS[args.workload]
    if args.spp > 0:
        SPP = args.spp
    with tempfile.TemporaryDirectory() as tmp:
        kind_s, payload = prebuild_arrays(args.workload, tmp)
        orc, kind, cores = make_cpu_oracle(kind_s, payload)
    ntri = orc.ntri
    # bounded sample per step: full resolution, a few of the frame's samples per pixel (per-sample cost is the
    # same for every sample index); sized from a 1-spp calibration frame so that a step takes ~4 s.
    secs0, ns0 = cpu_render(orc, kind, W, H, 1, B)
    spp = max(1, min(SPP, int(4.0 / max(secs0, 1e-3))))
    for _ in range(min(args.warmup, 1)):
        cpu_render(orc, kind, W, H, spp, B)
    t, ns = [], 0
    for _ in range(args.steps):
        secs, ns = cpu_render(orc, kind, W, H, spp, B)
        t.append(secs)
    ms = 1e3 * sum(t) / len(t)
    value = ns / (ms * 1e-3) * 1e-6
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(args, world, ntri),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": cpu_sample_text(W, H, spp, SPP, B, cores, kind)},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ---- configs[3]: closest-hit microbench -----------------------------------------------------------------
def c4_microbench(pt, torch, dist, dev, local_rank, rank, world, nrays, ncheck, peak, max_paths):
    """100M random rays (origin uniform in the mesh AABB inflated 10 %, direction uniform on the sphere) against the
    1M-triangle mesh (geometry only), rays resident in HBM, contiguous ray-range shards over the ranks.  Timed with CUDA
    events on the engine's stream (b2pt_stats.trace_seconds), max over ranks; hit ids and t bits of a stratified subset
    are compared with the reference CPU BVH on rank 0; a 64-bit hash of ALL ids makes runs comparable."""
    from path_tracer_ai_b200 import scenes
    ms = scenes.mesh_scene(1_000_000, seed=1234, room=False)
    order = pt.reference_order(ms["pos"])
    pos = ms["pos"][order]
    eng = pt.Engine(device=local_rank, max_paths=max_paths)
    eng.upload_scene(pos)
    build_ms = eng.stats()["build_seconds"] * 1e3
    BLK = 1_000_000
    nblk = max(1, nrays // BLK)
    nrays = nblk * BLK
    b0, b1 = (nblk * rank) // world, (nblk * (rank + 1)) // world
    n = (b1 - b0) * BLK
    lo = torch.tensor(np.asarray(ms["lo"], np.float64), device=dev)
    hi = torch.tensor(np.asarray(ms["hi"], np.float64), device=dev)
    ext = (hi - lo) * 0.1
    lo, hi = (lo - ext).float(), (hi + ext).float()
    o = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
    d = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    for b in range(b0, b1):   # block b of the batch is the same rays whatever the world size
        g.manual_seed(990000 + b)
        r = torch.rand((BLK, 5), generator=g, device=dev, dtype=torch.float32)
        s = slice((b - b0) * BLK, (b - b0 + 1) * BLK)
        o[s] = lo + r[:, 0:3] * (hi - lo)
        z = 1.0 - 2.0 * r[:, 3]
        phi = (2.0 * np.pi) * r[:, 4]
        rr = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
        d[s, 0] = rr * torch.cos(phi); d[s, 1] = rr * torch.sin(phi); d[s, 2] = z
    tri = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    tt = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    times, fallback = [], 0
    for rep in range(5):
        if n:
            eng.trace_closest_device(o.data_ptr(), d.data_ptr(), None, n, tri.data_ptr(), tt.data_ptr(), None)
            st = eng.stats()
            if rep >= 2:
                times.append(st["trace_seconds"])
            fallback = st["fallback_rays"]
        else:
            times.append(0.0)
    secs = float(np.median(times))
    # 64-bit hash of all ids: sum of id_i * (odd multiplier of the GLOBAL ray index), wrapping
    idx = torch.arange(b0 * BLK, b0 * BLK + n, device=dev, dtype=torch.int64)
    hsh = ((tri[:n].to(torch.int64) + 2) * (idx * 2654435761 + 1)).sum() if n else torch.zeros((), dtype=torch.int64, device=dev)
    hits = (tri[:n] >= 0).sum().to(torch.float64)
    red = torch.stack([torch.tensor(float(secs), dtype=torch.float64, device=dev), hits, torch.tensor(float(fallback), dtype=torch.float64, device=dev)])
    if world > 1:
        mx = red.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dist.all_reduce(hsh, op=dist.ReduceOp.SUM)
        secs, hits_total, fallback = float(mx[0]), float(sm[1]), float(sm[2])
    else:
        hits_total = float(hits)
    res = {"what": "BASELINE configs[3]: closest hit, random rays vs the 1M-triangle mesh (geometry only), rays resident in HBM",
           "triangles": int(len(pos)), "rays": int(nrays), "mrays_per_s": nrays / max(secs, 1e-12) * 1e-6, "ms": secs * 1e3,
           "hit_fraction": hits_total / nrays, "fallback_rays": int(fallback), "ids_hash64": f"{int(hsh.item()) & 0xFFFFFFFFFFFFFFFF:016x}",
           "build_ms": build_ms, "timing": "median of 3 after 2 warm-up passes, CUDA events on the engine's stream, max over ranks"}
    if rank == 0 and n:
        # fetch counts: instrumented counting build of the same kernel on the same BVH, 1/16 of this rank's rays
        ce = pt.Engine(device=local_rank, flags=pt.FLAG_COUNT_FETCHES, max_paths=max_paths)
        ce.upload_scene(pos)
        m = max(n // 16, 1)
        tri2 = torch.empty(m, dtype=torch.int32, device=dev); tt2 = torch.empty(m, dtype=torch.float32, device=dev)
        ce.trace_closest_device(o.data_ptr(), d.data_ptr(), None, m, tri2.data_ptr(), tt2.data_ptr(), None)
        cs = ce.stats(); info = ce.accel_info(); ce.close()
        npr, tpr = cs["node_fetches"] / m, cs["tri_fetches"] / m
        bpr = 32 + 16 + npr * info["wide_node_bytes"] + tpr * info["tri_bytes"]
        res.update({"nodes_per_ray": npr, "tris_per_ray": tpr, "bytes_per_ray": bpr, "wide_node_bytes": info["wide_node_bytes"],
                    "achieved_gbs": nrays * bpr / max(secs, 1e-12) * 1e-9, "frac_of_hbm_peak": nrays * bpr / max(secs, 1e-12) * 1e-9 / peak})
        if ncheck > 0:
            import oracle
            step = max(n // ncheck, 1)
            sel = torch.arange(0, n, step, device=dev)[:ncheck]
            ho, hd = o[sel].cpu().numpy(), d[sel].cpu().numpy()
            gt, gtt = tri[sel].cpu().numpy(), tt[sel].cpu().numpy()
            if oracle.ref_available():
                orc, kind, cores = oracle.RefOracle(ms["pos"]), "reference", oracle.RefOracle.max_threads()
                assert np.array_equal(orc.order(), order), "reference BVH order differs from b2pt_reference_order"
                c0 = time.perf_counter(); rt, rtt = orc.trace_closest(ho, hd); c1 = time.perf_counter()
            else:
                orc, kind, cores = oracle.PortOracle(ms["pos"]), "port", oracle.PortOracle.max_threads()
                c0 = time.perf_counter(); rt, rtt, _ = orc.trace_closest(ho, hd); c1 = time.perf_counter()
            res.update({"oracle_checked": int(len(sel)), "oracle_kind": kind, "oracle_id_mismatch": int((gt != rt).sum()),
                        "oracle_t_mismatch": int((gtt.view(np.uint32) != rtt.view(np.uint32)).sum()),
                        "cpu_mrays_per_s": len(sel) / (c1 - c0) * 1e-6, "cpu_threads": cores})
    eng.close()
    del o, d, tri, tt
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the configs[3] closest-hit microbench")
    ap.add_argument("--c4-rays", type=int, default=100_000_000)
    ap.add_argument("--c4-check", type=int, default=10_000_000, help="rays of the microbench compared with the reference CPU BVH")
    ap.add_argument("--max-paths", type=int, default=DEFAULT_MAX_PATHS)
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel (reported in config.spp)")
    ap.add_argument("--flags", type=int, default=0, help="B2PT_FLAG_* bits for the timed engine (experiments; 0 = product defaults)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return reference_arm(args, rank, world, emit)

    import torch
    import torch.distributed as dist

    import path_tracer_ai_b200 as pt
    from path_tracer_ai_b200 import distributed as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    W, H, SPP, B, _, desc = WORKLOADS[args.workload]
    if args.spp > 0:
        SPP = args.spp
    tmp = tempfile.TemporaryDirectory()
    kind_s, payload = prebuild_arrays(args.workload, tmp.name)
    sc = make_scene(kind_s, payload)
    cam = pt.Camera()
    eng = pt.Engine(device=local_rank, flags=args.flags, max_paths=args.max_paths)
    eng.upload_scene(sc.pos, sc.nrm, sc.mat, sc.materials8, sc.lights)
    build_s = eng.stats()["build_seconds"]

    if args.scaling == "weak":
        spp_total = SPP * world
        part = D.sample_partition(rank, world, SPP)
    else:
        spp_total = SPP
        part = D.tile_partition(rank, world, 32)

    d_rgb = torch.empty(W * H * 3, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    h_rgb = torch.empty(W * H * 3, dtype=torch.float32, pin_memory=True) if rank == 0 else None

    def step_device():
        flush.zero_()                      # L2 flush between iterations (on torch's stream; synchronised below)
        torch.cuda.synchronize(dev)
        # render (synchronous on the engine's stream) + the frame's one collective (NCCL sum-reduce to rank 0)
        return D.render_distributed(eng, cam.c, W, H, spp_total, B, d_rgb, part, seed=1234)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    agg = {"extend_rays": 0, "shadow_rays": 0, "samples": 0, "kernel_launches": 0, "gpu_seconds": 0.0, "extend_seconds": 0.0,
           "shadow_seconds": 0.0, "order_seconds": 0.0, "extend_launches": 0, "shadow_launches": 0, "fallback_rays": 0}
    for _ in range(args.steps):
        st = step_device()
        for k in agg:
            agg[k] += st[k]
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    # the combined frame (rank 0): 64-bit hash of its bytes — identical for every N under the tile partition
    frame_hash = None
    if rank == 0:
        h_rgb.copy_(d_rgb); torch.cuda.synchronize(dev)
        frame_hash = hashlib.blake2b(h_rgb.numpy().tobytes(), digest_size=8).hexdigest()

    wall = t1 - t0
    vals = torch.tensor([wall, agg["gpu_seconds"], float(agg["samples"]), float(agg["extend_rays"] + agg["shadow_rays"]),
                         float(agg["kernel_launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall, dev_s = float(mx[0]), float(mx[1])
        samples, rays, launches = float(sm[2]), float(sm[3]), float(sm[4])
    else:
        dev_s = agg["gpu_seconds"]
        samples, rays, launches = float(agg["samples"]), float(agg["extend_rays"] + agg["shadow_rays"]), float(agg["kernel_launches"])
    ms_per_step = 1e3 * wall / args.steps
    value = samples / wall * 1e-6

    # ---- e2e: the reference-facing call sequence with HOST buffers (uploadScene + render, as timed by
    # src/main.cpp:87-92), host->device scene copy and device->host framebuffer copy inside the timed region.
    # The step's inputs live in PINNED host memory (bench contract): the scene's triangle arrays are moved there once,
    # outside the timed region; b2pt_upload_scene copies straight from the caller's pointers.
    pinned = []
    for name in ("pos", "nrm", "mat"):
        t = torch.from_numpy(np.ascontiguousarray(getattr(sc, name))).pin_memory()
        pinned.append(t)
        setattr(sc, name, t.numpy())
    r = pt.B200Renderer(pt.Settings(width=W, height=H, samplesPerPixel=spp_total, maxBounces=B), device=local_rank, seed=1234,
                        max_paths=args.max_paths)
    r.initialize()
    r.uploadScene(sc); r.render(cam, part)   # warm-up (allocations)
    barrier()
    e0 = time.perf_counter()
    e_steps = max(1, args.steps)
    for _ in range(e_steps):
        r.uploadScene(sc)                       # host scene -> device (H2D inside the timed region)
        if world == 1:
            fb = r.render(cam, part)            # device -> host framebuffer (D2H inside the timed region)
        else:
            # each rank renders its share on the device, ONE NCCL sum-reduce, rank 0 copies the frame into pinned host memory
            D.render_distributed(r.engine, cam.c, W, H, spp_total, B, d_rgb, part, seed=1234)
            if rank == 0:
                h_rgb.copy_(d_rgb, non_blocking=True)
                torch.cuda.synchronize(dev)
    barrier()
    e1 = time.perf_counter()
    e_wall = e1 - e0
    if world > 1:
        tw = torch.tensor([e_wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e_wall = float(tw[0])
    e2e_value = (samples / args.steps) * e_steps / e_wall * 1e-6
    h2d = int(sc.pos.nbytes + sc.nrm.nbytes + sc.mat.nbytes + sc.materials8.nbytes + len(sc.lights) * 28)
    d2h = int(W * H * 3 * 4)
    r.engine.close()

    peak, peak_src = load_peaks()

    # ---- configs[3] in the same run (every rank traces its shard of the batch) --------------------------------
    c4 = None
    if not args.no_c4:
        eng.close()   # free the render engine's wavefront buffers first
        eng = None
        torch.cuda.empty_cache()
        c4 = c4_microbench(pt, torch, dist, dev, local_rank, rank, world, args.c4_rays, args.c4_check, peak, args.max_paths)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (measured live with CUDA events on the engine's stream) --------
    # Fetch counts per ray KIND from the instrumented counting build of the same kernels on a 1/16-size frame: one frame
    # with the lights (closest-hit + shadow rays) and the same frame without lights (the paths do not depend on the lights,
    # so its closest-hit rays are the same rays and there are no shadow rays); shadow = difference.
    ce = pt.Engine(device=local_rank, flags=pt.FLAG_COUNT_FETCHES, max_paths=args.max_paths)
    ce.upload_scene(sc.pos, sc.nrm, sc.mat, sc.materials8, sc.lights)
    ce.render(cam.c, W // 4, H // 4, 4, B, seed=1234)
    cst = ce.stats()
    info = ce.accel_info()
    ce.upload_scene(sc.pos, sc.nrm, sc.mat, sc.materials8, [])
    ce.render(cam.c, W // 4, H // 4, 4, B, seed=1234)
    cst0 = ce.stats()
    ce.close()
    assert cst0["extend_rays"] == cst["extend_rays"] and cst0["shadow_rays"] == 0
    shadow_dominant = agg["shadow_seconds"] >= agg["extend_seconds"]
    if shadow_dominant:
        nodes_per_ray = (cst["node_fetches"] - cst0["node_fetches"]) / max(cst["shadow_rays"], 1)
        tris_per_ray = (cst["tri_fetches"] - cst0["tri_fetches"]) / max(cst["shadow_rays"], 1)
    else:
        nodes_per_ray = cst0["node_fetches"] / max(cst0["extend_rays"], 1)
        tris_per_ray = cst0["tri_fetches"] / max(cst0["extend_rays"], 1)
    # every bounce is in hit-point order, so both ray kinds take the run-to-completion kernels (engine default flags)
    dominant = "k_shadow_rtc" if shadow_dominant else "k_extend_rtc"
    if shadow_dominant:
        # per shadow ray: read the queue entry (4 B) and the vertex it shares with the other lights
        # (g0+g1 = 32 B / nlight), write the visibility byte
        k_rays, k_secs, k_launches, io = agg["shadow_rays"], agg["shadow_seconds"], agg["shadow_launches"], 4 + 32 / max(len(sc.lights), 1) + 1
    else:
        # per extend ray: read origin + direction (32 B), write the hit record (16 B)
        k_rays, k_secs, k_launches, io = agg["extend_rays"], agg["extend_seconds"], agg["extend_launches"], 32 + 16
    bytes_per_ray = io + nodes_per_ray * info["wide_node_bytes"] + tris_per_ray * info["tri_bytes"]
    achieved = k_rays * bytes_per_ray / max(k_secs, 1e-12) * 1e-9
    frac = achieved / peak
    ms_per_launch = 1e3 * k_secs / max(k_launches, 1)
    traffic = traffic_src = issue = frac_hbm_measured = None
    bound = "hbm" if frac <= 1.0 else "issue"
    src_hash = kernel_source_hash()
    capture_note = "no ncu capture committed for this workload/kernel"
    if os.path.exists(TRAFFIC_JSON) and args.max_paths == DEFAULT_MAX_PATHS and args.spp == 0:
        tj_all = json.load(open(TRAFFIC_JSON))
        tj = tj_all.get(args.workload, {}).get(dominant)
        if tj and tj_all.get("source_hash") == src_hash:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            frac_hbm_measured = traffic / (ms_per_launch * 1e-3) * 1e-9 / peak
            # what actually bounds the kernel: warp-instruction issue (4 schedulers x 148 SMs x SM clock)
            issue = {"issue_active_pct_ncu": tj["issue_active_pct"], "active_lanes_per_instruction_ncu": tj["active_lanes_per_instruction"],
                     "warp_instructions_per_launch_ncu": tj["warp_instructions_per_launch"],
                     "peak_gwarp_inst_per_s": 148 * 4 * 1.965}
            bound = tj.get("bound", bound)
            capture_note = f"ncu capture of the same kernel sources ({src_hash})"
        elif tj:
            capture_note = f"ncu capture is from other kernel sources ({tj_all.get('source_hash')} != {src_hash}): not quoted"
    roofline = {"bound": bound, "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": frac,
                "traffic": traffic, "traffic_source": traffic_src, "frac_hbm_measured": frac_hbm_measured, "issue": issue,
                "algorithmic_bytes_per_launch": k_rays * bytes_per_ray / max(k_launches, 1),
                "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
                "tris_per_ray": tris_per_ray, "wide_node_bytes": info["wide_node_bytes"],
                "kernel_ms_per_launch": ms_per_launch, "kernel_launches": k_launches,
                "kernel_share_of_step": k_secs / max(dev_s, 1e-12), "kernel_source_hash": src_hash, "capture": capture_note,
                "note": "algorithmic bytes (SURVEY 8d) = per-ray queue I/O + mean wide-node fetches x node bytes + mean triangle fetches x 48 B, "
                        "fetch counts of this kernel's ray kind from the counting build of the same kernels on a 1/16-size frame.  The BVH of a 1M-triangle scene "
                        "(~100 MB) is L2-resident, so these fetches are mostly served by L1/L2: `frac` is the algorithmic traffic against "
                        "the HBM peak, `frac_hbm_measured` the DRAM bytes ncu counted per launch against the same peak, and `bound` what "
                        "limits the kernel (issue = warp-instruction issue slots, see `issue`)"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        orc, kind, cores = make_cpu_oracle(kind_s, payload)
        secs0, _ = cpu_render(orc, kind, W, H, 1, B)
        spp_c = max(1, min(SPP, int(12.0 / max(secs0, 1e-3))))
        secs, ns = cpu_render(orc, kind, W, H, spp_c, B)
        cpu_baseline = {"value": ns / secs * 1e-6, "unit": "Msamples/s", "cores": cores, "kind": kind,
                        "sample": cpu_sample_text(W, H, spp_c, SPP, B, cores, kind) + f" ({ns} samples, {secs:.1f} s)"}

    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(args, world, len(sc.pos)),
        "mrays_per_s": rays / wall * 1e-6, "rays_per_sample": rays / samples, "device_ms_per_step": 1e3 * dev_s / args.steps,
        "extend_ms_per_step": 1e3 * agg["extend_seconds"] / args.steps, "shadow_ms_per_step": 1e3 * agg["shadow_seconds"] / args.steps,
        "order_ms_per_step": 1e3 * agg["order_seconds"] / args.steps,
        "extend_mrays_per_s": agg["extend_rays"] / max(agg["extend_seconds"], 1e-12) * 1e-6,
        "shadow_mrays_per_s": agg["shadow_rays"] / max(agg["shadow_seconds"], 1e-12) * 1e-6,
        "fallback_rays_per_step": agg["fallback_rays"] / args.steps, "build_ms": 1e3 * build_s,
        "frame_hash": frame_hash,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps,
                "what": "B200Renderer.uploadScene + render with host buffers (the region src/main.cpp:87-92 times)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "c4": c4,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
