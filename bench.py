#!/usr/bin/env python
"""bench.py — headline benchmark of the per-pixel Monte-Carlo hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1|c2|c3|c5]
                    [--scaling weak|strong] [--spp S] [--max-paths P] [--no-cpu-baseline]

Workload (default, N=1): BASELINE.json configs[1] — the procedural Cornell box through Scene::loadFromObj,
1920x1080, 100 spp, 5 bounces, on one B200.  One "step" = one full frame.  metric = Msamples/s (camera paths per
second, whole job); Mrays/s is reported beside it.  c1 / c3 / c5 are configs[0] / [2] / [4]; configs[3] (the
closest-hit microbench) is tools/bench_trace.py.

* value   — frames rendered with the scene resident in HBM and the frame left on the device; timed between
            barrier + synchronize pairs (max over ranks), an L2 flush between iterations.
* e2e     — the reference-facing call sequence with HOST buffers: B200Renderer.uploadScene + render, i.e. the
            region src/main.cpp:87-92 times (scene H2D and framebuffer D2H inside the timed region).
* roofline— the dominant traversal kernel: algorithmic bytes (SURVEY 8d accounting, fetch counts from the
            instrumented build of the same kernels) over its CUDA-event time on the engine's stream, against the
            measured HBM peak; `traffic` / `issue` are the DRAM bytes and warp-instruction figures of the committed
            ncu capture (profiles/r01_traffic.json).
* cpu_baseline / --impl reference — the reference's own Renderer::render (oracle/_ref, unmodified headers) on all
            host threads, on a bounded sample of the same frame.

N>1: one process per GPU (torchrun), scene replicated.  WEAK scaling (default): the frame has 100*N spp, rank r
renders samples [100r, 100(r+1)) and the per-rank buffers are combined by ONE NCCL sum-reduce per frame inside the
timed region.  `--scaling strong` splits the fixed frame by interleaved 1024-pixel tiles instead (bit-identical image
for any N).
"""
from __future__ import annotations

import argparse
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
    # name: (width, height, spp, bounces, description)
    "c1": (800, 450, 10, 5, "BASELINE configs[0]: Cornell box 800x450 10 spp 5 bounces"),
    "c2": (1920, 1080, 100, 5, "BASELINE configs[1]: Cornell box 1920x1080 100 spp 5 bounces"),
    "c3": (1920, 1080, 256, 8, "BASELINE configs[2]: 1M-triangle mesh scene 1920x1080 256 spp 8 bounces"),
    "c5": (3840, 2160, 1024, 8, "BASELINE configs[4]: 10M-triangle dielectric-heavy scene 3840x2160 1024 spp 8 bounces"),
}
# (configs[3], the closest-hit microbench, is tools/bench_trace.py.)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def make_scene(workload: str):
    import path_tracer_ai_b200 as pt
    from path_tracer_ai_b200 import scenes
    sc = pt.Scene()
    if workload in ("c1", "c2"):
        with tempfile.TemporaryDirectory() as tmp:
            assert sc.loadFromObj(scenes.write_cornell_obj(tmp, seed=1234))
    elif workload == "c3":
        ms = scenes.mesh_scene(1_000_000, seed=1234)
        sc.setContents(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
    else:
        ms = scenes.mesh_scene(10_000_000, seed=1234, dielectric_fraction=0.6)
        sc.setContents(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
    return sc


def prebuild(sc):
    inv = np.empty_like(sc.order)
    inv[sc.order] = np.arange(len(sc.order), dtype=np.int32)
    return sc.pos[inv], sc.nrm[inv], sc.mat[inv]


def cpu_render(sc, W, H, spp, bounces, workload, window=None):
    """Reference CPU renderer on all host threads.  Returns (seconds, samples, kind, cores, rays or None)."""
    import oracle
    if oracle.ref_available() and window is None:
        R = oracle.RefOracle(*prebuild(sc), sc.materials8)
        _, secs = R.render(W, H, spp, bounces)
        return secs, W * H * spp, "reference", oracle.RefOracle.max_threads(), None
    P = oracle.PortOracle(*prebuild(sc), sc.materials8)
    x0, y0, x1, y1 = window if window else (0, 0, W, H)
    _, secs, rays = P.render(oracle.PortOracle.camera(), W, H, spp, bounces, seed=1, window=window)
    return secs, (x1 - x0) * (y1 - y0) * spp, "port", oracle.PortOracle.max_threads(), rays


def reference_arm(args, rank, world, emit):
    """--impl reference: the reference's own CPU path, rank 0 only."""
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
    W, H, SPP, B, desc = WORKLOADS[args.workload]
    sc = make_scene(args.workload)
    # bounded sample per step: full resolution, a few of the frame's samples per pixel (per-sample cost is the
    # same for every sample index); for the 1M-triangle scene a centre crop as well.
    window = None
    spp = 2 if args.workload != "c1" else SPP
    if args.workload in ("c3", "c5"):
        window = (W // 2 - 120, H // 2 - 68, W // 2 + 120, H // 2 + 68)
        spp = 1
    secs0, ns0, kind, cores, _ = cpu_render(sc, W, H, 1, B, args.workload, window)   # calibration (also warms caches)
    rate = ns0 / secs0
    target = 4.0   # seconds per step
    per_spp = (ns0 / 1)
    spp = max(1, min(SPP, int(rate * target / per_spp)))
    for _ in range(args.warmup):
        cpu_render(sc, W, H, spp, B, args.workload, window)
    t = []
    ns = 0
    for _ in range(args.steps):
        secs, ns, kind, cores, _ = cpu_render(sc, W, H, spp, B, args.workload, window)
        t.append(secs)
    ms = 1e3 * sum(t) / len(t)
    value = ns / (ms * 1e-3) * 1e-6
    sample = f"{W}x{H}" + (f" crop {window}" if window else "") + f", {spp} of {SPP} spp, {B} bounces per step; reference Renderer::render on {cores} threads"
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "width": W, "height": H, "spp": SPP, "bounces": B},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-paths", type=int, default=32 << 20)
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel (reported in config.spp)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 0)   # the contract asks for >= 3; honour the flag but it is the caller's call

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

    W, H, SPP, B, desc = WORKLOADS[args.workload]
    if args.spp > 0:
        SPP = args.spp
        desc += f" (spp overridden to {SPP})"
    sc = make_scene(args.workload)
    cam = pt.Camera()
    eng = pt.Engine(device=local_rank, max_paths=args.max_paths)
    eng.upload_scene(sc.pos, sc.nrm, sc.mat, sc.materials8, sc.lights)
    build_s = eng.stats()["build_seconds"]

    if args.scaling == "weak":
        spp_total = SPP * world
        part = D.sample_partition(rank, world, SPP)
        parallelism = f"sample ranges x{world}, scene replicated, 1 NCCL reduce/frame" if world > 1 else "single GPU"
    else:
        spp_total = SPP
        part = D.tile_partition(rank, world, 32)
        parallelism = f"interleaved 1024-pixel tiles x{world}, scene replicated, 1 NCCL reduce/frame" if world > 1 else "single GPU"

    d_rgb = torch.empty(W * H * 3, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

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
           "shadow_seconds": 0.0, "extend_launches": 0, "shadow_launches": 0, "fallback_rays": 0}
    for _ in range(args.steps):
        st = step_device()
        for k in agg:
            agg[k] += st[k]
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None

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
    r = pt.B200Renderer(pt.Settings(width=W, height=H, samplesPerPixel=spp_total, maxBounces=B), device=local_rank, seed=1234,
                        max_paths=args.max_paths)
    r.initialize()
    r.uploadScene(sc); r.render(cam, part)   # warm-up (allocations)
    barrier()
    e0 = time.perf_counter()
    e_steps = max(1, min(args.steps, 2))
    for _ in range(e_steps):
        r.uploadScene(sc)                       # host scene -> device (H2D inside the timed region)
        if world == 1:
            fb = r.render(cam, part)            # device -> host framebuffer (D2H inside the timed region)
        else:
            # each rank renders its share on the device, ONE NCCL sum-reduce, rank 0 copies the frame to the host
            D.render_distributed(r.engine, cam.c, W, H, spp_total, B, d_rgb, part, seed=1234)
            fb = d_rgb.cpu().numpy() if rank == 0 else None
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (measured live with CUDA events on the engine's stream) --------
    peak, peak_src = load_peaks()
    ce = pt.Engine(device=local_rank, flags=pt.FLAG_COUNT_FETCHES, max_paths=args.max_paths)
    ce.upload_scene(sc.pos, sc.nrm, sc.mat, sc.materials8, sc.lights)
    ce.render(cam.c, W // 4, H // 4, 4, B, seed=1234)      # instrumented counting build of the same kernels
    cst = ce.stats()
    info = ce.accel_info()
    ce.close()
    nrays_c = cst["extend_rays"] + cst["shadow_rays"]
    nodes_per_ray = cst["node_fetches"] / max(nrays_c, 1)
    tris_per_ray = cst["tri_fetches"] / max(nrays_c, 1)
    shadow_dominant = agg["shadow_seconds"] >= agg["extend_seconds"]
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
    traffic, traffic_src, issue = None, None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.max_paths == (32 << 20) and args.spp == 0:   # captured at this batch size
        tj = json.load(open(tpath)).get(args.workload, {}).get(dominant)
        if tj:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            # what actually bounds the kernel: warp-instruction issue (4 schedulers x 148 SMs x SM clock)
            issue = {"warp_instructions_per_launch": tj["warp_instructions_per_launch"], "issue_active_pct_ncu": tj["issue_active_pct"],
                     "active_lanes_per_instruction_ncu": tj["active_lanes_per_instruction"],
                     "achieved_gwarp_inst_per_s": tj["warp_instructions_per_launch"] / max(k_secs / max(k_launches, 1), 1e-12) * 1e-9,
                     "peak_gwarp_inst_per_s": 148 * 4 * 1.965}
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "issue": issue, "algorithmic_bytes_per_launch": k_rays * bytes_per_ray / max(k_launches, 1),
                "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
                "tris_per_ray": tris_per_ray, "kernel_ms_per_launch": 1e3 * k_secs / max(k_launches, 1), "kernel_launches": k_launches,
                "kernel_share_of_step": k_secs / max(dev_s, 1e-12),
                "note": "algorithmic bytes (SURVEY 8d) = per-ray queue I/O + mean wide-node fetches x 224 B + mean triangle fetches x 48 B, "
                        "fetch counts from the counting build of the same kernels on a 1/16-size frame. frac > 1 means the node/triangle "
                        "fetches are served by L1/L2, not HBM (the whole BVH of this workload is cache resident; compare `traffic`, the "
                        "DRAM bytes ncu measured per launch): the kernel is then bound by instruction issue, not by the memory roofline "
                        "(ncu: profiles/r01_ncu_c2_batch_v10.txt, issue slots 82 % busy; see `issue`)"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        secs0, ns0, kind, cores, _ = cpu_render(sc, W, H, 1, B, args.workload,
                                                (W // 2 - 120, H // 2 - 68, W // 2 + 120, H // 2 + 68) if args.workload in ("c3", "c5") else None)
        spp_c = max(1, min(SPP, int(15.0 / max(secs0, 1e-3))))
        window = (W // 2 - 120, H // 2 - 68, W // 2 + 120, H // 2 + 68) if args.workload in ("c3", "c5") else None
        secs, ns, kind, cores, _ = cpu_render(sc, W, H, spp_c, B, args.workload, window)
        cpu_baseline = {"value": ns / secs * 1e-6, "unit": "Msamples/s", "cores": cores, "kind": kind,
                        "sample": f"{W}x{H}" + (f" crop {window}" if window else "") + f", {spp_c} of {SPP} spp, {B} bounces ({ns} samples, {secs:.1f} s)"}

    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "width": W, "height": H, "spp": spp_total, "bounces": B, "triangles": int(len(sc.pos)),
                   "parallelism": parallelism, "max_paths_in_flight": args.max_paths,
                   "l2": "256 MB buffer written between timed iterations (L2 flush); per-batch path state (GBs) also exceeds L2"},
        "mrays_per_s": rays / wall * 1e-6, "rays_per_sample": rays / samples, "device_ms_per_step": 1e3 * dev_s / args.steps,
        "extend_ms_per_step": 1e3 * agg["extend_seconds"] / args.steps, "shadow_ms_per_step": 1e3 * agg["shadow_seconds"] / args.steps,
        "fallback_rays_per_step": agg["fallback_rays"] / args.steps, "build_ms": 1e3 * build_s,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "B200Renderer.uploadScene + render with host buffers (the region src/main.cpp:87-92 times)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
