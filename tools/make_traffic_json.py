"""profiles/r02_traffic.json from an `ncu --set full` capture of the traversal kernels of one whole wavefront batch.

    python tools/make_traffic_json.py /tmp/r02_c3_final.ncu-rep profiles/r02_ncu_c3_final.txt c3 [out.json]

Per kernel: mean over the captured launches (all depths of one batch at bench.py's batch size) of the DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum), executed warp instructions, issue-slot utilisation and active lanes per
instruction.  bench.py quotes them (roofline.traffic / roofline.issue) only while the kernel sources still hash to
`source_hash`.
"""
import csv, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_hash, TRAFFIC_JSON

rep, source, workload = sys.argv[1], sys.argv[2], sys.argv[3]
OUT = sys.argv[4] if len(sys.argv) > 4 else TRAFFIC_JSON   # (under gpurun only gpurun_out/ travels back)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
unit = dict(zip(hdr, units))

def num(d, k):
    v = float(d[k].replace(",", ""))
    u = unit[k].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return v * scale.get(u, 1)

per = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    for k in ("k_extend_rtc", "k_shadow_rtc", "k_shade", "k_hitinfo"):
        if k in name:
            per.setdefault(k, []).append(d)
out = json.load(open(TRAFFIC_JSON)) if os.path.exists(TRAFFIC_JSON) else {}
out["_what"] = ("Per-launch figures from one `ncu --set full` capture of the traversal kernels of a whole wavefront batch (all depths) at "
                "bench.py's batch size, mean over the captured launches: DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum), executed "
                "warp instructions, issue-slot utilisation, active lanes per instruction.  bench.py copies these into roofline.traffic / "
                "roofline.issue while the kernel sources hash to source_hash.")
out["source_hash"] = kernel_source_hash()
w = out.setdefault(workload, {})
for k, ds in per.items():
    n = len(ds)
    inst = sum(num(d, "smsp__inst_executed.sum") for d in ds)
    w[k] = {
        "dram_bytes_per_launch": int(sum(num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum") for d in ds) / n),
        "launches_captured": n,
        "mean_launch_ms_under_ncu": round(sum(num(d, "gpu__time_duration.sum") for d in ds) / n, 3),
        "warp_instructions_per_launch": int(inst / n),
        # instruction-weighted means over the launches
        "issue_active_pct": round(sum(num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active") * num(d, "smsp__inst_executed.sum") for d in ds) / inst, 1),
        "active_lanes_per_instruction": round(sum(num(d, "smsp__thread_inst_executed_per_inst_executed.ratio") * num(d, "smsp__inst_executed.sum") for d in ds) / inst, 1),
        "l1_hit_pct": round(sum(num(d, "l1tex__t_sector_hit_rate.pct") for d in ds) / n, 1),
        "l2_hit_pct": round(sum(num(d, "lts__t_sector_hit_rate.pct") for d in ds) / n, 1),
        "bound": "issue+latency (L1/L2-resident BVH; DRAM traffic is a few % of peak)" if k in ("k_extend_rtc", "k_shadow_rtc") else "hbm",
        "source": source,
    }
json.dump(out, open(OUT, "w"), indent=1)
print(json.dumps(out[workload], indent=1))
