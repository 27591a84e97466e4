"""Summaries of ncu artefacts for profiles/ (tracked).

    python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  profiles/r01_xxx.txt
    python tools/ncu_summary.py list gpurun_out/launches.csv  profiles/r01_launches_xxx.txt
"""
import csv, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]

def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for h, u in zip(hdr, units):
            if h in KEYS or "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                lines.append(f"  {h:78s} {d[h]:>16s} {u}")
    det = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
    keep = [l for l in det.splitlines() if any(k in l for k in ("Duration", "Registers Per", "Occupancy", "Ipc", "Issue Slots", "Active Threads Per Warp",
                                                              "Hit Rate", "Throughput", "Local", "Shared Memory", "Bank"))]
    open(out, "w").write("\n".join(lines) + "\n\n-- details page (selected) --\n" + "\n".join(keep) + "\n")
    print(open(out).read())

def lst(path, out):
    agg = collections.OrderedDict()
    rows = [r for r in csv.reader(open(path)) if len(r) > 14]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    total = 0.0
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("unnamed>::", "")
        t = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; total += t
    lines = [f"launch list: {path}  ({sum(a[0] for a in agg.values())} launches, {total*1e-6:.3f} ms total; per-launch times are cold-cache and serialised)",
             f"{'kernel':40s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:40s} {n:8d} {t*1e-6:10.3f} {t/n*1e-3:10.2f} {100*t/total:6.1f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))

if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
