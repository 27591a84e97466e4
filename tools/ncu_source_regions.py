"""Per-source-line roll-up of an ncu --import-source capture: share of executed warp instructions, mean active
lanes and stall samples by CUDA source line (needs -lineinfo).

    python tools/ncu_source_regions.py gpurun_out/prof.ncu-rep <kernel regex> [top N]
"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines, cur_file, hdr, done = [], None, None, False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if lines and hdr is not None and fn_seen != r[1]:
            pass
        fn_seen = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr is not None and r[0].isdigit() and len(r) > 8:
        ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ie]), int(r[it]), int(r[isamp])))
        except ValueError:
            pass
tot = sum(l[3] for l in lines) or 1
tots = sum(l[5] for l in lines) or 1
print(f"{fn_seen[:110]}\n total warp instructions {tot}, stall samples {tots}")
print(f"{'file:line':28s} {'inst%':>6s} {'lanes':>6s} {'smp%':>6s}  source")
for f, n, src, e, t, s in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{f + ':' + str(n):28s} {100 * e / tot:6.2f} {t / max(e, 1):6.1f} {100 * s / tots:6.2f}  {src[:90]}")
