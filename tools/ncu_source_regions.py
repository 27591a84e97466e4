"""Per-source-line roll-up of an ncu --import-source capture, summed over all captured launches of a kernel: share of
executed warp instructions, mean active lanes and stall samples by CUDA source line (needs -lineinfo).

    python tools/ncu_source_regions.py gpurun_out/prof.ncu-rep <kernel regex> [top N]
"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
agg = collections.defaultdict(lambda: [0,0,0,""])
cur_file=None; hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split("/")[-1]
    elif r[0]=="Line No": hdr=r
    elif hdr is not None and r[0].isdigit() and len(r)>8:
        ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        try:
            a=agg[(cur_file,int(r[0]))]; a[0]+=int(r[ie]); a[1]+=int(r[it]); a[2]+=int(r[isamp]); a[3]=r[1].strip()[:90]
        except ValueError: pass
tot=sum(a[0] for a in agg.values()); ts=sum(a[2] for a in agg.values())
print("total warp inst", tot, "samples", ts)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[3]) if len(sys.argv)>3 else 40]:
    print(f"{k[0]}:{k[1]:<5d} inst {100*a[0]/tot:5.2f}%  lanes {a[1]/max(a[0],1):5.1f}  smp {100*a[2]/ts:5.2f}%  {a[3]}")
# per-file sums
pf=collections.defaultdict(lambda:[0,0,0])
for k,a in agg.items():
    pf[k[0]][0]+=a[0]; pf[k[0]][1]+=a[1]; pf[k[0]][2]+=a[2]
for f,a in pf.items(): print(f, f"inst {100*a[0]/tot:5.1f}% lanes {a[1]/max(a[0],1):5.1f} smp {100*a[2]/ts:5.1f}%")
