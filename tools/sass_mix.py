"""Instruction mix of the kernels in libb2pt.so (cuobjdump -sass): per kernel the counts of the fp32 arithmetic, min/max,
memory and control mnemonics.  What it is for: the parity arithmetic must stay scalar, two-rounding fp32 — FMUL / FADD,
never a contracted FFMA on a value that decides a hit, and never the packed FFMA2 ptxas makes out of mul.rn.f32x2 +
add.rn.f32x2 (profiles/r02_experiments.md).

    python tools/sass_mix.py [path/to/libb2pt.so] > profiles/r02_sass_mix.txt
"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "path_tracer_ai_b200", "libb2pt.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, mix = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("b2pt::", "").replace("void ", "").split("(")[0]
        mix[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?", line)
    if m and kern:
        mix[kern][m.group(1)] += 1
cols = ["FMUL", "FADD", "FFMA", "FMUL2", "FADD2", "FFMA2", "FMNMX", "FMNMX3", "FSETP", "MUFU", "LDG", "LDL", "STL", "LDS", "STS", "BRA"]
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{c:>6s}" for c in cols))
for k, c in mix.items():
    if not k.startswith("k_"):
        continue
    print(f"{k[:58]:58s} {sum(c.values()):6d} " + " ".join(f"{c.get(x, 0):6d}" for x in cols))
