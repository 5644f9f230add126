"""Attribute the per-instruction counters of an ncu report (--page source --csv, SASS view) to CUDA source lines using
nvdisasm --print-line-info of the same cubin. usage: sass_lines.py <src.csv> <nvdisasm.txt> <mangled kernel name>"""
import csv, re, sys, collections
src_csv, sass_txt, kern = sys.argv[1:4]
# 1. nvdisasm: instruction offset -> (file, line) of the innermost location, plus the inline chain's outermost line
lines = open(sass_txt).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
loc = {}
cur = None; outer = None
for l in lines[start + 1:]:
    if l.startswith("\t.section") and ".text." in l: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        loc[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ci = {n: H.index(n) for n in ("Address", "Source", "Instructions Executed", "# Samples", "Thread Instructions Executed")}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    a = int(r[ci["Address"]], 16)
    if base is None: base = a
    k = loc.get(a - base, ("?", 0))
    v = [int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0)]
    for j in range(3):
        agg[k][j] += v[j]; tot[j] += v[j]
print("total warp-inst %d samples %d  lanes/inst %.1f" % (tot[0], tot[1], tot[2] / max(tot[0], 1)))
byfile = collections.defaultdict(lambda: [0, 0])
for k, v in agg.items():
    byfile[k[0]][0] += v[0]; byfile[k[0]][1] += v[1]
print({k: (round(100 * v[0] / tot[0], 1), round(100 * v[1] / tot[1], 1)) for k, v in byfile.items()})
print("%-22s %8s %8s %6s" % ("file:line", "inst%", "samples%", "lanes"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    print("%-22s %8.2f %8.2f %6.1f" % ("%s:%d" % k, 100 * v[0] / tot[0], 100 * v[1] / tot[1], v[2] / max(v[0], 1)))
# coarse phases by source range (tc_core.cuh)
phases = [("tc_core geometry 450-560", 450, 560), ("clip_line 602-642", 602, 642), ("setup bres/line2 658-698", 658, 698), ("fill_convex setup 700-780", 700, 780),
          ("polyline_setup 782-822", 782, 822), ("draw prims 825-910", 825, 910), ("plane helpers 576-600", 576, 600)]
print("--- tc_core.cuh by phase (inst%, samples%)")
for nm, lo, hi in phases:
    a = [v for k, v in agg.items() if k[0] == "tc_core.cuh" and lo <= k[1] <= hi]
    print("%-28s %6.2f %6.2f" % (nm, 100 * sum(v[0] for v in a) / tot[0], 100 * sum(v[1] for v in a) / tot[1]))
