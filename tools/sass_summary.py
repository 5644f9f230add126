"""Per-kernel resource usage and Blackwell-relevant SASS mnemonics of the built library (no GPU needed):
  python tools/sass_summary.py > profiles/r02_sass_summary.txt
REG / STACK / SHARED from `cuobjdump -res-usage`; from `cuobjdump -sass`: UBLKCP (TMA 1-D bulk copies, cp.async.bulk), SYNCS (mbarrier),
STG.E.EF.128 (128-bit evict-first observation stores), STL/LDL (local-memory spills), ATOMS (shared-memory atomics of the rasteriser), BAR.SYNC,
DFMA/DMUL/DADD (the float64 arithmetic parity needs), and the absence of UTMA* / UTC*MMA (no 2-D tile movement, no dense contraction on this path)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tinycarlo_b200 import _lib  # noqa: E402

so = _lib.LIB_PATH
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+ STACK:\d+ SHARED:\d+)", res):
    usage[m.group(1)] = m.group(2)
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
PAT = collections.OrderedDict([("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("STG.E.EF.128", r"\bSTG\.E\.EF\.128"), ("STG(all)", r"\bSTG\."), ("STL", r"\bSTL\b"), ("LDL", r"\bLDL"),
                               ("ATOMS", r"\bATOMS"), ("BAR", r"\bBAR\.SYNC"), ("F64 arith", r"\bD(FMA|MUL|ADD)\b"), ("UTMA*", r"\bUTMA"), ("UTC*MMA", r"\bUTC\w*MMA"), ("HMMA", r"\bHMMA")])
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and "/*" in ln:
        counts[cur]["instructions"] += 1
        for k, p in PAT.items():
            if re.search(p, ln):
                counts[cur][k] += 1
dem = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(so, ROOT)}  source hash {_lib.built_hash()}  (nvcc {' '.join(_lib.NVCC_FLAGS)})")
print(f"{'kernel':66s} {'resources':32s} {'instr':>6s} " + " ".join(f"{k:>12s}" for k in PAT))
for (fn, c), name in zip(counts.items(), dem):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"{name[:66]:66s} {usage.get(fn, ''):32s} {c['instructions']:6d} " + " ".join(f"{c[k]:12d}" for k in PAT))
