"""Summarises ncu outputs into the text files kept under profiles/ (run in the build container, no GPU needed).
  python tools/ncu_summary.py launches gpurun_out/launches_X.csv
  python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    body = rows[1:]
    first = next((i for i, r in enumerate(body) if r[ki].lstrip("void ").startswith("tc_")), 0)
    if first:   # launches before the library's first kernel are torch's zero-fills of the freshly allocated tensors: not part of a step
        print(f"# skipping {first} launches of the set-up (tensor allocation fills) before the first tc_* kernel")
        body = body[first:]
    for r in body:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)   # -> us
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot / 1e3:.3f} ms of device time (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'n':>4s} {'mean us':>10s} {'total us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:70]:70s} {len(v):4d} {sum(v) / len(v):10.1f} {sum(v):10.1f} {sum(v) / tot * 100:6.1f}%")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        print(f"\n## {r[hdr.index('Kernel Name')]}  (launch id {r[0]})")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
