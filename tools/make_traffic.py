"""profiles/traffic.json from the round's ncu --set full summaries (gpurun_out/r02_ncu_c*.txt, written by tools/gpu_r2_prof.sh):
DRAM bytes (read + write) of the dominant render kernel per env, per BASELINE config, tagged with the source hash of the library the
captures were made with. bench.py reports roofline.traffic from it only while the loaded library has that hash.
  python tools/make_traffic.py <so_hash> [dir]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so_hash = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ENVS = {"3": 16384, "2": 4096, "1": 8192, "5": 32768}


def kernels(path):
    out, cur = [], None
    for ln in open(path):
        m = re.match(r"## (.*)  \(launch id (\d+)\)", ln)
        if m:
            cur = {"kernel": m.group(1).strip(), "read": 0.0, "write": 0.0}
            out.append(cur)
            continue
        m = re.match(r"\s+(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum)\s+([\d.]+)\s+(\S+)", ln)
        if m and cur is not None:
            if m.group(2):
                cur[m.group(2)] = float(m.group(3)) * UNIT[m.group(4)]
            else:
                cur["duration_us"] = float(m.group(3)) * {"us": 1, "ms": 1e3, "ns": 1e-3}[m.group(4)]
    return out


cfgs = {}
for c, n in ENVS.items():
    p = os.path.join(src, f"r02_ncu_c{c}.txt")
    if not os.path.exists(p):
        continue
    ks = kernels(p)
    tot = sum(k["read"] + k["write"] for k in ks)
    cfgs[c] = {"so_hash": so_hash, "dram_bytes_per_env": tot / n, "envs": n,
               "capture": f"profiles/r02_ncu_c{c}.txt: " + " + ".join(f"{k['kernel']} ({k['duration_us']:.0f} us, dram read {k['read'] / 1e6:.1f} MB, write {k['write'] / 1e9:.3f} GB)" for k in ks)
               + "; ncu --set full --clock-control none of the bench command"}
if "3" in cfgs:
    cfgs["4"] = dict(cfgs["3"], capture="same kernel and frame size as config 3 (tc_render_classes_kernel<256,0> at 480x640): " + cfgs["3"]["capture"], envs=8192)
json.dump({"note": "DRAM traffic of the render kernel(s) of each BASELINE config from one ncu --set full capture; slightly BELOW the algorithmic "
                   "bytes where the tail of the output is still in the 126 MB L2 when the kernel ends", "configs": cfgs},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps({k: round(v["dram_bytes_per_env"]) for k, v in cfgs.items()}))
