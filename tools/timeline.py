"""Per-block timeline of the fused render kernel (tc_debug_set_timeline): phase durations and per-SM overlap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import subprocess
import numpy as np, torch, ctypes as C
from pair_util import make_config
from tinycarlo_b200 import _lib
# the timeline counters are compiled in only with -DTC_TIMELINE: build that variant next to the product library
VARIANT = os.path.join(_lib.LIB_DIR, "variant_timeline.so")
subprocess.check_call(["nvcc"] + _lib.NVCC_FLAGS + ["-DTC_TIMELINE", "-o", VARIANT, os.path.join(_lib.CSRC, "tc_api.cu")])
_lib.LIB_PATH = VARIANT
from tinycarlo_b200 import TinyCarloVecEnv

# usage: timeline.py [N] [map] [H] [W]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
MAP = sys.argv[2] if len(sys.argv) > 2 else "knuffingen"
RES = [int(sys.argv[3]), int(sys.argv[4])] if len(sys.argv) > 4 else [480, 640]
cfg = make_config(MAP, "classes", cam={"resolution": RES})
print(f"--- {MAP} {RES} N={N}")
env = TinyCarloVecEnv(cfg, N, device="cuda:0", autoreset="next_step")
env.reset(seed=0)
cc = torch.zeros((N, 2), device="cuda"); cc[:, 0] = 0.8
man = torch.zeros(N, dtype=torch.int32, device="cuda")
for _ in range(3):
    env.step({"car_control": cc, "maneuver": man})
tl = torch.zeros((N * env.n_classes, 10), dtype=torch.int64, device="cuda")
_lib.check(env._L.tc_debug_set_timeline(env._h, C.c_void_p(tl.data_ptr())), "timeline")
env.step({"car_control": cc, "maneuver": man})
torch.cuda.synchronize()
_lib.check(env._L.tc_debug_set_timeline(env._h, None), "timeline")
t = tl.cpu().numpy()
t = t[t[:, 5] != 0]   # small frames: one block per env (all classes), the other rows stay empty
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "timeline.npy"), t)
d = np.diff(t[:, 1:6], axis=1)
names = ["tma wait", "geometry", "raster", "store"]
for i, nm in enumerate(names):
    print(f"{nm:10s} cycles: mean {d[:, i].mean():9.0f}  p50 {np.median(d[:, i]):9.0f}  p90 {np.percentile(d[:, i], 90):9.0f}  max {d[:, i].max():9.0f}")
has = t[:, 6] > 0
print("planes with segments: %.3f, mean segs %.1f max %d" % (has.mean(), t[has, 6].mean(), t[:, 6].max()))
for nm, col in (("zero", 7), ("setup", 8), ("draw", 9)):
    v = t[has, col]
    print(f"  {nm:6s} (planes with segs) mean {v.mean():9.0f} p50 {np.median(v):9.0f} p90 {np.percentile(v, 90):9.0f} max {v.max():9.0f}")
print("  raster of planes with segs mean", d[has, 2].mean(), " without", d[~has, 2].mean())
print("block total mean", (t[:, 5] - t[:, 1]).mean())
# per SM: fraction of the SM's busy span during which at least one resident block is in its store phase
fr = []
for sm in np.unique(t[:, 0]):
    b = t[t[:, 0] == sm]
    lo, hi = b[:, 1].min(), b[:, 5].max()
    ev = sorted([(x, 1) for x in b[:, 4]] + [(x, -1) for x in b[:, 5]])
    cur, last, cov = 0, lo, 0
    for x, s in ev:
        if cur > 0:
            cov += x - last
        last = x
        cur += s
    fr.append(cov / (hi - lo))
print("fraction of SM time with >=1 block storing: mean %.3f min %.3f" % (np.mean(fr), np.min(fr)), " blocks/SM", len(t) / len(fr))
