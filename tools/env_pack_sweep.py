"""Small-frame render kernels side by side: tc_render_env_kernel (TC_ENV_PACK=0) against tc_render_envs_kernel with 1 / 2 / 4
envs per block, on BASELINE config 2 (simple_layout 84x84), the shipped Knuffingen 128x160 and Knuffingen 240x320. Each variant
runs in its own process (the choice is made when the handle is created). Prints per-step and per-kernel milliseconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("simple_layout", [84, 84], 4096, "random"), ("simple_layout", [84, 84], 65536, "random"), ("knuffingen", [128, 160], 32768, "stanley"),
         ("knuffingen", [240, 320], 8192, "stanley"), ("knuffingen", [84, 84], 32768, "stanley", 0.25)]
if os.environ.get("TC_SWEEP_CASES"):
    CASES = [CASES[int(i)] for i in os.environ["TC_SWEEP_CASES"].split(",")]

CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from tinycarlo_b200 import TinyCarloVecEnv
from tinycarlo_b200.config import make_config
case = json.loads(sys.argv[1])
m, res, n, pol = case[:4]
cam = {"resolution": res}
if len(case) > 4: cam["max_range"] = case[4]
cfg = make_config(m, "classes", cam=cam, car={"max_velocity": 0.15} if m == "simple_layout" else None)
env = TinyCarloVecEnv(cfg, n, device="cuda:0", autoreset="next_step")
env.reset(seed=0)
cc = torch.zeros((n, 2), device="cuda"); man = torch.zeros(n, dtype=torch.int32, device="cuda")
def one():
    if pol == "random":
        cc.uniform_(-1, 1); man.random_(0, 4)
    else:
        o = env.out
        cc[:, 0] = 0.8
        cc[:, 1] = (o["heading_error"] + torch.atan2(4 * o["cte"], torch.full_like(o["cte"], 0.8))) * (180 / np.pi / 30)
    env.step({"car_control": cc, "maneuver": man})
for _ in range(10): one()
torch.cuda.synchronize()
steps = 40
env.profile_begin(steps)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): one()
e1.record(); torch.cuda.synchronize()
k, ks = env.profile_end()
print(json.dumps({"ms_step": e0.elapsed_time(e1) / steps, "track": k["track"] / ks, "render": (k["project"] + k["raster"]) / ks, "cull": env.cull_info(), "ri": env.render_info(),
                  "sum": int(env.obs.sum(dtype=torch.int64).item() %% 1000003)}))
''' % ROOT

for case in CASES:
    for pack in sys.argv[1:] or ["0", "1", "2", "4"]:
        envv = dict(os.environ)
        if pack != "auto":
            envv["TC_ENV_PACK"] = pack
        if case[1][0] >= 240:
            envv["TC_FUSED_ALL"] = "1" if pack != "cls" else "0"
            if pack == "cls":
                envv.pop("TC_ENV_PACK")
        elif pack == "cls":
            continue
        r = subprocess.run([sys.executable, "-c", CHILD, json.dumps(case)], env=envv, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED " + r.stderr[-400:]
        try:
            d = json.loads(line)
            n = case[2]
            print(f"{case[0]:14s} {case[1]} N={n:6d} pack={pack:>3s}  step {d['ms_step']:.3f} ms  track {d['track']:.3f}  render {d['render']:.3f}  -> {n / d['ms_step'] / 1e3:7.2f} M env-steps/s"
                  f"  render-only {n / d['render'] / 1e3:7.2f} M  checksum {d['sum']}  cell nodes mean {d['cull']['mean_nodes']:.0f} max {d['cull']['max_nodes']} {d['ri']}", flush=True)
        except Exception:
            print(case, pack, line, flush=True)
