"""Parity soak: many envs x many steps of the CUDA path against the CPU oracle, counting every difference.
Usage: python tools/parity_soak.py [n_envs] [steps] [H] [W] [map]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
FMT = os.environ.get("TC_FMT", "classes")   # classes | classes_bits (decoded and compared with the same oracle frames)
import numpy as np, torch
from pair_util import make_config, oracle_env, stanley_actions
from tinycarlo_b200 import TinyCarloVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
W = int(sys.argv[4]) if len(sys.argv) > 4 else 160
mp = sys.argv[5] if len(sys.argv) > 5 else "knuffingen"
cfg = make_config(mp, "classes", cam={"resolution": [H, W]}, car={"max_velocity": 0.15} if mp == "simple_layout" else None)
env = TinyCarloVecEnv(cfg, n, device="cuda:0", obs_format=FMT)
FIELDS = ["cte", "heading_error", "velocity", "reward"] + [f"dist_{c}" for c in env.class_names]
per_field = {f: {"max_abs": 0.0, "max_rel": 0.0} for f in FIELDS}


def frames():
    o = env.obs.cpu().numpy()
    if FMT == "classes_bits":
        return np.unpackbits(o.view(np.uint32).view(np.uint8), axis=-1, bitorder="little")[..., : H * W].reshape(n, -1, H, W) * 255
    return o
oenv = oracle_env(cfg, n)
rng = np.random.default_rng(12345)
env.reset(seed=2024)
oenv.reset(env._spawn_nodes.cpu().numpy())
st = {"env_steps": 0, "bad_frames": 0, "bad_pixels": 0, "bad_index_rows": 0, "bad_flags": 0, "max_rel_state": 0.0, "max_rel_info": 0.0,
      "state_bit_equal_rows": 0, "resets": 0}
man = np.zeros(n, np.int32)
t0 = time.time()
for t in range(steps):
    cc = stanley_actions(oenv.cte.copy(), oenv.heading_error.copy(), cfg["car"]["max_steering_angle"])
    cc[:, 1] += rng.normal(0, 0.3, n).astype(np.float32)
    if t % 25 == 0:
        man = rng.integers(0, 4, n).astype(np.int32)
    if t % 40 == 20:
        cc[rng.random(n) < 0.05, 0] = -0.8
    env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
    oenv.step(cc.astype(np.float64), man)
    obs = frames()
    d = obs != oenv.obs
    bf = d.reshape(n, -1).any(axis=1)
    st["bad_frames"] += int(bf.sum()); st["bad_pixels"] += int(d.sum())
    s = env.state_dict(); sf = s["sf"].cpu().numpy(); si = s["si"].cpu().numpy()
    st["bad_index_rows"] += int((si[:, :10] != oenv.si[:, :10]).any(axis=1).sum()) + int((env.out["nearest_edge"].cpu().numpy() != oenv.nearest).any(axis=1).sum())
    st["bad_flags"] += int((env.out["terminated"].cpu().numpy() != oenv.terminated).sum() + (env.out["truncated"].cpu().numpy() != oenv.truncated).sum())
    st["state_bit_equal_rows"] += int((sf[:, :7] == oenv.sf[:, :7]).all(axis=1).sum())
    den = np.maximum(np.abs(oenv.sf[:, :7]), 1e-9)
    st["max_rel_state"] = max(st["max_rel_state"], float((np.abs(sf[:, :7] - oenv.sf[:, :7]) / den).max()))
    i64 = env.out["info_f64"].cpu().numpy()
    den = np.maximum(np.abs(oenv.info), 1e-6)
    st["max_rel_info"] = max(st["max_rel_info"], float((np.abs(i64 - oenv.info) / den).max()))
    for k, f in enumerate(FIELDS):   # per field: absolute error, and relative error where the value is not (almost) zero
        ad = np.abs(i64[:, k] - oenv.info[:, k])
        per_field[f]["max_abs"] = max(per_field[f]["max_abs"], float(ad.max()))
        big = np.abs(oenv.info[:, k]) > 1e-3
        if big.any():
            per_field[f]["max_rel"] = max(per_field[f]["max_rel"], float((ad[big] / np.abs(oenv.info[big, k])).max()))
    st["env_steps"] += n
    done = (oenv.terminated | oenv.truncated).astype(bool)
    if done.any():
        env.reset_done()
        oenv.reset(env._spawn_nodes.cpu().numpy(), mask=done)
        st["resets"] += int(done.sum())
        st["bad_frames"] += int((frames() != oenv.obs).reshape(n, -1).any(axis=1).sum())
st["seconds"] = round(time.time() - t0, 1)
st["info_error_per_field"] = per_field
st["max_rel_info_note"] = "max over all info fields of |gpu - oracle| / max(|oracle|, 1e-6): dominated by values near zero (see info_error_per_field: max_abs, and max_rel over |value| > 1e-3)"
st["kernels"] = env.render_info()
from tinycarlo_b200 import _lib
st["library_so_hash"] = _lib.build_info()
st["obs_format"] = FMT
st["config"] = f"{mp} {H}x{W} classes, {n} envs x {steps} steps, Stanley + noise, maneuver changes every 25 steps, 5% reverse bursts"
print(json.dumps(st))
