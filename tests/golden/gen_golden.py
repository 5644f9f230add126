"""Golden-vector generator: runs the UNMODIFIED reference (imported from /root/reference) and records traces.

Test infrastructure only. Run in the build container (the reference cannot travel to the GPU box):

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz

gymnasium is not installed in this image, so tests/golden/gym_stub provides the few names the reference
imports (seeding identical to gymnasium >= 0.26). Everything else (numpy 2.3.5, cv2 4.13.0) is the real thing.

Actions are passed as Python floats equal to float32 values (with float32 ndarrays numpy>=2 demotes the
reference's state to float32; Python floats keep the float64 path, which is what we match).

What is recorded per scenario (one .npz each):
  meta            json: config (map by name), wrappers, policy, versions
  act_cc[T,2] f64, act_man[T] i32              the actions fed at step t
  ev_kind[F] i8   0 = frame after reset, 1 = frame after step; ev_step[F] (step index, -1 for the first reset)
  reset_seed[F]   seed passed to reset (or -1 for reset() continuing the RNG stream / not a reset)
  spawn_node[F]   lanepath node drawn by that reset (-1 if not a reset)
  pos[F,2] rot[F] vel[F] steer[F] front[F,2]   car state when the frame was captured (f64)
  lp[F,4,2] i32 (-1 padded), lp_len[F], last_man[F]
  cte heading dist[F,C] velocity reward terminated truncated    (info / step results; zeros for reset frames)
  seg_off[F,C+1] i64, seg_i32[M,4] i32, seg_f64[M,4] f64        per-frame per-class projected segments (x0,y0,x1,y1)
  cls_bits[F, C*H*W/8] u8   np.packbits of the classes frame (>0)  (when classes are rendered)
  rgb_sha[F] S64, rgb[Fk,H,W,3] u8 + rgb_idx[Fk]                    RGB frames: sha256 for all, pixels for a subset
"""
import hashlib
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TINYCARLO_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "gym_stub"))
sys.path.insert(0, REF)

import cv2  # noqa: E402
import gymnasium as gym  # noqa: E402  (the stub)
import tinycarlo  # noqa: E402,F401  (registers tinycarlo-v2)
from tinycarlo.renderer import Renderer  # noqa: E402
from tinycarlo import wrapper as ref_wrappers  # noqa: E402

MAPS = os.path.join(REF, "examples", "maps")

CAR_SHIPPED = {"wheelbase": 0.0487, "track_width": 0.027, "max_velocity": 0.1, "max_steering_angle": 30,
               "steering_speed": 30, "max_acceleration": 0.1, "max_deceleration": 1.0}
CAM_SHIPPED = {"position": [0.0, -0.005, 0.04], "orientation": [22, 0, 0], "resolution": [128, 160], "fov": 80,
               "max_range": 0.5, "line_thickness": 2}
SPAWN_KNUFF = [156, 18, 217, 214, 325, 354, 176, 402, 339, 376, 385, 419, 396, 37, 149, 62, 240, 113, 98, 299, 2]
SPAWN_SIMPLE = [57, 143, 112, 121, 138, 157, 67, 46, 165, 124, 79, 33, 84, 21, 178, 7]
PPM = {"knuffingen": 222, "simple_layout": 450, "formula_student_track": 300, "formula_student_skidpad": 200}


def make_config(map_name, fmt, car=None, cam=None, spawn="default", fps=30):
    car_cfg = dict(CAR_SHIPPED)
    car_cfg.update(car or {})
    cam_cfg = json.loads(json.dumps(CAM_SHIPPED))
    cam_cfg.update(cam or {})
    map_cfg = {"map_name": map_name, "pixel_per_meter": PPM[map_name]}
    if spawn == "default":
        if map_name == "knuffingen":
            map_cfg["spawn_points"] = SPAWN_KNUFF
        elif map_name == "simple_layout":
            map_cfg["spawn_points"] = SPAWN_SIMPLE
    elif spawn is not None:
        map_cfg["spawn_points"] = list(spawn)
    return {"sim": {"fps": fps, "observation_space_format": fmt}, "car": car_cfg, "camera": cam_cfg, "map": map_cfg}


def ref_config(cfg):
    c = json.loads(json.dumps(cfg))
    c["map"]["json_path"] = os.path.join(MAPS, c["map"].pop("map_name") + ".json")
    return c


def f32(x):
    return float(np.float32(x))


class Recorder:
    """Hooks Renderer.render_camera_frame_rgb (always called first by Camera.capture_frame, camera.py:104)
    to grab the projected point pairs of the frame being drawn."""

    def __init__(self):
        self.last_points = None
        self._orig = Renderer.render_camera_frame_rgb
        rec = self

        def hooked(self_r, points, colors, resolution, line_thickness):
            rec.last_points = [[(np.array(a, dtype=np.float64), np.array(b, dtype=np.float64)) for a, b in layer]
                               for layer in points]
            return rec._orig(self_r, points, colors, resolution, line_thickness)

        Renderer.render_camera_frame_rgb = hooked

    def restore(self):
        Renderer.render_camera_frame_rgb = self._orig


def run_scenario(name, cfg, n_steps, policy, seed, wrappers=(), reset_mode="continue", rgb_keep_every=25,
                 cam_mutations=None, action_dtype="pyfloat"):
    """action_dtype: "pyfloat" feeds car_control as a list of Python floats (the float64 path every bit-exact test uses);
    "float32" feeds np.float32 ndarrays - what action_space.sample() yields - which makes numpy >= 2 keep parts of the
    reference's state in float32: recorded for the tolerance test of that case only."""
    rec = Recorder()
    try:
        env = gym.make("tinycarlo-v2", config=ref_config(cfg))
        for wname, wkw in wrappers:
            env = getattr(ref_wrappers, wname)(env, **wkw)
        base = env.unwrapped
        C = len(base.map.get_laneline_names())
        H, W = base.camera.resolution
        fmt = base.observation_space_format
        F = {k: [] for k in ("ev_kind", "ev_step", "reset_seed", "spawn_node", "pos", "rot", "vel", "steer", "front",
                             "lp", "lp_len", "last_man", "cte", "heading", "dist", "velocity", "reward", "terminated",
                             "truncated", "seg_off", "cls_bits", "rgb_sha", "E", "K")}
        seg_i32, seg_f64, rgb_frames, rgb_idx = [], [], [], []
        act_cc, act_man = [], []

        def record(kind, step, rseed, obs, info, reward=0.0, terminated=False, truncated=False):
            car = base.car
            f = len(F["ev_kind"])
            F["ev_kind"].append(kind)
            F["ev_step"].append(step)
            F["reset_seed"].append(-1 if rseed is None else rseed)
            F["spawn_node"].append(car.local_path[0][0] if kind == 0 else -1)
            F["pos"].append([float(car.position[0]), float(car.position[1])])
            F["rot"].append(float(car.rotation))
            F["vel"].append(float(car.velocity))
            F["steer"].append(float(car.steering_angle))
            F["front"].append([float(car.position_front[0]), float(car.position_front[1])])
            lp = np.full((4, 2), -1, np.int32)
            for i, e in enumerate(car.local_path[:4]):
                lp[i] = (int(e[0]), int(e[1]))
            F["lp"].append(lp)
            F["lp_len"].append(len(car.local_path))
            F["last_man"].append(int(car.last_maneuver))
            F["cte"].append(float(info["cte"]))
            F["heading"].append(float(info["heading_error"]))
            F["dist"].append([float(v) for v in info["laneline_distances"].values()])
            F["velocity"].append(float(info["velocity"]))
            F["reward"].append(float(reward))
            F["terminated"].append(bool(terminated))
            F["truncated"].append(bool(truncated))
            F["E"].append(np.array(base.camera.E, dtype=np.float64))
            F["K"].append(np.array(base.camera.K, dtype=np.float64))
            off = [len(seg_i32)]
            for layer in rec.last_points:
                for a, b in layer:
                    line = (a, b)
                    with np.errstate(invalid="ignore"):
                        q = np.int32([line])  # exactly the cast of renderer.py:43,50
                    seg_i32.append([q[0, 0, 0], q[0, 0, 1], q[0, 1, 0], q[0, 1, 1]])
                    seg_f64.append([a[0], a[1], b[0], b[1]])
                off.append(len(seg_i32))
            F["seg_off"].append(off)
            rgb = base.camera.get_last_frame_rgb()
            F["rgb_sha"].append(hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).hexdigest())
            if fmt == "classes":
                cls = base.camera.get_last_frame_classes()
                assert cls.shape == (C, H, W) and set(np.unique(cls)).issubset({0, 255})
                assert np.array_equal(obs, cls)
                F["cls_bits"].append(np.packbits(cls.reshape(-1) > 0))
            else:
                assert np.array_equal(obs, rgb)
            if f % rgb_keep_every == 0:
                rgb_frames.append(rgb.copy())
                rgb_idx.append(f)

        obs, info = env.reset(seed=seed)
        record(0, -1, seed, obs, info)
        rng = np.random.default_rng(seed + 1000)
        pol_state = {"noise": 0.0, "maneuver": 0}
        for t in range(n_steps):
            if cam_mutations and t in cam_mutations:
                cam = base.camera
                for k, v in cam_mutations[t].items():
                    setattr(cam, k, v)
                cam.update_params()
            cc, man = policy(t, info, rng, cfg, pol_state)
            cc = [f32(cc[0]), f32(cc[1])]
            act_cc.append(cc)
            act_man.append(int(man))
            fed = np.array(cc, dtype=np.float32) if action_dtype == "float32" else cc
            obs, reward, terminated, truncated, info = env.step({"car_control": fed, "maneuver": int(man)})
            record(1, t, None, obs, info, reward, terminated, truncated)
            if terminated or truncated:
                if reset_mode == "continue":
                    obs, info = env.reset()
                    record(0, t, None, obs, info)
                else:
                    rs = seed + 7919 * (t + 1)
                    obs, info = env.reset(seed=rs)
                    record(0, t, rs, obs, info)
        out = {k: np.array(v) for k, v in F.items() if len(v)}
        out["ev_kind"] = out["ev_kind"].astype(np.int8)
        out["seg_off"] = out["seg_off"].astype(np.int64)
        out["seg_i32"] = np.array(seg_i32, dtype=np.int32).reshape(-1, 4)
        out["seg_f64"] = np.array(seg_f64, dtype=np.float64).reshape(-1, 4)
        out["rgb_sha"] = np.array(F["rgb_sha"], dtype="S64")
        out["rgb"] = np.array(rgb_frames, dtype=np.uint8)
        out["rgb_idx"] = np.array(rgb_idx, dtype=np.int32)
        out["act_cc"] = np.array(act_cc, dtype=np.float64).reshape(-1, 2)
        out["act_man"] = np.array(act_man, dtype=np.int32)
        meta = {"name": name, "config": cfg, "wrappers": list(wrappers), "seed": seed, "reset_mode": reset_mode,
                "class_names": base.map.get_laneline_names(), "numpy": np.__version__, "cv2": cv2.__version__,
                "n_steps": n_steps,
                "cam_mutations": {str(k): v for k, v in (cam_mutations or {}).items()}}
        if action_dtype != "pyfloat":   # (only recorded when it is not the default, so that the older fixtures regenerate byte for byte)
            meta["action_dtype"] = action_dtype
        out["meta"] = np.array(json.dumps(meta))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        nres = int((out["ev_kind"] == 0).sum())
        print(f"{name}: {len(out['ev_kind'])} frames ({nres} resets), {len(seg_i32)} segments, "
              f"{os.path.getsize(path) / 1024:.0f} KiB")
    finally:
        rec.restore()


# ------------------------------------------------------------------ policies (workload drivers of the examples)
def pol_random(t, info, rng, cfg, st):
    """examples/random_control.py:11 — uniform car_control in [-1,1]^2, uniform maneuver."""
    return [rng.uniform(-1, 1), rng.uniform(-1, 1)], int(rng.integers(0, 4))


def pol_random_forward(t, info, rng, cfg, st):
    return [rng.uniform(0.2, 1), rng.uniform(-1, 1)], int(rng.integers(0, 4))


def make_stanley(speed=0.8, k=4.0, maneuver=0, noise_sigma=0.0, switch_every=0, reverse_prob=0.0):
    """examples/stanley_control.py:56-58 (+ Ornstein-Uhlenbeck noise of train_stanley_il.py:64)."""

    def pol(t, info, rng, cfg, st):
        if switch_every and t % switch_every == 0:
            st["maneuver"] = int(rng.integers(0, 4))
        man = st["maneuver"] if switch_every else maneuver
        cte, he = info["cte"], info["heading_error"]
        steer = (he + math.atan2(k * cte, speed)) * 180 / math.pi / cfg["car"]["max_steering_angle"]
        if noise_sigma:
            st["noise"] += 0.1 * (0.0 - st["noise"]) + noise_sigma * rng.standard_normal()
            steer += st["noise"]
        v = speed
        if reverse_prob and rng.uniform() < reverse_prob:
            v = -speed
        return [v, steer], man

    return pol


def main():
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    # A. Appendix-B smoke: shipped simple_layout config (rgb 128x160, max_velocity 0.15), seed 2, maneuver 3
    if want("simple_rgb_smoke"):
        run_scenario("simple_rgb_smoke", make_config("simple_layout", "rgb", car={"max_velocity": 0.15}), 40,
                     lambda t, i, r, c, s: ([0.4, 0.1], 3), seed=2, rgb_keep_every=4)
    # B. headline comparator: Knuffingen 480x640 classes, Stanley, maneuver 0 (SURVEY Appendix B values)
    if want("knuff_480_stanley"):
        run_scenario("knuff_480_stanley", make_config("knuffingen", "classes", cam={"resolution": [480, 640]}), 60,
                     make_stanley(0.8, 4.0, 0), seed=0, rgb_keep_every=1000)
    # C. Knuffingen 128x160 classes, mixed maneuvers incl. u-turn, OU noise, reverse commands, resets continue the RNG
    if want("knuff_mixed"):
        run_scenario("knuff_mixed", make_config("knuffingen", "classes"), 700,
                     make_stanley(0.8, 4.0, 0, noise_sigma=0.4, switch_every=40, reverse_prob=0.05), seed=11,
                     rgb_keep_every=100)
    # D. config-2 analog: simple_layout 84x84 classes, random control, resets
    if want("simple_84_random"):
        run_scenario("simple_84_random", make_config("simple_layout", "classes", car={"max_velocity": 0.15},
                                                     cam={"resolution": [84, 84]}), 600, pol_random, seed=5,
                     rgb_keep_every=100)
    # E. config-1: simple_layout 480x640 rgb, random control
    if want("simple_480_rgb_random"):
        run_scenario("simple_480_rgb_random", make_config("simple_layout", "rgb", car={"max_velocity": 0.15},
                                                          cam={"resolution": [480, 640]}), 120, pol_random_forward,
                     seed=0, rgb_keep_every=10)
    # F. thickness / camera-parameter variants (domain randomisation, train_stanley_il.py:52-57)
    for th in (1, 3, 6):
        if want(f"knuff_thick{th}"):
            run_scenario(f"knuff_thick{th}", make_config("knuffingen", "classes",
                                                         cam={"resolution": [96, 128], "line_thickness": th}), 120,
                         make_stanley(0.8, 4.0, 3, noise_sigma=0.2), seed=20 + th, rgb_keep_every=30)
    if want("knuff_camrand"):
        r = np.random.default_rng(99)
        muts = {}
        for t in range(0, 400, 10):
            muts[t] = {"orientation": [int(r.integers(5, 40)), int(r.integers(-5, 6)), int(r.integers(-30, 31))],
                       "fov": int(r.integers(60, 130)), "max_range": float(np.round(r.uniform(0.3, 2.0), 3)),
                       "position": [float(np.round(r.uniform(-0.01, 0.02), 4)), float(np.round(r.uniform(-0.01, 0.01), 4)),
                                    float(np.round(r.uniform(0.02, 0.08), 4))]}
        run_scenario("knuff_camrand", make_config("knuffingen", "classes", cam={"resolution": [120, 160]}), 400,
                     make_stanley(0.8, 4.0, 0, noise_sigma=0.3, switch_every=50), seed=3, cam_mutations=muts,
                     rgb_keep_every=100)
    # G. no rate limits (car defaults: steering_speed/max_acceleration None), spawn_points None, formula maps
    if want("fs_track_free"):
        run_scenario("fs_track_free", make_config("formula_student_track", "classes",
                                                  car={"steering_speed": None, "max_acceleration": None,
                                                       "max_deceleration": None, "max_velocity": 0.5},
                                                  cam={"resolution": [64, 96], "max_range": 1.5}, spawn=None), 300,
                     make_stanley(0.6, 4.0, 0, noise_sigma=0.3), seed=1, rgb_keep_every=100)
    if want("fs_skidpad_rgb"):
        run_scenario("fs_skidpad_rgb", make_config("formula_student_skidpad", "rgb",
                                                   cam={"resolution": [60, 80], "max_range": 1.0, "line_thickness": 3},
                                                   spawn=None), 200, pol_random_forward, seed=4, rgb_keep_every=20)
    # H. wrappers (stanley_control.py:41-43 stack, plus the laneline ones)
    if want("knuff_wrapped_cte"):
        run_scenario("knuff_wrapped_cte", make_config("knuffingen", "classes", cam={"resolution": [32, 48]}), 500,
                     make_stanley(0.8, 4.0, 0, noise_sigma=0.5, switch_every=60), seed=8,
                     wrappers=[("CTESparseRewardWrapper", {"min_cte": 0.01}),
                               ("CTETerminationWrapper", {"max_cte": 0.07, "number_of_steps": 5}),
                               ("CrashTerminationWrapper", {})], rgb_keep_every=1000)
    if want("knuff_wrapped_lane"):
        run_scenario("knuff_wrapped_lane", make_config("knuffingen", "classes", cam={"resolution": [32, 48]}), 500,
                     make_stanley(0.8, 4.0, 3, noise_sigma=0.5, switch_every=60), seed=9,
                     wrappers=[("CTELinearRewardWrapper", {"min_cte": 0.03, "max_reward": 1.0, "min_reward": -0.5}),
                               ("LanelineSparseRewardWrapper", {"sparse_rewards": {"outer": -10.0, "solid": -5.0}}),
                               ("LanelineLinearRewardWrapper", {"max_rewards": {"outer": -1.0, "dashed": 0.5,
                                                                                "solid": -0.5, "hold": 0.0,
                                                                                "area": 0.25}}),
                               ("LanelineCrossingTerminationWrapper", {"lanelines": ["outer", "solid"]})],
                     reset_mode="reseed", rgb_keep_every=1000)
    # J. float32 ndarray actions (action_space.sample() dtype): the reference then computes parts of the car state in float32
    if want("knuff_f32_actions"):
        run_scenario("knuff_f32_actions", make_config("knuffingen", "classes", cam={"resolution": [64, 96]}), 200,
                     make_stanley(0.8, 4.0, 0, noise_sigma=0.3, switch_every=50), seed=13, rgb_keep_every=1000,
                     action_dtype="float32")
    # I. spawn RNG parity: many seeds, draws per seed (map.py:51-69), with and without spawn_points
    if want("spawn_draws"):
        res = {}
        for map_name, spawn in (("knuffingen", "default"), ("knuffingen", None), ("simple_layout", "default"),
                                ("simple_layout", None)):
            cfg = make_config(map_name, "classes", cam={"resolution": [8, 8]}, spawn=spawn)
            env = gym.make("tinycarlo-v2", config=ref_config(cfg))
            draws = np.zeros((64, 12), np.int32)
            for s in range(64):
                env.reset(seed=s)
                draws[s, 0] = env.unwrapped.car.local_path[0][0]
                for k in range(1, 12):
                    env.reset()
                    draws[s, k] = env.unwrapped.car.local_path[0][0]
            res[f"{map_name}_{'default' if spawn else 'none'}"] = draws
        np.savez_compressed(os.path.join(HERE, "spawn_draws.npz"), **res)
        print("spawn_draws:", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
