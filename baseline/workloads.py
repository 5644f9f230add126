"""The five BASELINE.json configs as plain data, shared by both arms of bench.py (this repo's CUDA path and the CPU arms)
so that they describe — and print — the same workload. No product or oracle imports here.

  1  config_simple_layout.yaml, 480x640 rgb, random_control.py actions (the reference's own CPU-runnable case)
  2  simple_layout, 84x84 classes, random actions, 4096 envs in lockstep
  3  Knuffingen, 480x640 classes, Stanley controller on the lanepath CTE / heading info, 16384 envs   <- the metric's config
  4  Knuffingen, 480x640 classes, mixed maneuvers incl. u-turns every 100 steps, Stanley + OU noise (TD3-style rollout,
     examples/train_td3.py:42-44,143), CTESparseRewardWrapper(0.01) + CTETerminationWrapper (stanley_control.py:41-42),
     8192 envs per GPU (65536 on 8)
  5  Knuffingen, per-env camera pitch / fov / position, resolution from {84x84, 128x160, 240x320} in three groups, per-env car
     parameters (+-20 %), 32768 envs in total on any number of GPUs (strong scaling)"""
import numpy as np

CAR_KNUFF = {"wheelbase": 0.0487, "track_width": 0.027, "max_velocity": 0.1, "max_steering_angle": 30, "steering_speed": 30,
             "max_acceleration": 0.1, "max_deceleration": 1.0}
CAR_SIMPLE = dict(CAR_KNUFF, max_velocity=0.15)
CAM = {"position": [0.0, -0.005, 0.04], "orientation": [22, 0, 0], "resolution": [128, 160], "fov": 80, "max_range": 0.5, "line_thickness": 2}

WORKLOADS = {
    1: {"map": "simple_layout", "fmt": "rgb", "res": [480, 640], "car": CAR_SIMPLE, "policy": "random", "envs_per_gpu": 8192, "scaling": "weak",
        "wrappers": [], "metric": "env-steps/sec (480x640 rgb obs, simple_layout)",
        "workload": "simple_layout 480x640 rgb, random_control.py actions, auto-reset"},
    2: {"map": "simple_layout", "fmt": "classes", "res": [84, 84], "car": CAR_SIMPLE, "policy": "random", "envs_per_gpu": 4096, "scaling": "weak",
        "wrappers": [], "metric": "env-steps/sec (84x84 class obs, simple_layout)",
        "workload": "simple_layout 84x84 classes, random actions, auto-reset"},
    3: {"map": "knuffingen", "fmt": "classes", "res": [480, 640], "car": CAR_KNUFF, "policy": "stanley", "envs_per_gpu": 16384, "scaling": "weak",
        "wrappers": [], "metric": "env-steps/sec (480x640 class obs, Knuffingen)",
        "workload": "knuffingen 480x640 classes, Stanley actions + lanepath info, auto-reset"},
    4: {"map": "knuffingen", "fmt": "classes", "res": [480, 640], "car": CAR_KNUFF, "policy": "stanley_ou_mixed", "envs_per_gpu": 8192, "scaling": "weak",
        "wrappers": [("CTESparseRewardWrapper", {"min_cte": 0.01}), ("CTETerminationWrapper", {"max_cte": 0.07, "number_of_steps": 5})],
        "metric": "env-steps/sec (480x640 class obs, Knuffingen, mixed maneuvers + CTESparseRewardWrapper)",
        "workload": "knuffingen 480x640 classes, maneuvers uniform over {0,1,2,3} every 100 steps, Stanley + OU noise, CTESparseRewardWrapper + "
                    "CTETerminationWrapper, auto-reset, episode stats all-gathered every 100 steps"},
    5: {"map": "knuffingen", "fmt": "classes", "res": None, "groups": [[84, 84], [128, 160], [240, 320]], "car": CAR_KNUFF, "policy": "stanley",
        "envs_total": 32768, "scaling": "strong", "wrappers": [], "metric": "env-steps/sec (domain-randomised class obs 84x84/128x160/240x320, Knuffingen)",
        "workload": "knuffingen classes, per-env camera pitch [10,20) / fov [90,130) / position jitter / car parameters +-20 %, resolution groups "
                    "84x84 | 128x160 | 240x320 (a third of the envs each), Stanley actions, auto-reset"},
}
STANLEY_SPEED, STANLEY_K = 0.8, 4.0
OU_THETA, OU_SIGMA = 0.1, 0.4          # examples/train_td3.py:42-44
MANEUVER_PERIOD = 100


def obs_bytes(w, res=None):
    h, wd = res or w["res"]
    return (3 if w["fmt"] == "rgb" else 5) * h * wd


def group_sizes(total):
    """config 5: envs per resolution group (global, before sharding): about a third each, in multiples of 64 so that every rank of
    a 2 / 4 / 8-GPU job gets the same number of envs of every group"""
    a = total // 3
    if total >= 192:
        a = a // 64 * 64
    return [a, a, total - 2 * a]


def config5_params(total, seed=0):
    """Per-env domain randomisation of config 5 by GLOBAL env index (so that every sharding sees the same envs):
    camera pitch in [10,20) deg, fov in [90,130) deg (examples/train_stanley_il.py:53-54), camera position jitter of a few mm,
    wheelbase / max_velocity / max_steering_angle within +-20 %."""
    r = np.random.default_rng(seed)
    pitch = r.integers(10, 20, total).astype(np.float64)
    fov = r.integers(90, 130, total).astype(np.float64)
    pos = np.stack([r.uniform(-0.005, 0.005, total), -0.005 + r.uniform(-0.003, 0.003, total), 0.04 + r.uniform(-0.01, 0.01, total)], 1).round(4)
    car = {"wheelbase": (CAR_KNUFF["wheelbase"] * r.uniform(0.8, 1.2, total)).round(5),
           "max_velocity": (CAR_KNUFF["max_velocity"] * r.uniform(0.8, 1.2, total)).round(4),
           "max_steering_angle": (CAR_KNUFF["max_steering_angle"] * r.uniform(0.8, 1.2, total)).round(2)}
    return {"orientation": np.stack([pitch, np.zeros(total), np.zeros(total)], 1), "fov": fov, "position": pos, "car": car}
