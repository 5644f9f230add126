"""Device-loop timing of the other BASELINE.json configs (parity-test cases, not bench lines): prints one line each."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pair_util import make_config
from tinycarlo_b200 import TinyCarloVecEnv


def run(name, cfg, n, steps=30, policy="random", per_env=False, graph=False):
    env = TinyCarloVecEnv(cfg, n, device="cuda:0", autoreset="next_step")
    g = torch.Generator(device="cuda").manual_seed(0)
    if per_env:
        rng = np.random.default_rng(0)
        env.set_camera_params(orientation=np.stack([rng.uniform(10, 20, n), np.zeros(n), np.zeros(n)], 1).round(1),
                              fov=rng.integers(90, 130, n).astype(float))
        env.set_car_params(wheelbase=rng.uniform(0.04, 0.06, n), max_velocity=rng.uniform(0.08, 0.12, n))
    env.reset(seed=0)
    cc = torch.zeros((n, 2), device="cuda"); man = torch.zeros(n, dtype=torch.int32, device="cuda")

    def one():
        if policy == "random":
            cc.copy_(2 * torch.rand((n, 2), device="cuda", generator=g) - 1)
            man.copy_(torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32))
        else:
            o = env.out
            cc[:, 0] = 0.8
            cc[:, 1] = (o["heading_error"] + torch.atan2(4 * o["cte"], torch.full_like(o["cte"], 0.8))) * (180 / np.pi / 30)
        env.step({"car_control": cc, "maneuver": man})
    for _ in range(5):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ob = int(np.prod(env.obs_shape))
    print(f"{name:55s} N={n:6d} obs={ob:8d} B  {ms:8.3f} ms/step  {n / ms * 1e3 / 1e6:8.2f} M env-steps/s  {n * ob / ms / 1e6:8.1f} GB/s obs")
    env.close()


run("config2 simple_layout 84x84 classes random", make_config("simple_layout", "classes", cam={"resolution": [84, 84]}, car={"max_velocity": 0.15}), 4096)
run("config2 x16 envs", make_config("simple_layout", "classes", cam={"resolution": [84, 84]}, car={"max_velocity": 0.15}), 65536)
run("config1-like simple_layout 480x640 rgb random", make_config("simple_layout", "rgb", cam={"resolution": [480, 640]}, car={"max_velocity": 0.15}), 8192)
run("config3 knuffingen 480x640 classes stanley", make_config("knuffingen", "classes", cam={"resolution": [480, 640]}), 16384, policy="stanley")
run("config5-like knuffingen 128x160 per-env params", make_config("knuffingen", "classes", cam={"resolution": [128, 160]}), 32768, policy="stanley", per_env=True)
run("knuffingen 128x160 classes (shipped resolution)", make_config("knuffingen", "classes"), 32768, policy="stanley")


def run_fmt(fmt):
    cfg = make_config("knuffingen", "classes", cam={"resolution": [480, 640]})
    n = 16384
    env = TinyCarloVecEnv(cfg, n, device="cuda:0", autoreset="next_step", obs_format=fmt)
    env.reset(seed=0)
    cc = torch.zeros((n, 2), device="cuda"); cc[:, 0] = 0.8; man = torch.zeros(n, dtype=torch.int32, device="cuda")
    for _ in range(5):
        env.step({"car_control": cc, "maneuver": man})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        env.step({"car_control": cc, "maneuver": man})
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nb = env.obs[0].numel() * env.obs.element_size()
    print(f"knuffingen 480x640 obs_format={fmt:14s} N={n} obs={nb:8d} B {ms:8.3f} ms/step {n / ms * 1e3 / 1e6:8.2f} M env-steps/s {n * nb / ms / 1e6:8.1f} GB/s")
    env.close()


for f in ("classes", "classes_bits", "classes_bf16"):
    run_fmt(f)
