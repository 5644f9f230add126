"""CPU arm of bench.py: the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py) driven through its own
public API — gym.make("tinycarlo-v2", config=...), env.step(action), the reference's wrappers — on the host cores.

The reference is a single-process, single-env Python loop (tinycarlo/env.py:115-147), so "all the host threads it can use" is
P independent processes, one env each, as a user would launch them; the single-process figure is reported next to the
aggregate. Actions are Python floats equal to float32 values (the float64 path); info is consumed every step."""
import math
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "tinycarlo")) and os.path.isdir(os.path.join(REF_DIR, "gymnasium"))


def _ref_config(w, res, seed_index):
    import numpy as np
    sys.path.insert(0, HERE)
    import workloads as W
    cam = dict(W.CAM, resolution=list(res))
    car = dict(w["car"])
    if w.get("groups"):   # config 5: this process plays global env `seed_index`
        p = W.config5_params(w["envs_total"])
        i = seed_index % w["envs_total"]
        cam.update(orientation=[float(v) for v in p["orientation"][i]], fov=float(p["fov"][i]), position=[float(v) for v in p["position"][i]])
        for k, v in p["car"].items():
            car[k] = float(v[i])
    spawn = {"knuffingen": [156, 18, 217, 214, 325, 354, 176, 402, 339, 376, 385, 419, 396, 37, 149, 62, 240, 113, 98, 299, 2],
             "simple_layout": [57, 143, 112, 121, 138, 157, 67, 46, 165, 124, 79, 33, 84, 21, 178, 7]}[w["map"]]
    ppm = {"knuffingen": 222, "simple_layout": 450}[w["map"]]
    return {"sim": {"fps": 30, "observation_space_format": w["fmt"]}, "car": car, "camera": cam,
            "map": {"json_path": os.path.join(REF_DIR, "examples", "maps", w["map"] + ".json"), "pixel_per_meter": ppm, "spawn_points": spawn}}


def _worker(idx, w, res, n_warm, n_steps, barrier, q):
    try:
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[k] = "1"      # one process per core: no BLAS / OpenCV thread oversubscription
        sys.path.insert(0, REF_DIR)
        import contextlib
        import io
        import numpy as np
        import cv2
        cv2.setNumThreads(1)
        import gymnasium as gym      # the stand-in installed next to the reference (baseline/install_ref.py)
        import tinycarlo             # noqa: F401  registers tinycarlo-v2
        from tinycarlo import wrapper as ref_wrappers
        sys.path.insert(0, HERE)
        import workloads as W
        with contextlib.redirect_stdout(io.StringIO()):
            env = gym.make("tinycarlo-v2", config=_ref_config(w, res, idx))
        for name, kw in w["wrappers"]:
            env = getattr(ref_wrappers, name)(env, **kw)
        rng = np.random.default_rng(idx)
        max_steer = env.unwrapped.config["car"]["max_steering_angle"]
        obs, info = env.reset(seed=idx)
        state = {"ou": 0.0, "man": 0, "t": 0}

        def one():
            nonlocal obs, info
            pol = w["policy"]
            if pol == "random":          # examples/random_control.py:11
                cc = [float(np.float32(rng.uniform(-1, 1))), float(np.float32(rng.uniform(-1, 1)))]
                man = int(rng.integers(0, 4))
            else:                        # examples/stanley_control.py:56-58
                steer = (info["heading_error"] + math.atan2(W.STANLEY_K * info["cte"], W.STANLEY_SPEED)) * 180 / math.pi / max_steer
                man = 0
                if pol == "stanley_ou_mixed":
                    if state["t"] % W.MANEUVER_PERIOD == 0:
                        state["man"] = int(rng.integers(0, 4))
                    state["ou"] += W.OU_THETA * (0.0 - state["ou"]) + W.OU_SIGMA * rng.standard_normal()
                    steer = min(1.0, max(-1.0, steer + state["ou"]))
                    man = state["man"]
                cc = [float(np.float32(W.STANLEY_SPEED)), float(np.float32(steer))]
            obs, reward, terminated, truncated, info = env.step({"car_control": cc, "maneuver": man})
            state["t"] += 1
            if terminated or truncated:
                obs, info = env.reset()
        for _ in range(n_warm):
            one()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            one()
        dt = time.perf_counter() - t0
        q.put((idx, n_steps, dt, None))
    except Exception as e:   # pragma: no cover
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((idx, 0, 0.0, repr(e)))


def run(w, procs, n_warm, n_steps, res=None):
    """`procs` processes x n_steps env-steps each (after n_warm untimed ones), started together. Returns
    (aggregate env-steps/s = procs * n_steps / slowest process, slowest seconds, per-process env-steps/s list)."""
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(procs), ctx.Queue()
    reses = [res or w["res"] or w["groups"][i % len(w["groups"])] for i in range(procs)]
    ps = [ctx.Process(target=_worker, args=(i, w, reses[i], n_warm, n_steps, barrier, q)) for i in range(procs)]
    for p in ps:
        p.start()
    import queue
    out, deadline = [], time.time() + 1800
    while len(out) < len(ps):
        try:
            out.append(q.get(timeout=1.0))
        except queue.Empty:
            dead = [p for p in ps if p.exitcode not in (None, 0)]
            if dead or time.time() > deadline:   # a worker died before reporting (or the sample ran away): do not hang
                for p in ps:
                    if p.is_alive():
                        p.terminate()
                raise RuntimeError(f"reference worker exited with code {dead[0].exitcode}" if dead else "reference workers timed out")
    for p in ps:
        p.join(timeout=60)
    errs = [o[3] for o in out if o[3]]
    if errs:
        raise RuntimeError("reference worker failed: " + errs[0])
    slow = max(o[2] for o in out)
    return procs * n_steps / slow, slow, [o[1] / o[2] for o in out]


def versions():
    import cv2
    import numpy as np
    return {"numpy": np.__version__, "cv2": cv2.__version__, "python": sys.version.split()[0]}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"
