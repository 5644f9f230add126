"""Golden overview frames: runs the UNMODIFIED reference (imported from /root/reference, gymnasium stub as in gen_golden.py)
and records, for a few driven states, the car state and the image of Renderer.render_overview() (renderer.py:19-34).
Test infrastructure only; run in the build container:   python tests/golden/gen_overview.py   -> tests/golden/overview.npz"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import gym, make_config, ref_config, f32  # noqa: E402  (sets up the stub and the reference import)
from tinycarlo.renderer import Renderer  # noqa: E402


def main():
    out = {}
    meta = []
    for k, (map_name, ppm_view, bg, thick, names) in enumerate([("simple_layout", 150, None, 1, False), ("knuffingen", 100, (255, 255, 255), 2, True),
                                                               ("formula_student_skidpad", 60, (30, 60, 30), 1, False)]):
        cfg = make_config(map_name, "classes", spawn=None if "formula" in map_name else "default")
        env = gym.make("tinycarlo-v2", config=ref_config(cfg))
        base = env.unwrapped
        r = Renderer(base.map, base.car, ppm_view, bg, thick, names)
        rng = np.random.default_rng(k)
        env.reset(seed=5 + k)
        states, frames = [], []
        for t in range(40):
            act = {"car_control": [f32(rng.uniform(0.2, 1)), f32(rng.uniform(-1, 1))], "maneuver": int(rng.integers(0, 4))}
            _, _, term, trunc, _ = env.step(act)
            if term or trunc:
                env.reset()
            if t % 8 == 7:
                car = base.car
                lp = np.full((4, 2), -1, np.int32)
                lp[:len(car.local_path)] = np.array(car.local_path, np.int32).reshape(-1, 2)
                states.append([car.position[0], car.position[1], car.rotation, car.steering_angle])
                out[f"lp_{k}_{len(frames)}"] = lp
                frames.append(r.render_overview())
        out[f"static_{k}"] = Renderer(base.map, None, ppm_view, bg, thick, names).render_overview()
        out[f"states_{k}"] = np.array(states, np.float64)
        out[f"frames_{k}"] = np.packbits(np.stack(frames) > 0)          # occupancy bits keep the fixture small ...
        out[f"sums_{k}"] = np.array([int(f.astype(np.int64).sum()) for f in frames], np.int64)   # ... plus the pixel sums (colours)
        out[f"shape_{k}"] = np.array(np.stack(frames).shape, np.int64)
        meta.append({"config": cfg, "overview_pixel_per_meter": ppm_view, "background_color": bg, "line_thickness": thick, "node_names": names})
    out["meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(HERE, "overview.npz"), **out)
    print("wrote overview.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and k.startswith("frames")})


if __name__ == "__main__":
    main()
