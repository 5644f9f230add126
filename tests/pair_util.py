"""Builds (CUDA env, CPU oracle env) pairs on the same config for the GPU parity tests."""
import numpy as np

from oracle import oracle as orc

from tinycarlo_b200.config import CAM_SHIPPED, CAR_SHIPPED, PPM, SPAWN_KNUFF, SPAWN_SIMPLE, make_config  # noqa: E402,F401


def oracle_env(cfg, n, wrapped=False, cam_rows=None, thickness=None, car_rows=None):
    omap = orc.load_named_map(cfg["map"]["map_name"], cfg["map"]["pixel_per_meter"], cfg["map"].get("spawn_points"))
    cc = cfg["camera"]
    H, W = cc["resolution"]
    if cam_rows is None:
        E, K = orc.camera_matrices(cc["position"], cc["orientation"], cc["fov"], [H, W])
        cam_rows = orc.pack_cam(E, K, cc["max_range"])
    if car_rows is None:
        car_rows = orc.pack_car(cfg["car"], cfg["sim"].get("fps", 30))
    if thickness is None:
        thickness = cc["line_thickness"]
    return orc.OracleVecEnv(omap, n, car_rows, cam_rows, thickness, H, W, cfg["sim"]["observation_space_format"], wrapped=wrapped)


def stanley_actions(cte, heading, max_steer_deg, speed=0.8, k=4.0):
    """examples/stanley_control.py:56-58 on numpy arrays (float32-representable outputs)."""
    steer = (heading + np.arctan2(k * cte, speed)) * 180 / np.pi / max_steer_deg
    cc = np.stack([np.full_like(steer, speed), steer], axis=1).astype(np.float32)
    return cc
