"""Builds (CUDA env, CPU oracle env) pairs on the same config for the GPU parity tests."""
import numpy as np

from oracle import oracle as orc

CAR_SHIPPED = {"wheelbase": 0.0487, "track_width": 0.027, "max_velocity": 0.1, "max_steering_angle": 30, "steering_speed": 30,
               "max_acceleration": 0.1, "max_deceleration": 1.0}
CAM_SHIPPED = {"position": [0.0, -0.005, 0.04], "orientation": [22, 0, 0], "resolution": [128, 160], "fov": 80, "max_range": 0.5,
               "line_thickness": 2}
SPAWN_KNUFF = [156, 18, 217, 214, 325, 354, 176, 402, 339, 376, 385, 419, 396, 37, 149, 62, 240, 113, 98, 299, 2]
SPAWN_SIMPLE = [57, 143, 112, 121, 138, 157, 67, 46, 165, 124, 79, 33, 84, 21, 178, 7]
PPM = {"knuffingen": 222, "simple_layout": 450, "formula_student_track": 300, "formula_student_skidpad": 200}


def make_config(map_name, fmt, car=None, cam=None, spawn="default", fps=30):
    car_cfg = dict(CAR_SHIPPED)
    car_cfg.update(car or {})
    cam_cfg = {k: (list(v) if isinstance(v, list) else v) for k, v in CAM_SHIPPED.items()}
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
