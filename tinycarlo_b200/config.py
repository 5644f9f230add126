"""Configuration loading with the reference's schema and defaults (tinycarlo/env.py:27-45, car.py:12-18,
camera.py:16-21, map.py:13-17). A config is a path to a .yaml file, a directory holding config.yaml, or a dict with
the mandatory sections sim / car / camera / map (a missing section raises KeyError, as in the reference)."""
import math
import os
from typing import Any, Dict, Optional, Tuple, Union

MAPS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps")


def load_config(config: Optional[Union[str, Dict[str, Any]]]) -> Tuple[Dict[str, Any], Optional[str]]:
    """Returns (config dict, absolute path of the yaml file or None). env.py:27-35."""
    config_path = None
    if isinstance(config, str):
        import yaml
        if config.endswith(".yaml"):
            config_path = os.path.abspath(config)
        else:
            config_path = os.path.abspath(os.path.join(config, "config.yaml"))
        with open(config_path, "r") as stream:
            config = yaml.safe_load(stream)
    if config is None:
        raise TypeError("config must be a yaml path, a directory with config.yaml, or a dict")
    for section in ("sim", "car", "camera", "map"):
        config[section]  # KeyError like the reference
    return config, config_path


def resolve_map_path(map_config: Dict[str, Any], config_path: Optional[str]) -> str:
    """map.py:15-16: json_path is relative to the yaml's directory, or to the CWD for dict configs. As an extension,
    {"map_name": "knuffingen"} selects one of the maps bundled in tinycarlo_b200/maps/."""
    if "map_name" in map_config and "json_path" not in map_config:
        return os.path.join(MAPS_DIR, map_config["map_name"] + ".json")
    base = "./" if config_path is None else os.path.dirname(config_path)
    return os.path.join(base, map_config["json_path"])


def sim_params(config: Dict[str, Any]) -> Dict[str, Any]:
    sim = config["sim"]
    fps = sim.get("fps", 30)
    return {"fps": fps, "T": 1 / fps, "observation_space_format": sim.get("observation_space_format", "rgb"),
            "render_realtime": sim.get("render_realtime", False),
            "overview_pixel_per_meter": sim.get("overview_pixel_per_meter", 150),
            "render_node_names": sim.get("render_node_names", False)}


def car_param_row(car_config: Dict[str, Any], T: float):
    """One row of the per-env car parameter table (include/tinycarlo_b200.h TC_CP_*); None -> NaN (car.py:12-18)."""
    def g(key, default):
        v = car_config.get(key, default)
        return math.nan if v is None else float(v)
    if car_config.get("max_acceleration", None) is not None and car_config.get("max_deceleration", None) is None:
        # car.py:82 would raise TypeError on `None * dt`
        raise TypeError("car.max_deceleration must be set when car.max_acceleration is set")
    return [g("wheelbase", 0.08), g("track_width", 0.03), g("max_velocity", 1), g("max_steering_angle", 35),
            g("steering_speed", None), g("max_acceleration", None), g("max_deceleration", None), float(T)]


def camera_params(camera_config: Dict[str, Any]) -> Dict[str, Any]:
    """camera.py:16-21 defaults. max_range=None crashes the reference on the first frame (camera.py:80-82, SURVEY
    section 5); it is rejected here at construction."""
    cam = {"resolution": list(camera_config.get("resolution", [128, 160])),
           "position": list(camera_config.get("position", [0, 0, 0])),
           "orientation": list(camera_config.get("orientation", [0, 0, 0])),
           "fov": camera_config.get("fov", 90),
           "max_range": camera_config.get("max_range", None),
           "line_thickness": camera_config.get("line_thickness", 1)}
    if not cam["max_range"]:
        raise ValueError("camera.max_range must be a positive number (the reference raises on the first frame when it is None)")
    return cam
