"""Python front end of the CPU oracle (oracle/tc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module; the
product package tinycarlo_b200 never does.  It deliberately shares no code with the product: the map loader,
the camera-matrix construction and the spawn draw are restated here straight from the reference:

  map loading / px->m scaling      tinycarlo/map.py:9-37
  E and K                          tinycarlo/camera.py:145-178 (same cv2.Rodrigues / numpy calls)
  spawn draw                       tinycarlo/map.py:51-69 + gymnasium seeding (Generator(PCG64(SeedSequence(seed))))
  step / info / frame              oracle/tc_oracle.c (see its header for the file:line map)

Parity status: PINNED — tests/test_oracle_golden.py replays every trace in tests/golden/ (recorded from the
unmodified reference by tests/golden/gen_golden.py) through this oracle and demands bit-equality.
"""
import ctypes
import json
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libtc_oracle.so")
MAPS_DIR = os.path.join(HERE, "..", "tinycarlo_b200", "maps")  # data files only

SF_N, SI_N, CP_N, CAM_N = 8, 16, 8, 20

_c = ctypes
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_up = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile oracle/tc_oracle.c -> oracle/_build/libtc_oracle.so (gcc via oracle/Makefile)."""
    src = os.path.join(HERE, "tc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.orc_map_create.restype = _c.c_void_p
        L.orc_map_create.argtypes = [_c.c_int, _ip, _ip, _dp, _ip, _c.c_int, _c.c_int, _dp, _ip]
        L.orc_map_destroy.argtypes = [_c.c_void_p]
        L.orc_car_reset.restype = _c.c_int
        L.orc_car_reset.argtypes = [_c.c_void_p, _c.c_int, _dp, _dp, _ip]
        L.orc_car_step.restype = _c.c_int
        L.orc_car_step.argtypes = [_c.c_void_p, _dp, _dp, _ip, _c.c_double, _c.c_double, _c.c_int]
        L.orc_get_info.argtypes = [_c.c_void_p, _dp, _dp, _ip, _c.c_int, _dp, _dp, _dp, _dp, _ip, _dp, _ip]
        L.orc_camera_pose.argtypes = [_dp, _dp, _dp]
        L.orc_capture_frame.argtypes = [_c.c_void_p, _dp, _c.c_int, _c.c_int, _c.c_int, _up, _dp, _c.c_int, _up, _ip, _ip,
                                        _dp, _ip]
        L.orc_polyline.argtypes = [_up, _c.c_int, _c.c_int, _c.c_int, _up, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32,
                                   _c.c_int]
        L.orc_step_batch.argtypes = [_c.c_void_p, _c.c_int, _dp, _dp, _ip, _c.c_int, _c.c_int, _up, _c.c_int, _c.c_int, _dp,
                                     _ip, _dp, _ip, _up, _dp, _ip, _up, _up]
        L.orc_reset_batch.argtypes = [_c.c_void_p, _c.c_int, _up, _ip, _dp, _dp, _ip, _c.c_int, _c.c_int, _up, _c.c_int, _dp,
                                      _ip, _up]
        L.orc_clip_angle.restype = _c.c_double
        L.orc_clip_angle.argtypes = [_c.c_double]
        L.orc_layer_nearest_edge.restype = _c.c_int
        L.orc_layer_nearest_edge.argtypes = [_c.c_void_p, _dp]
        L.orc_layer_nearest_edge_with_orientation.restype = _c.c_int
        L.orc_layer_nearest_edge_with_orientation.argtypes = [_c.c_void_p, _dp, _c.c_double, _c.c_double]
        L.orc_layer_within_bounds.restype = _c.c_int
        L.orc_layer_within_bounds.argtypes = [_c.c_void_p, _dp, _c.c_int, _c.c_int]
        L.orc_layer_distance_to_edge.restype = _c.c_double
        L.orc_layer_distance_to_edge.argtypes = [_c.c_void_p, _dp, _c.c_int, _c.c_int]
        L.orc_layer_pick_node.restype = _c.c_int
        L.orc_layer_pick_node.argtypes = [_c.c_void_p, _c.c_int, _c.c_double, _ip, _c.c_int]
        L.orc_layer_nearest_connected_edge.restype = _c.c_int
        L.orc_layer_nearest_connected_edge.argtypes = [_c.c_void_p, _dp, _ip, _c.c_double, _ip]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


# ------------------------------------------------------------------------------------------------ map (map.py:9-37)
class OracleMap:
    def __init__(self, map_json, pixel_per_meter, spawn_points=None):
        if isinstance(map_json, str):
            with open(map_json) as f:
                data = json.load(f)
        else:
            data = map_json
        ppm = pixel_per_meter
        self.class_names = list(data["lanelines"].keys())
        self.colors = np.array([data["lanelines"][k]["layer_color"] for k in self.class_names], np.uint8).reshape(-1, 3)
        node_off, edge_off, nodes, edges = [0], [0], [], []
        for k in self.class_names:
            layer = data["lanelines"][k]
            nodes += [[n[0] / ppm, n[1] / ppm] for n in layer["nodes"]]
            edges += [[int(e[0]), int(e[1])] for e in layer["edges"]]
            node_off.append(len(nodes))
            edge_off.append(len(edges))
        self.node_off = np.array(node_off, np.int32)
        self.edge_off = np.array(edge_off, np.int32)
        self.nodes = np.array(nodes, np.float64).reshape(-1, 2)
        self.edges = np.array(edges, np.int32).reshape(-1, 2)
        self.lp_nodes = np.array([[n[0] / ppm, n[1] / ppm] for n in data["lanepath"]["nodes"]], np.float64).reshape(-1, 2)
        self.lp_edges = np.array(data["lanepath"]["edges"], np.int32).reshape(-1, 2)
        self.spawn_points = None if spawn_points is None else list(spawn_points)
        self.C = len(self.class_names)
        self.handle = lib().orc_map_create(self.C, _p(self.node_off, _ip), _p(self.edge_off, _ip), _p(self.nodes, _dp),
                                           _p(self.edges, _ip), len(self.lp_nodes), len(self.lp_edges),
                                           _p(self.lp_nodes, _dp), _p(self.lp_edges, _ip))

    def __del__(self):
        try:
            lib().orc_map_destroy(self.handle)
        except Exception:
            pass

    # map.py:51-69 — the draw; the pose part lives in orc_car_reset
    def sample_spawn_node(self, rng):
        while True:
            if self.spawn_points is None:
                idx = int(rng.integers(0, len(self.lp_nodes) - 1, size=1, dtype=int)[0])
            else:
                idx = int(rng.choice(self.spawn_points))
            if np.any(self.lp_edges[:, 0] == idx):
                return idx


def make_rng(seed):
    """gymnasium >= 0.26 seeding: Env.reset(seed=s) -> np.random.Generator(PCG64(SeedSequence(s)))."""
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def load_named_map(name, pixel_per_meter, spawn_points=None):
    return OracleMap(os.path.join(MAPS_DIR, name + ".json"), pixel_per_meter, spawn_points)


# ------------------------------------------------------------------------------------------------ camera.py:145-178
def camera_matrices(position, orientation, fov, resolution):
    import cv2
    angles_rad = np.radians(orientation + np.array([-90, 0, 90]))
    r_pr, _ = cv2.Rodrigues(np.array([1, 1, 0]) * angles_rad)
    r_y, _ = cv2.Rodrigues(np.array([0, 0, 1]) * angles_rad)
    t = np.column_stack((np.eye(3), -np.array(position)))
    E = r_pr @ r_y @ t
    fov_radians = np.radians(fov)
    fx = resolution[1] / (2 * np.tan(fov_radians / 2))
    fy = resolution[0] / (2 * np.tan(fov_radians / 2))
    K = np.array([[fx, 0, resolution[1] / 2], [0, fy, resolution[0] / 2], [0, 0, 1]])
    return np.asarray(E, np.float64), np.asarray(K, np.float64)


def pack_cam(E, K, max_range):
    cam = np.zeros(CAM_N, np.float64)
    cam[:12] = np.asarray(E, np.float64).reshape(-1)
    cam[12], cam[13], cam[14], cam[15] = K[0][0], K[1][1], K[0][2], K[1][2]
    cam[16] = max_range
    return cam


def pack_car(car_cfg, fps):
    def g(k, d):
        v = car_cfg.get(k, d)
        return math.nan if v is None else float(v)
    if car_cfg.get("max_acceleration", None) is not None and car_cfg.get("max_deceleration", None) is None:
        raise TypeError("max_deceleration must be set when max_acceleration is (car.py:82)")
    return np.array([g("wheelbase", 0.08), g("track_width", 0.03), g("max_velocity", 1), g("max_steering_angle", 35),
                     g("steering_speed", None), g("max_acceleration", None), g("max_deceleration", None), 1 / fps],
                    np.float64)


# ------------------------------------------------------------------------------------------------ batched env
class OracleVecEnv:
    """n independent reference envs stepped in lockstep on the CPU (each is the scalar restatement)."""

    def __init__(self, omap, n, car_params, cam_params, thickness, H, W, fmt="classes", wrapped=False):
        self.map, self.n, self.H, self.W = omap, n, int(H), int(W)
        self.fmt = 0 if fmt == "classes" else 1
        self.wrapped = bool(wrapped)
        self.car = np.array(np.broadcast_to(np.asarray(car_params, np.float64).reshape(-1, CP_N), (n, CP_N)), order="C")
        self.cam = np.array(np.broadcast_to(np.asarray(cam_params, np.float64).reshape(-1, CAM_N), (n, CAM_N)), order="C")
        self.thick = np.array(np.broadcast_to(np.asarray(thickness, np.int32).reshape(-1), (n,)), order="C")
        C = omap.C
        self.sf = np.zeros((n, SF_N), np.float64)
        self.si = np.full((n, SI_N), -1, np.int32)
        self.obs_shape = (C, self.H, self.W) if self.fmt == 0 else (self.H, self.W, 3)
        self.obs = np.zeros((n,) + self.obs_shape, np.uint8)
        self.info = np.zeros((n, 4 + C), np.float64)
        self.nearest = np.full((n, C), -1, np.int32)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)

    def reset(self, spawn_nodes, mask=None, render=True):
        sp = np.ascontiguousarray(spawn_nodes, np.int32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_reset_batch(self.map.handle, self.n, _p(m, _up), _p(sp, _ip), _p(self.car, _dp), _p(self.cam, _dp),
                              _p(self.thick, _ip), self.H, self.W, _p(self.map.colors, _up), self.fmt, _p(self.sf, _dp),
                              _p(self.si, _ip), _p(self.obs, _up) if render else None)

    def step(self, act_cc, act_man, render=True):
        cc = np.ascontiguousarray(act_cc, np.float64).reshape(self.n, 2)
        man = np.ascontiguousarray(act_man, np.int32).reshape(self.n)
        lib().orc_step_batch(self.map.handle, self.n, _p(self.car, _dp), _p(self.cam, _dp), _p(self.thick, _ip), self.H,
                             self.W, _p(self.map.colors, _up), self.fmt, int(self.wrapped), _p(cc, _dp), _p(man, _ip),
                             _p(self.sf, _dp), _p(self.si, _ip), _p(self.obs, _up) if render else None, _p(self.info, _dp),
                             _p(self.nearest, _ip), _p(self.terminated, _up), _p(self.truncated, _up))

    # convenient views of info[n, 4+C] = cte, heading_error, velocity, reward, dist[C]
    @property
    def cte(self):
        return self.info[:, 0]

    @property
    def heading_error(self):
        return self.info[:, 1]

    @property
    def velocity(self):
        return self.info[:, 2]

    @property
    def reward(self):
        return self.info[:, 3]

    @property
    def dist(self):
        return self.info[:, 4:]

    def segments(self, i):
        """Projected segments of env i at its current pose: (count[C], seg_i32[sumE,4], seg_f64[sumE,4], seg_edge[sumE])."""
        C, ne = self.map.C, int(self.map.edge_off[-1])
        cnt = np.zeros(C, np.int32)
        s32 = np.zeros((max(ne, 1), 4), np.int32)
        s64 = np.zeros((max(ne, 1), 4), np.float64)
        se = np.full(max(ne, 1), -1, np.int32)
        lib().orc_capture_frame(self.map.handle, _p(self.cam[i], _dp), self.H, self.W, int(self.thick[i]),
                                _p(self.map.colors, _up), _p(self.sf[i], _dp), self.fmt, None, _p(cnt, _ip), _p(s32, _ip),
                                _p(s64, _dp), _p(se, _ip))
        return cnt, s32, s64, se


def polyline(img, p0, p1, color, thickness):
    """cv2.polylines(img, np.int32([[p0, p1]]), False, color, thickness) restated (in place; img u8 [H,W] or [H,W,ch])."""
    assert img.dtype == np.uint8 and img.flags.c_contiguous
    H, W = img.shape[:2]
    nch = 1 if img.ndim == 2 else img.shape[2]
    col = np.zeros(4, np.uint8)
    col[:nch] = np.atleast_1d(np.asarray(color, np.uint8))[:nch]
    lib().orc_polyline(_p(img, _up), H, W, nch, _p(col, _up), int(p0[0]), int(p0[1]), int(p1[0]), int(p1[1]), int(thickness))
    return img
