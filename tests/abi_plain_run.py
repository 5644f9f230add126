"""(run by tests/test_gpu_abi_plain.py in a fresh interpreter) The drop-in boundary without torch: libtinycarlo_b200.so driven through ctypes with device memory from libcudart alone
(cudaMalloc / cudaMemcpy), as a C, Go or Java binding would do it. include/tinycarlo_b200.h has no torch types; this test
shows no torch is needed at run time either. Results are checked against the CPU oracle."""
import ctypes as C
import os

import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pair_util import make_config, oracle_env, stanley_actions
from tinycarlo_b200 import _lib
from tinycarlo_b200.camera_params import camera_row
from tinycarlo_b200.config import camera_params, car_param_row, resolve_map_path
from tinycarlo_b200.maptables import MapTables
from tinycarlo_b200.spawn import spawn_stream_states


class Rt:
    """the three libcudart calls a binding needs"""

    def __init__(self):
        self.L = None
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                self.L = C.CDLL(name)
                break
            except OSError:
                continue
        if self.L is None:
            print("SKIP: libcudart not found")
            sys.exit(77)
        self.L.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        self.L.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.L.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        self.L.cudaFree.argtypes = [C.c_void_p]
        self.bufs = []

    def alloc(self, nbytes):
        p = C.c_void_p()
        assert self.L.cudaMalloc(C.byref(p), max(int(nbytes), 16)) == 0
        assert self.L.cudaMemset(p, 0, max(int(nbytes), 16)) == 0
        self.bufs.append(p)
        return p

    def up(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.alloc(arr.nbytes)
        assert self.L.cudaMemcpy(p, arr.ctypes.data, arr.nbytes, 1) == 0
        return p

    def down(self, p, shape, dtype):
        out = np.empty(shape, dtype)
        assert self.L.cudaDeviceSynchronize() == 0
        assert self.L.cudaMemcpy(out.ctypes.data, p, out.nbytes, 2) == 0
        return out

    def free(self):
        for p in self.bufs:
            self.L.cudaFree(p)


def main():
    rt = Rt()
    L = _lib.lib()
    n = 96
    cfg = make_config("knuffingen", "classes")
    H, W = cfg["camera"]["resolution"]
    m = MapTables(resolve_map_path(cfg["map"], None), cfg["map"]["pixel_per_meter"], cfg["map"].get("spawn_points"))
    keep = [np.ascontiguousarray(a) for a in (m.ll_node_off, m.ll_edge_off, m.ll_nodes, m.ll_edges, m.colors, m.lp_nodes, m.lp_edges, m.lp_orient,
                                              m.lp_orient_rev)]
    desc = _lib.TcMapDesc(m.n_classes, *[a.ctypes.data for a in keep[:5]], len(m.lp_nodes), len(m.lp_edges), *[a.ctypes.data for a in keep[5:]])
    sim = _lib.TcSimDesc(H, W, _lib.TC_OBS_CLASSES)
    h = C.c_void_p()
    _lib.check(L.tc_create(C.byref(desc), C.byref(sim), n, 0, C.byref(h)), "tc_create")
    cc = camera_params(cfg["camera"])
    car = rt.up(np.tile(np.array(car_param_row(cfg["car"], 1 / 30), np.float64), (n, 1)))
    cam = rt.up(np.tile(camera_row(cc["position"], cc["orientation"], cc["fov"], cc["resolution"], cc["max_range"]), (n, 1)))
    thick = rt.up(np.full(n, cc["line_thickness"], np.int32))
    _lib.check(L.tc_set_car_params(h, car, None), "tc_set_car_params")
    _lib.check(L.tc_set_camera_params(h, cam, thick, None), "tc_set_camera_params")
    # spawn streams: env i = numpy's Generator(PCG64(SeedSequence(7 + i))), advanced on the device
    rng = rt.up(spawn_stream_states(n, 7))
    spawn_pts = rt.up(np.asarray(m.spawn_points, np.int32))
    last_spawn = rt.alloc(4 * n)
    _lib.check(L.tc_set_spawn_rng(h, rng, spawn_pts, len(m.spawn_points), last_spawn), "tc_set_spawn_rng")
    Cn = m.n_classes
    outs = _lib.TcOutputs()
    obs, info = rt.alloc(n * Cn * H * W), rt.alloc(8 * n * (4 + Cn))
    term, trunc, near = rt.alloc(n), rt.alloc(n), rt.alloc(4 * n * Cn)
    outs.obs, outs.info_f64, outs.terminated, outs.truncated, outs.nearest_edge = obs.value, info.value, term.value, trunc.value, near.value
    _lib.check(L.tc_reset(h, None, None, C.byref(outs), None), "tc_reset")            # NULL mask = all envs, NULL nodes = device draw
    nodes = rt.down(last_spawn, (n,), np.int32)
    gens = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(7 + i))) for i in range(n)]
    assert np.array_equal(nodes, [m.sample_spawn_node(g) for g in gens]), "device spawn draw vs numpy"
    oenv = oracle_env(cfg, n)
    oenv.reset(nodes)
    assert np.array_equal(rt.down(obs, (n, Cn, H, W), np.uint8), oenv.obs)
    d_cc, d_man = rt.alloc(16 * n), rt.alloc(4 * n)
    for t in range(5):
        cc_np = stanley_actions(oenv.cte.copy(), oenv.heading_error.copy(), cfg["car"]["max_steering_angle"]).astype(np.float64)
        man = np.full(n, t % 4, np.int32)
        assert rt.L.cudaMemcpy(d_cc, cc_np.ctypes.data, cc_np.nbytes, 1) == 0 and rt.L.cudaMemcpy(d_man, man.ctypes.data, man.nbytes, 1) == 0
        _lib.check(L.tc_step_f64(h, d_cc, d_man, C.byref(outs), None), "tc_step_f64")
        oenv.step(cc_np, man)
        assert np.array_equal(rt.down(obs, (n, Cn, H, W), np.uint8), oenv.obs), t
        np.testing.assert_allclose(rt.down(info, (n, 4 + Cn), np.float64), oenv.info, rtol=1e-9, atol=1e-11)
        assert np.array_equal(rt.down(near, (n, Cn), np.int32), oenv.nearest)
        assert np.array_equal(rt.down(term, (n,), np.uint8), oenv.terminated) and np.array_equal(rt.down(trunc, (n,), np.uint8), oenv.truncated)
    # error behaviour of the boundary: codes + message, no exceptions across the ABI
    assert L.tc_step(h, None, d_man, C.byref(outs), None) == -1 and b"null" in L.tc_last_error()
    _lib.check(L.tc_destroy(h), "tc_destroy")
    rt.free()
    assert "torch" not in sys.modules, "torch was imported somewhere on this path"
    print("OK: C ABI driven without torch,", n, "envs x 5 steps equal to the oracle")


if __name__ == "__main__":
    main()
