"""Pins the CPU oracle (oracle/tc_oracle.c) against traces of the UNMODIFIED reference (tests/golden/*.npz).

Everything is compared bit-for-bit: float64 state/info with ==, int32 segments, frames. CPU-only."""
import hashlib

import numpy as np
import pytest

from golden_util import SCENARIOS, Golden
from oracle import oracle as orc


def replay(name, check_frames=True):
    g = Golden(name)
    cfg = g.cfg
    omap = orc.load_named_map(cfg["map"]["map_name"], cfg["map"]["pixel_per_meter"], cfg["map"].get("spawn_points"))
    assert omap.class_names == g.class_names
    car = orc.pack_car(cfg["car"], cfg["sim"].get("fps", 30))
    E0, K0 = orc.camera_matrices(cfg["camera"]["position"], cfg["camera"]["orientation"], cfg["camera"]["fov"], [g.H, g.W])
    # E/K restated from camera.py must equal what the reference computed
    assert np.array_equal(E0, g["E"][0]) and np.array_equal(K0, g["K"][0])
    env = orc.OracleVecEnv(omap, 1, car, orc.pack_cam(E0, K0, cfg["camera"]["max_range"]), cfg["camera"]["line_thickness"],
                           g.H, g.W, g.fmt, wrapped=g.wrapped)
    mr = g.max_range_per_frame()
    rgb_keep = {int(f): i for i, f in enumerate(g["rgb_idx"])}
    n_seg = 0
    for f in range(g.F):
        env.cam[0] = orc.pack_cam(g["E"][f], g["K"][f], mr[f])
        if g["ev_kind"][f] == 0:
            env.reset([int(g["spawn_node"][f])])
        else:
            t = int(g["ev_step"][f])
            env.step(g["act_cc"][t][None], g["act_man"][t][None])
            assert bool(env.truncated[0]) == bool(g["truncated"][f]), (name, f)
            if not g.wrapped:
                assert env.reward[0] == g["reward"][f], (name, f)
                assert bool(env.terminated[0]) == bool(g["terminated"][f]), (name, f)
            assert env.cte[0] == g["cte"][f], (name, f, env.cte[0], g["cte"][f])
            assert env.heading_error[0] == g["heading"][f], (name, f)
            assert env.velocity[0] == g["velocity"][f], (name, f)
            assert np.array_equal(env.dist[0], g["dist"][f]), (name, f, env.dist[0], g["dist"][f])
        sf, si = env.sf[0], env.si[0]
        assert sf[0] == g["pos"][f][0] and sf[1] == g["pos"][f][1], (name, f, sf[:2], g["pos"][f])
        assert sf[2] == g["rot"][f] and sf[3] == g["steer"][f] and sf[4] == g["vel"][f], (name, f)
        assert sf[5] == g["front"][f][0] and sf[6] == g["front"][f][1], (name, f)
        L = int(g["lp_len"][f])
        assert si[0] == L and si[1] == g["last_man"][f], (name, f)
        assert np.array_equal(si[2:2 + 2 * L].reshape(L, 2), g["lp"][f][:L]), (name, f)
        # projected segments
        cnt, s32, s64, _ = env.segments(0)
        gi, gf = g.segments(f)
        for c in range(g.C):
            o = int(omap.edge_off[c])
            assert cnt[c] == len(gi[c]), (name, f, c)
            assert np.array_equal(s32[o:o + cnt[c]], gi[c]), (name, f, c)
            assert np.array_equal(s64[o:o + cnt[c]], gf[c], equal_nan=True), (name, f, c)
            n_seg += int(cnt[c])
        if check_frames:
            if g.fmt == "classes":
                assert np.array_equal(env.obs[0], g.classes_frame(f)), (name, f)
            else:
                assert hashlib.sha256(env.obs[0].tobytes()).hexdigest().encode() == g["rgb_sha"][f], (name, f)
                if f in rgb_keep:
                    assert np.array_equal(env.obs[0], g["rgb"][rgb_keep[f]])
    return n_seg


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_replays_reference_trace(name):
    assert replay(name) > 0


def test_oracle_rgb_of_classes_scenarios():
    """The reference renders the RGB frame even in classes mode (camera.py:104); its sha256 is in the goldens."""
    g = Golden("knuff_thick3")
    cfg = g.cfg
    omap = orc.load_named_map(cfg["map"]["map_name"], cfg["map"]["pixel_per_meter"], cfg["map"].get("spawn_points"))
    env = orc.OracleVecEnv(omap, 1, orc.pack_car(cfg["car"], 30), orc.pack_cam(g["E"][0], g["K"][0], cfg["camera"]["max_range"]),
                           cfg["camera"]["line_thickness"], g.H, g.W, "rgb")
    for f in range(g.F):
        if g["ev_kind"][f] == 0:
            env.reset([int(g["spawn_node"][f])])
        else:
            t = int(g["ev_step"][f])
            env.step(g["act_cc"][t][None], g["act_man"][t][None])
        assert hashlib.sha256(env.obs[0].tobytes()).hexdigest().encode() == g["rgb_sha"][f], f


SPAWN_KNUFF = [156, 18, 217, 214, 325, 354, 176, 402, 339, 376, 385, 419, 396, 37, 149, 62, 240, 113, 98, 299, 2]
SPAWN_SIMPLE = [57, 143, 112, 121, 138, 157, 67, 46, 165, 124, 79, 33, 84, 21, 178, 7]


def test_spawn_draws():
    """map.py:51-69 with gymnasium seeding: 64 seeds x 12 consecutive resets, with and without spawn_points."""
    import os
    from golden_util import GOLDEN_DIR
    d = np.load(os.path.join(GOLDEN_DIR, "spawn_draws.npz"))
    for key, (mname, ppm, sp) in {"knuffingen_default": ("knuffingen", 222, SPAWN_KNUFF), "knuffingen_none": ("knuffingen", 222, None),
                                  "simple_layout_default": ("simple_layout", 450, SPAWN_SIMPLE),
                                  "simple_layout_none": ("simple_layout", 450, None)}.items():
        omap = orc.load_named_map(mname, ppm, sp)
        want = d[key]
        for s in range(want.shape[0]):
            rng = orc.make_rng(s)
            got = [omap.sample_spawn_node(rng) for _ in range(want.shape[1])]
            assert got == list(want[s]), (key, s)
