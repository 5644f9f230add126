"""CPU checks of the arithmetic the CUDA kernels run (tinycarlo_b200/csrc/tc_core.cuh built for the host, one lane per
group; tests/hosttest_util.py): the node-parallel clip passes, the closed-form rasteriser and the CSR-based tracking are
replayed against the reference's golden traces and fuzzed against cv2 / the oracle. The product never uses this build."""
import hashlib
import os

import numpy as np
import pytest

from golden_util import SCENARIOS, Golden
from hosttest_util import HostCore, P, ht, polyline
from tinycarlo_b200.camera_params import camera_row
from tinycarlo_b200.config import car_param_row, resolve_map_path
from tinycarlo_b200.maptables import MapTables

cv2 = pytest.importorskip("cv2")


def same_segments(tables, cnt_a, seg_a, cnt_b, seg_b):
    """per class, the first count entries (the arrays keep stale rows behind them)"""
    if not np.array_equal(cnt_a, cnt_b):
        return False
    for c in range(tables.n_classes):
        o = int(tables.ll_edge_off[c])
        for i in np.nonzero(cnt_a[:, c])[0]:
            if not np.array_equal(seg_a[i, o:o + cnt_a[i, c]], seg_b[i, o:o + cnt_a[i, c]]):
                return False
    return True


def cam_row_from(E, K, mr):
    r = np.zeros(20)
    r[:12] = np.asarray(E).reshape(-1)
    r[12], r[13], r[14], r[15], r[16] = K[0][0], K[1][1], K[0][2], K[1][2], mr
    return r


@pytest.mark.parametrize("name", SCENARIOS)
def test_host_core_replays_reference_trace(name):
    g = Golden(name)
    cfg = g.cfg
    tables = MapTables(resolve_map_path(cfg["map"], None), cfg["map"]["pixel_per_meter"], cfg["map"].get("spawn_points"))
    cc = cfg["camera"]
    row0 = camera_row(cc["position"], cc["orientation"], cc["fov"], cc["resolution"], cc["max_range"])
    assert np.array_equal(row0[:12].reshape(3, 4), g["E"][0]) and row0[12] == g["K"][0][0][0] and row0[13] == g["K"][0][1][1]
    env = HostCore(tables, 1, car_param_row(cfg["car"], 1 / cfg["sim"].get("fps", 30)), row0, cc["line_thickness"], g.H, g.W, g.fmt,
                   wrapped=g.wrapped, rows_per_band=32 if g.H > 64 else 0)
    mr = g.max_range_per_frame()
    radius_of = {}
    for f in range(g.F):
        env.cam[0] = cam_row_from(g["E"][f], g["K"][f], mr[f])
        if g["ev_kind"][f] == 0:
            env.reset([int(g["spawn_node"][f])])
        else:
            t = int(g["ev_step"][f])
            env.step(g["act_cc"][t][None], g["act_man"][t][None])
            assert bool(env.truncated[0]) == bool(g["truncated"][f]), (name, f)
            if not g.wrapped:
                assert bool(env.terminated[0]) == bool(g["terminated"][f]), (name, f)
                np.testing.assert_allclose(env.info[0, 3], g["reward"][f], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(env.info[0, 0], g["cte"][f], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(env.info[0, 1], g["heading"][f], rtol=1e-12, atol=1e-15)
            assert env.info[0, 2] == g["velocity"][f]
            np.testing.assert_allclose(env.info[0, 4:], g["dist"][f], rtol=1e-12, atol=1e-15)
        sf, si = env.sf[0], env.si[0]
        np.testing.assert_allclose(sf[:7], [*g["pos"][f], g["rot"][f], g["steer"][f], g["vel"][f], *g["front"][f]], rtol=1e-13, atol=1e-15)
        L = int(g["lp_len"][f])
        assert si[0] == L and si[1] == g["last_man"][f], (name, f)
        assert np.array_equal(si[2:2 + 2 * L].reshape(L, 2), g["lp"][f][:L]), (name, f)
        gi, _ = g.segments(f)
        for c in range(g.C):
            o = int(tables.ll_edge_off[c])
            k = int(env.seg_count[0, c])
            assert k == len(gi[c]), (name, f, c)
            assert np.array_equal(env.seg[0, o:o + k], gi[c]), (name, f, c)
        if g.fmt == "classes":
            assert np.array_equal(env.obs[0], g.classes_frame(f)), (name, f)
        else:
            assert hashlib.sha256(env.obs[0].tobytes()).hexdigest().encode() == g["rgb_sha"][f], (name, f)
        # the visible-set tables of the block-per-env kernel give the reference's segments too (tables per camera row,
        # as tc_set_camera_params rebuilds them)
        if f % 5 == 0 or g["ev_kind"][f] == 0:
            key = env.cam[0].tobytes()
            if key not in radius_of:
                radius_of[key] = ht().ht_cull_radius_of(P(env.cam[0]), g.H, g.W)
            cnt, seg, cell_nodes, info = env.project_culled(radius_of[key])
            assert same_segments(tables, cnt, seg, env.seg_count, env.seg), (name, f, "culled camera pass")


@pytest.mark.parametrize("H,W,spread,n,nlanes", [(24, 32, 12, 4000, 1), (24, 32, 12, 3000, 32), (24, 32, 12, 3000, -1), (48, 64, 300, 2500, -1),
                                                  (480, 640, 200, 120, -1), (48, 64, 300, 2500, 32),
                                                  (24, 32, 12, 6000, -2), (48, 64, 300, 4000, -2), (84, 84, 40, 3000, -2), (480, 640, 200, 200, -2), (128, 160, 2000, 3000, -2),
                                                  (84, 84, 40, 1500, 32), (480, 640, 200, 120, 32)])
def test_bitplane_rasteriser_vs_cv2(H, W, spread, n, nlanes):
    rng = np.random.default_rng(H * 7 + W + nlanes)
    bad = []
    for _ in range(n):
        t = int(rng.integers(1, 9))
        p0 = (int(rng.integers(-spread, W + spread)), int(rng.integers(-spread, H + spread)))
        p1 = (int(rng.integers(-spread, W + spread)), int(rng.integers(-spread, H + spread)))
        a = np.zeros((H, W), np.uint8)
        cv2.polylines(a, np.int32([[p0, p1]]), False, 255, t)
        if not np.array_equal(a, polyline(H, W, p0, p1, t, nlanes=nlanes)):
            bad.append((p0, p1, t))
    assert not bad, bad[:5]


@pytest.mark.parametrize("mag", [10**6, 10**9, 2**31 - 1])
def test_bitplane_rasteriser_far_endpoints(mag):
    rng = np.random.default_rng(mag % 9973)
    H, W = 48, 64
    bad = []
    for _ in range(1500):
        t = int(rng.integers(1, 7))
        p0 = (int(rng.integers(-5, W + 5)), int(rng.integers(-5, H + 5)))
        p1 = (int(rng.integers(-mag, mag + 1)), int(rng.integers(-mag, mag + 1)))
        if rng.random() < 0.5:
            p0, p1 = p1, p0
        a = np.zeros((H, W), np.uint8)
        cv2.polylines(a, np.int32([[p0, p1]]), False, 255, t)
        if not np.array_equal(a, polyline(H, W, p0, p1, t, nlanes=32)) or not np.array_equal(a, polyline(H, W, p0, p1, t, nlanes=-2)):
            bad.append((p0, p1, t))
    assert not bad, bad[:5]
    for t in (1, 2, 3):
        for p0, p1 in [((10, 10), (-2**31, -2**31)), ((-2**31, 5), (20, 20)), ((30, -2**31), (30, 40)), ((-2**31, 20), (2**31 - 1, 20))]:
            a = np.zeros((H, W), np.uint8)
            cv2.polylines(a, np.int32([[p0, p1]]), False, 255, t)
            assert np.array_equal(a, polyline(H, W, p0, p1, t, nlanes=32)), (p0, p1, t)


def test_bitplane_bands_compose_to_full_frame():
    """A band block only owns rows [y_lo, y_hi): the union of the bands must be the full-frame raster."""
    rng = np.random.default_rng(5)
    H, W = 96, 80
    for _ in range(300):
        t = int(rng.integers(1, 7))
        p0 = (int(rng.integers(-30, W + 30)), int(rng.integers(-30, H + 30)))
        p1 = (int(rng.integers(-30, W + 30)), int(rng.integers(-30, H + 30)))
        full = polyline(H, W, p0, p1, t, nlanes=32)
        parts = np.zeros_like(full)
        for y in range(0, H, 32):
            parts |= polyline(H, W, p0, p1, t, y_lo=y, y_hi=min(H, y + 32), nlanes=32)
        assert np.array_equal(full, parts), (p0, p1, t)


def test_device_spawn_streams_equal_numpy_and_reference_draws():
    """tc_pcg_bounded / tc_spawn_draw (the code the reset paths of the tracking kernel run, host build) against numpy's
    Generator itself and against the spawn nodes recorded from the reference (map.py:51-69 under gymnasium seeding)."""
    import ctypes as C
    from hosttest_util import HostCore, P, ht
    from tinycarlo_b200.pcg64 import VecPCG64
    from tinycarlo_b200.spawn import spawn_stream_states
    from pair_util import SPAWN_KNUFF, SPAWN_SIMPLE
    from golden_util import GOLDEN_DIR
    seeds = np.array(list(range(0, 64)) + [2**32 - 1, 2**32 + 5, 2**63 + 7, 2**64 - 1], dtype=np.uint64)
    st = VecPCG64(seeds).device_rows()
    gens = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(s)))) for s in seeds]
    for hi in [21, 16, 429, 428, 2**31 + 1, 3, 2**32 - 1, 1, 7, 100000, 21]:   # 2^31+1: ~50 % rejection (uint32 argument)
        out = np.zeros((len(seeds), 3), np.uint32)
        ht().ht_pcg_bounded(len(seeds), P(st), hi, 3, P(out))
        want = np.array([[int(g.integers(0, hi)) for _ in range(3)] for g in gens], dtype=np.uint32)
        assert np.array_equal(out, want), hi
    # the state rows round-trip through the host model (checkpoints)
    v = VecPCG64()
    v.load_device_rows(st)
    assert np.array_equal(v.bounded(1000), np.array([int(g.integers(0, 1000)) for g in gens]))

    d = np.load(os.path.join(GOLDEN_DIR, "spawn_draws.npz"))
    for key, (mname, ppm, sp) in {"knuffingen_default": ("knuffingen", 222, SPAWN_KNUFF), "knuffingen_none": ("knuffingen", 222, None),
                                  "simple_layout_default": ("simple_layout", 450, SPAWN_SIMPLE), "simple_layout_none": ("simple_layout", 450, None)}.items():
        t = MapTables(resolve_map_path({"map_name": mname}, None), ppm, sp)
        want = d[key]                      # [64 seeds, 12 consecutive resets]
        core = HostCore(t, 1, np.zeros(8), np.zeros(20), 1, 8, 8)
        st = spawn_stream_states(40, 10, env_index_offset=3)   # env i = the reference's stream for seed 13 + i
        out = np.zeros((40, 12), np.int32)
        spp = None if sp is None else np.asarray(sp, np.int32)
        ht().ht_spawn_draws(core.h, 40, P(st), P(spp), 0 if sp is None else len(sp), 12, P(out))
        assert np.array_equal(out, want[13:53]), key


def _poses(E, xy, rot):
    """camera.py:61 with car.py:159-165 in plain numpy (both camera passes under test take the same pose)"""
    n = len(rot)
    c, s_ = np.cos(-rot), np.sin(-rot)
    M = np.zeros((n, 4, 4))
    M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1], M[:, 2, 2], M[:, 3, 3] = c, -s_, s_, c, 1, 1
    T = np.tile(np.eye(4), (n, 1, 1))
    T[:, 0, 3], T[:, 1, 3] = -xy[:, 0], -xy[:, 1]
    return np.ascontiguousarray((E[None] @ (M @ T)).reshape(n, 12))


@pytest.mark.parametrize("map_name,ppm,res,cam", [
    ("knuffingen", 222, [128, 160], {}), ("knuffingen", 222, [480, 640], {"max_range": 1.5}), ("knuffingen", 222, [84, 84], {"fov": 125, "orientation": [10, 0, 0]}),
    ("knuffingen", 222, [128, 160], {"max_range": 0.2, "position": [0.02, -0.01, 0.02], "orientation": [35, 4, -6]}),
    ("simple_layout", 450, [84, 84], {}), ("simple_layout", 450, [84, 84], {"max_range": 0.15}),
    ("knuffingen", 222, [96, 128], {"orientation": [1, 0, 0], "max_range": 1.0}),                       # looking almost horizontally
    ("knuffingen", 222, [96, 128], {"orientation": [88, 0, 0], "position": [0, 0, 0.3], "fov": 150}),   # looking almost straight down, very wide
    ("simple_layout", 450, [64, 64], {"orientation": [22, 30, 90], "position": [0.05, 0.03, 0.01]}),    # rolled, yawed, 1 cm above the ground
    ("formula_student_track", 100, [128, 160], {"max_range": 3.0}), ("formula_student_skidpad", 100, [96, 128], {"max_range": 1.0})])
def test_visible_set_tables_never_change_the_segments(map_name, ppm, res, cam):
    """tc_cull.h: the camera pass on the sub-graph of the camera's ground cell emits exactly the segments of the pass over the
    whole map - for cameras on the track, next to it, on cell borders, far outside the map, at any yaw."""
    from pair_util import make_config
    from tinycarlo_b200.config import camera_params
    cfg = make_config(map_name, "classes", cam=dict({"resolution": res}, **cam), spawn=None)
    ppm = cfg["map"]["pixel_per_meter"]
    tables = MapTables(resolve_map_path(cfg["map"], None), ppm, None)
    cc = camera_params(cfg["camera"])
    row = camera_row(cc["position"], cc["orientation"], cc["fov"], cc["resolution"], cc["max_range"])
    H, W = res
    rng = np.random.default_rng(len(map_name) + H)
    nodes = np.asarray(tables.ll_nodes, np.float64).reshape(-1, 2)
    lo, hi = nodes.min(0), nodes.max(0)
    n = 6000
    near = nodes[rng.integers(0, len(nodes), n // 2)] + rng.normal(0, 0.15, (n // 2, 2))      # on / next to the lines
    anywhere = rng.uniform(lo - 1.5, hi + 1.5, (n // 4, 2))
    on_nodes = nodes[rng.integers(0, len(nodes), n - n // 2 - n // 4)]                          # degenerate: camera base on a node
    xy = np.concatenate([near, anywhere, on_nodes])
    rot = rng.uniform(-np.pi, np.pi, n)
    rot[::7] = np.round(rot[::7] / (np.pi / 2)) * (np.pi / 2)                                    # axis-aligned views: edges parallel to the clip planes
    env = HostCore(tables, n, np.zeros(8), row, 2, H, W)
    env.pose[:] = _poses(row[:12].reshape(3, 4), xy, rot)
    env.obs = None
    env.render()
    radius = ht().ht_cull_radius_of(P(row), H, W)
    assert radius > 0
    for cell in (0.25, 0.1):
        cnt, seg, cell_nodes, info = env.project_culled(radius, cell=cell)
        assert same_segments(tables, cnt, seg, env.seg_count, env.seg), (map_name, cell)
    assert env.seg_count.sum() > n, "the poses should see something"
    if info[0] > 0:   # culling active: the sub-graphs are smaller than the map
        assert info[3] <= 0.85 * len(nodes) and cell_nodes.max() <= info[3]


def test_cull_radius_rejects_cameras_outside_the_argument():
    cc = {"position": [0, -0.005, 0.04], "orientation": [22, 0, 0], "fov": 80, "resolution": [128, 160], "max_range": 0.5}
    row = camera_row(cc["position"], cc["orientation"], cc["fov"], cc["resolution"], cc["max_range"])
    r = ht().ht_cull_radius_of(P(row), 128, 160)
    fx, fy, cx, cy = row[12:16]
    assert abs(r - 0.5 * np.sqrt(1 + (max(cx, 160 - cx) / fx) ** 2 + (max(cy, 128 - cy) / fy) ** 2)) < 1e-12
    bad = row.copy(); bad[1] *= 1.01                                   # extrinsics no longer a rotation
    assert ht().ht_cull_radius_of(P(bad), 128, 160) < 0
    ground = camera_row([0, 0, 0.0], cc["orientation"], 80, [128, 160], 0.5)   # camera in the ground plane
    assert ht().ht_cull_radius_of(P(ground), 128, 160) < 0
    inf = row.copy(); inf[16] = np.inf
    assert ht().ht_cull_radius_of(P(inf), 128, 160) < 0


@pytest.mark.parametrize("map_name", ["knuffingen", "simple_layout", "formula_student_track", "formula_student_skidpad"])
def test_nearest_laneline_index_equals_the_full_scan(map_name):
    """tc_build_near: the candidate list of a ground cell contains every edge that can be the arg-min of d(p,n0)+d(p,n1)
    (layer.py:33-44) for a point of the cell, so the indexed search returns what the scan over all edges returns - on the
    lines, between them, on cell borders, on nodes (ties -> first edge) and outside the grid (plain scan)."""
    from pair_util import make_config
    cfg = make_config(map_name, "classes", spawn=None)
    tables = MapTables(resolve_map_path(cfg["map"], None), cfg["map"]["pixel_per_meter"], None)
    env = HostCore(tables, 1, np.zeros(8), np.zeros(20), 1, 8, 8)
    rng = np.random.default_rng(len(map_name))
    nodes = np.asarray(tables.ll_nodes, np.float64).reshape(-1, 2)
    lo, hi = nodes.min(0), nodes.max(0)
    n = 20000
    pts = np.concatenate([nodes[rng.integers(0, len(nodes), n // 4)] + rng.normal(0, 0.05, (n // 4, 2)),     # next to the lines
                          rng.uniform(lo - 1.2, hi + 1.2, (n // 4, 2)),                                    # anywhere, also outside the grid
                          nodes[rng.integers(0, len(nodes), n // 4)],                                      # on nodes: exact ties between adjacent edges
                          0.5 * (nodes[rng.integers(0, len(nodes), n // 4)] + nodes[rng.integers(0, len(nodes), n // 4)])])
    pts = np.ascontiguousarray(pts)
    C = tables.n_classes
    full = np.zeros((len(pts), C), np.int32)
    ht().ht_nearest(env.h, len(pts), P(pts), 0, 1, P(full), None, None)
    info = np.zeros(3)
    cand = np.zeros(len(pts), np.int32)
    for nlanes in (1, 32):
        got = np.zeros_like(full)
        ht().ht_nearest(env.h, len(pts), P(pts), 1, nlanes, P(got), P(cand), P(info))
        assert np.array_equal(got, full), (map_name, nlanes, int((got != full).sum()))
    assert info[0] > 1000 and (cand >= 0).mean() > 0.7, "the index should cover the map"
    m_total = int(tables.ll_edge_off[-1])
    assert info[1] < 0.25 * m_total / C + 8, ("lists should be short", info)


def test_table_builders_survive_degenerate_maps():
    """Visible-set tables and nearest-laneline index on maps with an empty class, a class of one isolated node, coincident
    nodes and a self-loop edge: the builders must not crash and the culled camera pass / indexed search must still equal the
    plain ones."""
    data = {"height": 300, "width": 300,
            "lanelines": {"empty": {"layer_color": [1, 2, 3], "nodes": [], "edges": []},
                          "dot": {"layer_color": [9, 9, 9], "nodes": [[50, 50]], "edges": []},
                          "twin": {"layer_color": [255, 0, 0], "nodes": [[100, 100], [100, 100], [140, 100], [140, 160]], "edges": [[0, 1], [1, 2], [2, 3], [3, 3]]},
                          "line": {"layer_color": [0, 255, 0], "nodes": [[10 + 20 * i, 200] for i in range(12)], "edges": [[i, i + 1] for i in range(11)]}},
            "lanepath": {"layer_color": [0, 0, 0], "nodes": [[20, 150], [120, 150], [220, 150]], "edges": [[0, 1], [1, 2]]}}
    tables = MapTables(data, 100)
    cc = {"position": [0, -0.005, 0.04], "orientation": [22, 0, 0], "fov": 80, "resolution": [96, 128], "max_range": 0.6}
    row = camera_row(cc["position"], cc["orientation"], cc["fov"], cc["resolution"], cc["max_range"])
    rng = np.random.default_rng(0)
    n = 3000
    xy = rng.uniform(-0.5, 3.5, (n, 2))
    rot = rng.uniform(-np.pi, np.pi, n)
    env = HostCore(tables, n, np.zeros(8), row, 2, 96, 128)
    env.pose[:] = _poses(row[:12].reshape(3, 4), xy, rot)
    env.obs = None
    env.render()
    radius = ht().ht_cull_radius_of(P(row), 96, 128)
    cnt, seg, cell_nodes, info = env.project_culled(radius, cell=0.2)
    assert same_segments(tables, cnt, seg, env.seg_count, env.seg)
    assert env.seg_count.sum() > 0
    pts = np.ascontiguousarray(rng.uniform(-0.5, 3.5, (4000, 2)))
    full = np.zeros((len(pts), tables.n_classes), np.int32)
    got = np.zeros_like(full)
    ht().ht_nearest(env.h, len(pts), P(pts), 0, 1, P(full), None, None)
    ht().ht_nearest(env.h, len(pts), P(pts), 1, 32, P(got), None, None)
    assert np.array_equal(got, full)
    assert (full[:, 0] == -1).all() and (full[:, 1] == -1).all()   # classes without edges have no nearest edge
