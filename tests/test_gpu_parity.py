"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
 (1) the golden traces recorded from the unmodified reference (tests/golden/), and
 (2) the CPU oracle on seeded batches of envs.
Bars (BASELINE.json north_star): class masks, lanepath / segment indices bit-exact; pose, CTE, heading error, distances
within 1e-5 relative in the exported fp32 tensors (the f64 mirrors are checked much tighter, see RTOL64)."""
import hashlib
import zlib

import numpy as np
import pytest
import torch

from golden_util import SCENARIOS, Golden
from pair_util import make_config, oracle_env, stanley_actions

pytestmark = pytest.mark.gpu

RTOL32 = 1e-5   # the north-star tolerance, on the fp32 outputs
RTOL64 = 1e-9   # what the float64 state actually achieves (device libm differs from glibc by ulps only)


FAR = 1 << 20


def _check_segments(got, want, where, stats):
    """Projected int32 endpoints. Coordinates of ordinary magnitude must be identical. Endpoints that the near-plane
    fix-up (camera.py:112-122) sent ~1e9 px away carry the rounding of z ~ -1e-7 (relative 1e-10): a 1-ulp difference
    between the device libm and glibc upstream moves them by O(0.1-10) px at 1e9, which cannot change a mask pixel
    (checked separately, bit for bit); they are compared to 1e-7 relative."""
    got = got.astype(np.int64)
    want = want.astype(np.int64)
    near = (np.abs(want) < FAR) & (np.abs(got) < FAR)
    stats["near"] += int(near.sum())
    stats["far"] += int((~near).sum())
    stats["far_diff"] += int(((got != want) & ~near).sum())
    assert np.array_equal(got[near], want[near]), (where, got, want)
    if (~near).any():
        special = (want == -2**31) | (got == -2**31)
        assert np.array_equal(got[~near & special], want[~near & special]), (where, "INT_MIN endpoints")
        m = ~near & ~special
        np.testing.assert_allclose(got[m], want[m], rtol=1e-7, err_msg=str(where))


def _vec(cfg, n, **kw):
    from tinycarlo_b200 import TinyCarloVecEnv
    return TinyCarloVecEnv(cfg, n, device="cuda:0", **kw)


def cam_row_from(E, K, mr):
    r = np.zeros(20)
    r[:12] = np.asarray(E).reshape(-1)
    r[12], r[13], r[14], r[15], r[16] = K[0][0], K[1][1], K[0][2], K[1][2], mr
    return r


@pytest.mark.parametrize("fused", [False, True], ids=["unfused+segments", "fused"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_cuda_replays_reference_trace(name, fused):
    """One env driven with the recorded actions must reproduce the reference's trace - through the unfused kernels, which
    also export the projected segments (tc_project_kernel + tc_raster_*), and through the fused render kernels that every
    production step uses (tc_render_classes_kernel / tc_render_env_kernel / tc_render_env_banded_kernel)."""
    g = Golden(name)
    env = _vec(g.cfg, 1, debug_segments=not fused)
    if g.wrapped:
        env.set_wrapped(True)
    mr = g.max_range_per_frame()
    off = env.map.ll_edge_off
    last_cam = None
    seg_stats = {"near": 0, "far": 0, "far_diff": 0}
    for f in range(g.F):
        cam = cam_row_from(g["E"][f], g["K"][f], mr[f])
        if last_cam is None or not np.array_equal(cam, last_cam):
            env.set_camera_rows(cam[None])
            last_cam = cam
        if g["ev_kind"][f] == 0:
            obs, info = env.reset(spawn_nodes=torch.tensor([int(g["spawn_node"][f])], dtype=torch.int32))
        else:
            t = int(g["ev_step"][f])
            act = {"car_control": torch.tensor(g["act_cc"][t][None], dtype=torch.float32, device="cuda:0"),
                   "maneuver": torch.tensor(g["act_man"][t][None], dtype=torch.int32, device="cuda:0")}
            obs, reward, term, trunc, info = env.step(act)
            i64 = env.out["info_f64"][0].cpu().numpy()
            assert bool(trunc[0]) == bool(g["truncated"][f]), (name, f)
            if not g.wrapped:
                assert bool(term[0]) == bool(g["terminated"][f]), (name, f)
                np.testing.assert_allclose(i64[3], g["reward"][f], rtol=RTOL64, atol=1e-12)
                np.testing.assert_allclose(float(reward[0]), g["reward"][f], rtol=RTOL32, atol=1e-7)
            np.testing.assert_allclose(i64[0], g["cte"][f], rtol=RTOL64, atol=1e-12)
            np.testing.assert_allclose(i64[1], g["heading"][f], rtol=RTOL64, atol=1e-12)
            np.testing.assert_allclose(i64[2], g["velocity"][f], rtol=RTOL64, atol=1e-12)
            np.testing.assert_allclose(i64[4:], g["dist"][f], rtol=RTOL64, atol=1e-12)
            np.testing.assert_allclose(float(info["cte"][0]), g["cte"][f], rtol=RTOL32, atol=1e-7)
            np.testing.assert_allclose(float(info["heading_error"][0]), g["heading"][f], rtol=RTOL32, atol=1e-7)
            np.testing.assert_allclose(info["laneline_distances"][0].cpu().numpy(), g["dist"][f], rtol=RTOL32, atol=1e-7)
        st = env.state_dict()
        sf, si = st["sf"][0].cpu().numpy(), st["si"][0].cpu().numpy()
        np.testing.assert_allclose(sf[:7], [*g["pos"][f], g["rot"][f], g["steer"][f], g["vel"][f], *g["front"][f]], rtol=RTOL64, atol=1e-12)
        L = int(g["lp_len"][f])
        assert si[0] == L and si[1] == g["last_man"][f], (name, f)
        assert np.array_equal(si[2:2 + 2 * L].reshape(L, 2), g["lp"][f][:L]), (name, f)
        o = obs[0].cpu().numpy()
        if g.fmt == "classes":
            assert np.array_equal(o, g.classes_frame(f)), (name, f)
            if f % 10 == 0:
                rgb = env.render_rgb()[0].cpu().numpy()
                assert hashlib.sha256(rgb.tobytes()).hexdigest().encode() == g["rgb_sha"][f], (name, f)
        else:
            assert hashlib.sha256(o.tobytes()).hexdigest().encode() == g["rgb_sha"][f], (name, f)
        if fused:
            continue
        gi, _ = g.segments(f)
        cnt = env.out["seg_count"][0].cpu().numpy()
        seg = env.out["seg_i32"][0].cpu().numpy()
        for c in range(g.C):
            assert cnt[c] == len(gi[c]), (name, f, c)
            _check_segments(seg[off[c]:off[c] + cnt[c]], gi[c], (name, f, c), seg_stats)
    print(f"{name}: segment coordinates {seg_stats}")
    env.close()


def _compare_batch(env, oenv, step, stats):
    st = env.state_dict()
    sf, si = st["sf"].cpu().numpy(), st["si"].cpu().numpy()
    np.testing.assert_allclose(sf[:, :7], oenv.sf[:, :7], rtol=RTOL64, atol=1e-11, err_msg=f"state step {step}")
    assert np.array_equal(si[:, :10], oenv.si[:, :10]), f"local path step {step}"
    i64 = env.out["info_f64"].cpu().numpy()
    np.testing.assert_allclose(i64, oenv.info, rtol=RTOL64, atol=1e-11, err_msg=f"info step {step}")
    assert np.array_equal(env.out["nearest_edge"].cpu().numpy(), oenv.nearest), f"nearest laneline edge step {step}"
    assert np.array_equal(env.out["terminated"].cpu().numpy(), oenv.terminated), f"terminated step {step}"
    assert np.array_equal(env.out["truncated"].cpu().numpy(), oenv.truncated), f"truncated step {step}"
    # fp32 exports at the north-star tolerance
    np.testing.assert_allclose(env.out["cte"].cpu().numpy(), oenv.cte, rtol=RTOL32, atol=1e-7)
    np.testing.assert_allclose(env.out["heading_error"].cpu().numpy(), oenv.heading_error, rtol=RTOL32, atol=1e-7)
    np.testing.assert_allclose(env.out["laneline_distances"].cpu().numpy(), oenv.dist, rtol=RTOL32, atol=1e-7)
    np.testing.assert_allclose(env.out["position"].cpu().numpy(), oenv.sf[:, :2], rtol=RTOL32, atol=1e-7)
    obs = env.obs.cpu().numpy()
    bad = int((obs != oenv.obs).reshape(obs.shape[0], -1).any(axis=1).sum())
    stats["frames"] += obs.shape[0]
    stats["bad_frames"] += bad
    assert bad == 0, f"{bad} of {obs.shape[0]} frames differ from the oracle at step {step}"


@pytest.mark.parametrize("map_name,fmt,res,n,steps,policy", [
    ("knuffingen", "classes", [128, 160], 1024, 40, "stanley"),
    ("knuffingen", "classes", [480, 640], 64, 25, "stanley_mixed"),
    ("simple_layout", "classes", [84, 84], 2048, 30, "random"),
    ("simple_layout", "rgb", [96, 128], 512, 25, "random"),
    ("formula_student_track", "classes", [64, 96], 512, 30, "random_forward"),
])
def test_cuda_batch_matches_oracle(map_name, fmt, res, n, steps, policy):
    """Lockstep batches with auto-reset: every step, every env, against the oracle."""
    cfg = make_config(map_name, fmt, cam={"resolution": res, "max_range": 0.5 if "formula" not in map_name else 1.5},
                      car={"max_velocity": 0.15} if map_name == "simple_layout" else None,
                      spawn="default" if "formula" not in map_name else None)
    env = _vec(cfg, n)
    oenv = oracle_env(cfg, n)
    rng = np.random.default_rng(zlib.crc32(f"{map_name}{fmt}{n}".encode()))
    env.reset(seed=123)
    spawn = env._spawn_nodes.cpu().numpy()
    oenv.reset(spawn)
    stats = {"frames": 0, "bad_frames": 0}
    assert np.array_equal(env.obs.cpu().numpy(), oenv.obs)
    man = rng.integers(0, 4, n).astype(np.int32) if policy != "stanley" else np.zeros(n, np.int32)
    for t in range(steps):
        if policy.startswith("stanley"):
            cc = stanley_actions(oenv.cte.copy(), oenv.heading_error.copy(), cfg["car"]["max_steering_angle"])
            cc[:, 1] += rng.normal(0, 0.2, n).astype(np.float32)
            if policy == "stanley_mixed" and t % 8 == 0:
                man = rng.integers(0, 4, n).astype(np.int32)
        elif policy == "random":
            cc = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
            man = rng.integers(0, 4, n).astype(np.int32)
        else:
            cc = np.stack([rng.uniform(0.2, 1, n), rng.uniform(-1, 1, n)], 1).astype(np.float32)
            man = rng.integers(0, 4, n).astype(np.int32)
        env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        oenv.step(cc.astype(np.float64), man)
        _compare_batch(env, oenv, t, stats)
        done = (oenv.terminated | oenv.truncated).astype(bool)
        if done.any():
            env.reset_done()
            sp = env._spawn_nodes.cpu().numpy()
            oenv.reset(sp, mask=done)
            assert np.array_equal(env.obs.cpu().numpy(), oenv.obs), f"frames after reset, step {t}"
            st = env.state_dict()
            np.testing.assert_allclose(st["sf"].cpu().numpy()[:, :7], oenv.sf[:, :7], rtol=RTOL64, atol=1e-11)
    assert stats["bad_frames"] == 0
    env.close()


def test_cuda_per_env_params_match_oracle():
    """Config-5 style domain randomisation: per-env camera pose / fov / range / thickness and car parameters."""
    n = 512
    cfg = make_config("knuffingen", "classes", cam={"resolution": [120, 160]})
    rng = np.random.default_rng(77)
    env = _vec(cfg, n)
    pos = np.stack([rng.uniform(-0.01, 0.02, n), rng.uniform(-0.01, 0.01, n), rng.uniform(0.02, 0.08, n)], 1).round(4)
    ori = np.stack([rng.integers(5, 40, n), rng.integers(-5, 6, n), rng.integers(-30, 31, n)], 1).astype(np.float64)
    fov = rng.integers(60, 130, n).astype(np.float64)
    mr = rng.uniform(0.3, 2.0, n).round(3)
    th = rng.integers(1, 7, n).astype(np.int32)
    env.set_camera_params(position=pos, orientation=ori, fov=fov, max_range=mr, line_thickness=th)
    wb = rng.uniform(0.04, 0.06, n)
    mv = rng.uniform(0.08, 0.3, n)
    ms = rng.uniform(24, 36, n)
    env.set_car_params(wheelbase=wb, max_velocity=mv, max_steering_angle=ms)
    oenv = oracle_env(cfg, n, cam_rows=env._cam_rows.copy(), thickness=th, car_rows=env._car_rows.copy())
    env.reset(seed=5)
    oenv.reset(env._spawn_nodes.cpu().numpy())
    assert np.array_equal(env.obs.cpu().numpy(), oenv.obs)
    stats = {"frames": 0, "bad_frames": 0}
    for t in range(25):
        cc = stanley_actions(oenv.cte.copy(), oenv.heading_error.copy(), ms)
        cc[:, 1] += rng.normal(0, 0.3, n).astype(np.float32)
        man = rng.integers(0, 4, n).astype(np.int32) if t % 6 == 0 else np.zeros(n, np.int32)
        env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        oenv.step(cc.astype(np.float64), man)
        _compare_batch(env, oenv, t, stats)
    env.close()


def test_cuda_layer_known_answers():
    """The reference's own unit-test answers (test/test_layer.py, test/test_helper.py) on the DEVICE functions."""
    import layer_kats as K
    from tinycarlo_b200 import TinyCarloVecEnv

    def env_for(nodes, edges):
        data = {"width": 100, "height": 100,
                "lanelines": {"test": {"layer_color": [0, 0, 0], "nodes": [list(map(float, p)) for p in nodes], "edges": [list(e) for e in edges]}},
                "lanepath": {"layer_color": [0, 0, 0], "nodes": [[0.0, 0.0], [1.0, 0.0]], "edges": [[0, 1]]}}
        import json, tempfile, os
        d = tempfile.mkdtemp()
        with open(os.path.join(d, "m.json"), "w") as fh:
            json.dump(data, fh)
        cfg = {"sim": {"observation_space_format": "classes"}, "car": {}, "camera": {"resolution": [8, 8], "max_range": 1.0},
               "map": {"json_path": os.path.join(d, "m.json"), "pixel_per_meter": 1}}
        return TinyCarloVecEnv(cfg, 1, device="cuda:0")

    for nodes, edges, cases in K.NEAREST_EDGE:
        env = env_for(nodes, edges)
        for pos, want in cases:
            assert env.debug_layer_query(0, pos)[0] == want, (pos, want)
    for nodes, edges, cases in K.NEAREST_EDGE_ORIENT:
        env = env_for(nodes, edges)
        for (pos, o), want in cases:
            assert env.debug_layer_query(1, pos, angle=o)[0] == (-1 if want is None else want), (pos, o, want)
    for nodes, edges, cases in K.WITHIN_BOUNDS:
        env = env_for(nodes, edges)
        for pos, want in cases:
            assert bool(env.debug_layer_query(2, pos, edge=edges[0])[0]) == want, (nodes, edges, pos)
    for nodes, edges, cases in K.DISTANCE_TO_EDGE_EXACT:
        env = env_for(nodes, edges)
        for pos, want in cases:
            assert env.debug_layer_query(3, pos, edge=edges[0])[1] == want
    for nodes, edges, cases in K.DISTANCE_TO_EDGE_CLOSE:
        env = env_for(nodes, edges)
        for pos, want in cases:
            assert abs(env.debug_layer_query(3, pos, edge=edges[0])[1] - want) < 1e-5
    env = env_for([(0, 0), (1, 0)], [(0, 1)])
    for a, want in K.CLIP_ANGLE:
        assert env.debug_layer_query(4, angle=a)[1] == want


def test_cuda_autoreset_next_step_matches_oracle():
    """autoreset="next_step": the step after an env finished resets it inside the kernel (action ignored, reward 0,
    empty info, reset observation). Emulated on the oracle with explicit resets on the same spawn nodes."""
    n, steps = 1024, 60
    cfg = make_config("simple_layout", "classes", cam={"resolution": [84, 84]}, car={"max_velocity": 0.15})
    env = _vec(cfg, n, autoreset="next_step")
    oenv = oracle_env(cfg, n)
    rng = np.random.default_rng(321)
    env.reset(seed=9)
    # the device streams against their host model (numpy-compatible PCG64 per env): column k = node of an env's k-th reset
    from tinycarlo_b200.spawn import SpawnSampler
    host_draws = SpawnSampler(env.map, n, table_len=32).seed(9)
    n_drawn = np.ones(n, np.int64)
    assert np.array_equal(env._spawn_nodes.cpu().numpy(), host_draws[:, 0])
    oenv.reset(host_draws[:, 0])
    done = np.zeros(n, bool)
    n_resets = 0
    for t in range(steps):
        cc = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        man = rng.integers(0, 4, n).astype(np.int32)
        nodes = host_draws[np.arange(n), n_drawn]   # what a finished env must draw inside this step
        n_drawn[done] += 1
        obs, reward, term, trunc, info = env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        # oracle: step the live envs, reset the finished ones
        keep_sf, keep_si, keep_obs = oenv.sf.copy(), oenv.si.copy(), oenv.obs.copy()
        oenv.step(cc.astype(np.float64), man)
        if done.any():
            oenv.sf[done], oenv.si[done], oenv.obs[done] = keep_sf[done], keep_si[done], keep_obs[done]
            oenv.reset(nodes, mask=done)
            oenv.info[done] = 0
            oenv.terminated[done] = 0
            oenv.truncated[done] = 0
            n_resets += int(done.sum())
        assert np.array_equal(obs.cpu().numpy(), oenv.obs), t
        np.testing.assert_allclose(env.out["info_f64"].cpu().numpy(), oenv.info, rtol=RTOL64, atol=1e-11)
        assert np.array_equal(term.cpu().numpy(), oenv.terminated.astype(bool)) and np.array_equal(trunc.cpu().numpy(), oenv.truncated.astype(bool))
        st = env.state_dict()
        np.testing.assert_allclose(st["sf"].cpu().numpy()[:, :7], oenv.sf[:, :7], rtol=RTOL64, atol=1e-11)
        assert np.array_equal(st["si"].cpu().numpy()[:, :10], oenv.si[:, :10])
        done = (oenv.terminated | oenv.truncated).astype(bool)
        assert np.array_equal(env.done_flags.cpu().numpy().astype(bool), done)
        assert np.array_equal(env._spawn_nodes.cpu().numpy(), host_draws[np.arange(n), n_drawn - 1]), "device spawn draw"
    assert n_resets > 0
    env.close()


def test_cuda_full_size_properties():
    """BASELINE.json's headline size (16384 envs, Knuffingen, 5x480x640 classes = 25 GB of observations), checked through
    size-independent properties plus a sample of envs against the oracle."""
    n, reps = 16384, 256
    cfg = make_config("knuffingen", "classes", cam={"resolution": [480, 640]})
    env = _vec(cfg, n)
    rng = np.random.default_rng(2024)
    # env i and env i + k*reps share spawn node and actions -> must stay identical (replication invariance)
    spawn_pool = np.array(cfg["map"]["spawn_points"], np.int32)
    spawn = np.tile(spawn_pool[rng.integers(0, len(spawn_pool), reps)], n // reps)
    env.reset(seed=0, spawn_nodes=torch.from_numpy(spawn))
    sample = np.arange(0, reps, 8)                                   # 32 distinct envs checked against the oracle
    oenv = oracle_env(cfg, len(sample))
    oenv.reset(spawn[sample])
    man = np.tile(rng.integers(0, 4, reps).astype(np.int32), n // reps)
    for t in range(6):
        cc = np.tile(np.stack([rng.uniform(0.3, 1, reps), rng.uniform(-1, 1, reps)], 1).astype(np.float32), (n // reps, 1))
        obs, reward, term, trunc, info = env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        oenv.step(cc[sample].astype(np.float64), man[sample])
        assert np.array_equal(obs[torch.from_numpy(sample).cuda()].cpu().numpy(), oenv.obs), t
        np.testing.assert_allclose(env.out["info_f64"][torch.from_numpy(sample).cuda()].cpu().numpy(), oenv.info, rtol=RTOL64, atol=1e-11)
    # only 0 / 255 and replicas identical — in chunks of `reps` envs, so that the temporaries stay small next to the 25 GB tensor
    first = obs[:reps]
    total_px = 0
    for k in range(n // reps):
        chunk = obs[k * reps:(k + 1) * reps]
        assert torch.equal(chunk, first), f"replica block {k} diverged"
        if k % 8 == 0:
            assert bool(((chunk == 0) | (chunk == 255)).all())
            total_px += int((chunk > 0).sum())
    assert total_px > 0
    # re-rendering the same poses is idempotent, and the RGB render marks exactly the union of the class masks
    keep = obs[:512].clone()
    env.render_obs()
    assert torch.equal(obs[:512], keep) and torch.equal(obs[n - reps:], first)
    rgb = env.render_rgb()[:64]
    assert torch.equal((rgb > 0).any(dim=3), (obs[:64] > 0).any(dim=1))
    env.close()


def test_cuda_checkpoint_restore_resumes_bit_for_bit():
    """checkpoint() / restore(): a rollout resumed in a fresh env continues exactly (state, spawn streams, autoreset flags)."""
    n = 2048
    cfg = make_config("simple_layout", "classes", cam={"resolution": [84, 84]}, car={"max_velocity": 0.15})
    rng = np.random.default_rng(5)
    acts = [(torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32)).cuda(), torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).cuda())
            for _ in range(90)]
    env = _vec(cfg, n, autoreset="next_step")
    env.reset(seed=77)
    for cc, man in acts[:40]:
        env.step({"car_control": cc, "maneuver": man})
    ck = env.checkpoint()
    tail_a = []
    force = {k: torch.from_numpy(rng.random(n) < 0.1).cuda() for k in (3, 4, 5, 20, 33)}   # forced episode ends -> spawn draws
    for k, (cc, man) in enumerate(acts[40:]):
        if k in force:
            env.mark_done(force[k])
        obs, r, te, tr, _ = env.step({"car_control": cc, "maneuver": man})
        tail_a.append((obs.clone(), r.clone(), te.clone(), tr.clone(), env.out["info_f64"].clone()))
    env.close()
    env2 = _vec(cfg, n, autoreset="next_step")
    env2.reset(seed=1)          # different streams, then overwritten by the checkpoint
    env2.restore(ck)
    for k, ((cc, man), want) in enumerate(zip(acts[40:], tail_a)):
        if k in force:
            env2.mark_done(force[k])
        obs, r, te, tr, _ = env2.step({"car_control": cc, "maneuver": man})
        assert torch.equal(env2.out["info_f64"], want[4]), "info"
        assert torch.equal(te, want[2]) and torch.equal(tr, want[3]), "flags"
        assert torch.equal(r, want[1]), "reward"
        assert torch.equal(obs, want[0]), "obs"

    assert not torch.equal(env2._rng_state.cpu(), ck["spawn_rng"]), "no spawn draw happened after the checkpoint"
    env2.close()


def test_cuda_grouped_resolutions_match_oracle():
    """BASELINE config 5: per-env resolution via resolution groups (one dense tensor per group), per-env camera pitch / fov
    and car parameters on top; global env indices (hence spawn seeds) run through the groups."""
    from tinycarlo_b200 import TinyCarloGroupedVecEnv
    groups = [(192, [84, 84]), (128, [128, 160]), (64, [240, 320])]
    n = sum(g[0] for g in groups)
    cfg = make_config("knuffingen", "classes")
    rng = np.random.default_rng(31)
    env = TinyCarloGroupedVecEnv(cfg, groups, device="cuda:0")
    pitch = rng.integers(10, 20, n).astype(np.float64)
    fov = rng.integers(90, 130, n).astype(np.float64)
    env.set_camera_params(orientation=np.stack([pitch, np.zeros(n), np.zeros(n)], 1), fov=fov)
    wb = rng.uniform(0.04, 0.06, n)
    env.set_car_params(wheelbase=wb)
    obs_list, info = env.reset(seed=3)
    oenvs = []
    for e, (cnt, res) in zip(env.envs, groups):
        c = make_config("knuffingen", "classes", cam={"resolution": res})
        o = oracle_env(c, cnt, cam_rows=e._cam_rows.copy(), car_rows=e._car_rows.copy())
        o.reset(e._spawn_nodes.cpu().numpy())
        oenvs.append(o)
    # a sharded draw equals one flat draw: group g's env i is global env offsets[g] + i
    from tinycarlo_b200.spawn import SpawnSampler
    flat = SpawnSampler(env.envs[0].map, n, table_len=1).seed(3)[:, 0]
    assert np.array_equal(np.concatenate([e._spawn_nodes.cpu().numpy() for e in env.envs]), flat)
    for t in range(15):
        cte = np.concatenate([o.cte for o in oenvs])
        he = np.concatenate([o.heading_error for o in oenvs])
        cc = stanley_actions(cte, he, cfg["car"]["max_steering_angle"])
        cc[:, 1] += rng.normal(0, 0.2, n).astype(np.float32)
        man = np.zeros(n, np.int32)
        obs_list, reward, term, trunc, info = env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        torch.cuda.synchronize()
        for o, sl, obs in zip(oenvs, env._slices(), obs_list):
            o.step(cc[sl].astype(np.float64), man[sl])
            assert np.array_equal(obs.cpu().numpy(), o.obs), t
        np.testing.assert_allclose(info["cte"].cpu().numpy(), np.concatenate([o.cte for o in oenvs]), rtol=RTOL32, atol=1e-7)
        assert np.array_equal(trunc.cpu().numpy(), np.concatenate([o.truncated for o in oenvs]).astype(bool))
    env.close()


@pytest.mark.parametrize("res,n", [([128, 160], 512), ([480, 640], 64), ([84, 84], 256)])
def test_cuda_policy_formats_equal_the_u8_masks(res, n):
    """obs formats beyond the reference: "classes_bits" and "classes_bf16" carry exactly the u8 class masks."""
    cfg = make_config("knuffingen", "classes", cam={"resolution": res})
    envs = {f: _vec(cfg, n, obs_format=f) for f in ("classes", "classes_bf16")}
    oenv = oracle_env(cfg, n)
    if (res[0] * res[1]) % 32 == 0:
        envs["classes_bits"] = _vec(cfg, n, obs_format="classes_bits")
    rng = np.random.default_rng(8)
    for e in envs.values():
        e.reset(seed=4)
    oenv.reset(envs["classes"]._spawn_nodes.cpu().numpy())
    H, W = res
    for t in range(6):
        cc = torch.from_numpy(np.stack([rng.uniform(0.3, 1, n), rng.uniform(-1, 1, n)], 1).astype(np.float32)).cuda()
        man = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).cuda()
        for e in envs.values():
            e.step({"car_control": cc, "maneuver": man})
        oenv.step(cc.cpu().numpy().astype(np.float64), man.cpu().numpy())
        ref = envs["classes"].obs
        assert int((ref > 0).sum()) > 0 and np.array_equal(ref.cpu().numpy(), oenv.obs)
        # every format against the ORACLE's masks (not against this library's own u8 output)
        bf = envs["classes_bf16"].obs
        assert bf.dtype == torch.bfloat16
        assert np.array_equal(bf.float().cpu().numpy(), (oenv.obs > 0).astype(np.float32)), "bf16 must be exactly 0.0 / 1.0 where the oracle's mask is 0 / 255"
        if "classes_bits" in envs:
            words = envs["classes_bits"].obs.cpu().numpy().view(np.uint32)           # [n, C, H*W/32]
            bits = np.unpackbits(words.view(np.uint8), axis=-1, bitorder="little")[..., : H * W].reshape(n, -1, H, W)
            assert np.array_equal(bits * 255, oenv.obs)
    for e in envs.values():
        e.close()


def test_cuda_visible_set_tables_do_not_change_frames():
    """Block-per-env kernel: frames rendered through the per-cell visible-set tables equal those of the whole-graph camera pass
    (TC_CULL=0), for shipped and for per-env randomised cameras, over a driven rollout with auto-reset."""
    import os
    n = 4096
    cfg = make_config("knuffingen", "classes", cam={"resolution": [128, 160]}, car={"max_velocity": 0.3})
    rng = np.random.default_rng(11)
    env = _vec(cfg, n, autoreset="next_step")
    info = env.cull_info()
    assert info["radius"] > 0 and info["max_nodes"] < 0.6 * len(env.map.ll_nodes), info
    os.environ["TC_CULL"] = "0"
    try:
        ref = _vec(cfg, n, autoreset="next_step")
        assert ref.cull_info()["radius"] == -1.0
        for phase in range(2):
            if phase == 1:   # config-5 style cameras: the tables are rebuilt for the largest reach
                kw = dict(orientation=np.stack([rng.uniform(10, 30, n), rng.uniform(-3, 3, n), rng.uniform(-20, 20, n)], 1).round(1),
                          fov=rng.integers(70, 120, n).astype(float), max_range=rng.uniform(0.3, 0.8, n).round(2))
                ref.set_camera_params(**kw)
                del os.environ["TC_CULL"]
                env.set_camera_params(**kw)
                os.environ["TC_CULL"] = "0"
                assert env.cull_info()["radius"] > info["radius"] and ref.cull_info()["radius"] == -1.0
            env.reset(seed=3)
            ref.reset(seed=3)
            assert torch.equal(env.obs, ref.obs)
            for t in range(30):
                cc = torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32)).cuda()
                man = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).cuda()
                env.step({"car_control": cc, "maneuver": man})
                ref.step({"car_control": cc, "maneuver": man})
                assert torch.equal(env.obs, ref.obs), (phase, t)
            assert env.obs.any()
        ref.close()
    finally:
        os.environ.pop("TC_CULL", None)
    env.close()


def test_cuda_step_is_graph_capturable():
    """step() enqueues kernels on torch's current stream and neither allocates nor synchronises, so a rollout loop (policy ops +
    step, in-kernel autoreset included) can be captured in a CUDA graph; replays equal eager stepping bit for bit."""
    n = 1024
    cfg = make_config("simple_layout", "classes", cam={"resolution": [84, 84]}, car={"max_velocity": 0.15})
    envs = [_vec(cfg, n, autoreset="next_step") for _ in range(2)]
    for e in envs:
        e.reset(seed=4)
    cc = [torch.zeros((n, 2), device="cuda") for _ in range(2)]
    man = torch.zeros(n, dtype=torch.int32, device="cuda")

    def policy_and_step(e, c):   # Stanley controller on the previous step's info (examples/stanley_control.py:56-58)
        o = e.out
        c[:, 0] = 0.9
        c[:, 1] = (o["heading_error"] + torch.atan2(4.0 * o["cte"], torch.full_like(o["cte"], 0.8))) * (180.0 / np.pi / 30.0)
        e.step({"car_control": c, "maneuver": man})
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):          # warm-up on the side stream, as torch's capture rules ask
            policy_and_step(envs[0], cc[0])
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        policy_and_step(envs[1], cc[1])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        policy_and_step(envs[0], cc[0])
    policy_and_step(envs[1], cc[1])   # the captured call did not execute: one eager step for the twin ... and one replay
    g.replay()
    for t in range(60):
        g.replay()
        policy_and_step(envs[1], cc[1])
        if t % 10 == 9:
            torch.cuda.synchronize()
            assert torch.equal(envs[0].obs, envs[1].obs) and torch.equal(envs[0].out["info_f64"], envs[1].out["info_f64"]), t
            assert torch.equal(envs[0].done_flags, envs[1].done_flags)
    for e in envs:
        e.close()


@pytest.mark.parametrize("thickness", [2, 1])
def test_cuda_banded_env_kernel_with_many_segments(thickness):
    """Large RGB / bit-packed frames go through the banded block-per-env kernel. With a long camera range a frame shows more
    segments than the kernel keeps set up at once (48): the per-band rounds must give the oracle's frames too. Thickness 1
    exercises the Bresenham primitives in bands."""
    n, steps = 48, 8
    cam = {"resolution": [480, 640], "max_range": 3.0, "orientation": [30, 0, 0], "line_thickness": thickness}
    cfg_rgb = make_config("simple_layout", "rgb", cam=cam, car={"max_velocity": 0.15})
    cfg_cls = make_config("simple_layout", "classes", cam=cam, car={"max_velocity": 0.15})
    env_rgb, env_bits = _vec(cfg_rgb, n), _vec(cfg_cls, n, obs_format="classes_bits")
    env_seg = _vec(cfg_cls, n, debug_segments=True)      # unfused path: exports the segment counts
    o_rgb, o_cls = oracle_env(cfg_rgb, n), oracle_env(cfg_cls, n)
    rng = np.random.default_rng(3)
    for e in (env_rgb, env_bits, env_seg):
        e.reset(seed=6)
    for o in (o_rgb, o_cls):
        o.reset(env_rgb._spawn_nodes.cpu().numpy())
    most = 0
    for t in range(steps):
        cc = np.stack([rng.uniform(0.3, 1, n), rng.uniform(-1, 1, n)], 1).astype(np.float32)
        man = rng.integers(0, 4, n).astype(np.int32)
        for e in (env_rgb, env_bits, env_seg):
            e.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        for o in (o_rgb, o_cls):
            o.step(cc.astype(np.float64), man)
        assert np.array_equal(env_rgb.obs.cpu().numpy(), o_rgb.obs), ("rgb", t)
        words = env_bits.obs.cpu().numpy().view(np.uint32)
        bits = np.unpackbits(words.view(np.uint8), axis=-1, bitorder="little")[..., : 480 * 640].reshape(n, -1, 480, 640)
        assert np.array_equal(bits * 255, o_cls.obs), ("bits", t)
        most = max(most, int(env_seg.out["seg_count"].sum(1).max()))
    assert most > 48, most
    for e in (env_rgb, env_bits, env_seg):
        e.close()


@pytest.mark.parametrize("fmt,res", [("classes", [45, 71]), ("rgb", [45, 71]), ("classes", [33, 35]), ("rgb", [481, 643]), ("classes", [481, 643])])
def test_cuda_odd_resolutions_match_oracle(fmt, res):
    """Frames whose rows and per-env strides are not multiples of 16 bytes (unaligned heads / tails of the vector stores, partial
    plane words), small and large, through every render path."""
    n = 24 if res[0] > 400 else 192
    cfg = make_config("knuffingen", fmt, cam={"resolution": res})
    env, oenv = _vec(cfg, n), oracle_env(cfg, n)
    rng = np.random.default_rng(res[0])
    env.reset(seed=2)
    oenv.reset(env._spawn_nodes.cpu().numpy())
    assert np.array_equal(env.obs.cpu().numpy(), oenv.obs)
    for t in range(6):
        cc = np.stack([rng.uniform(0.3, 1, n), rng.uniform(-1, 1, n)], 1).astype(np.float32)
        man = rng.integers(0, 4, n).astype(np.int32)
        env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        oenv.step(cc.astype(np.float64), man)
        assert np.array_equal(env.obs.cpu().numpy(), oenv.obs), (fmt, res, t)
    assert oenv.obs.any()
    env.close()
