"""GPU parity tests of the BASELINE.json configs that sit above the plain step: the config-4 stack (tensor reward /
termination wrappers + mark_done + in-kernel autoreset + mixed maneuvers incl. u-turns) against the oracle stepped env by env
with the SCALAR wrappers, and the config-5 resolution groups sharded over ranks against the unsharded job. Plus the smaller
surface added with them: host-buffer steps with observations, graph capture, float32-ndarray actions, argument hardening."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from golden_util import Golden
from pair_util import make_config, oracle_env, stanley_actions

pytestmark = pytest.mark.gpu
RTOL64, RTOL32 = 1e-9, 1e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _vec(cfg, n, **kw):
    from tinycarlo_b200 import TinyCarloVecEnv
    return TinyCarloVecEnv(cfg, n, device="cuda:0", **kw)


class _ScalarFeed:
    """A scalar env whose step() returns what the oracle computed for one env: the base of a scalar wrapper stack
    (the wrappers restate tinycarlo/wrapper/reward.py, termination.py and are pinned to the reference's recorded rewards /
    terminations in tests/test_wrappers_cpu.py and tests/test_gpu_single_env.py)."""

    class _Car:
        def __init__(self, tw):
            self.track_width = tw

    def __init__(self, track_width):
        self.wrapped = False
        self.car = self._Car(track_width)
        self.next = None

    @property
    def unwrapped(self):
        return self

    def step(self, action):
        return self.next

    def reset(self, **kw):
        return None, {}


def test_cuda_config4_stack_matches_oracle_with_scalar_wrappers():
    """BASELINE config 4 on the device: Knuffingen, maneuvers uniform over {0,1,2,3} resampled every 25 steps (u-turn transitions,
    car.py:130), Stanley + Ornstein-Uhlenbeck noise (train_td3.py:42-44,143), CTESparseRewardWrapper(0.01) +
    CTETerminationWrapper(0.07, 5) as TENSOR wrappers feeding mark_done, in-kernel next-step autoreset with device spawn draws.
    Reference side: the oracle stepped in lockstep, every env's result passed through its own SCALAR wrapper stack, finished envs
    reset explicitly on the node the host model of the spawn streams predicts. 320 steps, every env, every step: rewards,
    terminations, truncations, frames (reset frames included), info, spawn draws."""
    from tinycarlo_b200 import wrapper as W
    from tinycarlo_b200.spawn import SpawnSampler
    n, steps = 256, 320
    cfg = make_config("knuffingen", "classes", cam={"resolution": [96, 128]}, car={"max_velocity": 0.25})
    base = _vec(cfg, n, autoreset="next_step")
    env = W.CTETerminationWrapper(W.CTESparseRewardWrapper(base, min_cte=0.01), max_cte=0.07, number_of_steps=5)
    assert base.wrapped is True
    oenv = oracle_env(cfg, n, wrapped=True)
    feeds = [_ScalarFeed(cfg["car"]["track_width"]) for _ in range(n)]
    stacks = [W.CTETerminationWrapper(W.CTESparseRewardWrapper(f, min_cte=0.01), max_cte=0.07, number_of_steps=5) for f in feeds]
    assert all(f.wrapped for f in feeds)
    rng = np.random.default_rng(44)
    env.reset(seed=21)
    draws = SpawnSampler(base.map, n, table_len=64).seed(21)
    n_drawn = np.ones(n, np.int64)
    assert np.array_equal(base._spawn_nodes.cpu().numpy(), draws[:, 0])
    oenv.reset(draws[:, 0])
    assert np.array_equal(base.obs.cpu().numpy(), oenv.obs)
    done = np.zeros(n, bool)
    man = np.zeros(n, np.int32)
    ou = np.zeros(n)
    seen = {"resets": 0, "terminated": 0, "truncated": 0, "uturn": 0, "reward": 0.0}
    for t in range(steps):
        if t % 25 == 0:
            man = rng.integers(0, 4, n).astype(np.int32)
        ou += 0.1 * (0.0 - ou) + 0.4 * rng.standard_normal(n)
        cc = stanley_actions(oenv.cte.copy(), oenv.heading_error.copy(), cfg["car"]["max_steering_angle"])
        cc[:, 1] = np.clip(cc[:, 1] + ou, -1, 1).astype(np.float32)
        nodes = draws[np.arange(n), n_drawn]          # what a finished env must draw inside this step
        n_drawn[done] += 1
        seen["uturn"] += int(((man == 2) & (oenv.si[:, 1] != 2) & ~done).sum())
        obs, reward, term, trunc, info = env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        keep_sf, keep_si, keep_obs = oenv.sf.copy(), oenv.si.copy(), oenv.obs.copy()
        oenv.step(cc.astype(np.float64), man)
        want_r, want_term, want_trunc = np.zeros(n), np.zeros(n, bool), np.zeros(n, bool)
        if done.any():     # the envs that finished at the previous step are reset by this one: action ignored, reward 0, empty info
            oenv.sf[done], oenv.si[done], oenv.obs[done] = keep_sf[done], keep_si[done], keep_obs[done]
            oenv.reset(nodes, mask=done)
            oenv.info[done] = 0
            oenv.terminated[done] = 0
            oenv.truncated[done] = 0
            seen["resets"] += int(done.sum())
        for i in np.nonzero(~done)[0]:
            feeds[i].next = (None, float(oenv.reward[i]), bool(oenv.terminated[i]), bool(oenv.truncated[i]),
                             {"cte": float(oenv.cte[i]), "heading_error": float(oenv.heading_error[i]), "velocity": float(oenv.velocity[i])})
            _, r, te, tr, _ = stacks[i].step(None)
            want_r[i], want_term[i], want_trunc[i] = r, te, tr
        assert np.array_equal(base.reset_mask.cpu().numpy(), done), t
        assert np.array_equal(term.cpu().numpy(), want_term), f"terminated at step {t}"
        assert np.array_equal(trunc.cpu().numpy(), want_trunc), f"truncated at step {t}"
        np.testing.assert_allclose(reward.cpu().numpy(), want_r, rtol=RTOL32, atol=1e-7, err_msg=f"reward at step {t}")
        assert np.array_equal(obs.cpu().numpy(), oenv.obs), f"frames at step {t}"
        np.testing.assert_allclose(base.out["info_f64"].cpu().numpy(), oenv.info, rtol=RTOL64, atol=1e-11)
        st = base.state_dict()
        np.testing.assert_allclose(st["sf"].cpu().numpy()[:, :7], oenv.sf[:, :7], rtol=RTOL64, atol=1e-11)
        assert np.array_equal(st["si"].cpu().numpy()[:, :10], oenv.si[:, :10]), f"local path at step {t}"
        assert np.array_equal(base._spawn_nodes.cpu().numpy(), draws[np.arange(n), n_drawn - 1]), "device spawn draw"
        done = want_term | want_trunc
        assert np.array_equal(base.done_flags.cpu().numpy().astype(bool), done), f"autoreset flags at step {t}"
        seen["terminated"] += int(want_term.sum())
        seen["truncated"] += int(want_trunc.sum())
        seen["reward"] += float(want_r.sum())
    print("config-4 stack:", seen)
    assert seen["resets"] > 50 and seen["terminated"] > 20 and seen["uturn"] > 100 and seen["reward"] > 0, seen
    assert n_drawn.max() < 64
    base.close()


def test_cuda_config5_sharded_groups_equal_unsharded():
    """BASELINE config 5: per-env camera pitch / fov / position and car parameters by GLOBAL env index, three resolution groups,
    every rank holding a contiguous slice of every group (distributed.shard_groups). Two ranks' envs (built side by side on this
    GPU) must equal the unsharded job env for env - spawn draws, frames, info - and the unsharded job must equal the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import workloads as WL
    from tinycarlo_b200 import TinyCarloGroupedVecEnv
    from tinycarlo_b200.distributed import shard_groups
    total, world = 384, 2
    w = WL.WORKLOADS[5]
    sizes = WL.group_sizes(total)
    p = WL.config5_params(total)
    cfg = make_config("knuffingen", "classes", car=w["car"])

    def build(rank, wsize):
        local, offs = shard_groups(sizes, rank, wsize)
        e = TinyCarloGroupedVecEnv(cfg, list(zip(local, w["groups"])), device="cuda:0", group_index_offsets=offs, autoreset="next_step")
        for sub, n, o in zip(e.envs, local, offs):
            sl = slice(o, o + n)
            sub.set_camera_params(position=p["position"][sl], orientation=p["orientation"][sl], fov=p["fov"][sl])
            sub.set_car_params(**{k: v[sl] for k, v in p["car"].items()})
        return e, local, offs
    full, _, full_offs = build(0, 1)
    shards = [build(r, world) for r in range(world)]
    full.reset(seed=7)
    for e, _, _ in shards:
        e.reset(seed=7)
    oenvs = []
    for sub, res in zip(full.envs, w["groups"]):
        o = oracle_env(make_config("knuffingen", "classes", car=w["car"], cam={"resolution": res}), sub.num_envs,
                       cam_rows=sub._cam_rows.copy(), car_rows=sub._car_rows.copy())
        o.reset(sub._spawn_nodes.cpu().numpy())
        oenvs.append(o)
    rng = np.random.default_rng(5)
    ms = p["car"]["max_steering_angle"]
    alive = [np.ones(o.n, bool) for o in oenvs]
    checked = 0
    for t in range(16):
        # global action arrays; group g of the full job owns global indices [full_offs[g], full_offs[g] + sizes[g])
        cte = np.concatenate([o.cte for o in oenvs])
        he = np.concatenate([o.heading_error for o in oenvs])
        cc = stanley_actions(cte, he, ms)
        cc[:, 1] += rng.normal(0, 0.2, total).astype(np.float32)
        man = (rng.integers(0, 4, total) * (t % 4 == 0)).astype(np.int32)
        obs_full, r_full, te_full, tr_full, info_full = full.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
        torch.cuda.synchronize()
        for g, (o, sl, obs) in enumerate(zip(oenvs, full._slices(), obs_full)):
            o.step(cc[sl].astype(np.float64), man[sl])
            live = alive[g]       # envs that have not finished yet (afterwards the in-kernel autoreset takes over; the oracle
            assert np.array_equal(obs.cpu().numpy()[live], o.obs[live]), ("unsharded vs oracle", t)   # side of that is the config-4 test)
            assert np.array_equal(te_full[sl].cpu().numpy()[live], o.terminated.astype(bool)[live])
            alive[g] = live & ~(o.terminated | o.truncated).astype(bool)
            checked += int(live.sum())
        for e, local, offs in shards:
            idx = np.concatenate([np.arange(o, o + n) for o, n in zip(offs, local)])       # global ids of this rank's envs, group by group
            obs_s, r_s, te_s, tr_s, info_s = e.step({"car_control": torch.from_numpy(cc[idx]).cuda(), "maneuver": torch.from_numpy(man[idx]).cuda()})
            torch.cuda.synchronize()
            for g, (o, n) in enumerate(zip(offs, local)):
                lo = o - full_offs[g]
                assert torch.equal(obs_s[g], obs_full[g][lo:lo + n]), ("frames", t, g)
            ti = torch.from_numpy(idx).cuda()
            assert torch.equal(info_s["cte"], info_full["cte"][ti]) and torch.equal(r_s, r_full[ti]) and torch.equal(tr_s, tr_full[ti])
    assert checked > 10 * total and sum(int((~a).sum()) for a in alive) > 0   # some envs finished: the shards went through autoresets too
    # spawn draws: the shards drew exactly what the unsharded job drew for the same global envs
    flat = np.concatenate([e._spawn_nodes.cpu().numpy() for e in full.envs])
    for e, local, offs in shards:
        idx = np.concatenate([np.arange(o, o + n) for o, n in zip(offs, local)])
        assert np.array_equal(np.concatenate([s._spawn_nodes.cpu().numpy() for s in e.envs]), flat[idx])
    full.close()
    for e, _, _ in shards:
        e.close()


def test_cuda_float32_ndarray_actions_within_tolerance():
    """fed float32 ndarrays (what action_space.sample() yields) the reference keeps parts of its car state in float32
    under numpy >= 2, so its trace differs from its own float64 trace at ~1e-6 relative. This library always computes in float64
    on the float32 action VALUES; against a reference trace recorded with float32 ndarray actions (tests/golden/knuff_f32_actions.npz)
    it stays within the north star's 1e-5 relative on pose / CTE / heading / distances, with identical local paths. Masks may then
    differ by single boundary pixels (the poses differ at 1e-7), which is reported."""
    g = Golden("knuff_f32_actions")
    assert g.meta["action_dtype"] == "float32"
    env = _vec(g.cfg, 1)
    worst = {"pos": 0.0, "cte": 0.0, "heading": 0.0, "dist": 0.0}
    diff_px = frames = 0
    for f in range(g.F):
        if g["ev_kind"][f] == 0:
            obs, _ = env.reset(spawn_nodes=torch.tensor([int(g["spawn_node"][f])], dtype=torch.int32))
        else:
            t = int(g["ev_step"][f])
            act = {"car_control": torch.tensor(g["act_cc"][t][None], dtype=torch.float32, device="cuda:0"),
                   "maneuver": torch.tensor(g["act_man"][t][None], dtype=torch.int32, device="cuda:0")}
            obs, reward, term, trunc, info = env.step(act)
            i64 = env.out["info_f64"][0].cpu().numpy()
            for key, got, want in (("cte", i64[0], g["cte"][f]), ("heading", i64[1], g["heading"][f])):
                np.testing.assert_allclose(got, want, rtol=RTOL32, atol=2e-6, err_msg=f"{key} frame {f}")
                worst[key] = max(worst[key], abs(got - want))
            np.testing.assert_allclose(i64[4:], g["dist"][f], rtol=RTOL32, atol=2e-6)
            worst["dist"] = max(worst["dist"], float(np.abs(i64[4:] - g["dist"][f]).max()))
            assert bool(trunc[0]) == bool(g["truncated"][f])
        sf, si = (x[0].cpu().numpy() for x in env.state_dict().values())
        np.testing.assert_allclose(sf[:2], g["pos"][f], rtol=RTOL32, atol=2e-6)
        worst["pos"] = max(worst["pos"], float(np.abs(sf[:2] - g["pos"][f]).max()))
        L = int(g["lp_len"][f])
        assert si[0] == L and np.array_equal(si[2:2 + 2 * L].reshape(L, 2), g["lp"][f][:L]), f
        frames += 1
        diff_px += int((obs[0].cpu().numpy() != g.classes_frame(f)).sum())
    print(f"float32-ndarray actions: max abs differences {worst}, {diff_px} differing mask pixels over {frames} frames")
    assert diff_px <= 40 * frames // 100 + 40   # boundary pixels only
    env.close()


def test_cuda_action_arguments_are_hardened():
    """non-contiguous float64 actions, step before reset, NaN actions, degenerate car parameters."""
    from tinycarlo_b200 import TinyCarloError
    n = 64
    cfg = make_config("knuffingen", "classes", cam={"resolution": [32, 48]})
    env, twin = _vec(cfg, n), _vec(cfg, n)
    man = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    with pytest.raises(TinyCarloError, match="reset"):
        env.step({"car_control": torch.zeros((n, 2), device="cuda:0"), "maneuver": man})
    for e in (env, twin):
        e.reset(seed=3)
    rng = np.random.default_rng(0)
    wide = torch.from_numpy(rng.uniform(-1, 1, (n, 5))).cuda()          # float64 [n,5]; the action is a strided slice of it
    for _ in range(5):
        env.step({"car_control": wide[:, 1:3], "maneuver": man})         # non-contiguous float64
        twin.step({"car_control": wide[:, 1:3].contiguous(), "maneuver": man})
        assert torch.equal(env.out["info_f64"], twin.out["info_f64"]) and torch.equal(env.obs, twin.obs)
    assert env.out["velocity"].abs().max() > 0
    # a NaN action propagates like np.clip does (env.py:118) and poisons only its own env; nothing crashes or hangs
    cc = torch.zeros((n, 2), device="cuda:0")
    cc[:, 0] = 0.5
    cc[3, 0] = float("nan")
    cc[4, 1] = float("inf")      # clipped to 1
    for _ in range(3):
        env.step({"car_control": cc, "maneuver": torch.full((n,), 7, dtype=torch.int32, device="cuda:0")})   # maneuver outside 0..3 too
    torch.cuda.synchronize()
    v = env.out["velocity"].cpu().numpy()
    assert np.isnan(v[3]) and np.isfinite(np.delete(v, 3)).all()
    with pytest.raises(ValueError, match="wheelbase"):
        env.set_car_params(wheelbase=0.0)
    with pytest.raises(ValueError, match="max_velocity"):
        env.set_car_params(max_velocity=float("nan"))
    env.step({"car_control": cc, "maneuver": man})       # the rejected parameters were not uploaded
    torch.cuda.synchronize()
    env.close()
    twin.close()


def test_cuda_step_host_with_observations_and_capture():
    """tc_step_host_obs (bit-packed frames to pinned host memory in one copy) and TinyCarloVecEnv.capture() (policy + step in a
    CUDA graph) against plain stepping."""
    n = 256
    cfg = make_config("knuffingen", "classes", cam={"resolution": [96, 128]})
    a, b = _vec(cfg, n, autoreset="next_step", obs_format="classes_bits"), _vec(cfg, n, autoreset="next_step", obs_format="classes_bits")
    for e in (a, b):
        e.reset(seed=2)
    pin = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt).pin_memory()   # noqa: E731
    h = [pin(n, 2), pin(n, dt=torch.int32), pin(n), pin(n, dt=torch.uint8), pin(n, dt=torch.uint8), pin(n), pin(n)]
    h_obs = torch.zeros(a.obs.shape, dtype=a.obs.dtype).pin_memory()
    rng = np.random.default_rng(1)
    for t in range(8):
        h[0].numpy()[:] = rng.uniform(-1, 1, (n, 2))
        h[1].numpy()[:] = rng.integers(0, 4, n)
        a.step_host(*h, obs_host=h_obs)
        _, r, te, tr, info = b.step({"car_control": h[0].cuda(), "maneuver": h[1].cuda()})
        assert torch.equal(h_obs, b.obs.cpu()) and torch.equal(h[2], r.cpu()) and torch.equal(h[5], info["cte"].cpu())
        assert torch.equal(h[3].bool(), te.cpu()) and torch.equal(h[4].bool(), tr.cpu())
    assert h_obs.any()
    with pytest.raises(ValueError):
        a.step_host(h[0].cuda(), *h[1:])
    # graph capture: 4 steps per replay with an on-device policy
    cc = torch.zeros((n, 2), device="cuda:0")
    cc2 = torch.zeros((n, 2), device="cuda:0")
    man = torch.zeros(n, dtype=torch.int32, device="cuda:0")

    def make_policy(buf):
        def policy(e):
            o = e.out
            buf[:, 0] = 0.9
            buf[:, 1] = (o["heading_error"] + torch.atan2(4.0 * o["cte"], torch.full_like(o["cte"], 0.8))) * (180.0 / np.pi / 30.0)
            return {"car_control": buf, "maneuver": man}
        return policy
    pa, pb = make_policy(cc), make_policy(cc2)
    for e in (a, b):
        e.reset(seed=9)
    graph = a.capture(pa, steps=4, warmup=3)       # warm-up advanced `a` by 3 steps
    for _ in range(3):
        b.step(pb(b))
    for k in range(10):
        graph.replay()
        for _ in range(4):
            b.step(pb(b))
        torch.cuda.synchronize()
        assert torch.equal(a.obs, b.obs) and torch.equal(a.out["info_f64"], b.out["info_f64"]), k
    a.close()
    b.close()


def test_gym_make_with_a_gymnasium_on_the_path():
    """gym.make("tinycarlo-v2", config=...) (tinycarlo/__init__.py:3): gymnasium is not installed in this image, so the check
    runs in a fresh interpreter with the ~100-line stand-in of tests/golden/gym_stub on the path (the same one the golden
    generator imports the reference with): registration on import, make(), reset(seed), step, spaces."""
    code = (
        "import numpy as np, gymnasium as gym, tinycarlo_b200\n"
        "from tinycarlo_b200.config import make_config\n"
        "from tinycarlo_b200 import gym_compat\n"
        "assert gym_compat.HAVE_GYMNASIUM\n"
        "assert gym.envs.registration.registry['tinycarlo-v2'] == 'tinycarlo_b200.env:TinyCarloEnv'\n"
        "env = gym.make('tinycarlo-v2', config=make_config('simple_layout', 'rgb', cam={'resolution': [32, 48]}))\n"
        "assert isinstance(env, gym.Env)\n"
        "obs, info = env.reset(seed=0)\n"
        "assert obs.shape == (32, 48, 3) and info['cte'] == 0\n"
        "obs, r, te, tr, info = env.step({'car_control': [0.5, 0.1], 'maneuver': 0})\n"
        "assert obs.shape == (32, 48, 3) and isinstance(r, float) and len(info['local_path']) == 4\n"
        "assert env.action_space['car_control'].shape == (2,) and env.observation_space.shape == (32, 48, 3)\n"
        "env.close()\nprint('gym.make ok')\n")
    envv = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "golden", "gym_stub"), ROOT]))
    out = subprocess.run([sys.executable, "-c", code], env=envv, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "gym.make ok" in out.stdout, out.stderr[-2000:]


def test_cuda_camera_randomisation_reuses_tables():
    """Per-episode camera randomisation (examples/train_stanley_il.py:52-57): the visible-set tables are built once per reach step
    and kept, so alternating between camera settings neither rebuilds on the host nor synchronises; frames stay the oracle's."""
    n = 96
    cfg = make_config("knuffingen", "classes", cam={"resolution": [96, 128]})
    env = _vec(cfg, n)
    rng = np.random.default_rng(3)
    settings = [dict(fov=80.0, max_range=0.5, pitch=22.0), dict(fov=120.0, max_range=0.9, pitch=12.0), dict(fov=95.0, max_range=0.55, pitch=18.0)]
    builds_after_first_cycle = None
    for ep in range(9):
        s_ = settings[ep % 3]
        env.set_camera_params(orientation=np.array([s_["pitch"], 0.0, 0.0]), fov=s_["fov"], max_range=s_["max_range"])
        oenv = oracle_env(cfg, n, cam_rows=env._cam_rows.copy())
        env.reset(seed=ep)
        oenv.reset(env._spawn_nodes.cpu().numpy())
        assert np.array_equal(env.obs.cpu().numpy(), oenv.obs), ep
        for t in range(3):
            cc = np.stack([rng.uniform(0.3, 1, n), rng.uniform(-1, 1, n)], 1).astype(np.float32)
            man = np.zeros(n, np.int32)
            env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
            oenv.step(cc.astype(np.float64), man)
            assert np.array_equal(env.obs.cpu().numpy(), oenv.obs), (ep, t)
        if ep == 2:
            builds_after_first_cycle = env.cull_stats()["builds"]
    st = env.cull_stats()
    print("cull stats:", st, env.cull_info())
    assert st["builds"] == builds_after_first_cycle <= 4, st     # the whole-graph set of construction + at most one per setting
    assert st["cache_hits"] >= 4, st
    env.close()


@pytest.mark.parametrize("case", ["packed128", "oneenv128", "packed84", "rgb96", "env240", "classes480", "rgb480", "bits480", "bits480_overflow", "bf16_128"])
def test_cuda_guard_bands_determinism_and_full_overwrite(case):
    """compute-sanitizer is closed on this GPU pool, so the memory-safety evidence is made here, per render path:
    * guard bands: the observation and every per-env output live inside larger buffers whose surroundings are filled with a sentinel;
      after stepping, the sentinels must be intact (an out-of-bounds store of a render / tracking kernel would hit them);
    * full overwrite: the observation buffer is poisoned with 0xAB before every step and must come back holding only legal values and
      exactly the oracle's frame (a store the kernel skipped, or one fed from uninitialised shared memory, shows up here);
    * determinism: the same step from the same state, 6 times over, gives bit-identical frames and outputs (the kernels overlay
      shared-memory regions and draw with atomicOr: a missing barrier would make results depend on timing)."""
    import os
    setenv, kw, n = {}, {}, 24
    if case == "packed128":
        cfg = make_config("knuffingen", "classes", cam={"resolution": [128, 160]}); n = 37
    elif case == "oneenv128":
        cfg = make_config("knuffingen", "classes", cam={"resolution": [128, 160]}); setenv = {"TC_ENV_PACK": "0"}; n = 19
    elif case == "packed84":
        cfg = make_config("simple_layout", "classes", cam={"resolution": [84, 84]}, car={"max_velocity": 0.15}); n = 41
    elif case == "rgb96":
        cfg = make_config("simple_layout", "rgb", cam={"resolution": [96, 128]}, car={"max_velocity": 0.15}); n = 21
    elif case == "env240":
        cfg = make_config("knuffingen", "classes", cam={"resolution": [240, 320]}); n = 9
    elif case == "classes480":
        cfg = make_config("knuffingen", "classes", cam={"resolution": [480, 640]}); n = 5
    elif case == "rgb480":
        cfg = make_config("simple_layout", "rgb", cam={"resolution": [480, 640]}, car={"max_velocity": 0.15}); n = 5
    elif case == "bits480":
        cfg = make_config("knuffingen", "classes", cam={"resolution": [480, 640]}); kw = {"obs_format": "classes_bits"}; n = 7
    elif case == "bits480_overflow":
        cfg = make_config("simple_layout", "classes", cam={"resolution": [480, 640], "max_range": 3.0, "orientation": [30, 0, 0]}, car={"max_velocity": 0.15})
        kw = {"obs_format": "classes_bits"}; n = 5
    else:
        cfg = make_config("knuffingen", "classes", cam={"resolution": [128, 160]}); kw = {"obs_format": "classes_bf16"}; n = 11
    old = {k: os.environ.get(k) for k in setenv}
    os.environ.update(setenv)
    try:
        env = _vec(cfg, n, **kw)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    oenv = oracle_env(cfg, n)
    H, W = cfg["camera"]["resolution"]
    # ---- move every output into a guarded buffer: [guard | payload | guard], sentinel bytes 0x5C
    G = 4096
    guards = []

    def guarded(t):
        nb = t.numel() * t.element_size()
        pad = (-nb) % 16
        buf = torch.full((G + nb + pad + G,), 0x5C, dtype=torch.uint8, device=t.device)
        view = buf[G:G + nb].view(t.dtype).view(t.shape)
        view.copy_(t)
        guards.append((buf, nb))
        return view
    env.obs = guarded(env.obs)
    for k in list(env.out):
        env.out[k] = guarded(env.out[k])
    env._outs = env._make_outputs(with_obs=True)
    env._outs_noobs = env._make_outputs(with_obs=False)

    def decode(obs):
        if kw.get("obs_format") == "classes_bf16":
            f = obs.float().cpu().numpy()
            assert set(np.unique(f)).issubset({0.0, 1.0})
            return (f * 255).astype(np.uint8)
        o = obs.cpu().numpy()
        if kw.get("obs_format") == "classes_bits":
            return np.unpackbits(o.view(np.uint32).view(np.uint8), axis=-1, bitorder="little")[..., : H * W].reshape(n, -1, H, W) * 255
        return o
    rng = np.random.default_rng(11)
    env.obs.view(torch.uint8).fill_(0xAB)
    env.reset(seed=8)
    oenv.reset(env._spawn_nodes.cpu().numpy())
    assert np.array_equal(decode(env.obs), oenv.obs), "reset frames"
    for t in range(4):
        cc = torch.from_numpy(np.stack([rng.uniform(0.3, 1, n), rng.uniform(-1, 1, n)], 1).astype(np.float32)).cuda()
        man = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).cuda()
        st0 = {k: v.clone() for k, v in env.state_dict().items()}
        first = None
        for rep in range(6):            # the same step six times from the same state
            env.load_state_dict(st0)
            env.obs.view(torch.uint8).fill_(0xAB)
            env.step({"car_control": cc, "maneuver": man})
            snap = (env.obs.clone(), env.out["info_f64"].clone(), env.out["nearest_edge"].clone(), env.out["truncated"].clone())
            if first is None:
                first = snap
            else:
                assert all(torch.equal(a, b) for a, b in zip(first, snap)), f"step {t} repetition {rep} differs: the kernels are not deterministic"
        oenv.step(cc.cpu().numpy().astype(np.float64), man.cpu().numpy())
        assert np.array_equal(decode(env.obs), oenv.obs), f"frames at step {t}"
        done = (oenv.terminated | oenv.truncated).astype(bool)
        if done.any():
            env.reset_done()
            oenv.reset(env._spawn_nodes.cpu().numpy(), mask=done)
            assert np.array_equal(decode(env.obs), oenv.obs), f"frames after the reset at step {t}"
    torch.cuda.synchronize()
    for buf, nb in guards:
        assert bool((buf[:G] == 0x5C).all()) and bool((buf[G + nb + ((-nb) % 16):] == 0x5C).all()), "a kernel wrote outside an output buffer"
    env.close()


def test_cuda_episode_stats_kernel_matches_torch():
    """EpisodeStats.update on CUDA tensors is one launch of tc_episode_stats; it must count what the torch formulation counts
    (bool and uint8 flags, sizes that are not multiples of the block, accumulation over calls)."""
    from tinycarlo_b200.distributed import EpisodeStats
    torch.manual_seed(0)
    st = EpisodeStats(torch.device("cuda:0"))
    want = torch.zeros(4, dtype=torch.float64)
    for n, as_bool in ((1, True), (255, True), (4096, False), (100003, True)):
        reward = torch.randn(n, device="cuda:0")
        term = torch.rand(n, device="cuda:0") < 0.1
        trunc = torch.rand(n, device="cuda:0") < 0.05
        if not as_bool:
            term, trunc = term.to(torch.uint8), trunc.to(torch.uint8)
        st.update(reward, term, trunc)
        want += torch.tensor([float((term.bool() | trunc.bool()).sum()), float(trunc.sum()), float(reward.double().sum()), float(n)], dtype=torch.float64)
    got = st.local.cpu()
    assert torch.equal(got[[0, 1, 3]], want[[0, 1, 3]]), (got, want)
    assert abs(float(got[2] - want[2])) < 1e-6 * max(1.0, abs(float(want[2]))) + 1e-3, (got, want)
    assert st.gather().shape == (1, 4)
