"""CPU tests of the reward / termination wrappers (tinycarlo_b200/wrapper): the recorded `info` stream of the
reference (tests/golden/knuff_wrapped_*.npz) is replayed through the same wrapper stacks and must reproduce the
reference's rewards and terminations — for the scalar (single-env) form and for the tensor (vectorised) form."""
import numpy as np
import pytest
import torch

from golden_util import Golden
from tinycarlo_b200 import wrapper as W
from tinycarlo_b200.gym_compat import Env


class _Car:
    def __init__(self, tw):
        self.track_width = tw


class ReplayEnv(Env):
    """Feeds the recorded step results of the unwrapped reference env (reward 0 / terminated False because wrapped)."""

    def __init__(self, g: Golden):
        self.g = g
        self.f = 0
        self.wrapped = False
        self.car = _Car(g.cfg["car"]["track_width"])

    def _info(self, f):
        g = self.g
        return {"cte": float(g["cte"][f]), "heading_error": float(g["heading"][f]), "velocity": float(g["velocity"][f]),
                "laneline_distances": {n: float(g["dist"][f][k]) for k, n in enumerate(g.class_names)}}

    def reset(self, seed=None, options=None):
        assert self.g["ev_kind"][self.f] == 0
        info = self._info(self.f)
        self.f += 1
        return None, info

    def step(self, action):
        assert self.g["ev_kind"][self.f] == 1
        f = self.f
        self.f += 1
        return None, 0, False, bool(self.g["truncated"][f]), self._info(f)


class ReplayVecEnv:
    """The same stream as n identical envs with CPU tensors (the wrappers only use torch ops)."""
    is_vector_env = True
    autoreset = None

    def __init__(self, g: Golden, n=3):
        self.g, self.f, self.num_envs, self.device = g, 0, n, torch.device("cpu")
        self.class_names = g.class_names
        self.track_width = g.cfg["car"]["track_width"]
        self.wrapped = False

    def set_wrapped(self, w):
        self.wrapped = w

    def _info(self, f):
        g, n = self.g, self.num_envs
        rep = lambda v: torch.tensor(np.repeat(np.asarray(v, np.float64)[None], n, 0))  # noqa: E731  (float64 keeps the comparisons exact)
        return {"cte": rep(g["cte"][f]), "heading_error": rep(g["heading"][f]), "velocity": rep(g["velocity"][f]),
                "laneline_distances": rep(g["dist"][f])}

    def reset(self, *a, **kw):
        info = self._info(self.f)
        self.f += 1
        return None, info

    def step(self, action):
        f = self.f
        self.f += 1
        n = self.num_envs
        return None, torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.bool), \
            torch.full((n,), bool(self.g["truncated"][f])), self._info(f)


def build_stack(env, wrappers):
    for name, kw in wrappers:
        env = getattr(W, name)(env, **kw)
    return env


@pytest.mark.parametrize("name", ["knuff_wrapped_cte", "knuff_wrapped_lane"])
def test_scalar_wrappers_reproduce_reference(name):
    g = Golden(name)
    base = ReplayEnv(g)
    env = build_stack(base, g.meta["wrappers"])
    assert base.wrapped is True
    env.reset()
    n_term = 0
    while base.f < g.F:
        f = base.f
        if g["ev_kind"][f] == 0:
            env.reset()
            continue
        _, reward, terminated, truncated, _ = env.step(None)
        assert reward == g["reward"][f], (f, reward, g["reward"][f])
        assert bool(terminated) == bool(g["terminated"][f]), f
        assert bool(truncated) == bool(g["truncated"][f]), f
        n_term += bool(terminated)
    assert n_term > 0


@pytest.mark.parametrize("name", ["knuff_wrapped_cte", "knuff_wrapped_lane"])
def test_vector_wrappers_reproduce_reference(name):
    g = Golden(name)
    base = ReplayVecEnv(g)
    env = build_stack(base, g.meta["wrappers"])
    assert base.wrapped is True
    env.reset()
    while base.f < g.F:
        f = base.f
        if g["ev_kind"][f] == 0:
            env.reset()
            continue
        _, reward, terminated, truncated, _ = env.step(None)
        np.testing.assert_allclose(reward.numpy(), g["reward"][f], rtol=1e-12, atol=1e-15)
        assert bool(terminated.all()) == bool(g["terminated"][f]) and bool(terminated.any()) == bool(g["terminated"][f]), f


def test_utils_known_values():
    assert W.sparse_reward({"a": True, "b": False, "c": True}, {"a": 1.5, "b": 2.0}) == 1.5
    assert W.linear_reward(0.0, 0.03, 1.0) == 1.0
    assert W.linear_reward(0.03, 0.03, 1.0) == 0.0
    assert W.linear_reward(0.06, 0.03, 1.0, -0.5) == -0.5
    assert W.linear_reward(0.06, 0.03, -1.0) == 0.0   # negative max_reward: min(y, min_reward)
    assert W.linear_reward(-0.015, 0.03, 1.0) == 0.5  # |x|
