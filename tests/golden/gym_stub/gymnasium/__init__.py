"""Test-only gymnasium stand-in (gymnasium is not installed in this image).

Just enough surface to import and run the *unmodified* reference from /root/reference when
generating golden vectors (tests/golden/gen_golden.py). Seeding follows gymnasium >= 0.26:
``Env.reset(seed=s)`` installs ``np.random.Generator(np.random.PCG64(np.random.SeedSequence(s)))``.
Not product code; the product's own soft-dependency layer is tinycarlo_b200/gym_compat.py.
"""
import importlib
import numpy as np
from . import spaces  # noqa: F401
from .envs import registration
from .envs.registration import register  # noqa: F401


def _make_rng(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class Env:
    metadata = {}
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = _make_rng(None)
        return self._np_random

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = _make_rng(seed)

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)


def make(id, **kwargs):
    mod, cls = registration.registry[id].split(":")
    return getattr(importlib.import_module(mod), cls)(**kwargs)
