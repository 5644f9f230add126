"""Soft dependency on gymnasium. When gymnasium is importable its Env / Wrapper / spaces are used (so
gym.make("tinycarlo-v2", ...) and third-party wrappers work); otherwise minimal stand-ins with the same seeding
semantics (Env.reset(seed=s) -> np.random.Generator(PCG64(SeedSequence(s))), gymnasium >= 0.26) keep the single-env
drop-in and the wrappers usable without it."""
import numpy as np

try:  # pragma: no cover - depends on the installation
    import gymnasium as _gym
    from gymnasium import spaces
    Env, Wrapper = _gym.Env, _gym.Wrapper
    HAVE_GYMNASIUM = True
except Exception:
    HAVE_GYMNASIUM = False

    def _make_rng(seed):
        return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    class Env:
        metadata = {}
        render_mode = None
        _np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = _make_rng(None)
            return self._np_random

        @np_random.setter
        def np_random(self, value):
            self._np_random = value

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

        def render(self):
            return self.env.render()

        def close(self):
            return self.env.close()

    class _Space:
        def seed(self, seed=None):
            self._rng = _make_rng(seed)

        @property
        def rng(self):
            if getattr(self, "_rng", None) is None:
                self._rng = _make_rng(None)
            return self._rng

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=dtype)
            self.high = np.full(self.shape, high, dtype=dtype)

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return self.rng.integers(self.low, self.high, endpoint=True).astype(self.dtype)
            return self.rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Discrete(_Space):
        def __init__(self, n):
            self.n = int(n)

        def sample(self):
            return int(self.rng.integers(0, self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class Dict(_Space):
        def __init__(self, spaces_dict):
            self.spaces = dict(spaces_dict)

        def __getitem__(self, key):
            return self.spaces[key]

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    class spaces:  # noqa: N801 - mirrors `gymnasium.spaces`
        Box, Discrete, Dict = Box, Discrete, Dict
