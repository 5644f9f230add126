"""Termination wrappers with the reference's names, arguments and logic (tinycarlo/wrapper/termination.py:4-70), for the
single-env drop-in (Python scalars) and for TinyCarloVecEnv (per-env counters as device tensors). With the vectorised
env's autoreset="next_step", terminations raised here are ORed into the env's done flags so the next step resets them."""
from typing import List, Union

import torch

from ..gym_compat import Wrapper
from .reward import _VecWrapper, _make


def _feed_autoreset(vec_wrapper, terminated):
    u = vec_wrapper.unwrapped
    if getattr(u, "autoreset", None):
        u.mark_done(terminated)


# ------------------------------------------------------------------------------------------------ laneline crossing
class _LanelineCrossingScalar(Wrapper):
    def __init__(self, env, lanelines: Union[List[str], str]):
        super().__init__(env)
        self.unwrapped.wrapped = True
        self.lanelines = lanelines if isinstance(lanelines, list) else [lanelines]

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        for name in self.lanelines:
            if info["laneline_distances"][name] <= self.unwrapped.car.track_width / 2:
                terminated = True
        return observation, reward, terminated, truncated, info


class _LanelineCrossingVec(_VecWrapper):
    def __init__(self, env, lanelines: Union[List[str], str]):
        super().__init__(env)
        self.lanelines = lanelines if isinstance(lanelines, list) else [lanelines]
        names = self.unwrapped.class_names
        self._cols = [names.index(n) for n in self.lanelines]   # ValueError for an unknown laneline (reference: KeyError)

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        u = self.unwrapped
        hit = (info["laneline_distances"][:, self._cols] <= u.track_width / 2).any(dim=1)
        if getattr(u, "autoreset", None):
            hit = hit & ~u.reset_mask   # the empty info of a reset step (all distances 0) is not a crossing
        terminated = terminated | hit
        _feed_autoreset(self, terminated)
        return observation, reward, terminated, truncated, info


class LanelineCrossingTerminationWrapper:
    """terminated when any of the given lanelines is closer than track_width/2 (termination.py:4-22)."""
    __new__ = _make(_LanelineCrossingScalar, _LanelineCrossingVec)


# ------------------------------------------------------------------------------------------------ CTE
class _CTETerminationScalar(Wrapper):
    def __init__(self, env, max_cte: float, number_of_steps: int = 1):
        super().__init__(env)
        self.unwrapped.wrapped = True
        self.max_cte = max_cte
        self.number_of_steps = number_of_steps
        self.steps_true = 0

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        if abs(info["cte"]) > self.max_cte:
            self.steps_true += 1
            if self.steps_true >= self.number_of_steps:
                terminated = True
                self.steps_true = 0
        else:
            self.steps_true = 0
        return observation, reward, terminated, truncated, info


class _CounterVec(_VecWrapper):
    """terminated after `number_of_steps` consecutive steps with the condition true; the counter restarts after firing
    and whenever the condition is false — exactly the scalar logic above, per env."""

    def __init__(self, env, number_of_steps: int):
        super().__init__(env)
        self.number_of_steps = number_of_steps
        u = self.unwrapped
        self.steps_true = torch.zeros(u.num_envs, dtype=torch.int32, device=u.device)

    def _update(self, cond: torch.Tensor, terminated: torch.Tensor) -> torch.Tensor:
        cnt = torch.where(cond, self.steps_true + 1, torch.zeros_like(self.steps_true))
        u = self.unwrapped
        if getattr(u, "autoreset", None):
            # a step that reset the env is not a step of the wrapped env (a caller of the reference calls reset() there,
            # which leaves steps_true alone)
            cnt = torch.where(u.reset_mask, self.steps_true, cnt)
            fire = (cnt >= self.number_of_steps) & ~u.reset_mask
        else:
            fire = cnt >= self.number_of_steps
        self.steps_true = torch.where(fire, torch.zeros_like(cnt), cnt)
        terminated = terminated | fire
        _feed_autoreset(self, terminated)
        return terminated


class _CTETerminationVec(_CounterVec):
    def __init__(self, env, max_cte: float, number_of_steps: int = 1):
        super().__init__(env, number_of_steps)
        self.max_cte = max_cte

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        return observation, reward, self._update(info["cte"].abs() > self.max_cte, terminated), truncated, info


class CTETerminationWrapper:
    """terminated after number_of_steps consecutive steps with |cte| > max_cte (termination.py:24-48)."""
    __new__ = _make(_CTETerminationScalar, _CTETerminationVec)


# ------------------------------------------------------------------------------------------------ crash
class _CrashScalar(Wrapper):
    def __init__(self, env, velcoity_threshold: float = 0.005, number_of_steps: int = 10):
        super().__init__(env)
        self.unwrapped.wrapped = True
        self.velcoity_threshold = velcoity_threshold   # (sic) the reference's argument name
        self.number_of_steps = number_of_steps
        self.steps_true = 0

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        if abs(info["velocity"]) < self.velcoity_threshold:
            self.steps_true += 1
            if self.steps_true >= self.number_of_steps:
                terminated = True
                self.steps_true = 0
        else:
            self.steps_true = 0
        return observation, reward, terminated, truncated, info


class _CrashVec(_CounterVec):
    def __init__(self, env, velcoity_threshold: float = 0.005, number_of_steps: int = 10):
        super().__init__(env, number_of_steps)
        self.velcoity_threshold = velcoity_threshold

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        return observation, reward, self._update(info["velocity"].abs() < self.velcoity_threshold, terminated), truncated, info


class CrashTerminationWrapper:
    """terminated after number_of_steps consecutive steps with |velocity| below the threshold (termination.py:50-70)."""
    __new__ = _make(_CrashScalar, _CrashVec)


for _pub, _impls in ((LanelineCrossingTerminationWrapper, (_LanelineCrossingScalar, _LanelineCrossingVec)),
                     (CTETerminationWrapper, (_CTETerminationScalar, _CTETerminationVec)), (CrashTerminationWrapper, (_CrashScalar, _CrashVec))):
    for _i in _impls:
        _i.__name__ = _pub.__name__
