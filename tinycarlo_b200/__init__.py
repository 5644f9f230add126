"""tinycarlo_b200 — B200-native batched implementation of tinycarlo's per-step hot path.

Public surface:
  TinyCarloVecEnv   N envs in lockstep on one GPU, CUDA tensors in and out (tinycarlo_b200/vec_env.py)
  TinyCarloGroupedVecEnv  several resolution groups presented as one env (per-env resolution, BASELINE config 5)
  TinyCarloEnv      single-env drop-in for the reference's gymnasium env (tinycarlo_b200/env.py), a batch of one
  tinycarlo_b200.wrapper   the reference's reward / termination wrappers, for both of the above
`gym.make("tinycarlo-v2", config=...)` is registered when gymnasium is importable (soft dependency)."""
from ._lib import TinyCarloError  # noqa: F401


def __getattr__(name):
    if name == "TinyCarloVecEnv":
        from .vec_env import TinyCarloVecEnv
        return TinyCarloVecEnv
    if name == "TinyCarloGroupedVecEnv":
        from .grouped_env import TinyCarloGroupedVecEnv
        return TinyCarloGroupedVecEnv
    if name == "TinyCarloEnv":
        from .env import TinyCarloEnv
        return TinyCarloEnv
    raise AttributeError(name)


def _register():
    try:
        from gymnasium.envs.registration import register
    except Exception:
        return
    try:
        register(id="tinycarlo-v2", entry_point="tinycarlo_b200.env:TinyCarloEnv")
    except Exception:
        pass


_register()
