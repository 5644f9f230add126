"""sparse_reward / linear_reward of the reference (tinycarlo/wrapper/utils.py:3-37) for Python scalars, and their
tensor forms for the vectorised env."""
from typing import Dict

import torch


def sparse_reward(conditions: Dict[str, bool], sparse_rewards: Dict[str, float]) -> float:
    reward = 0.0
    for name, condition in conditions.items():
        if condition and name in sparse_rewards:
            reward += sparse_rewards[name]
    return reward


def linear_reward(x: float, max_x: float, max_reward: float = 1.0, min_reward: float = 0.0) -> float:
    y = (-max_reward / max_x) * abs(x) + max_reward
    return max(y, min_reward) if max_reward > 0 else min(y, min_reward)


def linear_reward_tensor(x: torch.Tensor, max_x, max_reward: float = 1.0, min_reward: float = 0.0) -> torch.Tensor:
    y = (-max_reward / max_x) * x.abs() + max_reward
    return torch.clamp(y, min=min_reward) if max_reward > 0 else torch.clamp(y, max=min_reward)
