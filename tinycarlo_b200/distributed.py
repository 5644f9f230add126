"""Multi-GPU plumbing. Environments are independent, so a job of `total_envs` is sharded by env index with NO per-step
collective: rank r owns the contiguous range shard_range(total, r, world), seeds its envs with their GLOBAL index
(TinyCarloVecEnv(env_index_offset=...)), and the only communication is an all-gather of a few episode statistics at log
cadence (NCCL over NVLink on the GPUs, gloo in the CPU tests)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the envs owned by `rank`; the first total % world ranks get one extra env."""
    base, extra = divmod(int(total_envs), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_groups(group_sizes, rank: int, world: int):
    """Resolution groups (TinyCarloGroupedVecEnv) of a multi-GPU job: every rank takes a contiguous slice of EVERY group.
    group_sizes: global envs per group (group g owns the global indices [sum(sizes[:g]), sum(sizes[:g+1]))).
    -> (local sizes, global index of each local group's env 0): pass them as the group sizes and group_index_offsets."""
    sizes, offsets, base = [], [], 0
    for n in group_sizes:
        lo, hi = shard_range(int(n), rank, world)
        sizes.append(hi - lo)
        offsets.append(base + lo)
        base += int(n)
    return sizes, offsets


class EpisodeStats:
    """Per-rank running episode statistics on the device: episodes finished, truncations, reward sum, env-steps."""
    FIELDS = ("episodes", "truncated", "reward_sum", "env_steps")

    def __init__(self, device):
        self.local = torch.zeros(len(self.FIELDS), dtype=torch.float64, device=device)

    def update(self, reward: torch.Tensor, terminated: torch.Tensor, truncated: torch.Tensor):
        if reward.is_cuda and reward.dtype == torch.float32 and reward.is_contiguous() and terminated.is_contiguous() and truncated.is_contiguous() \
                and terminated.element_size() == 1 and truncated.element_size() == 1:
            # one launch (tc_episode_stats) instead of eight tiny torch kernels per step
            import ctypes as C
            from . import _lib
            with torch.cuda.device(reward.device):
                _lib.check(_lib.lib().tc_episode_stats(C.c_void_p(reward.data_ptr()), C.c_void_p(terminated.data_ptr()), C.c_void_p(truncated.data_ptr()),
                                                       reward.numel(), C.c_void_p(self.local.data_ptr()),
                                                       C.c_void_p(torch.cuda.current_stream(reward.device).cuda_stream)), "tc_episode_stats")
            return
        self.local[0] += (terminated | truncated).sum()
        self.local[1] += truncated.sum()
        self.local[2] += reward.sum()
        self.local[3] += reward.numel()

    def gather(self) -> torch.Tensor:
        """[world, 4] on every rank (all_gather_into_tensor: NCCL over NVLink / NVSwitch on GPUs)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.local[None].clone()
        out = torch.zeros(dist.get_world_size() * len(self.FIELDS), dtype=torch.float64, device=self.local.device)
        dist.all_gather_into_tensor(out, self.local)
        return out.view(dist.get_world_size(), len(self.FIELDS))
