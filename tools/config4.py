"""BASELINE.json configs[3] as a device loop: Knuffingen 128x160 classes (the shipped camera), 8192 envs per GPU sharded by env
index (65536 on 8 GPUs), maneuvers uniform over {0,1,2,3} resampled every 100 steps (u-turns included), Stanley controller +
Ornstein-Uhlenbeck noise (theta 0.1, sigma 0.4, examples/train_td3.py:42-44,143), CTESparseRewardWrapper(min_cte=0.01)
(examples/stanley_control.py:41) and CTETerminationWrapper, in-kernel next-step autoreset, per-rank episode statistics
all-gathered over NCCL every 100 steps. One line per job on rank 0.

  python tools/config4.py                                   # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from pair_util import make_config
from tinycarlo_b200 import TinyCarloVecEnv
from tinycarlo_b200.distributed import EpisodeStats
from tinycarlo_b200.wrapper import CTESparseRewardWrapper, CTETerminationWrapper

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("TC_ENVS_PER_GPU", 8192))
H, W = (int(v) for v in os.environ.get("TC_RES", "128x160").split("x"))
cfg = make_config("knuffingen", "classes", cam={"resolution": [H, W]})
base = TinyCarloVecEnv(cfg, N, device=dev, env_index_offset=rank * N, autoreset="next_step")
env = CTETerminationWrapper(CTESparseRewardWrapper(base, min_cte=0.01), max_cte=0.1, number_of_steps=5)
stats = EpisodeStats(dev)
gen = torch.Generator(device=dev).manual_seed(1000 + rank)
man = torch.zeros(N, dtype=torch.int32, device=dev)
cc = torch.zeros((N, 2), dtype=torch.float32, device=dev)
ou = torch.zeros(N, device=dev)
max_steer = float(cfg["car"]["max_steering_angle"])
env.reset(seed=0)
steps, warm = int(os.environ.get("TC_STEPS", 300)), 20
gathered = None


def one(t):
    global gathered
    if t % 100 == 0:
        man.copy_(torch.randint(0, 4, (N,), device=dev, generator=gen, dtype=torch.int32))
    o = base.out
    ou.add_(-0.1 * ou + 0.4 * torch.randn(N, device=dev, generator=gen))               # OU: x += theta * (0 - x) + sigma * N(0,1)
    cc[:, 0] = 0.8
    cc[:, 1] = ((o["heading_error"] + torch.atan2(4.0 * o["cte"], torch.full_like(o["cte"], 0.8))) * (180.0 / np.pi / max_steer) + ou).clamp_(-1, 1)
    _, reward, terminated, truncated, _ = env.step({"car_control": cc, "maneuver": man})
    base.mark_done(terminated)                                                           # the wrapper's terminations feed the in-kernel autoreset
    stats.update(reward, terminated, truncated)
    if t % 100 == 99:
        gathered = stats.gather()                                                        # the only collective


for t in range(warm):
    one(t)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(steps):
    one(t)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    g = gathered.sum(0).tolist()
    print(json.dumps({"config": f"configs[3]: knuffingen {H}x{W} classes, {N} envs/GPU x {world} GPUs, mixed maneuvers, Stanley + OU noise, CTESparseRewardWrapper, "
                      "CTETerminationWrapper, autoreset, stats all-gather every 100 steps", "n_gpus": world, "envs_total": N * world, "ms_per_step": float(ms.item()),
                      "env_steps_per_s": N * world / (float(ms.item()) * 1e-3), "episodes": g[0], "truncated": g[1], "reward_sum": g[2], "env_steps_counted": g[3]}))
if world > 1:
    dist.destroy_process_group()
