"""A few steps of a small-frame config for ncu: usage small_prof.py [map] [H] [W] [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tinycarlo_b200.config import make_config
from tinycarlo_b200 import TinyCarloVecEnv
MAP = sys.argv[1] if len(sys.argv) > 1 else "simple_layout"
RES = [int(sys.argv[2]), int(sys.argv[3])] if len(sys.argv) > 3 else [84, 84]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
env = TinyCarloVecEnv(make_config(MAP, "classes", cam={"resolution": RES}), n, device="cuda:0", autoreset="next_step")
env.reset(seed=0)
cc = torch.zeros((n, 2), device="cuda"); cc[:, 0] = 0.8; man = torch.zeros(n, dtype=torch.int32, device="cuda")
for _ in range(6):
    env.step({"car_control": cc, "maneuver": man})
env.profile_begin(10)
for _ in range(10):
    env.step({"car_control": cc, "maneuver": man})
torch.cuda.synchronize()
print(env.profile_end())
