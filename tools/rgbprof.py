import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from pair_util import make_config
from tinycarlo_b200 import TinyCarloVecEnv
n=8192
env=TinyCarloVecEnv(make_config("simple_layout","rgb",cam={"resolution":[480,640]},car={"max_velocity":0.15}),n,device="cuda:0",autoreset="next_step")
env.reset(seed=0)
cc=torch.zeros((n,2),device="cuda"); cc[:,0]=0.8; man=torch.zeros(n,dtype=torch.int32,device="cuda")
for _ in range(3): env.step({"car_control":cc,"maneuver":man})
env.profile_begin(10)
for _ in range(10): env.step({"car_control":cc,"maneuver":man})
print(env.profile_end())
print("bands", env.H, env.W)
