"""simple_layout 480x640 rgb (BASELINE config 1 on the device), 8192 envs, random actions: step and render-kernel times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tinycarlo_b200 import TinyCarloVecEnv
from tinycarlo_b200.config import make_config
n = 8192
env = TinyCarloVecEnv(make_config("simple_layout", "rgb", cam={"resolution": [480, 640]}, car={"max_velocity": 0.15}), n, device="cuda:0", autoreset="next_step")
env.reset(seed=0)
cc = torch.zeros((n, 2), device="cuda"); man = torch.zeros(n, dtype=torch.int32, device="cuda")
def one():
    cc.uniform_(-1, 1); man.random_(0, 4)
    env.step({"car_control": cc, "maneuver": man})
for _ in range(10): one()
torch.cuda.synchronize()
env.profile_begin(40)
for _ in range(40): one()
torch.cuda.synchronize()
k, ks = env.profile_end()
r = (k["project"] + k["raster"]) / ks
print(f"lib={os.environ.get('TC_LIB', 'main')[-14:]} rgb 480x640 N={n}: track {k['track'] / ks:.3f} render {r:.3f} ms -> {n * 921600 / r / 1e6:.0f} GB/s  checksum {int(env.obs.sum(dtype=torch.int64).item() % 1000003)}")
