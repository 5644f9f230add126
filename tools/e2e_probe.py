import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from pair_util import make_config
from tinycarlo_b200 import TinyCarloVecEnv
N = 16384
cfg = make_config("knuffingen", "classes", cam={"resolution": [480, 640]})
env = TinyCarloVecEnv(cfg, N, device="cuda:0", autoreset="next_step")
env.reset(seed=0)
pin = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt).pin_memory()
h_cc, h_man, h_rew, h_term, h_trunc, h_cte, h_head = pin(N, 2), pin(N, dt=torch.int32), pin(N), pin(N, dt=torch.uint8), pin(N, dt=torch.uint8), pin(N), pin(N)
h_cc[:, 0] = 0.8
cc_np, cte_np, head_np = h_cc.numpy(), h_cte.numpy(), h_head.numpy()
ts = []
for i in range(40):
    t0 = time.perf_counter()
    cc_np[:, 1] = (head_np + np.arctan2(4.0 * cte_np, 0.8)) * (180.0 / np.pi / 30.0)
    t1 = time.perf_counter()
    env.step_host(h_cc, h_man, h_rew, h_term, h_trunc, h_cte, h_head)
    t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1))
ts = np.array(ts) * 1e3
print("numpy ms", ts[5:, 0].mean(), "step_host ms mean", ts[5:, 1].mean(), "min", ts[5:, 1].min(), "max", ts[5:, 1].max())
print(np.round(ts[:, 1], 2))
