"""Config 2 (simple_layout 84x84, random actions) eager vs CUDA graph: the step is launch-bound at small batch sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from pair_util import make_config
from tinycarlo_b200 import TinyCarloVecEnv

cfg = make_config("simple_layout", "classes", cam={"resolution": [84, 84]}, car={"max_velocity": 0.15})
for n in (256, 1024, 4096, 16384):
    env = TinyCarloVecEnv(cfg, n, device="cuda:0", autoreset="next_step")
    env.reset(seed=0)
    gen = torch.Generator(device="cuda").manual_seed(0)
    cc = torch.zeros((n, 2), device="cuda"); man = torch.zeros(n, dtype=torch.int32, device="cuda")
    pool_cc = 2 * torch.rand((64, n, 2), device="cuda", generator=gen) - 1      # pre-drawn random actions (graph replays cannot advance a torch generator cheaply)
    pool_man = torch.randint(0, 4, (64, n), device="cuda", generator=gen, dtype=torch.int32)
    idx = torch.zeros((), dtype=torch.long, device="cuda")

    def one():
        cc.copy_(pool_cc.index_select(0, idx.view(1))[0]); man.copy_(pool_man.index_select(0, idx.view(1))[0])
        idx.add_(1).remainder_(64)
        env.step({"car_control": cc, "maneuver": man})

    def timed(fn, steps=300):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            one()
    torch.cuda.current_stream().wait_stream(s)
    eager = timed(one)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        one()
    graph = timed(g.replay)
    print(f"config2 N={n:6d}: eager {eager * 1e3:7.1f} us/step {n / eager / 1e3:7.2f} M env-steps/s | graph {graph * 1e3:7.1f} us/step {n / graph / 1e3:7.2f} M env-steps/s")
    env.close()
