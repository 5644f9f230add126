"""GPU diagnostic: compares the CUDA env with the oracle on a small batch and prints where they differ."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pair_util import make_config, oracle_env, stanley_actions
from tinycarlo_b200 import TinyCarloVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
res = [int(sys.argv[2]), int(sys.argv[3])] if len(sys.argv) > 3 else [128, 160]
cfg = make_config("knuffingen", "classes", cam={"resolution": res})
env = TinyCarloVecEnv(cfg, n, device="cuda:0", debug_segments=True)
oenv = oracle_env(cfg, n)
env.reset(seed=0)
oenv.reset(env._spawn_nodes.cpu().numpy())
off = env.map.ll_edge_off

def report(tag):
    obs = env.obs.cpu().numpy()
    bad = np.nonzero((obs != oenv.obs).reshape(n, -1).any(axis=1))[0]
    st = env.state_dict()
    sfd = np.abs(st["sf"].cpu().numpy()[:, :7] - oenv.sf[:, :7]).max()
    cnt = env.out["seg_count"].cpu().numpy(); seg = env.out["seg_i32"].cpu().numpy()
    nseg_bad = 0; nseg = 0; maxd = 0
    for i in range(n):
        oc, o32, o64, _ = oenv.segments(i)
        for c in range(env.n_classes):
            if cnt[i, c] != oc[c]:
                print(tag, "env", i, "class", c, "count differs", cnt[i, c], oc[c]); continue
            a = seg[i, off[c]:off[c] + cnt[i, c]].astype(np.int64); b = o32[off[c]:off[c] + oc[c]].astype(np.int64)
            d = np.abs(a - b)
            nseg += len(a); nseg_bad += int((d.max(axis=1) > 0).sum()) if len(a) else 0
            if len(a) and d.max() > 0:
                maxd = max(maxd, d.max())
                k = np.argmax(d.max(axis=1))
                if nseg_bad < 6: print(tag, "env", i, "class", c, "seg", k, "cuda", a[k], "oracle", b[k], "f64", o64[off[c] + k])
    print(tag, f"bad frames {len(bad)}/{n}  max|state diff| {sfd:.3e}  differing segments {nseg_bad}/{nseg} max int diff {maxd}")
    for i in bad[:4]:
        d = obs[i] != oenv.obs[i]
        for c in range(env.n_classes):
            if d[c].any():
                ys, xs = np.nonzero(d[c])
                print(tag, "  env", i, "class", c, "px differ", d[c].sum(), "cuda set", int((obs[i, c] > 0).sum()), "oracle set", int((oenv.obs[i, c] > 0).sum()),
                      "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())

report("reset")
for t in range(3):
    cc = stanley_actions(oenv.cte.copy(), oenv.heading_error.copy(), cfg["car"]["max_steering_angle"])
    man = np.zeros(n, np.int32)
    env.step({"car_control": torch.from_numpy(cc).cuda(), "maneuver": torch.from_numpy(man).cuda()})
    oenv.step(cc.astype(np.float64), man)
    report(f"step{t}")
