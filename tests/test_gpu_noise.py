"""GPU test of the blob-noise kernel (tc_noise_blobs / NoiseObservationWrapper): the reference's operation
(tinycarlo/wrapper/observation.py:14-27, with cv2.circle) re-run on the CPU with the same Philox draws must give the same masks."""
import numpy as np
import pytest
import torch

from pair_util import make_config

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def philox4x32_10(c, k):
    c = [int(x) & 0xFFFFFFFF for x in c]
    k = [int(x) & 0xFFFFFFFF for x in k]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def reference_noise(obs, env_global, seed, step, n_blobs, max_radius):
    """observation.py:14-27 with the draws replaced by the documented Philox contract"""
    Cn, H, W = obs.shape
    for c in range(Cn):
        for k in range(n_blobs):
            r = philox4x32_10([env_global, step, c * n_blobs + k, 0], [seed & 0xFFFFFFFF, seed >> 32])
            x, y = r[0] % W, r[1] % H
            radius = 1 + r[2] % (max_radius - 1) if max_radius > 1 else 1
            if (r[3] & 0xFFFF) < 19661:
                mask = np.zeros(obs[c].shape, dtype=np.uint8)
                cv2.circle(mask, (x, y), radius, 255, -1)
                mask = cv2.bitwise_and(obs[(r[3] >> 16) % Cn], mask, mask=mask)
                obs[c] = cv2.bitwise_or(obs[c], mask)
            else:
                cv2.circle(obs[c], (x, y), radius, 0, -1)
    return obs


@pytest.mark.parametrize("res,max_radius,n_blobs", [([128, 160], 100, 10), ([84, 84], 30, 6), ([96, 128], 2, 3)])
def test_noise_wrapper_matches_the_reference_operation(res, max_radius, n_blobs):
    from tinycarlo_b200 import TinyCarloVecEnv
    from tinycarlo_b200.wrapper import NoiseObservationWrapper
    n, seed, off = 48, 0x1234567890, 1000
    cfg = make_config("knuffingen", "classes", cam={"resolution": res})
    base = TinyCarloVecEnv(cfg, n, device="cuda:0", env_index_offset=off)
    clean = TinyCarloVecEnv(cfg, n, device="cuda:0", env_index_offset=off)
    env = NoiseObservationWrapper(base, blob_max_radius=max_radius, n_blobs=n_blobs, seed=seed)
    assert base.wrapped
    env.reset(seed=1)
    clean.reset(seed=1)
    rng = np.random.default_rng(0)
    changed = 0
    for t in range(4):
        act = {"car_control": torch.from_numpy(np.stack([rng.uniform(0.3, 1, n), rng.uniform(-1, 1, n)], 1).astype(np.float32)).cuda(),
               "maneuver": torch.zeros(n, dtype=torch.int32, device="cuda")}
        obs, *_ = env.step(act)
        cobs, *_ = clean.step(act)
        got, c = obs.cpu().numpy(), cobs.cpu().numpy()
        for i in range(n):
            want = reference_noise(c[i].copy(), off + i, seed, t, n_blobs, max_radius)
            assert np.array_equal(got[i], want), (t, i)
        changed += int((got != c).sum())
    assert changed > 0
    base.close()
    clean.close()


def test_noise_wrapper_on_the_single_env_drop_in():
    from tinycarlo_b200 import TinyCarloEnv
    from tinycarlo_b200.wrapper import NoiseObservationWrapper
    cfg = make_config("knuffingen", "classes")
    env = NoiseObservationWrapper(TinyCarloEnv(config=cfg), blob_max_radius=60, n_blobs=8, seed=5)
    obs, info = env.reset(seed=0)
    o1, *_ = env.step({"car_control": [0.8, 0.0], "maneuver": 0})
    clean = env.unwrapped._vec
    assert o1.shape == (5, 128, 160) and set(np.unique(o1)).issubset({0, 255})
    want = reference_noise(TinyCarloEnvClean(cfg), 0, 5, 0, 8, 60)
    assert np.array_equal(o1, want)
    env.close()


def TinyCarloEnvClean(cfg):
    from tinycarlo_b200 import TinyCarloEnv
    e = TinyCarloEnv(config=cfg)
    e.reset(seed=0)
    o, *_ = e.step({"car_control": [0.8, 0.0], "maneuver": 0})
    e.close()
    return o.copy()
