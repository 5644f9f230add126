"""world_size-2 gloo test (CPU) of the multi-GPU host logic: env-index sharding, global seeding (a sharded job draws
the same spawn nodes as a single-process job) and the episode-statistics all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tinycarlo_b200.config import resolve_map_path
from tinycarlo_b200.distributed import EpisodeStats, shard_range
from tinycarlo_b200.maptables import MapTables
from tinycarlo_b200.spawn import SpawnSampler

SPAWN_KNUFF = [156, 18, 217, 214, 325, 354, 176, 402, 339, 376, 385, 419, 396, 37, 149, 62, 240, 113, 98, 299, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    t = MapTables(resolve_map_path({"map_name": "knuffingen"}, None), 222, SPAWN_KNUFF)
    tab = SpawnSampler(t, hi - lo, table_len=6, env_index_offset=lo).seed(42)
    st = EpisodeStats("cpu")
    st.update(torch.full((hi - lo,), float(rank + 1)), torch.zeros(hi - lo, dtype=torch.bool), torch.ones(hi - lo, dtype=torch.bool))
    g = st.gather()
    q.put((rank, lo, hi, tab, g.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for total, world in ((65536, 8), (10, 3), (7, 8), (32768, 4)):
        r = [shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_job_matches_single_process():
    total, world = 11, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = MapTables(resolve_map_path({"map_name": "knuffingen"}, None), 222, SPAWN_KNUFF)
    single = SpawnSampler(t, total, table_len=6).seed(42)
    sharded = np.concatenate([r[3] for r in res])
    assert np.array_equal(single, sharded)            # results do not depend on the sharding
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 6, 6, 11)
    for r in res:                                      # every rank holds the gathered statistics of all ranks
        g = r[4]
        assert g.shape == (2, 4)
        assert list(g[:, 0]) == [6, 5] and list(g[:, 1]) == [6, 5] and list(g[:, 2]) == [6.0, 10.0] and list(g[:, 3]) == [6, 5]


def _worker_groups(rank, world, port, sizes, q):
    """config 5's sharding on the host side: every rank takes a contiguous slice of EVERY resolution group; env seeds and per-env
    parameters are functions of the global env index"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
    import workloads as WL
    from tinycarlo_b200.distributed import shard_groups
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local, offs = shard_groups(sizes, rank, world)
    t = MapTables(resolve_map_path({"map_name": "knuffingen"}, None), 222, SPAWN_KNUFF)
    p = WL.config5_params(sum(sizes))
    draws, fov = [], []
    for n, o in zip(local, offs):
        draws.append(SpawnSampler(t, n, table_len=3, env_index_offset=o).seed(7))
        fov.append(p["fov"][o:o + n])
    counts = torch.tensor([float(sum(local))], dtype=torch.float64)
    dist.all_reduce(counts)
    q.put((rank, local, offs, draws, fov, float(counts.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_resolution_groups_match_single_process():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
    import workloads as WL
    sizes, world = [7, 5, 9], 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_groups, args=(r, world, port, sizes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = MapTables(resolve_map_path({"map_name": "knuffingen"}, None), 222, SPAWN_KNUFF)
    single = SpawnSampler(t, sum(sizes), table_len=3).seed(7)
    fov = WL.config5_params(sum(sizes))["fov"]
    base = np.cumsum([0] + sizes)
    for g in range(len(sizes)):     # group g of the job = the ranks' slices of it, in rank order
        got = np.concatenate([r[3][g] for r in res])
        assert np.array_equal(got, single[base[g]:base[g + 1]]), g
        assert np.array_equal(np.concatenate([r[4][g] for r in res]), fov[base[g]:base[g + 1]])
    assert res[0][5] == sum(sizes) and [r[1] for r in res] == [[4, 3, 5], [3, 2, 4]]   # the first total % world ranks of a group get one extra env
