"""Spawn-node draws with the reference's RNG semantics (tinycarlo/map.py:51-69 under gymnasium seeding).

gymnasium's Env.reset(seed=s) installs np.random.Generator(PCG64(SeedSequence(s))); each reset then makes one bounded
draw (choice(spawn_points), or integers(0, n_nodes-1) without spawn points) and redraws while the node has no
successor. The sequence of spawn nodes of an env therefore depends only on (seed, number of resets), so the host
pre-draws K resets ahead per env into a table that the device consumes with a per-env cursor.
Env i of a vectorised env is seeded with seed + global_index(i), like gymnasium.vector."""
from typing import Optional

import numpy as np

from .maptables import MapTables


def make_generator(seed: Optional[int]) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class SpawnSampler:
    def __init__(self, tables: MapTables, num_envs: int, table_len: int = 16, env_index_offset: int = 0):
        self.tables, self.n, self.K, self.offset = tables, int(num_envs), int(table_len), int(env_index_offset)
        self.rngs = None
        self.table = None  # int32 [n, K]: the next K spawn nodes of every env

    def seed(self, seed: Optional[int]):
        base = None if seed is None else int(seed) + self.offset
        self.rngs = [make_generator(None if base is None else base + i) for i in range(self.n)]
        tab = np.empty((self.n, self.K), np.int32)
        draw = self.tables.sample_spawn_node
        for i, rng in enumerate(self.rngs):
            for k in range(self.K):
                tab[i, k] = draw(rng)
        self.table = tab
        return tab

    def advance(self, consumed: np.ndarray):
        """Env i used its first consumed[i] entries: shift them out and draw as many new ones at the end."""
        tab, K = self.table, self.K
        draw = self.tables.sample_spawn_node
        for i in np.nonzero(consumed)[0]:
            c = int(min(consumed[i], K))
            tab[i, :K - c] = tab[i, c:]
            rng = self.rngs[i]
            for k in range(K - c, K):
                tab[i, k] = draw(rng)
        return tab
