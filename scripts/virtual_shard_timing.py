"""Per-shard cost of the sharded (peer-table) code path vs the single-GPU path, on one GPU (phase-by-phase emulation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m
from bench import observations
T = 41
ys = observations(T)
for log2_shard in (23, 21):
    for world in (1, 2, 4):
        n = (1 << log2_shard) * world
        m.parity.virtual_shards(m.lgssm4(), n, world, ys[:3], seed=1)
        m.parity.virtual_shards(m.lgssm4(), n, world, ys, seed=1)
        print(f"shard 2^{log2_shard} world {world}: {m.parity.virtual_shards.last_loop_ms / T / world * 1e3:.1f} us per shard-step (3 host syncs per step included)")
