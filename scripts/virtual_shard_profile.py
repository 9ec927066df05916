import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modppl_b200 as m
from bench import observations
world = int(sys.argv[1]); log2_shard = int(sys.argv[2])
ys = observations(5)
m.parity.virtual_shards(m.lgssm4(), (1 << log2_shard) * world, world, ys, seed=1)
print("ok", m.parity.virtual_shards.last_loop_ms)
