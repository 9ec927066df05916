"""Short config-2 / config-3 run for ncu: one importance-sampling batch per model (2^20 proposals) and a few MH sweeps over 2^20
chains; also times the pieces of importance_resampling (the exact running sum on very peaked weights)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m
from bench import regression_data

xs, ys_line, ys_hier = regression_data()
n = 1 << 20
for name, model, ys in (("line", m.line_model(xs), ys_line), ("hierarchical", m.hierarchical_model(xs), ys_hier)):
    m.importance_sampling(model, ys, n, seed=0, batch=0, return_traces=False)
    t0 = time.perf_counter(); lat, lnw, lml = m.importance_sampling(model, ys, n, seed=0, batch=1); t1 = time.perf_counter()
    m.importance_resampling(model, ys, n, 1 << 10, seed=0, batch=1); t2 = time.perf_counter()
    p = np.exp(lnw)
    t3 = time.perf_counter(); m.parity.cumsum_sequential(p); t4 = time.perf_counter()
    print(f"{name}: importance_sampling (all traces to host) {1e3 * (t1 - t0):.2f} ms, importance_resampling {1e3 * (t2 - t1):.2f} ms, "
          f"exact running sum alone (incl. 16 MB of copies) {1e3 * (t4 - t3):.2f} ms; weights: {np.count_nonzero(p)} non-zero of {n}, lml {lml:.4f}")
ch = m.Chains(m.hierarchical_model(xs), ys_hier, n, seed=2)
acc, ms = m.hierarchical_full_sweeps(ch, 8, timed=True)
print(f"mh: 8 sweeps of 18 moves over 2^20 chains: {ms:.2f} ms = {n * 8 * 18 / (ms * 1e-3):.3g} chain-steps/s, acceptance {acc / (n * 8 * 18):.3f}")
