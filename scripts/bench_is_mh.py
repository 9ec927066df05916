"""Configs 2 and 3 of BASELINE.json on one B200: batched importance sampling (2^20 proposals per batch) and many-chain MH
(2^20 chains).  These paths are register-resident (FP64 / SFU bound), so the figure of merit is proposals/s and
chain-steps/s, not an HBM fraction."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m

xs = np.arange(-5.0, 6.0)
rng = np.random.default_rng(2)
ys_line = 0.5 * xs - 1.0 + 0.1 * rng.normal(size=11)
ys_hier = 0.3 + 0.4 * xs + 0.5 * xs * xs + 0.1 * rng.normal(size=11)
out = {}
n = 1 << 20
for name, model, ys in (("line", m.line_model(xs), ys_line), ("hierarchical", m.hierarchical_model(xs), ys_hier)):
    m.importance_sampling(model, ys, n, seed=0, batch=0)
    t0 = time.perf_counter()
    lmls = [m.importance_sampling(model, ys, n, seed=0, batch=b)[2] for b in range(16)]
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    lmls2 = [m.importance_sampling(model, ys, n, seed=0, batch=b, return_traces=False)[2] for b in range(64)]
    dt2 = time.perf_counter() - t0
    assert lmls2[:16] == lmls
    out[f"is_{name}"] = {"proposals_per_s_e2e_incl_d2h_of_all_traces": 16 * n / dt, "proposals_per_s_lml_only": 64 * n / dt2,
                         "lml_mean": float(np.mean(lmls2)), "lml_std": float(np.std(lmls2))}
ch = m.Chains(m.hierarchical_model(xs), ys_hier, n, seed=2)
m.hierarchical_sweeps(ch, 2)
acc, ms = m.hierarchical_sweeps(ch, 100, timed=True)      # 1400 moves per chain
st = ch.read()
out["mh_hierarchical"] = {"chains": n, "moves_per_chain": 1400, "chain_steps_per_s": n * 1400 / (ms * 1e-3), "ms": ms, "acceptance": acc / (n * 1400),
                          "posterior_mean": {"is_linear": float(st[0].mean()), "a": float(st[1].mean()), "b": float(st[2].mean()), "c": float(st[3][st[0] == 0].mean())}}
bounds, cov = [-5.0, 5.0, -5.0, 5.0], [1.0, -0.6, -0.6, 2.0]
pc = m.Chains(m.pointed_model(bounds, cov), [0.0, 0.0], n, seed=3)
t0 = time.perf_counter(); acc = m.mh(pc, m.mh.__globals__["POINTED_DRIFT"], 0.5, 1000); dt = time.perf_counter() - t0
out["mh_pointed"] = {"chains": n, "chain_steps_per_s_wall": n * 1000 / dt, "acceptance": acc / (n * 1000)}
print(json.dumps(out, indent=1))
