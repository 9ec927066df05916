"""Small end-to-end pass over every kernel family for compute-sanitizer (sizes chosen so that ragged tails, heavy tiles,
multi-chunk expansions and the parallel exact cumsum are all exercised)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m

rng = np.random.default_rng(0)
ys = rng.normal(size=(6, 2))
for dtype in ("f32", "f64"):
    for n in (1, 777, 4097, 70001):
        for scheme in (m.MULTINOMIAL, m.SYSTEMATIC, m.SYSTEMATIC_FIXED, m.MULTINOMIAL_FIXED):
            ps = m.ParticleSystem(m.lgssm4(), n, seed=3, dtype=dtype)
            ps.init_step(ys[0]); ps.resample(scheme)
            for t in range(1, 4):
                ps.step(ys[t]); ps.effective_sample_size(False); ps.resample(scheme)
            ps.step(ys[4]); _ = ps.traces, ps.log_weights, ps.parents, ps.log_marginal_likelihood_estimate()
            ps.close()
# degenerate weights: heavy warp tiles, overflow pass
th = 0.3 * np.arange(5) + 0.5
sp = np.stack([0.4 * np.cos(th), 0.4 * np.sin(th)], 1)
ps = m.ParticleSystem(m.spiral_model(), 200000, seed=1, dtype="f32")
ps.init_step(sp[0]); ps.resample(m.SYSTEMATIC_FIXED)
for t in range(1, 5):
    ps.step(sp[t]); ps.resample(m.SYSTEMATIC_FIXED)
ps.close()
lw = np.full(300000, -80.0, dtype=np.float32); lw[123456] = 0.0
anc, _, _ = m.parity.fixed_resample(lw, scheme=2, seed=3, t=9)
assert np.all(anc == 123456)
anc, _, _ = m.parity.fixed_resample(lw, scheme=2, seed=3, t=9)     # second system: overflow pass launched from the start? (new handle: no)
# ESS-triggered device loop + hmm + sv
sv = m.ParticleSystem(m.stochastic_volatility(), 50000, seed=2, dtype="f32")
sv.upload_observations(rng.normal(size=(12, 1)) * 0.5); sv.run(0, 12, m.SYSTEMATIC_FIXED, ess_threshold=0.5); sv.traces; sv.close()
hm = m.ParticleSystem(m.hmm([0.2, 0.3, 0.5], [[0.1, 0.2, 0.7], [0.2, 0.7, 0.1], [0.7, 0.2, 0.1]], [[0.4, 0.4, 0.2], [0.2, 0.3, 0.5], [0.9, 0.05, 0.05]]), 10000, seed=1)
hm.init_step([0]); hm.step([1]); hm.resample(m.MULTINOMIAL); hm.close()
# virtual shards (multi-GPU kernels on one device)
m.parity.virtual_shards(m.lgssm4(), 1 << 15, 4, ys[:5], dtype="f32", seed=17)
# exact cumsum (parallel path) and searches
p = rng.random(20011); p /= p.sum()
m.parity.resample_indices(p, rng.random(5000)); m.parity.resample_indices(p, [0.3], n_draws=20011, scheme=1); m.parity.cumsum_sequential(p)
m.parity.logsumexp_stats(rng.normal(size=100003)); m.parity.logsumexp_stats(rng.normal(size=1001).astype(np.float32))
# importance sampling and MH
xs = np.arange(-5.0, 6.0); yl = 0.5 * xs - 1
m.importance_sampling(m.line_model(xs), yl, 10001); m.importance_resampling(m.hierarchical_model(xs), yl, 10001, 100)
m.importance_sampling(m.pointed_model([-5, 5, -5, 5], [1, -.6, -.6, 2]), [0, 0], 5000)
ch = m.Chains(m.hierarchical_model(xs), yl, 3001, seed=1)
m.mh(ch, 0, 0.1, 3); m.mh(ch, 1, 0.025, 3); m.regen_mh(ch, 7, 3); m.hierarchical_sweeps(ch, 2); ch.read(); ch.close()
pc = m.Chains(m.pointed_model([-5, 5, -5, 5], [1, -.6, -.6, 2]), [0, 0], 2000, seed=1); m.mh(pc, 3, 0.5, 5); pc.close()
print("sanitize pass complete")
