"""BASELINE.json configs[4] shape on ONE B200: stochastic-volatility particle filter, N = 2^26, fp32, systematic
resampling when the fresh ESS drops below N/2 (decision on the GPU, device-resident loop).  Also config 1 (spiral, N = 1000,
T = 100, fp64, reference multinomial scheme) through the call-per-step API."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m

out = {}
T = 200
rng = np.random.default_rng(5)
x, ys = -1.024, []
for t in range(T):
    x = -1.024 + 0.9702 * (x + 1.024) + 0.178 * rng.normal()
    ys.append([math.exp(x / 2) * rng.normal()])
ys = np.array(ys)
for log2n, scheme, name in ((26, m.SYSTEMATIC_NESTED, "nested"), (26, m.SYSTEMATIC_FIXED, "single-level"), (24, m.SYSTEMATIC_NESTED, "nested")):
    n = 1 << log2n
    ps = m.ParticleSystem(m.stochastic_volatility(), n, seed=5, dtype="f32")
    ps.upload_observations(ys)
    ps.run(0, 20, scheme, ess_threshold=0.5)
    r0 = ps.num_resamples()
    ms = ps.run(20, T - 20, scheme, ess_threshold=0.5)
    nres = ps.num_resamples() - r0
    steps = T - 20
    # 16 B/particle on steps without a resample, 24 B with (SURVEY 8d, D = 1 fp32)
    bytes_alg = n * (16.0 * (steps - nres) + 24.0 * nres)
    out[f"sv_2^{log2n}_{name}"] = {"ms_per_step": ms / steps, "particle_steps_per_s": n * steps / (ms * 1e-3), "resampled_steps": int(nres), "of": steps,
                            "algorithmic_GBps": bytes_alg / (ms * 1e-3) / 1e9, "frac_of_6504": bytes_alg / (ms * 1e-3) / 1e9 / 6504.1, "lml": ps.log_marginal_likelihood_estimate()}
    ps.close()
# config 1
th = 2 * math.pi * np.arange(100) / 100 + 0.7
obs = np.stack([0.4 * np.cos(th), 0.4 * np.sin(th)], 1)
f = m.ParticleSystem(m.spiral_model(), 1000, seed=1, dtype="f64")
f.init_step(obs[0]); f.resample(m.MULTINOMIAL)
f.sync(); t0 = time.perf_counter()
for t in range(1, 100):
    f.step(obs[t]); f.resample(m.MULTINOMIAL)
f.sync(); dt = time.perf_counter() - t0
out["config1_spiral_N1000_T100_f64_multinomial"] = {"particle_steps_per_s": 1000 * 99 / dt, "ms_per_step": dt / 99 * 1e3, "lml": f.log_marginal_likelihood_estimate()}
print(json.dumps(out, indent=1))
