"""Time the reference's multinomial scheme (sequential-rounding cumsum, bit-exact) and systematic-exact at N = 2^24."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modppl_b200 as m
from bench import observations
ys = observations(8)
out = {}
for scheme, name in ((m.MULTINOMIAL, "multinomial_exact"), (m.SYSTEMATIC, "systematic_exact"), (m.MULTINOMIAL_FIXED, "multinomial_fixed"), (m.SYSTEMATIC_FIXED, "systematic_fixed")):
    ps = m.ParticleSystem(m.lgssm4(), 1 << 24, seed=1, dtype="f32")
    ps.init_step(ys[0]); ps.resample(scheme)
    ps.step(ys[1]); ps.resample(scheme)
    ps.profile_enable(True)
    for t in range(2, 6):
        ps.step(ys[t]); ps.resample(scheme, sync=False)
    ps.sync()
    names = ("extend", "weight_reduce", "normalize", "cumsum_exact", "search", "fixed_reduce", "fixed_scan", "fixed_cumsum", "fixed_search")
    out[name] = {k: round(ps.profile_get(k)[0] / max(1, ps.profile_get(k)[1]), 4) for k in names if ps.profile_get(k)[1]}
    out[name]["lml"] = ps.log_marginal_likelihood_estimate()
    ps.close()
print(json.dumps(out, indent=1))
