"""Short config-4 run for ncu: init + a few step/resample rounds at N = 2^24 (no CPU baseline, no e2e leg)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m
from bench import observations

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--log2n", type=int, default=24)
ap.add_argument("--scheme", type=int, default=m.SYSTEMATIC_FIXED)
ap.add_argument("--dtype", default="f32")
a = ap.parse_args()
ys = observations(a.steps + 2)
ps = m.ParticleSystem(m.lgssm4(), 1 << a.log2n, seed=1, dtype=a.dtype)
ps.upload_observations(ys)
ms = ps.run(0, a.steps + 1, a.scheme)
print("ms per step", ms / (a.steps + 1), "lml", ps.log_marginal_likelihood_estimate())
