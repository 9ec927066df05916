"""A model of your own, written as a spec and compiled at run time (NVRTC) into the engine's kernels -- what `dyngen!` +
`DynUnfold` are to the reference (modppl-macros/src/lib.rs:20-114, modppl/src/modeling/dynunfold.rs:41-100), for the restricted
class of Unfold-style state-space models.  No rebuild of the library.

AR(1) with Gaussian observations; the filter's log marginal likelihood is compared with the exact Kalman value.

    python examples/custom_model.py [log2_particles] [T]
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m

PHI, S, R, X0 = 0.9, 0.3, 0.5, 1.0

SPEC = {
    "name": "ar1", "state_dim": 1, "obs_dim": 1,
    "params": {"phi": PHI, "q": S, "r": R, "x0": X0},              # (x, y, t, z, u, s are reserved names)
    "init": [{"dist": "normal", "args": ["0", "x0"]}],                              # x_0 ~ N(0, x0)
    "step": [{"dist": "normal", "args": ["phi * x[0]", "q"]}],                      # x_t ~ N(phi x_{t-1}, q)
    "observe": [{"dist": "normal", "value": "y[0]", "args": ["x[0]", "r"]}],        # y_t ~ N(x_t, r), constrained
}


def kalman_log_ml(ys):
    mean, var, ll = 0.0, X0 * X0, 0.0
    for t, y in enumerate(ys):
        if t > 0:
            mean, var = PHI * mean, PHI * PHI * var + S * S
        sy = var + R * R
        ll += -0.5 * (math.log(2 * math.pi * sy) + (y - mean) ** 2 / sy)
        gain = var / sy
        mean, var = mean + gain * (y - mean), (1 - gain) * var
    return ll


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    rng = np.random.default_rng(0)
    x, ys = rng.normal() * X0, []
    for t in range(T):
        if t > 0:
            x = PHI * x + S * rng.normal()
        ys.append(x + R * rng.normal())
    model = m.compile_model(SPEC)
    # the reference's loop (tests/smc.rs:63-85), one call per step ...
    f = m.ParticleSystem(model, 1 << log2n, seed=1, dtype="f32")
    f.init_step([ys[0]]); f.resample(m.SYSTEMATIC_NESTED)
    for y in ys[1:]:
        f = f.step([y]); f.resample(m.SYSTEMATIC_NESTED)
    lml_calls = f.log_marginal_likelihood_estimate()
    # ... and the same filter as one device-resident run
    g = m.ParticleSystem(model, 1 << log2n, seed=1, dtype="f32")
    g.upload_observations(np.asarray(ys).reshape(T, 1))
    ms = g.run(0, T, m.SYSTEMATIC_NESTED)
    lml_run = g.log_marginal_likelihood_estimate()
    truth = kalman_log_ml(ys)
    print(f"2^{log2n} particles, {T} steps: log-ML {lml_calls:.4f} (call per step), {lml_run:.4f} (device-resident run, {ms / T * 1e3:.1f} us per step); Kalman {truth:.4f}")
    assert lml_calls == lml_run, "the two ways of driving the filter take the same arithmetic path"
    assert abs(lml_run - truth) < 0.05 + 20.0 / math.sqrt(1 << log2n) * math.sqrt(T)
    f.close(); g.close()


if __name__ == "__main__":
    main()
