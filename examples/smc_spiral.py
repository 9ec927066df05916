"""The reference's SMC demo (modppl/tests/smc.rs:48-92) on the B200 engine: spiral model, N particles, T steps,
resample after every step, and the same JSON dumps `visualization/visualizer.py` reads (`../data/smc_*.json`), plus
full trajectories rebuilt from the ancestor log.

    python examples/smc_spiral.py [out_dir] [N] [T]
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m


def simulate_loop(rng, bounds, timesteps):
    """tests/smc.rs:17-46: a deformed circle of observations."""
    xmin, xmax, ymin, ymax = bounds
    init_angle = rng.random() * 2 * math.pi
    center = np.array([(xmax - xmin) / 2 + xmin, (ymax - ymin) / 2 + ymin])
    radius = max(xmax - xmin, ymax - ymin) / 5.0
    perturb = [t for t in range(timesteps) if rng.random() < 0.3]
    obs = []
    for t in range(timesteps):
        deformation = sum(math.exp(-((t - p) ** 2 + math.log(2 * math.pi)) / 2) for p in perturb)   # normal.logpdf(t; p, 1).exp()
        r = radius + deformation
        a = 2 * math.pi * t / timesteps
        obs.append(center + r * np.array([math.cos(a + init_angle), math.sin(a + init_angle)]))
    return np.array(obs)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "data"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500          # tests/smc.rs:54
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 20           # tests/smc.rs:53
    os.makedirs(out, exist_ok=True)
    obs = simulate_loop(np.random.default_rng(1), (-1.0, 1.0, -1.0, 1.0), T)
    json.dump(obs.tolist(), open(os.path.join(out, "smc_obs.json"), "w"))
    f = m.ParticleSystem(m.spiral_model(), n, seed=1, dtype="f64")
    f.enable_history(T)
    for t in range(T):
        if t == 0:
            f.init_step(obs[0])
        else:
            f = f.step(obs[t])
        json.dump(f.traces.T.tolist(), open(os.path.join(out, f"smc_traces_before_resample_{t}.json"), "w"))
        f.resample(m.MULTINOMIAL)
        json.dump(f.traces.T.tolist(), open(os.path.join(out, f"smc_traces_{t}.json"), "w"))
    traj = f.trajectories(np.arange(min(n, 50)))                  # [ids, T, (r, theta)]
    json.dump(traj.tolist(), open(os.path.join(out, "smc_trajectories.json"), "w"))
    print("log-ML estimate", f.log_marginal_likelihood_estimate(), "distinct ancestors at t=0 among 50 lineages:", len({tuple(x) for x in traj[:, 0].round(12)}))


if __name__ == "__main__":
    main()
