"""BASELINE.json configs[4] on N GPUs of one box: stochastic-volatility particle filter, N = 2^26 particles sharded over the
ranks, fp32, systematic resampling on integer weights when the fresh ESS drops below N/2 -- the decision is taken on the
GPUs inside the device-resident loop (every rank reaches the same one from the exchanged integer totals).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config5_multi.py [--check]
--check: rank 0 first runs the same filter unsharded and the sharded result must reproduce it (resample count, state and
log-weight sums)."""
import argparse, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import modppl_b200 as m
from modppl_b200.distributed import ShardedParticleSystem, max_over_ranks

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=26)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--check", action="store_true")
ap.add_argument("--scheme", default="systematic", choices=["systematic", "nested"])
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
scheme = m.SYSTEMATIC_FIXED if a.scheme == "systematic" else m.SYSTEMATIC_NESTED
T, n = a.steps, 1 << a.log2n
rng = np.random.default_rng(5)
x, ys = -1.024, []
for t in range(T):
    x = -1.024 + 0.9702 * (x + 1.024) + 0.178 * rng.normal()
    ys.append([math.exp(x / 2) * rng.normal()])
ys = np.array(ys)
ref = None
if a.check and rank == 0:
    one = m.ParticleSystem(m.stochastic_volatility(), n, seed=5, dtype="f32", device=local)
    one.upload_observations(ys)
    one.run(0, T, scheme, ess_threshold=0.5)
    ref = (one.num_resamples(), float(np.sum(one.traces)), float(np.sum(one.log_weights)))
    one.close()
dist.barrier()
ps = ShardedParticleSystem(m.stochastic_volatility(), n, rank, world, seed=5, dtype="f32", device=local)
ps.upload_observations(ys)
dist.barrier(); torch.cuda.synchronize()
ps.run(0, 20, scheme, ess_threshold=0.5)
ps.sync(); dist.barrier(); torch.cuda.synchronize()
r0 = ps.num_resamples()
ms = max_over_ranks(ps.run(20, T - 20, scheme, ess_threshold=0.5))
ps.sync(); dist.barrier()
nres = ps.num_resamples()
sums = torch.tensor([float(np.sum(ps.traces)), float(np.sum(ps.log_weights))], dtype=torch.float64)
dist.all_reduce(sums)
err = ps.peer_error() if world > 1 else 0
if rank == 0:
    steps = T - 20
    bytes_alg = n * (16.0 * (steps - (nres - r0)) + 24.0 * (nres - r0))
    out = {"n_gpus": world, "particles": f"2^{a.log2n}", "scheme": a.scheme, "ms_per_step": ms / steps, "particle_steps_per_s": n * steps / (ms * 1e-3),
           "resampled_steps": int(nres - r0), "of": steps, "frac_of_roofline": bytes_alg / (ms * 1e-3) / 1e9 / (6504.1 * world), "peer_wait_timeouts": err}
    if ref is not None:
        out["single_gpu"] = {"resamples": ref[0], "sum_state": ref[1], "sum_log_weights": ref[2]}
        out["sharded"] = {"resamples": int(nres), "sum_state": float(sums[0]), "sum_log_weights": float(sums[1])}
        out["matches_single_gpu"] = bool(ref[0] == nres and abs(ref[1] - float(sums[0])) <= 1e-9 * abs(ref[1]) and abs(ref[2] - float(sums[1])) <= 1e-9 * abs(ref[2]))
    print(json.dumps(out))
ps.close()
dist.destroy_process_group()
