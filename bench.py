#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on its named workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5]

--config 4 (default; BASELINE.json configs[3], the configuration the metric is quoted on): 4-D linear-Gaussian
  state-space model, N = 2^24 particles (sharded over the GPUs: strong scaling), fp32 state and log-weights, bootstrap
  proposal, resample after every step.  A "step" is one `step(obs_t); resample()` over all particles; metric =
  particle-steps/s.  The line carries the Kalman filter's exact log-ML next to the estimate, the ms/step of every
  resampling scheme incl. the reference's own multinomial routine, the roofline of the dominant kernel and of the
  whole step, and the CPU baseline (the reference's single-threaded O(N^2) algorithm, with the multi-threaded O(N)
  port beside it).
--config 1: spiral model (tests/dyngenfns/unfold.rs), N = 1000, T = 100, fp64, the reference's multinomial scheme,
  call-per-step API (launch-bound by construction).
--config 2: importance_sampling / importance_resampling, 2^20 proposals per batch, 64 batches; proposals/s.
--config 3: hierarchical model, 2^20 chains x 10 000 moves (the sweep of tests/mh.rs:93-106 plus regen_mh); chain-steps/s.
--config 5: stochastic volatility, N = 2^26, ESS-triggered systematic resampling, sharded over the GPUs.

The timed product path never touches oracle/: the oracle is only executed for `cpu_baseline` and `--impl reference`.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_PARTICLES = 24
BYTES_PER_PARTICLE_STEP = 48      # 2*D*s + 2*w + 8 with D = 4, s = w = 4
METRIC4 = "particle-steps/sec (SMC step incl. resample)"


# --------------------------------------------------------------------------------------------------------------- data
def observations(T, seed=4):
    """config 4: simulated from the model itself (q = 0.1, r = 0.5, x0 = 1)"""
    rng = np.random.default_rng(seed)
    A = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], float)
    x = rng.normal(size=4)
    ys = np.empty((T, 2))
    for t in range(T):
        if t > 0:
            x = A @ x + 0.1 * rng.normal(size=4)
        ys[t] = x[:2] + 0.5 * rng.normal(size=2)
    return ys


def kalman_log_ml(ys, q=0.1, r=0.5, x0=1.0):
    """Exact log marginal likelihood of config 4's model (host fp64; plain NumPy, nothing from oracle/):
    x_0 ~ N(0, x0^2 I), x_t = A x_{t-1} + N(0, q^2 I), y_t = (x_t)_{0,1} + N(0, r^2 I)."""
    A = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], float)
    H = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], float)
    m, P = np.zeros(4), np.eye(4) * x0 * x0
    lml = 0.0
    for t, y in enumerate(np.asarray(ys, float).reshape(-1, 2)):
        if t > 0:
            m, P = A @ m, A @ P @ A.T + q * q * np.eye(4)
        S = H @ P @ H.T + r * r * np.eye(2)
        v = y - H @ m
        Si = np.linalg.inv(S)
        lml += -0.5 * (2 * math.log(2 * math.pi) + math.log(np.linalg.det(S)) + v @ Si @ v)
        K = P @ H.T @ Si
        m, P = m + K @ v, P - K @ H @ P
    return float(lml)


def sv_observations(T, seed=5):
    rng = np.random.default_rng(seed)
    x, ys = -1.024, []
    for _ in range(T):
        x = -1.024 + 0.9702 * (x + 1.024) + 0.178 * rng.normal()
        ys.append([math.exp(x / 2) * rng.normal()])
    return np.array(ys)


def spiral_observations(T):
    th = 2 * math.pi * np.arange(T) / T + 0.7
    return np.stack([0.4 * np.cos(th), 0.4 * np.sin(th)], 1)


def regression_data():
    xs = np.arange(-5.0, 6.0)
    rng = np.random.default_rng(2)
    ys_line = 0.5 * xs - 1.0 + 0.1 * rng.normal(size=11)                      # tests/importance.rs:61-69
    ys_hier = 0.3 + 0.4 * xs + 0.5 * xs * xs + 0.1 * rng.normal(size=11)      # tests/importance.rs:98-106
    return xs, ys_line, ys_hier


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_constants():
    """Per-launch figures that only a profiler can give (DRAM bytes, instructions), copied from the committed ncu summaries by
    scripts/ncu_summary.py.  They describe the same kernels at the same sizes; they are provenance, not a measurement of
    this run, and every use names the file."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_constants.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def host_threads():
    return max(1, len(os.sched_getaffinity(0))) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


ORACLE_SCHEME = {"systematic": 2, "multinomial": 3, "nested": 4, "reference": 0}


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    return O


# ------------------------------------------------------------------------------------------------------ CPU baselines
def cpu_baseline_config4(max_steps=4, budget_s=12.0, fair_log2n=20, fair_steps=6, scheme=4):
    """`value`: the reference's algorithm as written -- inference/particle_filter.rs:73-116 with the per-draw clone + sum +
    linear scan of modeling/dists/categorical.rs:22-32, i.e. O(N^2) per resample -- on ONE thread (the reference is
    single-threaded: ThreadRng is !Send), at the largest N whose step takes about a second, with the fitted O(N^2)
    extrapolation to the benchmarked N beside it.  `fair`: the same filter with an O(N) cumsum + search resampler and its
    per-particle loops spread over every host thread -- what a competent CPU implementation of the same estimator does."""
    O = _oracle()
    ys = observations(max(max_steps, fair_steps) + 2)
    O.L.mo_set_threads(1)
    t_start = time.perf_counter()
    log2n, rate, per_step, samples = 11, None, None, []
    while True:   # grow N until a step takes ~1 s
        n = 1 << log2n
        ps = O.OraclePS("lgssm4", [0.1, 0.5, 1.0], n, dtype="f64", seed=1)
        ps.init_step(ys[0]); ps.resample_faithful_cost()
        t0 = time.perf_counter()
        ps.step(ys[1]); ps.resample_faithful_cost()
        dt = time.perf_counter() - t0
        samples.append((log2n, dt))
        if dt >= 0.7 or log2n >= 17 or time.perf_counter() - t_start > budget_s:
            steps = max(1, min(max_steps, int(3.0 / max(dt, 1e-3))))
            t0 = time.perf_counter()
            for k in range(steps):
                ps.step(ys[2 + k % (len(ys) - 2)]); ps.resample_faithful_cost()
            dts = time.perf_counter() - t0
            rate, per_step = n * steps / dts, dts / steps
            break
        log2n += 1
    n_f = 1 << log2n
    # fitted exponent of the step time in N over the sizes tried (expected -> 2)
    if len(samples) >= 3:
        xs_, ys_ = np.log2([1 << s for s, _ in samples[-3:]]), np.log2([d for _, d in samples[-3:]])
        expo = float(np.polyfit(xs_, ys_, 1)[0])
    else:
        expo = 2.0
    n_big = 1 << LOG2_PARTICLES
    extrap = rate * (n_f / n_big)          # rate = N / t(N), t ~ c N^2  =>  rate ~ 1 / N
    # fair port
    threads_all = host_threads()
    n = 1 << fair_log2n
    rates = {}
    for th in sorted({1, threads_all}):
        O.L.mo_set_threads(th)
        ps = O.OraclePS("lgssm4", [0.1, 0.5, 1.0], n, dtype="f32", seed=1)
        ps.init_step(ys[0]); ps.resample(scheme)
        t0 = time.perf_counter()
        for t in range(1, fair_steps + 1):
            ps.step(ys[t]); ps.resample(scheme)
        rates[th] = n * fair_steps / (time.perf_counter() - t0)
    O.L.mo_set_threads(1)
    best = max(rates, key=rates.get)
    scheme_name = {4: "nested systematic", 2: "systematic", 3: "multinomial"}.get(scheme, str(scheme)) + " resampling on integer weights (O(N))"
    return {"value": rate, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"oracle restatement of ParticleSystem::step + resample with the reference's O(N^2) multinomial (categorical.rs:22-32: clone, sum, "
                      f"linear scan per draw), lgssm4 fp64, N = 2^{log2n} (the largest N whose step takes ~1 s: {per_step:.2f} s/step), 1 thread",
            "particles": n_f, "seconds_per_step": per_step, "fitted_time_exponent_in_N": expo,
            "extrapolated_to_2p24": {"value": extrap, "unit": "particle-steps/s", "how": "rate ~ 1/N for an O(N^2) step"},
            "fair": {"value": rates[best], "unit": "particle-steps/s", "cores": best, "kind": "port",
                     "sample": f"same filter, {scheme_name}, flat SoA state, lgssm4 fp32, N = 2^{fair_log2n}, {fair_steps} steps; "
                               + "; ".join(f"{th} thread{'s' if th > 1 else ''}: {r:.3g}/s" for th, r in sorted(rates.items()))},
            "note": "neither figure includes the reference's trie / regex / Arc overhead per particle: real modppl is slower than `value`",
            "seconds": time.perf_counter() - t_start}


def cpu_baseline_config1():
    """config 1 is CPU-sized: the oracle runs the reference's algorithm literally (O(N^2) multinomial, one thread) at the full size"""
    O = _oracle()
    T, n = 100, 1000
    ys = spiral_observations(T)
    O.L.mo_set_threads(1)
    r = O.OraclePS("spiral", [0.1, 0.4, 0.2, 0.001], n, dtype="f64", seed=1)
    r.init_step(ys[0]); r.resample_faithful_cost()
    t0 = time.perf_counter()
    for t in range(1, T):
        r.step(ys[t]); r.resample_faithful_cost()
    dt = time.perf_counter() - t0
    return {"value": n * (T - 1) / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": "oracle restatement, spiral model fp64, N = 1000, T = 100, the reference's O(N^2) multinomial resampling, 1 thread (the whole config)",
            "seconds": dt}


def cpu_baseline_config2(n=1 << 18):
    O = _oracle()
    xs, ys_line, ys_hier = regression_data()
    out = {}
    t_all = time.perf_counter()
    for name, ys in (("line", ys_line), ("hierarchical", ys_hier)):
        rates = {}
        for th in sorted({1, host_threads()}):
            O.L.mo_set_threads(th)
            t0 = time.perf_counter()
            O.importance_sampling(name, xs, ys, n, seed=0, batch=0)
            rates[th] = n / (time.perf_counter() - t0)
        out[name] = rates
    O.L.mo_set_threads(1)
    best = max(out["line"], key=out["line"].get)
    return {"value": out["line"][best], "unit": "proposals/s", "cores": best, "kind": "port",
            "sample": f"oracle importance_sampling (importance.rs:12-28), line model, {n} proposals of one batch; per thread count: "
                      + json.dumps({k: {str(t): float(f"{r:.4g}") for t, r in v.items()} for k, v in out.items()}),
            "seconds": time.perf_counter() - t_all}


def cpu_baseline_config3(n=1 << 14, sweeps=20):
    O = _oracle()
    xs, _, ys_hier = regression_data()
    rates = {}
    t_all = time.perf_counter()
    for th in sorted({1, host_threads()}):
        O.L.mo_set_threads(th)
        ch = O.OracleChains("hierarchical", xs, ys_hier, n, seed=2)
        t0 = time.perf_counter()
        for _ in range(sweeps):
            ch.move(1, 0.025, 0, 1); ch.move(0, 0.1, 0, 3); ch.move(0, 0.01, 0, 10)
            for mask in (8, 1, 2, 4):
                ch.move(2, 1.0, mask, 1)
        rates[th] = n * sweeps * 18 / (time.perf_counter() - t0)
    O.L.mo_set_threads(1)
    best = max(rates, key=rates.get)
    return {"value": rates[best], "unit": "chain-steps/s", "cores": best, "kind": "port",
            "sample": f"oracle mh / regen_mh (mh.rs:9-67), hierarchical model, {n} chains x {sweeps} sweeps of 18 moves; "
                      + "; ".join(f"{th} thread{'s' if th > 1 else ''}: {r:.3g}/s" for th, r in sorted(rates.items())),
            "seconds": time.perf_counter() - t_all}


def cpu_baseline_config5(log2n=20, steps=12):
    O = _oracle()
    ys = sv_observations(steps + 1)
    n = 1 << log2n
    rates = {}
    t_all = time.perf_counter()
    for th in sorted({1, host_threads()}):
        O.L.mo_set_threads(th)
        ps = O.OraclePS("sv", [-1.024, 0.9702, 0.178], n, dtype="f32", seed=5)
        ps.init_step(ys[0])
        t0 = time.perf_counter()
        for t in range(1, steps + 1):
            ps.step(ys[t])
            if ps.effective_sample_size(False) < 0.5 * n:
                ps.resample(2)
        rates[th] = n * steps / (time.perf_counter() - t0)
    O.L.mo_set_threads(1)
    best = max(rates, key=rates.get)
    return {"value": rates[best], "unit": "particle-steps/s", "cores": best, "kind": "port",
            "sample": f"oracle PF, stochastic volatility fp32, N = 2^{log2n}, {steps} steps, systematic resampling on integer weights when ESS < N/2 (O(N) port; the "
                      "reference's own O(N^2) multinomial is infeasible at this size); " + "; ".join(f"{th} thread{'s' if th > 1 else ''}: {r:.3g}/s" for th, r in sorted(rates.items())),
            "seconds": time.perf_counter() - t_all}


CPU_BASELINES = {1: cpu_baseline_config1, 2: cpu_baseline_config2, 3: cpu_baseline_config3, 4: cpu_baseline_config4, 5: cpu_baseline_config5}
METRICS = {1: (METRIC4, "particle-steps/s"), 2: ("proposals/sec (importance_sampling, 2^20 proposals per batch)", "proposals/s"),
           3: ("chain-steps/sec (mh + regen_mh, 2^20 chains)", "chain-steps/s"), 4: (METRIC4, "particle-steps/s"), 5: (METRIC4, "particle-steps/s")}
WORKLOADS = {1: "config 1: spiral model particle filter, N = 1000, T = 100, fp64, reference multinomial resampling every step",
             2: "config 2: Bayesian regression by importance_sampling / importance_resampling, 2^20 proposals per batch, 64 batches",
             3: "config 3: hierarchical model, 2^20 chains x 10 000 moves: sweeps of tests/mh.rs:93-106 (1 add/remove + 3 drift(.1) + 10 drift(.01)) + 4 regen_mh",
             4: "config 4: lgssm4 (4-D linear-Gaussian SSM) bootstrap particle filter, resample every step",
             5: "config 5: stochastic-volatility particle filter, N = 2^26, systematic resampling when ESS < N/2"}


def run_reference(args):
    """The reference's own CPU implementation of the path, restated (the reference is Rust; no toolchain in this image), timed
    on the box's host cores on a bounded sample of the same workload.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    base = CPU_BASELINES[args.config]()
    metric, unit = METRICS[args.config]
    line = {"impl": "reference", "metric": metric, "value": base["value"], "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if args.config in (1, 2, 3, 4) else "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config] + " -- CPU restatement of modppl's algorithm (oracle/), bounded sample: " + base["sample"]},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------ config 4, one GPU
PROFILE_KERNELS = ("extend", "fixed_reduce", "fixed_scan", "fixed_overflow", "fixed_cumsum", "fixed_search", "nested_quantise", "nested_sections", "nested_level1",
                   "nested_plan", "nested_expand", "nested_heavy", "weight_reduce", "normalize", "cumsum_exact", "search")


def time_scheme(ps, m, ys, t_first, scheme, steps, warm=3):
    """ms per step of `steps` device-resident steps with another resampling scheme, continuing the same filter"""
    ps.run(t_first, warm, scheme)
    ms = ps.run(t_first + warm, steps, scheme)
    return ms / steps, t_first + warm + steps


def run_config4_single(args):
    import modppl_b200 as m
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_global = 1 << args.log2_particles
    K, W = args.steps, args.warmup
    k_alt = max(3, min(K, 10))           # steps timed per alternative scheme
    T = 1 + W + K + K + K + 2 * (3 + k_alt) + 4
    ys = observations(T)
    schemes = {"systematic": m.SYSTEMATIC_FIXED, "multinomial": m.MULTINOMIAL_FIXED, "nested": m.SYSTEMATIC_NESTED, "reference": m.MULTINOMIAL}
    scheme = schemes[args.scheme]

    ps = m.ParticleSystem(m.lgssm4(), n_global, seed=1, dtype="f32", device=local_rank)
    ps.upload_observations(ys)
    ps.run(0, 1 + W, scheme)                                   # init + resample + warm-up steps (untimed)
    sampler = ClockSampler(local_rank).start()
    time.sleep(0.3)
    l0 = ps.launch_count()
    ms = ps.run(1 + W, K, scheme)                              # timed: K steps, CUDA events on the launching stream
    launches = ps.launch_count() - l0
    value = n_global * K / (ms * 1e-3)
    tr = ps.device_trace()                                     # %globaltimer stamps of the last timed step (scripts/step_timeline.py)
    ext_in_loop_ms = (tr[14] - tr[13]) * 1e-6 if tr[14] > tr[13] > 0 else None
    lml = ps.log_marginal_likelihood_estimate()                # estimate after 1 + W + K observations ...
    lml_truth = kalman_log_ml(ys[:1 + W + K])                  # ... and the exact value for the same observations

    # e2e: the reference-facing calls, one host round trip per step (observation in, log total weight out)
    t_first = 1 + W + K
    ps.sync()
    t0 = time.perf_counter()
    for k in range(K):
        ps.step_resample(ys[t_first + k], scheme)              # observation: host -> device; returns log_total_weight: device -> host
    ps.sync()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # roofline of the dominant kernel: per-kernel CUDA-event times over a fresh pass
    t_prof = t_first + K
    ps.profile_enable(True)
    for k in range(K):
        ps.step_resample(ys[t_prof + k], scheme, sync=False)   # (the same kernels as the timed loop, each between its own pair of events)
    prof = {k: ps.profile_get(k) for k in PROFILE_KERNELS}
    ps.profile_enable(False)
    t_next = t_prof + K
    # every scheme at this size, incl. the reference's own (scheme 0: normalised f64 weights, sequential-f64 cumsum reproduced
    # bit for bit, one Philox uniform + binary search per draw) -- the parity path, not the throughput path
    scheme_ms = {args.scheme: ms / K}
    for name in ("nested", "systematic", "reference"):
        if name in scheme_ms:
            continue
        scheme_ms[name], t_next = time_scheme(ps, m, ys, t_next, schemes[name], k_alt)
    peak, peak_src = measured_peak()
    consts = ncu_constants()
    ext = consts.get("extend_2p24_" + args.scheme, {})
    ext_bytes = int(ext.get("algorithmic_bytes_per_particle", 40))
    ext_ms = prof["extend"][0] / max(1, prof["extend"][1])
    achieved = ext_bytes * n_global / (ext_ms * 1e-3) / 1e9
    step_gbs = BYTES_PER_PARTICLE_STEP * value / 1e9
    base = cpu_baseline_config4(scheme=ORACLE_SCHEME.get(args.scheme, 4) or 4)
    in_loop = None
    if ext_in_loop_ms:
        g = ext_bytes * n_global / (ext_in_loop_ms * 1e-3) / 1e9
        in_loop = {"extend_ms": ext_in_loop_ms, "achieved": g, "frac": g / peak,
                   "source": "device %globaltimer stamps inside the timed loop (extend block 0 past its dependency wait -> last block done); the CUDA-event "
                             "figure above times the kernel between its own pair of events, outside the loop's programmatic dependent launch"}
    line = {
        "metric": METRIC4, "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[4], "particles": f"2^{args.log2_particles}", "T_timed": K,
                   "resampling": {"nested": "nested systematic on integer weights (engine-defined; absent upstream)", "systematic": "single-level systematic on integer weights (engine-defined)",
                                  "multinomial": "multinomial on integer weights (engine-defined)", "reference": "the reference's multinomial (categorical.rs:22-32), bit-exact"}[args.scheme],
                   "l2": "inputs exceed L2 (2 x 256 MiB state buffers stream every step)",
                   "log_ml": lml, "log_ml_truth": lml_truth, "log_ml_abs_err": abs(lml - lml_truth), "log_ml_steps": 1 + W + K,
                   "log_ml_truth_source": "Kalman filter, host fp64 NumPy (bench.py: kalman_log_ml), same observations"},
        "schemes": {"unit": "ms per 2^%d-particle step (extend + resample), device-resident loop" % args.log2_particles,
                    "nested": scheme_ms.get("nested"), "systematic": scheme_ms.get("systematic"), "reference_multinomial": scheme_ms.get("reference"),
                    "note": "nested / systematic are engine-defined integer-weight schemes (bit-exact against their oracle restatement, not against upstream); "
                            "reference_multinomial is the reference's own routine, bit-exact on injected weights and uniforms"},
        "e2e": {"value": n_global * K / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 24,
                "note": "mpl_ps_step_resample(host obs -> host log total weight) per step: the observation travels as a kernel argument, the log total weight is posted "
                        "into mapped pinned host memory (tagged words) and polled there; particles stay in HBM by design"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": ext.get("kernel", "pf_extend_kernel<Lgssm4<float>, GATHER>"), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ext.get("dram_bytes_per_launch") if args.log2_particles == 24 else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_particle": ext_bytes, "algorithmic_bytes_per_launch": ext_bytes * n_global,
                     "traffic_source": ext.get("source", "none committed for this scheme") + " -- provenance (ncu --set full of the same kernel at the same size), not a measurement of this run",
                     "whole_step": {"bytes_per_particle": BYTES_PER_PARTICLE_STEP, "achieved": step_gbs, "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0},
                     "kernel_ms": {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items() if v[1]},
                     "in_loop": in_loop},
        "cpu_baseline": base,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------ config 4 / 5, several GPUs
def init_dist(local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    # control plane only (handle exchange, barriers, max-over-ranks of timings): gloo keeps NCCL off the critical path; the data
    # path is NVLink peer memory inside the kernels
    dist.init_process_group("gloo")
    return torch, dist


def run_config4_multi(args):
    """strong scaling of the config-4 workload (2^24 particles in total), one process per GPU"""
    import modppl_b200 as m
    from modppl_b200.distributed import ShardedParticleSystem, max_over_ranks
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch, dist = init_dist(local_rank)
    n_global = 1 << args.log2_particles
    K, W = args.steps, args.warmup
    ys = observations(1 + W + 3 * K + 8)
    scheme = {"systematic": m.SYSTEMATIC_FIXED, "nested": m.SYSTEMATIC_NESTED}.get(args.scheme)
    if scheme is None:
        raise SystemExit("sharded runs: --scheme nested | systematic")
    if args.variant == "island":
        return run_config4_islands(args, m, torch, dist, rank, world, local_rank, ys, scheme)
    ps = ShardedParticleSystem(m.lgssm4(), n_global, rank, world, seed=1, dtype="f32", device=local_rank)
    ps.upload_observations(ys)
    dist.barrier(); torch.cuda.synchronize()
    ps.run(0, 1 + W, scheme)
    ps.sync()
    dist.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    if sampler:
        time.sleep(0.5)      # nvidia-smi needs a moment to enumerate 8 GPUs
    l0 = ps.launch_count()
    nv0 = ps.nvlink_bytes()
    dist.barrier(); torch.cuda.synchronize()
    ps.device_barrier()      # the host barrier releases the ranks hundreds of us apart: the GPUs rendezvous themselves in front of the timed steps
    ms = ps.run(1 + W, K, scheme)
    ps.sync()
    dist.barrier(); torch.cuda.synchronize()
    ms = max_over_ranks(ms)
    launches = ps.launch_count() - l0
    nv = [None] * world
    dist.all_gather_object(nv, (ps.nvlink_bytes() - nv0) / K)
    phases = ps.phase_times()
    all_phases = [None] * world
    dist.all_gather_object(all_phases, phases)
    lml = ps.log_marginal_likelihood_estimate()     # same point of the run as the single-GPU arm: identical by construction
    lml_truth = kalman_log_ml(ys[:1 + W + K])
    # e2e: one host round trip per step on every rank
    t_first = 1 + W + K
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        ps.step_resample(ys[t_first + k], scheme)
    ps.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    # per-kernel CUDA-event times (includes time spent waiting for peers inside the kernels)
    ps.profile_enable(True)
    t_prof = t_first + K
    for k in range(min(K, len(ys) - t_prof)):
        ps.step_resample(ys[t_prof + k], scheme, sync=False)
    prof = {k: ps.profile_get(k) for k in PROFILE_KERNELS}
    ps.profile_enable(False)
    kernel_ms = {k: (v[0] / v[1]) for k, v in prof.items() if v[1]}
    err = ps.peer_error()
    dist.barrier()
    if rank == 0:
        value = n_global * K / (ms * 1e-3)
        peak, peak_src = measured_peak()
        step_gbs = BYTES_PER_PARTICLE_STEP * value / 1e9
        line = {
            "metric": METRIC4, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[4], "particles": f"2^{args.log2_particles} in total, sharded", "T_timed": K,
                       "resampling": f"global {args.scheme} resampling on integer weights (same ancestors as on one GPU); NVLink peer loads/stores inside the kernels, no NCCL on the data path",
                       "l2": "per-GPU state buffers stream every step", "log_ml": lml, "log_ml_truth": lml_truth, "log_ml_abs_err": abs(lml - lml_truth), "log_ml_steps": 1 + W + K,
                       "peer_wait_timeouts": err, "variant": "global",
                       "nvlink_payload_bytes_per_step": {"total": float(sum(nv)), "per_rank": [float(v) for v in nv],
                                                         "what": "parent states gathered across shard edges, integer weights / chunk records of chunks that own another shard's slots, section records"}},
            "e2e": {"value": n_global * K / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 24},
            "gpu_launches": int(launches) * world, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": step_gbs, "peak": peak * world, "unit": "GB/s", "frac": step_gbs / (peak * world), "traffic": None,
                         "peak_source": peak_src + f" x {world} GPUs", "kernel_ms_rank0": kernel_ms},
            "phase_ns_per_rank": all_phases,
        }
        print(json.dumps(line))
    ps.close()
    dist.destroy_process_group()


def run_config4_islands(args, m, torch, dist, rank, world, local_rank, ys, scheme):
    """the local-resample variant next to the global scheme (north_star): one island of N / G particles per GPU, local
    resampling every step, island weights compared every `--exchange-every` steps (island-level resampling when their ESS has
    dropped).  A different estimator: its log-ML is reported against the same Kalman value."""
    from modppl_b200.distributed import IslandParticleSystem, max_over_ranks
    n_global = 1 << args.log2_particles
    K, W, S = args.steps, args.warmup, args.exchange_every
    isl = IslandParticleSystem(m.lgssm4(), n_global, rank, world, seed=1, dtype="f32", device=local_rank)
    isl.upload_observations(ys)
    dist.barrier(); torch.cuda.synchronize()
    isl.run(0, 1 + W, scheme, exchange_every=S)
    isl.sync()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    if sampler:
        time.sleep(0.5)
    l0 = isl.launch_count()
    x0 = isl.bytes_exchanged
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms_dev = isl.run(1 + W, K, scheme, exchange_every=S)          # device time of the segments; the comparisons in between are host work
    isl.sync()
    dist.barrier(); torch.cuda.synchronize()
    wall = max_over_ranks(time.perf_counter() - t0)               # everything: segments, island comparisons, barriers
    ms_dev = max_over_ranks(ms_dev)
    launches = isl.launch_count() - l0
    lml = isl.log_marginal_likelihood_estimate()
    lml_truth = kalman_log_ml(ys[:1 + W + K])
    t_first = 1 + W + K
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        isl.step_resample(ys[t_first + k], scheme, exchange_every=S)
    isl.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    ex = [None] * world
    dist.all_gather_object(ex, (isl.bytes_exchanged - x0, isl.n_island_resamplings))
    if rank == 0:
        value = n_global * K / wall
        peak, peak_src = measured_peak()
        step_gbs = BYTES_PER_PARTICLE_STEP * value / 1e9
        line = {
            "metric": METRIC4, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": wall * 1e3 / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[4], "particles": f"2^{args.log2_particles} in total, {world} islands of 2^{args.log2_particles}/{world}", "T_timed": K,
                       "variant": "island", "resampling": f"LOCAL {args.scheme} resampling on each island every step; island weights compared every {S} steps, islands resampled when "
                                                          "their ESS < G/2 -- a different estimator from the global scheme (not the same ancestors)",
                       "l2": "per-GPU state buffers stream every step", "log_ml": lml, "log_ml_truth": lml_truth, "log_ml_abs_err": abs(lml - lml_truth), "log_ml_steps": 1 + W + K,
                       "island_resamplings": int(ex[0][1]), "device_ms_per_step": ms_dev / K,
                       "nvlink_payload_bytes_per_step": {"total": float(sum(e[0] for e in ex)) / K, "what": "whole islands copied at island-level resamplings (none when the island weights stay balanced)"},
                       "host_bytes_per_comparison": 16 * world},
            "e2e": {"value": n_global * K / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 24},
            "gpu_launches": int(launches) * world, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": step_gbs, "peak": peak * world, "unit": "GB/s", "frac": step_gbs / (peak * world), "traffic": None,
                         "peak_source": peak_src + f" x {world} GPUs"},
        }
        print(json.dumps(line))
    isl.close()
    dist.destroy_process_group()


def run_config5(args):
    """stochastic volatility, N = 2^26 (sharded when --gpus > 1), fp32, systematic resampling when the fresh ESS < N/2; the decision is
    taken on the GPU(s) inside the device-resident loop"""
    import modppl_b200 as m
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    log2n = args.log2_particles if args.log2_particles != LOG2_PARTICLES else 26
    n = 1 << log2n
    K, W = args.steps, max(args.warmup, 3)
    ys = sv_observations(1 + W + K)
    scheme = m.SYSTEMATIC_NESTED if args.scheme == "nested" else m.SYSTEMATIC_FIXED
    if world > 1:
        from modppl_b200.distributed import ShardedParticleSystem, max_over_ranks
        torch, dist = init_dist(local_rank)
        ps = ShardedParticleSystem(m.stochastic_volatility(), n, rank, world, seed=5, dtype="f32", device=local_rank)
    else:
        ps = m.ParticleSystem(m.stochastic_volatility(), n, seed=5, dtype="f32", device=local_rank)
    ps.upload_observations(ys)
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    ps.run(0, 1 + W, scheme, ess_threshold=0.5)
    ps.sync()
    r0 = ps.num_resamples()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
        ps.device_barrier()
    l0 = ps.launch_count()
    ms = ps.run(1 + W, K, scheme, ess_threshold=0.5)
    ps.sync()
    if world > 1:
        dist.barrier()
        ms = max_over_ranks(ms)
    launches = ps.launch_count() - l0
    nres = ps.num_resamples() - r0
    clocks = sampler.stop() if sampler else None
    lml = ps.log_marginal_likelihood_estimate() if world == 1 else None
    err = ps.peer_error() if world > 1 else 0
    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_alg = n * (16.0 * (K - nres) + 24.0 * nres)      # SURVEY 8d, D = 1 fp32: 16 B/particle without, 24 B with a resample
        gbs = bytes_alg / (ms * 1e-3) / 1e9
        value = n * K / (ms * 1e-3)
        line = {"metric": METRIC4, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOADS[5], "particles": f"2^{log2n}" + (" in total, sharded" if world > 1 else ""), "T_timed": K, "resampled_steps": int(nres),
                           "resampling": args.scheme + " on integer weights, ESS-triggered on the device", "l2": "inputs exceed L2", "log_ml": lml, "peer_wait_timeouts": err},
                "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "device-resident loop only: with an ESS trigger the call-per-step API needs the ESS on the host every step; all observations are uploaded once (8 B per step)"},
                "gpu_launches": int(launches) * world, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": gbs, "peak": peak * world, "unit": "GB/s", "frac": gbs / (peak * world), "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes": "16 B/particle on steps without a resample, 24 B with"},
                "cpu_baseline": cpu_baseline_config5()}
        print(json.dumps(line))
    ps.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------ configs 1, 2, 3 (one GPU)
def run_config1(args):
    import modppl_b200 as m
    T, n = 100, 1000
    ys = spiral_observations(T)
    reps = max(1, args.steps // 20)

    def one_run(seed):
        f = m.ParticleSystem(m.spiral_model(), n, seed=seed, dtype="f64")
        f.init_step(ys[0]); f.resample(m.MULTINOMIAL)
        f.sync(); t0 = time.perf_counter()
        for t in range(1, T):
            f.step(ys[t]); f.resample(m.MULTINOMIAL)      # host obs in, host log total weight out: this IS the e2e path
        f.sync()
        dt = time.perf_counter() - t0
        lc, lml = f.launch_count(), f.log_marginal_likelihood_estimate()
        f.close()
        return dt, lc, lml
    for _ in range(max(3, args.warmup) // 3):
        one_run(0)
    sampler = ClockSampler(0).start()
    runs = [one_run(1 + r) for r in range(reps)]
    clocks = sampler.stop()
    dt = float(np.mean([r[0] for r in runs]))
    value = n * (T - 1) / dt
    line = {"metric": METRIC4, "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": (T - 1) * reps, "warmup": T - 1, "ms_per_step": dt / (T - 1) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[1], "particles": n, "T": T, "resampling": "the reference's multinomial (categorical.rs:22-32), bit-exact", "l2": "56 KB of state: L2/L1 resident, launch-bound by construction",
                       "log_ml": runs[0][2]},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 128, "note": "call-per-step API; the timed region is the e2e path"},
            "gpu_launches": int(runs[0][1]), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": 56 * value / 1e9, "peak": measured_peak()[0], "unit": "GB/s", "frac": 56 * value / 1e9 / measured_peak()[0], "traffic": None,
                         "note": "N = 1000: a step is a handful of ~3 us launches; the roofline is not the bound, launch latency is"},
            "cpu_baseline": cpu_baseline_config1()}
    print(json.dumps(line))


def run_config2(args):
    import modppl_b200 as m
    xs, ys_line, ys_hier = regression_data()
    n, batches, n_ret = 1 << 20, 64, 1 << 10
    out = {}
    consts = ncu_constants()
    sampler = ClockSampler(0).start()
    for name, model, ys in (("line", m.line_model(xs), ys_line), ("hierarchical", m.hierarchical_model(xs), ys_hier)):
        for b in range(3):
            m.importance_sampling(model, ys, n, seed=0, batch=1000 + b, return_traces=False)
        t0 = time.perf_counter()
        lmls = [m.importance_sampling(model, ys, n, seed=0, batch=b, return_traces=False)[2] for b in range(batches)]
        dt_lml = time.perf_counter() - t0
        t0 = time.perf_counter()
        for b in range(8):
            m.importance_sampling(model, ys, n, seed=0, batch=b)                      # all traces and weights back to the host, as the reference returns them
        dt_full = (time.perf_counter() - t0) / 8
        t0 = time.perf_counter()
        for b in range(8):
            m.importance_resampling(model, ys, n, n_ret, seed=0, batch=b)
        dt_res = (time.perf_counter() - t0) / 8
        out[name] = {"proposals_per_s_lml_only": batches * n / dt_lml, "proposals_per_s_all_traces_to_host": n / dt_full, "importance_resampling_ms_per_batch": dt_res * 1e3,
                     "lml_mean_over_batches": float(np.mean(lmls)), "lml_std_over_batches": float(np.std(lmls)), "batches": batches}
    clocks = sampler.stop()
    value = out["line"]["proposals_per_s_lml_only"]
    L = 2
    line = {"metric": METRICS[2][0], "value": value, "unit": "proposals/s", "n_gpus": 1, "steps": batches, "warmup": 3, "ms_per_step": n / value * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[2], "proposals_per_batch": n, "models": out, "n_ret": n_ret, "l2": "register-resident proposals; the 8 MB weight pass is L2-resident"},
            "e2e": {"value": out["line"]["proposals_per_s_all_traces_to_host"], "unit": "proposals/s", "h2d_bytes_per_step": 88, "d2h_bytes_per_step": (L + 1) * n * 8 + 8,
                    "note": "importance_sampling as the reference returns it: every latent and normalised log-weight copied to the host; `value` returns the log-ML estimate only"},
            "gpu_launches": 3 * batches, "clocks": clocks,
            "roofline": dict({"bound": "issue slots (register-resident, fp64: one thread per proposal, 11 observation log-densities each; not an HBM path)", "peak": 100.0, "traffic": None,
                              "note": "achieved = issue-slot utilisation of is_kernel under ncu (provenance: profiles/ncu_constants.json); the batch also runs weight_reduce (fp64 exp per weight) and the normalisation"},
                             **consts.get("is_kernel", {})),
            "cpu_baseline": cpu_baseline_config2()}
    print(json.dumps(line))


def run_config3(args):
    import modppl_b200 as m
    xs, _, ys_hier = regression_data()
    n = 1 << 20
    moves_target = 10000
    sweeps = moves_target // 18 + 1          # 556 sweeps x 18 moves = 10 008 moves per chain
    ch = m.Chains(m.hierarchical_model(xs), ys_hier, n, seed=2)
    m.hierarchical_sweeps(ch, 2)
    for mask in (m.MASK_IS_LINEAR, m.MASK_A, m.MASK_B, m.MASK_C):
        m.regen_mh(ch, mask, 1)
    sampler = ClockSampler(0).start()
    acc, ms = m.hierarchical_full_sweeps(ch, sweeps, timed=True)
    clocks = sampler.stop()
    moves = sweeps * 18
    st = ch.read()
    value = n * moves / (ms * 1e-3)
    consts = ncu_constants()
    line = {"metric": METRICS[3][0], "value": value, "unit": "chain-steps/s", "n_gpus": 1, "steps": moves, "warmup": 32, "ms_per_step": ms / moves,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[3], "chains": n, "moves_per_chain": moves, "acceptance": acc / (n * moves), "l2": "register-resident chains (40 B of state each)",
                       "posterior_mean": {"is_linear": float(st[0].mean()), "a": float(st[1].mean()), "b": float(st[2].mean()), "c": float(st[3][st[0] == 0].mean())},
                       "generating_values": {"a": 0.3, "b": 0.4, "c": 0.5}},
            "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                    "note": "chains live in HBM between calls by design (like the reference's traces live in the caller's memory); one launch runs all sweeps, the accept count comes back"},
            "gpu_launches": 1, "clocks": clocks,
            "roofline": dict({"bound": "issue slots (register-resident, fp64: one chain per thread, fused propose / score / accept; not an HBM path)", "peak": 100.0, "traffic": None},
                             **consts.get("mh_sweep_kernel", {})),
            "cpu_baseline": cpu_baseline_config3()}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5], help="BASELINE.json configs[config - 1]; 4 is the one the metric is quoted on")
    ap.add_argument("--scheme", default=None, choices=["systematic", "multinomial", "nested", "reference"],
                    help="config 4/5 resampler: nested systematic on integer weights (config 4's default: fastest when every step resamples), single-level "
                         "systematic (config 5's default: fastest when the ESS trigger fires on one step in ten), multinomial on integer weights, or the "
                         "reference's own multinomial (bit-exact parity path)")
    ap.add_argument("--log2-particles", type=int, default=LOG2_PARTICLES)
    ap.add_argument("--variant", default="global", choices=["global", "island"],
                    help="several GPUs, config 4: global resampling over all shards (default; the same ancestors as on one GPU) or one island per GPU "
                         "with local resampling and occasional island-level resampling (a different estimator)")
    ap.add_argument("--exchange-every", type=int, default=50, help="island variant: steps between two comparisons of the island weights")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.scheme is None:
        args.scheme = "systematic" if args.config == 5 else "nested"
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if args.config == 4:
        return run_config4_multi(args) if world > 1 else run_config4_single(args)
    if args.config == 5:
        return run_config5(args)
    if world > 1:   # independent batches / chains: replicas only, no collective to measure
        raise SystemExit("configs 1-3 are single-GPU benchmarks (independent chains / batches shard as replicas)")
    return {1: run_config1, 2: run_config2, 3: run_config3}[args.config](args)


if __name__ == "__main__":
    main()
