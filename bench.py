#!/usr/bin/env python
"""bench.py -- particle-steps/sec of one SMC step (extend + normalise + resample) on the config-4 workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

workload (BASELINE.json configs[3]): 4-D linear-Gaussian state-space model, N = 2^24 particles (sharded over the
GPUs: strong scaling), fp32 state and log-weights, bootstrap proposal, resample after every step.  A "step" is one
`step(obs_t); resample()` over all particles.
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
EXTEND_BYTES_PER_PARTICLE = 40    # ancestor 4 + parent state 16 + child state 16 + log-weight 4
EXTEND_DRAM_BYTES_NCU = 571104000   # dram__bytes_read.sum + dram__bytes_write.sum of one pf_extend_kernel<.., GATHER, NESTED> launch at N = 2^24 (profiles/r1_ncu_summary_nested.txt)


def observations(T, seed=4):
    rng = np.random.default_rng(seed)
    A = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], float)
    x = rng.normal(size=4)
    ys = np.empty((T, 2))
    for t in range(T):
        if t > 0:
            x = A @ x + 0.1 * rng.normal(size=4)
        ys[t] = x[:2] + 0.5 * rng.normal(size=2)
    return ys


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


ORACLE_SCHEME = {"systematic": 2, "multinomial": 3, "nested": 4}


def cpu_baseline_port(steps=6, log2n=20, scheme=4):
    """The oracle's particle filter (restates inference/particle_filter.rs) on a bounded sample of the workload.  The
    reference itself is single-threaded (ThreadRng is !Send); the port is timed on one host thread AND with its
    per-particle loops spread over all host threads available to this process -- the better of the two is `value`.
    Resampling uses cumsum + search, i.e. the reference's algorithm without its O(N^2) per-draw clone-and-sum; the
    faithful cost is reported separately."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    n = 1 << log2n
    ys = observations(steps + 1)
    threads_all = max(1, len(os.sched_getaffinity(0))) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rates, total_s = {}, 0.0
    for th in sorted({1, threads_all}):
        O.L.mo_set_threads(th)
        ps = O.OraclePS("lgssm4", [0.1, 0.5, 1.0], n, dtype="f32", seed=1)
        ps.init_step(ys[0]); ps.resample(scheme)
        t0 = time.perf_counter()
        for t in range(1, steps + 1):
            ps.step(ys[t]); ps.resample(scheme)
        dt = time.perf_counter() - t0
        rates[th] = n * steps / dt
        total_s += dt
    O.L.mo_set_threads(1)
    best = max(rates, key=rates.get)
    # faithful reference cost (categorical.rs:22-32: clone + sum + linear scan per draw) at N = 2^12: O(N^2)
    nf = 1 << 12
    pf = O.OraclePS("lgssm4", [0.1, 0.5, 1.0], nf, dtype="f64", seed=1)
    pf.init_step(ys[0])
    t0 = time.perf_counter()
    pf.resample_faithful_cost(); pf.step(ys[1]); pf.resample_faithful_cost()
    dtf = time.perf_counter() - t0
    scheme_name = ("nested " if scheme == 4 else "") + ("multinomial" if scheme == 3 else "systematic")
    return {"value": rates[best], "unit": "particle-steps/s", "cores": best, "kind": "port",
            "sample": f"oracle PF, lgssm4 f32, N=2^{log2n}, {steps} steps, {scheme_name} resampling on integer weights (O(N)); "
                      + "; ".join(f"{th} thread{'s' if th > 1 else ''}: {r:.3g}/s" for th, r in sorted(rates.items()))
                      + f"; faithful O(N^2) reference resample at N=2^12: {nf * 2 / dtf:.3g} particle-steps/s (extrapolates to ~{nf * 2 / dtf * nf / (1 << 24):.2g}/s at N=2^24)",
            "seconds": total_s + dtf}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline_port(steps=max(2, min(args.steps, 8)), scheme=ORACLE_SCHEME[args.scheme])
    line = {"impl": "reference", "metric": "particle-steps/sec (SMC step incl. resample)", "value": base["value"], "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "lgssm4 bootstrap particle filter, resample every step (CPU port of modppl's ParticleSystem; reference is Rust, no toolchain here)",
                       "particles": "2^20 sample of 2^24", "T": args.steps},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import modppl_b200 as m
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    n_global = 1 << args.log2_particles
    K, W = args.steps, args.warmup
    T = 1 + W + K + K + 8
    ys = observations(T)
    scheme = {"systematic": m.SYSTEMATIC_FIXED, "multinomial": m.MULTINOMIAL_FIXED, "nested": m.SYSTEMATIC_NESTED}[args.scheme]
    if world > 1:
        from modppl_b200 import distributed as D
        return D.bench_multi(args, ys, scheme, rank, world, local_rank)

    ps = m.ParticleSystem(m.lgssm4(), n_global, seed=1, dtype="f32", device=local_rank)
    ps.upload_observations(ys)
    ps.run(0, 1 + W, scheme)                                   # init + resample + warm-up steps (untimed)
    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.3)
    l0 = ps.launch_count()
    ms = ps.run(1 + W, K, scheme)                              # timed: K steps, CUDA events on the launching stream
    launches = ps.launch_count() - l0
    value = n_global * K / (ms * 1e-3)
    tr = ps.device_trace()                                     # %globaltimer stamps of the last timed step (scripts/step_timeline.py)
    ext_in_loop_ms = (tr[14] - tr[13]) * 1e-6 if tr[14] > tr[13] > 0 else None

    # e2e: the reference-facing calls, one host round trip per step (observation in, log total weight out)
    t_first = 1 + W + K
    ps.sync()
    t0 = time.perf_counter()
    for k in range(K):
        ps.step_resample(ys[t_first + k], scheme)              # observation: host -> device; returns log_total_weight: device -> host
    ps.sync()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    lml = ps.log_marginal_likelihood_estimate()

    # roofline of the dominant kernel: per-kernel CUDA-event times over a fresh pass
    ps2_first = t_first + K
    ps.profile_enable(True)
    steps_prof = min(K, T - ps2_first)
    for k in range(steps_prof):
        ps.step_resample(ys[ps2_first + k], scheme, sync=False)   # (the same kernels as the timed loop, each between its own pair of events)
    prof = {k: ps.profile_get(k) for k in ("extend", "fixed_reduce", "fixed_scan", "fixed_overflow", "fixed_cumsum", "fixed_search", "nested_quantise", "nested_sections", "nested_level1", "nested_scan")}
    ps.profile_enable(False)
    peak, peak_src = measured_peak()
    ext_ms = prof["extend"][0] / max(1, prof["extend"][1])
    achieved = EXTEND_BYTES_PER_PARTICLE * n_global / (ext_ms * 1e-3) / 1e9
    step_gbs = BYTES_PER_PARTICLE_STEP * value / 1e9
    base = cpu_baseline_port(scheme=ORACLE_SCHEME[args.scheme])
    line = {
        "metric": "particle-steps/sec (SMC step incl. resample)", "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "lgssm4 (4-D linear-Gaussian SSM) bootstrap particle filter, resample every step", "particles": f"2^{args.log2_particles}",
                   "T_timed": K, "resampling": args.scheme + " on integer weights", "l2": "inputs exceed L2 (2 x 256 MiB state buffers stream every step)",
                   "log_ml": lml},
        "e2e": {"value": n_global * K / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 24,
                "note": "mpl_ps_step_resample(host obs -> host log total weight) per step: the observation travels as a kernel argument, the log total weight is written by the level-1 kernel into mapped pinned host memory (tagged words) and polled there, so the call returns as soon as the value exists and the next step queues behind the running expansion; particles stay in HBM by design"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "pf_extend_kernel<Lgssm4<float>, GATHER" + (", NESTED>" if args.scheme == "nested" else ">"), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": EXTEND_DRAM_BYTES_NCU if (args.log2_particles == 24 and args.scheme == "nested") else None, "peak_source": peak_src, "algorithmic_bytes_per_particle": EXTEND_BYTES_PER_PARTICLE,
                     "algorithmic_bytes_per_launch": EXTEND_BYTES_PER_PARTICLE * n_global, "traffic_source": "profiles/r1_ncu_summary_nested.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch; nested scheme)",
                     "whole_step": {"bytes_per_particle": BYTES_PER_PARTICLE_STEP, "achieved": step_gbs, "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0},
                     "kernel_ms": {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()},
                     "in_loop": None if not ext_in_loop_ms else {
                         "extend_ms": ext_in_loop_ms, "achieved": EXTEND_BYTES_PER_PARTICLE * n_global / (ext_in_loop_ms * 1e-3) / 1e9,
                         "frac": EXTEND_BYTES_PER_PARTICLE * n_global / (ext_in_loop_ms * 1e-3) / 1e9 / peak,
                         "source": "device %globaltimer stamps inside the timed loop (extend block 0 past its dependency wait -> last block done); "
                                   "the CUDA-event figure above times the kernel between its own pair of events, outside the loop's programmatic dependent launch"}},
        "cpu_baseline": base,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scheme", default="nested", choices=["systematic", "multinomial", "nested"],
                    help="resampler on integer weights: nested systematic (default, fastest), single-level systematic, multinomial")
    ap.add_argument("--log2-particles", type=int, default=LOG2_PARTICLES)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
