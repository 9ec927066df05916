"""Sharded particle systems: one process per GPU (torchrun), particles split into equal contiguous shards.

torch.distributed is plumbing only: it carries the one-off all-gather of CUDA-IPC handle blobs, barriers and the
max-over-ranks of the timings.  The per-step exchange (weight statistics, integer weight totals, cross-shard ancestors
and parents) happens inside the kernels over NVLink peer memory (modppl_b200/csrc/multi_gpu.cu).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import lib, check
from .particle_filter import ParticleSystem, SYSTEMATIC_FIXED

PEER_BLOB_BYTES = 256


def shard_range(n_global, rank, world):
    """Equal contiguous shards: rank r owns global particle ids [r * n/world, (r+1) * n/world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if n_global % world:
        raise ValueError(f"n_global={n_global} is not divisible by world={world}")
    n_loc = n_global // world
    return rank * n_loc, n_loc


def all_gather_bytes(blob: bytes, dist=None):
    """All-gather one fixed-size byte string per rank (rank order) over whatever backend the group has."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]


def max_over_ranks(value: float, dist=None):
    import torch
    if dist is None:
        import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class ShardedParticleSystem(ParticleSystem):
    """This rank's shard of a particle system of `n_global` particles; same calls as ParticleSystem.  Every rank must
    issue the same sequence of step / resample calls (only MPL_RESAMPLE_SYSTEMATIC_FIXED is supported sharded)."""

    def __init__(self, model, n_global, rank, world, seed=0, dtype="f32", device=-1, dist=None):
        off, n_loc = shard_range(n_global, rank, world)
        super().__init__(model, n_loc, seed=seed, dtype=dtype, device=device, gid_offset=off, n_global=n_global)
        self.rank, self.world, self.n_global = rank, world, n_global
        if world > 1:
            blob = C.create_string_buffer(PEER_BLOB_BYTES)
            check(lib.mpl_ps_peer_export(self._h, blob))
            blobs = all_gather_bytes(blob.raw, dist)
            joined = C.create_string_buffer(b"".join(blobs), PEER_BLOB_BYTES * world)
            check(lib.mpl_ps_peer_attach(self._h, rank, world, joined))

    def trace(self):
        buf = (C.c_longlong * 16)()
        check(lib.mpl_ps_trace(self._h, buf))
        return list(buf)

    def peer_error(self):
        e = C.c_int()
        check(lib.mpl_ps_peer_error(self._h, C.byref(e)))
        return e.value

    def close(self):
        if getattr(self, "_h", None) and getattr(self, "world", 1) > 1:
            lib.mpl_ps_peer_detach(self._h)
        super().close()


def bench_multi(args, ys, scheme, rank, world, local_rank):
    """bench.py's N > 1 arm: strong scaling of the config-4 workload (2^24 particles in total)."""
    import json
    import time
    import torch
    import torch.distributed as dist
    import bench as B
    import modppl_b200 as m

    torch.cuda.set_device(local_rank)
    # control plane only (handle exchange, barriers, max-over-ranks of timings): gloo keeps stdout clean and NCCL off
    # the critical path; the data path is NVLink peer memory inside the kernels
    dist.init_process_group("gloo")
    n_global = 1 << args.log2_particles
    K, W = args.steps, args.warmup
    ps = ShardedParticleSystem(m.lgssm4(), n_global, rank, world, seed=1, dtype="f32", device=local_rank)
    ps.upload_observations(ys)
    dist.barrier(); torch.cuda.synchronize()
    ps.run(0, 1 + W, scheme)
    ps.sync()
    dist.barrier(); torch.cuda.synchronize()
    sampler = B.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.5)      # nvidia-smi needs a moment to enumerate 8 GPUs
    l0 = ps.launch_count()
    dist.barrier(); torch.cuda.synchronize()
    ms = ps.run(1 + W, K, scheme)
    ps.sync()
    dist.barrier(); torch.cuda.synchronize()
    ms = max_over_ranks(ms)
    launches = ps.launch_count() - l0
    tr = ps.trace()     # last step of the timed run, this rank's clock (ns)
    if scheme == m.SYSTEMATIC_NESTED:   # stamps: see scripts/step_timeline.py
        phases = {"extend_gate_wait": tr[1] - tr[0], "extend_start_to_last_block": tr[14] - tr[13], "to_section_pass": tr[9] - tr[14], "section_phase_a": tr[5] - tr[9],
                  "collect_records_and_top_level": tr[7] - tr[6], "to_level1": tr[10] - tr[7], "level1_to_expansion": tr[11] - tr[10],
                  "expansion_to_done_signal": tr[8] - tr[11]}
    else:
        phases = {"extend_gate_wait": tr[1] - tr[0], "extend_body": tr[2] - tr[1], "to_reduce_gate": tr[3] - tr[2], "reduce_gate_wait": tr[4] - tr[3],
                  "reduce_body": tr[5] - tr[4], "to_scan_gate": tr[6] - tr[5], "scan_gate_wait": tr[7] - tr[6], "scan_to_signal": tr[8] - tr[7]}
    all_phases = [None] * world
    dist.all_gather_object(all_phases, phases)
    # e2e: one host round trip per step on every rank
    t_first = 1 + W + K
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        ps.step_resample(ys[t_first + k], scheme)
    ps.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    lml = ps.log_marginal_likelihood_estimate()     # same point of the run as the single-GPU arm: identical by construction
    # per-kernel CUDA-event times (includes time spent waiting for peers inside the kernels)
    ps.profile_enable(True)
    t_prof = t_first + K
    for k in range(min(K, len(ys) - t_prof)):
        ps.step_resample(ys[t_prof + k], scheme, sync=False)
    prof = {k: ps.profile_get(k) for k in ("extend", "fixed_reduce", "fixed_scan", "fixed_overflow", "nested_quantise", "nested_sections", "nested_level1", "nested_scan")}
    ps.profile_enable(False)
    kernel_ms = {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()}
    err = ps.peer_error()
    dist.barrier()
    if rank == 0:
        value = n_global * K / (ms * 1e-3)
        peak, peak_src = B.measured_peak()
        step_gbs = B.BYTES_PER_PARTICLE_STEP * value / 1e9
        line = {
            "metric": "particle-steps/sec (SMC step incl. resample)", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "lgssm4 (4-D linear-Gaussian SSM) bootstrap particle filter, resample every step", "particles": f"2^{args.log2_particles} in total, sharded",
                       "T_timed": K, "resampling": f"global {args.scheme} resampling on integer weights; NVLink peer loads/stores inside the kernels, no NCCL on the data path",
                       "l2": "per-GPU state buffers stream every step", "log_ml": lml, "peer_wait_timeouts": err},
            "e2e": {"value": n_global * K / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 24},
            "gpu_launches": int(launches) * world, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": step_gbs, "peak": peak * world, "unit": "GB/s", "frac": step_gbs / (peak * world), "traffic": None,
                         "peak_source": peak_src + f" x {world} GPUs", "kernel_ms_rank0": kernel_ms},
            "phase_ns_per_rank": all_phases,
        }
        print(json.dumps(line))
    ps.close()
    dist.destroy_process_group()
