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

PEER_BLOB_BYTES = 1024


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
    issue the same sequence of step / resample calls.  Sharded resampling: the integer-weight systematic schemes
    (SYSTEMATIC_FIXED: any equal shards; SYSTEMATIC_NESTED: equal shards of whole 2^17-particle sections)."""

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

    def phase_times(self):
        """ns spent in each phase of the last step on this rank (device %globaltimer stamps; diagnostics)"""
        tr = self.trace()
        if tr[9] > 0:   # nested scheme (stamps: scripts/step_timeline.py)
            return {"extend_gate_wait": tr[1] - tr[0], "extend_start_to_last_block": tr[14] - tr[13], "to_section_pass": tr[9] - tr[14], "section_phase_a": tr[5] - tr[9],
                    "collect_records_and_top_level": tr[7] - tr[6], "to_plan_pass": tr[10] - tr[7], "plan_pass_to_expansion": tr[11] - tr[10], "expansion": tr[8] - tr[11]}
        return {"extend_gate_wait": tr[1] - tr[0], "extend_body": tr[2] - tr[1], "to_reduce_gate": tr[3] - tr[2], "reduce_gate_wait": tr[4] - tr[3],
                "reduce_body": tr[5] - tr[4], "to_scan_gate": tr[6] - tr[5], "scan_gate_wait": tr[7] - tr[6], "scan_to_signal": tr[8] - tr[7]}

    def device_barrier(self):
        """queue a device-side rendezvous of all ranks' streams: whatever is queued after it starts on all GPUs together"""
        check(lib.mpl_ps_peer_barrier(self._h))

    def nvlink_bytes(self):
        """payload bytes this rank has requested from its peers' memory so far (remote parents, weights, records)"""
        n = C.c_uint64()
        check(lib.mpl_ps_nvlink_bytes(self._h, C.byref(n)))
        return n.value

    def peer_error(self):
        e = C.c_int()
        check(lib.mpl_ps_peer_error(self._h, C.byref(e)))
        return e.value

    def close(self):
        if getattr(self, "_h", None) and getattr(self, "world", 1) > 1:
            lib.mpl_ps_peer_detach(self._h)
        super().close()


# ----------------------------------------------------------------------------------------------------------------- islands
def island_seed(seed, island):
    """every island draws from its own Philox key"""
    return (seed + 0x9E3779B97F4A7C15 * (island + 1)) & 0xFFFFFFFFFFFFFFFF


def island_resampling_plan(deltas, epoch, seed, ess_fraction=0.5):
    """Island-level decision, identical on every rank (pure function of what all ranks know).

    deltas[g]: island g's log-ML increment since the last island-level resampling = the log of its weight.  Returns
    (resample?, ancestors[g], log_mean_weight, island_ess).  When the ESS of the island weights is below ess_fraction * G the
    islands are resampled systematically by weight (ancestors[g] = the island that island g continues from) and their
    weights start again from 1; log_mean_weight then goes into the running log-ML."""
    d = np.asarray(deltas, dtype=np.float64)
    G = d.size
    mx = d.max()
    w = np.exp(d - mx)
    ess = w.sum() ** 2 / np.square(w).sum()
    log_mean = mx + np.log(w.mean())
    if G == 1 or ess >= ess_fraction * G:
        return False, np.arange(G), log_mean, ess
    u = np.random.default_rng([int(seed) & 0xFFFFFFFF, int(epoch)]).random()
    cum = np.cumsum(w / w.sum())
    cum[-1] = 1.0
    anc = np.searchsorted(cum, (u + np.arange(G)) / G, side="right").clip(0, G - 1)
    # keep survivors in place: an island that is selected stays where it is, so only the others are overwritten
    counts = np.bincount(anc, minlength=G)
    out = np.arange(G)
    spare = [g for g in range(G) for _ in range(counts[g] - 1) if counts[g] > 1]
    for g in range(G):
        if counts[g] == 0:
            out[g] = spare.pop()
    return True, out, log_mean, ess


class _IslandBase:
    """Local-resample variant: G islands of N / G particles, each a complete particle filter of its own that
    resamples locally every step -- no per-step exchange, no rank ever waits for another inside a step.  Every `exchange_every`
    steps the islands' weights (their log-ML increments) are compared; when their ESS has dropped below half the number of
    islands, whole islands are resampled (copied over NVLink).  log-ML estimate: the sum over those epochs of log mean island
    weight -- unbiased for the likelihood, but NOT the estimator of the global scheme (other ancestors, other variance)."""

    def _init_island_state(self, seed):
        self.seed = seed
        self.log_ml_base = 0.0          # log-ML accumulated at past island-level resamplings
        self.lml_ref = 0.0              # this island's own running log-ML at the last island-level resampling
        self.epoch = 0
        self.n_island_resamplings = 0
        self.bytes_exchanged = 0

    def _exchange(self, deltas_all):
        raise NotImplementedError


class IslandParticleSystem(_IslandBase):
    """one island per process / GPU (torchrun); torch.distributed (gloo) carries 16 bytes per rank per comparison"""

    def __init__(self, model, n_global, rank, world, seed=0, dtype="f32", device=-1, dist=None):
        import torch.distributed as tdist
        self.dist = dist or tdist
        off, n_loc = shard_range(n_global, rank, world)
        self.rank, self.world, self.n_global, self.n_loc = rank, world, n_global, n_loc
        self.ps = ParticleSystem(model, n_loc, seed=island_seed(seed, rank), dtype=dtype, device=device)
        self._elem_bytes = 4 if dtype in ("f32", 0) else 8
        self._steps_since_compare = 0
        self._init_island_state(seed)
        if world > 1:
            blob = C.create_string_buffer(PEER_BLOB_BYTES)
            check(lib.mpl_ps_island_export(self.ps._h, blob))
            blobs = all_gather_bytes(blob.raw, self.dist)
            joined = C.create_string_buffer(b"".join(blobs), PEER_BLOB_BYTES * world)
            check(lib.mpl_ps_island_attach(self.ps._h, rank, world, joined))

    def upload_observations(self, obs):
        self.ps.upload_observations(obs)

    def _gather(self, values):
        import torch
        t = torch.tensor(values, dtype=torch.float64)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.numpy() for o in out])

    def compare_islands(self):
        """the island-level step: compare weights, resample islands if their ESS has dropped"""
        mine = self.ps.log_marginal_likelihood_estimate() - self.lml_ref       # (one small device-to-host read)
        table = self._gather([mine, 0.0])
        resample, anc, log_mean, ess = island_resampling_plan(table[:, 0], self.epoch, self.seed)
        self.epoch += 1
        if resample:
            live = C.c_int()
            check(lib.mpl_ps_live_buffer(self.ps._h, C.byref(live)))          # applies this island's pending local resample
            table = self._gather([mine, float(live.value)])                   # (doubles as the barrier: every island is at rest)
            src = int(anc[self.rank])
            if src != self.rank:
                check(lib.mpl_ps_island_copy_from(self.ps._h, src, int(table[src, 1])))
                self.bytes_exchanged += self.n_loc * self.ps.state_dim * self._elem_bytes
            self.dist.barrier()                                   # nobody continues before every copy has been read
            self.log_ml_base += log_mean
            self.lml_ref = self.ps.log_marginal_likelihood_estimate()
            self.n_island_resamplings += 1
        return ess

    def run(self, first_step, n_steps, scheme, exchange_every=50):
        """n_steps x (step; local resample) in segments of `exchange_every` steps with an island comparison after each; returns the
        summed device time of the segments (ms) -- the comparisons are host work between them (time the whole call for e2e)."""
        ms, t = 0.0, first_step
        while t < first_step + n_steps:
            k = min(exchange_every, first_step + n_steps - t)
            ms += self.ps.run(t, k, scheme)
            t += k
            if self.world > 1:
                self.compare_islands()
        return ms

    def step_resample(self, constraints, scheme, exchange_every=50):
        """the call-per-step API: `step(obs); resample()` on this island, with an island comparison every `exchange_every` calls"""
        lse = self.ps.step_resample(constraints, scheme)
        self._steps_since_compare += 1
        if self.world > 1 and self._steps_since_compare >= exchange_every:
            self._steps_since_compare = 0
            self.compare_islands()
        return lse

    def log_marginal_likelihood_estimate(self):
        mine = self.ps.log_marginal_likelihood_estimate() - self.lml_ref
        d = self._gather([mine, 0.0])[:, 0] if self.world > 1 else np.array([mine])
        mx = d.max()
        return self.log_ml_base + mx + float(np.log(np.mean(np.exp(d - mx))))

    def launch_count(self):
        return self.ps.launch_count()

    def sync(self):
        self.ps.sync()

    def close(self):
        self.ps.close()


class VirtualIslands(_IslandBase):
    """the same island filter with all islands in ONE process on one GPU (tests; also a way to run G independent filters whose
    estimates are combined): exchanges are device-to-device copies"""

    def __init__(self, model, n_global, n_islands, seed=0, dtype="f32", device=-1):
        assert n_global % n_islands == 0
        self.G, self.n_loc = n_islands, n_global // n_islands
        self.islands = [ParticleSystem(model, self.n_loc, seed=island_seed(seed, g), dtype=dtype, device=device) for g in range(n_islands)]
        self._init_island_state(seed)
        self.lml_ref = np.zeros(n_islands)

    def upload_observations(self, obs):
        for ps in self.islands:
            ps.upload_observations(obs)

    def compare_islands(self):
        lmls = np.array([ps.log_marginal_likelihood_estimate() for ps in self.islands])
        resample, anc, log_mean, ess = island_resampling_plan(lmls - self.lml_ref, self.epoch, self.seed)
        self.epoch += 1
        if resample:
            for g in range(self.G):
                if anc[g] != g:
                    check(lib.mpl_ps_copy_state(self.islands[g]._h, self.islands[int(anc[g])]._h))
                    self.bytes_exchanged += self.n_loc * self.islands[g].state_dim * 4
            self.log_ml_base += log_mean
            self.lml_ref = np.array([ps.log_marginal_likelihood_estimate() for ps in self.islands])
            self.n_island_resamplings += 1
        return ess

    def run(self, first_step, n_steps, scheme, exchange_every=50):
        ms, t = 0.0, first_step
        while t < first_step + n_steps:
            k = min(exchange_every, first_step + n_steps - t)
            ms += max(ps.run(t, k, scheme) for ps in self.islands)
            t += k
            self.compare_islands()
        return ms

    def log_marginal_likelihood_estimate(self):
        d = np.array([ps.log_marginal_likelihood_estimate() for ps in self.islands]) - self.lml_ref
        mx = d.max()
        return self.log_ml_base + mx + float(np.log(np.mean(np.exp(d - mx))))

    def close(self):
        for ps in self.islands:
            ps.close()
