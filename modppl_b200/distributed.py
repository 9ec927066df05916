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

    def peer_error(self):
        e = C.c_int()
        check(lib.mpl_ps_peer_error(self._h, C.byref(e)))
        return e.value

    def close(self):
        if getattr(self, "_h", None) and getattr(self, "world", 1) > 1:
            lib.mpl_ps_peer_detach(self._h)
        super().close()
