"""Many-chain Metropolis-Hastings: reference modppl/src/inference/mh.rs:9-76, one independent chain per GPU thread.

`mh(model, trace, proposal, proposal_args)` is generic over the proposal (mh.rs:9-14): here a proposal is selected by the
name of the reference's fixture -- the model registers its proposals as device functors (csrc/is_mh.cu)."""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import lib, check, check_handle, Move

# proposals registered by the built-in static models, under the reference fixtures' names
HIER_DRIFT = "hierarchical_drift_proposal"          # tests/dyngenfns/hierarchical.rs:63-71
HIER_ADD_REMOVE = "add_or_remove_param_proposal"    # tests/dyngenfns/hierarchical.rs:48-61
POINTED_DRIFT = "pointed_2d_drift_proposal"         # tests/pointed_model/proposal.rs, tests/dyngenfns/simple.rs:33-39
MASK_A, MASK_B, MASK_C, MASK_IS_LINEAR = 1, 2, 4, 8
MOVE_MH, MOVE_REGEN = 0, 1


class Chains:
    """n independent traces, each initialised as `model.generate(args, observations).0` (tests/mh.rs:34,61,91)."""

    def __init__(self, model, constraints, n_chains, seed=0, chain_offset=0, device=-1):
        a = np.ascontiguousarray(np.asarray(constraints, dtype=np.float64).ravel())
        self.n = int(n_chains)
        self.model = model
        self._h = check_handle(lib.mpl_chains_new(model._h, a.ctypes.data_as(_lib.c_double_p), a.size, self.n, seed, chain_offset, device))
        self.slots = lib.mpl_chains_num_slots(self._h)

    def read(self):
        out = np.empty((self.slots, self.n), dtype=np.float64)
        check(lib.mpl_chains_read(self._h, out.ctypes.data_as(_lib.c_double_p), out.nbytes))
        return out

    def write(self, st):
        a = np.ascontiguousarray(np.asarray(st, dtype=np.float64).reshape(self.slots, self.n))
        check(lib.mpl_chains_write(self._h, a.ctypes.data_as(_lib.c_double_p), a.nbytes))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.mpl_chains_destroy(h)

    def __del__(self):
        self.close()


def proposals(model):
    """names of the proposals registered for a static model"""
    return [lib.mpl_model_proposal_name(model._h, i).decode() for i in range(lib.mpl_model_num_proposals(model._h))]


def metropolis_hastings(chains, proposal, proposal_args, n_steps=1):
    """mh.rs:9-40 applied `n_steps` times to every chain with the named proposal; returns the number of accepted transitions."""
    acc = C.c_uint64()
    check(lib.mpl_mh(chains._h, proposal.encode(), float(proposal_args), n_steps, C.byref(acc)))
    return acc.value


mh = metropolis_hastings


def regenerative_metropolis_hastings(chains, mask, n_steps=1):
    """mh.rs:54-67 with `mask` a bit set over {coeffs/a, coeffs/b, coeffs/c, is_linear}."""
    acc = C.c_uint64()
    check(lib.mpl_regen_mh(chains._h, mask, n_steps, C.byref(acc)))
    return acc.value


regen_mh = regenerative_metropolis_hastings


def run_schedule(chains, moves, n_sweeps=1, timed=False):
    """`moves`: the body of the caller's MCMC loop as a list of ("mh", proposal_name, arg, repeat) / ("regen_mh", mask, repeat)
    entries, run `n_sweeps` times per chain in ONE launch (the chain's state stays in registers)."""
    arr = (Move * len(moves))()
    for k, mv in enumerate(moves):
        if mv[0] == "mh":
            idx = lib.mpl_model_proposal_index(chains.model._h, mv[1].encode())
            check(min(idx, 0))
            arr[k] = Move(MOVE_MH, idx, float(mv[2]), 0, int(mv[3]))
        elif mv[0] == "regen_mh":
            arr[k] = Move(MOVE_REGEN, -1, 1.0, int(mv[1]), int(mv[2]))
        else:
            raise ValueError(f"unknown move {mv[0]!r}")
    acc, ms = C.c_uint64(), C.c_float()
    check(lib.mpl_mh_schedule(chains._h, arr, len(moves), n_sweeps, C.byref(acc), C.byref(ms) if timed else None))
    return (acc.value, ms.value) if timed else acc.value


# tests/mh.rs:93-106: per sweep 1 add/remove(.025) + 3 drift(.1) + 10 drift(.01)
HIER_SWEEP = [("mh", HIER_ADD_REMOVE, 0.025, 1), ("mh", HIER_DRIFT, 0.1, 3), ("mh", HIER_DRIFT, 0.01, 10)]
# config 3 of BASELINE.json: the same sweep followed by regen_mh on {is_linear}, {coeffs/a}, {coeffs/b}, {coeffs/c}
HIER_FULL_SWEEP = HIER_SWEEP + [("regen_mh", MASK_IS_LINEAR, 1), ("regen_mh", MASK_A, 1), ("regen_mh", MASK_B, 1), ("regen_mh", MASK_C, 1)]


def hierarchical_sweeps(chains, n_sweeps, timed=False):
    """n_sweeps x the schedule of tests/mh.rs:93-106 (14 moves), fused in one launch."""
    return run_schedule(chains, HIER_SWEEP, n_sweeps, timed)


def hierarchical_full_sweeps(chains, n_sweeps, timed=False):
    """n_sweeps x (the 14 proposal moves + 4 regen_mh moves) = 18 moves per sweep, fused in one launch."""
    return run_schedule(chains, HIER_FULL_SWEEP, n_sweeps, timed)
