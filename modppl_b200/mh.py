"""Many-chain Metropolis-Hastings: reference modppl/src/inference/mh.rs:9-76, one independent chain per GPU thread."""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import lib, check, check_handle

HIER_DRIFT, HIER_ADD_REMOVE, HIER_REGEN, POINTED_DRIFT = 0, 1, 2, 3
MASK_A, MASK_B, MASK_C, MASK_IS_LINEAR = 1, 2, 4, 8


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


def metropolis_hastings(chains, proposal, proposal_args, n_steps=1):
    """mh.rs:9-40 applied `n_steps` times to every chain; returns the number of accepted transitions."""
    acc = C.c_uint64()
    check(lib.mpl_mh(chains._h, proposal, float(proposal_args), n_steps, C.byref(acc)))
    return acc.value


mh = metropolis_hastings


def regenerative_metropolis_hastings(chains, mask, n_steps=1):
    """mh.rs:54-67 with `mask` a bit set over {coeffs/a, coeffs/b, coeffs/c, is_linear}."""
    acc = C.c_uint64()
    check(lib.mpl_regen_mh(chains._h, mask, n_steps, C.byref(acc)))
    return acc.value


regen_mh = regenerative_metropolis_hastings


def hierarchical_sweeps(chains, n_sweeps, timed=False):
    """n_sweeps x (1 add/remove + 3 drift(.1) + 10 drift(.01)): the schedule of tests/mh.rs:93-106, fused in one launch."""
    acc, ms = C.c_uint64(), C.c_float()
    check(lib.mpl_mh_hier_sweeps(chains._h, n_sweeps, C.byref(acc), C.byref(ms) if timed else None))
    return (acc.value, ms.value) if timed else acc.value
