"""ParticleSystem: host mirror of reference modppl/src/inference/particle_filter.rs:8-121 over the C ABI."""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import lib, check, check_handle, PfConfig

F32, F64 = 0, 1
MULTINOMIAL, SYSTEMATIC, SYSTEMATIC_FIXED, MULTINOMIAL_FIXED, SYSTEMATIC_NESTED = 0, 1, 2, 3, 4
_DTYPES = {"f32": F32, "f64": F64, F32: F32, F64: F64}


def _obs(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
    return a, a.ctypes.data_as(_lib.c_double_p), a.size


class ParticleSystem:
    """Basic particle filter for Unfold-style models (particle_filter.rs:7).

    `seed` replaces the ThreadRng argument of `ParticleSystem::new` (:44).  `step` returns self, standing in for
    the reference's `step(self) -> Self` (:73)."""

    def __init__(self, model, num_particles, seed=0, dtype="f64", device=-1, gid_offset=0, n_global=0):
        self.model = model
        self.num_particles = int(num_particles)
        cfg = PfConfig(_DTYPES[dtype], device, seed, gid_offset, n_global)
        self._h = check_handle(lib.mpl_particle_system_new(model._h, self.num_particles, C.byref(cfg)))
        self.state_dim = model.state_dim

    def init_step(self, constraints):                      # :60-70
        a, p, n = _obs(constraints)
        check(lib.mpl_ps_init_step(self._h, p, n))
        self._steps_done = 1

    def step(self, constraints):                           # :73-95
        a, p, n = _obs(constraints)
        check(lib.mpl_ps_step(self._h, p, n))
        self._steps_done = getattr(self, "_steps_done", 0) + 1
        return self

    def effective_sample_size(self, stale_like_reference=True):   # :98-100 (quirk Q1: the reference value is stale)
        out = C.c_double()
        check(lib.mpl_ps_effective_sample_size(self._h, int(bool(stale_like_reference)), C.byref(out)))
        return out.value

    def resample(self, scheme=MULTINOMIAL, sync=True):     # :103-116
        out = C.c_double()
        check(lib.mpl_ps_resample(self._h, scheme, C.byref(out) if sync else None))
        return out.value if sync else None

    def step_resample(self, constraints, scheme=MULTINOMIAL, sync=True):
        """`filter = filter.step(..); filter.resample()` (tests/smc.rs:78-81) as one call: same results, and the nested scheme
        on fp32 gets its weights quantised by the extend kernel."""
        a, p, n = _obs(constraints)
        out = C.c_double()
        check(lib.mpl_ps_step_resample(self._h, p, n, scheme, C.byref(out) if sync else None))
        self._steps_done = getattr(self, "_steps_done", 0) + 1
        return out.value if sync else None

    def checkpoint(self):
        """-> bytes: the filter's resumable state (native precision); see `restore`."""
        n = C.c_uint64()
        check(lib.mpl_ps_checkpoint_size(self._h, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        check(lib.mpl_ps_checkpoint(self._h, buf.ctypes.data_as(C.c_void_p), n.value))
        return buf.tobytes()

    def restore(self, blob):
        """Continue from a checkpoint taken from a filter with the same model, particle count, precision and seed."""
        buf = np.frombuffer(blob, dtype=np.uint8)
        check(lib.mpl_ps_restore(self._h, buf.ctypes.data_as(C.c_void_p), buf.size))
        self._steps_done = int(np.frombuffer(blob, dtype=np.int64, count=1, offset=64)[0])   # CkptHeader.t
        return self

    def device_trace(self):
        """%globaltimer stamps (ns) left by the kernels of the last step (diagnostics; include/modppl_b200.h: mpl_ps_trace)."""
        buf = (C.c_longlong * 16)()
        check(lib.mpl_ps_trace(self._h, buf))
        return list(buf)

    def log_marginal_likelihood_estimate(self):            # :119-121
        out = C.c_double()
        check(lib.mpl_ps_log_marginal_likelihood_estimate(self._h, C.byref(out)))
        return out.value

    # `pub traces` (:13): the live state of every particle, SoA [D, N]
    @property
    def traces(self):
        out = np.empty((self.state_dim, self.num_particles), dtype=np.float64)
        check(lib.mpl_ps_read(self._h, 0, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    @property
    def log_weights(self):
        out = np.empty(self.num_particles, dtype=np.float64)
        check(lib.mpl_ps_read(self._h, 1, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    @property
    def parents(self):
        out = np.empty(self.num_particles, dtype=np.int64)
        check(lib.mpl_ps_read(self._h, 2, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def write_state(self, state):
        a = np.ascontiguousarray(np.asarray(state, dtype=np.float64).reshape(self.state_dim, self.num_particles))
        check(lib.mpl_ps_write(self._h, 0, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def write_log_weights(self, lw):
        a = np.ascontiguousarray(np.asarray(lw, dtype=np.float64).ravel())
        check(lib.mpl_ps_write(self._h, 1, a.ctypes.data_as(C.c_void_p), a.nbytes))

    # device-resident loop
    def upload_observations(self, obs):
        a = np.ascontiguousarray(np.asarray(obs, dtype=np.float64))
        a = a.reshape(a.shape[0], -1)
        check(lib.mpl_ps_upload_observations(self._h, a.ctypes.data_as(_lib.c_double_p), a.shape[0], a.shape[1]))

    def run(self, first_step, n_steps, scheme=SYSTEMATIC_FIXED, ess_threshold=0.0, timed=True):
        ms = C.c_float(0.0)
        check(lib.mpl_ps_run(self._h, first_step, n_steps, scheme, ess_threshold, C.byref(ms) if timed else None))
        self._steps_done = first_step + n_steps
        return ms.value

    def enable_history(self, max_steps):
        """Log per-step states and ancestors so that `trajectories` can rebuild `traces[i].retv` (dynunfold.rs:91-92)."""
        check(lib.mpl_ps_history_enable(self._h, int(max_steps)))
        self._hist_cap = int(max_steps)

    def trajectories(self, ids):
        """-> array [len(ids), T, D]: the lineage of each particle, oldest step first."""
        ids = np.ascontiguousarray(np.asarray(ids, dtype=np.int64).ravel())
        T = C.c_uint64()
        steps = self._steps_done        # logged steps == init_step + step calls so far
        out = np.empty((ids.size, steps, self.state_dim), dtype=np.float64)
        check(lib.mpl_ps_trajectories(self._h, ids.ctypes.data_as(_lib.c_i64_p), ids.size, out.ctypes.data_as(_lib.c_double_p), out.nbytes, C.byref(T)))
        return out

    def num_resamples(self):
        n = C.c_uint64()
        check(lib.mpl_ps_num_resamples(self._h, C.byref(n)))
        return n.value

    def sync(self):
        check(lib.mpl_ps_sync(self._h))

    def profile_enable(self, on=True):
        check(lib.mpl_ps_profile_enable(self._h, int(on)))

    def profile_get(self, kernel):
        ms, n = C.c_double(), C.c_uint64()
        check(lib.mpl_ps_profile_get(self._h, kernel.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        n = C.c_uint64()
        check(lib.mpl_ps_launch_count(self._h, C.byref(n)))
        return n.value

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.mpl_ps_destroy(h)

    def __del__(self):
        self.close()
