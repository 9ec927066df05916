"""importance_sampling / importance_resampling: reference modppl/src/inference/importance.rs:12-51."""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import lib, check


def _obs(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
    return a, a.ctypes.data_as(_lib.c_double_p), a.size


def importance_sampling(model, constraints, num_samples, seed=0, batch=0, return_traces=True):
    """-> (traces [L, n] latents, log normalised weights [n], log-ML estimate)   (importance.rs:12-28)
    return_traces=False: the proposals stay on the GPU and only the log-ML estimate comes back -> (None, None, lml)."""
    a, p, n = _obs(constraints)
    lml = C.c_double()
    if not return_traces:
        check(lib.mpl_importance_sampling(model._h, p, n, num_samples, seed, batch, None, None, C.byref(lml)))
        return None, None, lml.value
    lat = np.empty((model.num_latents, num_samples), dtype=np.float64)
    lnw = np.empty(num_samples, dtype=np.float64)
    check(lib.mpl_importance_sampling(model._h, p, n, num_samples, seed, batch, lat.ctypes.data_as(_lib.c_double_p), lnw.ctypes.data_as(_lib.c_double_p), C.byref(lml)))
    return lat, lnw, lml.value


def importance_resampling(model, constraints, num_samples, num_ret_samples, seed=0, batch=0):
    """-> (all traces [L, n], resampled indices [n_ret], log-ML estimate)   (importance.rs:37-51, quirk Q10)"""
    a, p, n = _obs(constraints)
    lat = np.empty((model.num_latents, num_samples), dtype=np.float64)
    idx = np.empty(num_ret_samples, dtype=np.int64)
    lml = C.c_double()
    check(lib.mpl_importance_resampling(model._h, p, n, num_samples, num_ret_samples, seed, batch, lat.ctypes.data_as(_lib.c_double_p), idx.ctypes.data_as(_lib.c_i64_p), C.byref(lml)))
    return lat, idx, lml.value
