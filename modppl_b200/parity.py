"""Parity hooks: injected inputs, no RNG (include/modppl_b200.h, 'parity hooks')."""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import lib, check


def resample_indices(probs, uniforms, n_draws=None, scheme=0):
    """categorical.rs:22-32 / particle_filter.rs:37-41 on the device, sequential-f64-cumsum semantics."""
    p = np.ascontiguousarray(probs, dtype=np.float64)
    u = np.ascontiguousarray(np.atleast_1d(uniforms), dtype=np.float64)
    n_draws = int(u.size if n_draws is None else n_draws)
    out = np.empty(n_draws, dtype=np.int64)
    check(lib.mpl_resample_indices(p.ctypes.data_as(_lib.c_double_p), u.ctypes.data_as(_lib.c_double_p), p.size, n_draws, scheme, out.ctypes.data_as(_lib.c_i64_p)))
    return out


def cumsum_sequential(probs):
    p = np.ascontiguousarray(probs, dtype=np.float64)
    out = np.empty_like(p)
    check(lib.mpl_cumsum_sequential(p.ctypes.data_as(_lib.c_double_p), p.size, out.ctypes.data_as(_lib.c_double_p)))
    return out


def logsumexp_stats(lw):
    a = np.ascontiguousarray(lw)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    lse, ess, mx = C.c_double(), C.c_double(), C.c_double()
    check(lib.mpl_logsumexp_stats(a.ctypes.data_as(C.c_void_p), a.size, 1 if a.dtype == np.float64 else 0, C.byref(lse), C.byref(ess), C.byref(mx)))
    return lse.value, ess.value, mx.value


def fixed_resample(lw, scheme=2, seed=0, t=0):
    a = np.ascontiguousarray(lw, dtype=np.float32)
    anc = np.empty(a.size, dtype=np.int32)
    lse, W = C.c_double(), C.c_uint64()
    check(lib.mpl_fixed_resample(a.ctypes.data_as(_lib.c_float_p), a.size, scheme, seed, t, anc.ctypes.data_as(_lib.c_i32_p), C.byref(lse), C.byref(W)))
    return anc, lse.value, W.value


def logpdf(dist, x, params):
    xx = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    if xx.size < 2:
        xx = np.concatenate([xx, [0.0]])
    pp = np.ascontiguousarray(np.atleast_1d(params), dtype=np.float64)
    out = C.c_double()
    check(lib.mpl_logpdf(dist.encode(), xx.ctypes.data_as(_lib.c_double_p), pp.ctypes.data_as(_lib.c_double_p), pp.size, C.byref(out)))
    return out.value


def virtual_shards(model, n_global, world, obs, dtype="f32", seed=0, scheme=2):
    """Shards emulated on one GPU (include/modppl_b200.h: mpl_test_virtual_shards) -> (state [D, N], log-weights [N], log-ML)."""
    ys = np.ascontiguousarray(np.asarray(obs, dtype=np.float64))
    ys = ys.reshape(ys.shape[0], -1)
    D = model.state_dim
    st = np.empty((D, n_global), dtype=np.float64)
    lw = np.empty(n_global, dtype=np.float64)
    lml, ms = C.c_double(), C.c_double()
    check(lib.mpl_test_virtual_shards_scheme(model._h, n_global, world, 1 if dtype == "f64" else 0, seed, int(scheme), ys.ctypes.data_as(_lib.c_double_p), ys.shape[0], ys.shape[1],
                                      st.ctypes.data_as(_lib.c_double_p), lw.ctypes.data_as(_lib.c_double_p), C.byref(lml), C.byref(ms)))
    virtual_shards.last_loop_ms = ms.value
    return st, lw, lml.value


def set_inline_level1(mode):
    """test hook (include/modppl_b200.h: mpl_test_set_inline_level1): -1 by shard size, 0 plan pass, 1 inside the expansion"""
    check(lib.mpl_test_set_inline_level1(int(mode)))
