"""ctypes binding of the CPU oracle (oracle/libmodppl_oracle.so).  Test infrastructure: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libmodppl_oracle.so")


def build():
    src = os.path.join(ORACLE_DIR, "modppl_oracle.cpp")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


build()
L = C.CDLL(SO)
dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def _f(name, res, *args):
    fn = getattr(L, name)
    fn.restype = res
    fn.argtypes = list(args)
    return fn


_f("mo_logsumexp", C.c_double, dp, C.c_size_t)
_f("mo_normal_logpdf", C.c_double, C.c_double, C.c_double, C.c_double)
_f("mo_bernoulli_logpdf", C.c_double, C.c_int, C.c_double)
_f("mo_uniform_logpdf", C.c_double, C.c_double, C.c_double, C.c_double)
_f("mo_uniform_discrete_logpdf", C.c_double, C.c_int64, C.c_int64, C.c_int64)
_f("mo_geometric_logpdf", C.c_double, C.c_int64, C.c_double)
_f("mo_poisson_logpdf", C.c_double, C.c_int64, C.c_double)
_f("mo_beta_logpdf", C.c_double, C.c_double, C.c_double, C.c_double)
_f("mo_gamma_logpdf", C.c_double, C.c_double, C.c_double, C.c_double)
_f("mo_uniform2d_logpdf", C.c_double, C.c_double, C.c_double, dp)
_f("mo_mvnormal_logpdf", C.c_double, dp, dp, dp, C.c_int)
_f("mo_categorical_random", C.c_int64, dp, C.c_size_t, C.c_double)
_f("mo_categorical_logpdf", C.c_double, C.c_int64, dp, C.c_size_t)
_f("mo_cumsum_sequential", None, dp, C.c_size_t, dp)
_f("mo_resample_indices_faithful", C.c_int, dp, dp, C.c_size_t, C.c_size_t, C.c_int, i64p)
_f("mo_resample_indices", C.c_int, dp, dp, C.c_size_t, C.c_size_t, C.c_int, i64p)
_f("mo_philox4x32_10", None, u32p, u32p, u32p)
_f("mo_u01_f64", C.c_double, C.c_uint32, C.c_uint32)
_f("mo_u01_f32", C.c_float, C.c_uint32)
_f("mo_exp2_poly", C.c_float, C.c_float)
_f("mo_fixed_weight", C.c_uint64, C.c_float, C.c_int)
_f("mo_fixed_kbits", C.c_int, C.c_uint64)
_f("mo_set_threads", None, C.c_int)
_f("mo_get_threads", C.c_int)
_f("mo_fixed_systematic", C.c_uint64, fp, C.c_size_t, C.c_uint64, i32p, dp)
_f("mo_fixed_multinomial", C.c_uint64, fp, C.c_size_t, C.c_uint64, C.c_uint32, i32p, dp)
_f("mo_nested_systematic", C.c_uint64, fp, C.c_size_t, C.c_uint64, i32p, dp)
_f("mo_ps_new", C.c_void_p, C.c_char_p, dp, C.c_size_t, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64)
_f("mo_ps_free", None, C.c_void_p)
_f("mo_ps_init_step", C.c_int, C.c_void_p, dp, C.c_size_t)
_f("mo_ps_step", C.c_int, C.c_void_p, dp, C.c_size_t)
_f("mo_ps_effective_sample_size", C.c_double, C.c_void_p, C.c_int)
_f("mo_ps_resample", C.c_double, C.c_void_p, C.c_int)
_f("mo_ps_resample_faithful_cost", C.c_double, C.c_void_p)
_f("mo_ps_log_marginal_likelihood_estimate", C.c_double, C.c_void_p)
_f("mo_ps_state_dim", C.c_int, C.c_void_p)
_f("mo_ps_read_state", None, C.c_void_p, dp)
_f("mo_ps_read_log_weights", None, C.c_void_p, dp)
_f("mo_ps_read_parents", None, C.c_void_p, i64p)
_f("mo_ps_write_state", None, C.c_void_p, dp)
_f("mo_ps_write_log_weights", None, C.c_void_p, dp)
_f("mo_importance_sampling", C.c_int, C.c_char_p, dp, C.c_size_t, dp, C.c_size_t, C.c_uint32, C.c_uint64, C.c_uint64, dp, dp, dp)
_f("mo_importance_resampling_indices", C.c_int, dp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, i64p)
_f("mo_is_num_latents", C.c_int, C.c_char_p)
_f("mo_chains_new", C.c_void_p, C.c_char_p, dp, C.c_size_t, dp, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64)
_f("mo_chains_free", None, C.c_void_p)
_f("mo_chains_move", C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_uint32, C.c_uint32, u64p)
_f("mo_chains_num_slots", C.c_int, C.c_void_p)
_f("mo_chains_read", None, C.c_void_p, dp)
_f("mo_chains_write", None, C.c_void_p, dp)
_f("mo_hier_mh_alpha", C.c_double, dp, dp, C.c_size_t, dp, dp, C.c_int, C.c_double, dp)
_f("mo_hier_logjp", C.c_double, dp, dp, C.c_size_t, dp)
_f("mo_hmm_forward", C.c_double, dp, dp, dp, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int)
_f("mo_kalman_lml_lgssm4", C.c_double, C.c_double, C.c_double, C.c_double, dp, C.c_int)
_f("mo_line_model_lml", C.c_double, dp, dp, C.c_int)
_f("mo_hier_model_lml", C.c_double, dp, dp, C.c_int, dp)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def logsumexp(xs):
    a, p = _d(xs)
    return L.mo_logsumexp(p, a.size)


def normal_logpdf(x, mu, sd):
    return L.mo_normal_logpdf(x, mu, sd)


def logpdf(dist, x, params):
    """The remaining built-in distributions by name (same names as modppl_b200.parity.logpdf)."""
    if dist == "uniform_discrete":
        return L.mo_uniform_discrete_logpdf(int(x), int(params[0]), int(params[1]))
    if dist == "geometric":
        return L.mo_geometric_logpdf(int(x), float(params[0]))
    if dist == "poisson":
        return L.mo_poisson_logpdf(int(x), float(params[0]))
    if dist == "beta":
        return L.mo_beta_logpdf(float(x), float(params[0]), float(params[1]))
    if dist == "gamma":
        return L.mo_gamma_logpdf(float(x), float(params[0]), float(params[1]))
    if dist == "categorical":
        pr = np.ascontiguousarray(params, dtype=np.float64)
        return L.mo_categorical_logpdf(int(x), pr.ctypes.data_as(dp), pr.size)
    raise ValueError(dist)


def mvnormal_logpdf(x, mu, cov):
    xa, xp = _d(x)
    ma, mp = _d(mu)
    ca, cp = _d(np.asarray(cov).ravel())
    return L.mo_mvnormal_logpdf(xp, mp, cp, xa.size)


def uniform2d_logpdf(x, y, bounds):
    ba, bp = _d(bounds)
    return L.mo_uniform2d_logpdf(x, y, bp)


def categorical_random(probs, u):
    a, p = _d(probs)
    return L.mo_categorical_random(p, a.size, u)


def cumsum_sequential(probs):
    a, p = _d(probs)
    out = np.empty_like(a)
    L.mo_cumsum_sequential(p, a.size, out.ctypes.data_as(dp))
    return out


def resample_indices(probs, uniforms, n_draws=None, scheme=0, faithful=False):
    a, p = _d(probs)
    u, up = _d(np.atleast_1d(uniforms))
    n_draws = int(u.size if n_draws is None else n_draws)
    out = np.empty(n_draws, dtype=np.int64)
    fn = L.mo_resample_indices_faithful if faithful else L.mo_resample_indices
    rc = fn(p, up, a.size, n_draws, scheme, out.ctypes.data_as(i64p))
    assert rc == 0
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    L.mo_philox4x32_10(c, k, o)
    return list(o)


P_MODEL, P_RESAMPLE_U, P_RESAMPLE_OFFSET, P_IS, P_MH, P_IS_RESAMPLE, P_MH_INIT = range(7)


def stream_block(seed, ident, t, purpose, blk=0):
    return philox([ident & 0xFFFFFFFF, ident >> 32, t, (purpose << 24) | blk], [seed & 0xFFFFFFFF, seed >> 32])


def resample_offset_word(seed, t):
    x = stream_block(seed, 0, t, P_RESAMPLE_OFFSET)
    return (x[0] << 32) | x[1]


def fixed_systematic(lw, rand_word):
    a = np.ascontiguousarray(lw, dtype=np.float32)
    anc = np.empty(a.size, dtype=np.int32)
    lse = C.c_double()
    W = L.mo_fixed_systematic(a.ctypes.data_as(fp), a.size, rand_word, anc.ctypes.data_as(i32p), C.byref(lse))
    return anc, lse.value, W


def nested_systematic(lw, rand_word):
    a = np.ascontiguousarray(lw, dtype=np.float32)
    anc = np.empty(a.size, dtype=np.int32)
    lse = C.c_double()
    W = L.mo_nested_systematic(a.ctypes.data_as(fp), a.size, rand_word, anc.ctypes.data_as(i32p), C.byref(lse))
    return anc, lse.value, W


def fixed_multinomial(lw, seed, t):
    a = np.ascontiguousarray(lw, dtype=np.float32)
    anc = np.empty(a.size, dtype=np.int32)
    lse = C.c_double()
    W = L.mo_fixed_multinomial(a.ctypes.data_as(fp), a.size, seed, t, anc.ctypes.data_as(i32p), C.byref(lse))
    return anc, lse.value, W


class OraclePS:
    """oracle particle system (restates inference/particle_filter.rs)"""

    def __init__(self, model, params, n, dtype="f64", seed=0, gid_offset=0, n_global=0):
        pa, pp = _d(params)
        self.n = n
        self._h = L.mo_ps_new(model.encode(), pp, pa.size, n, 1 if dtype == "f64" else 0, seed, gid_offset, n_global)
        assert self._h, "unknown oracle model"
        self.D = L.mo_ps_state_dim(self._h)

    def init_step(self, obs):
        a, p = _d(obs)
        assert L.mo_ps_init_step(self._h, p, a.size) == 0

    def step(self, obs):
        a, p = _d(obs)
        assert L.mo_ps_step(self._h, p, a.size) == 0
        return self

    def effective_sample_size(self, stale=True):
        return L.mo_ps_effective_sample_size(self._h, int(stale))

    def resample(self, scheme=0):
        return L.mo_ps_resample(self._h, scheme)

    def resample_faithful_cost(self):
        return L.mo_ps_resample_faithful_cost(self._h)

    def log_marginal_likelihood_estimate(self):
        return L.mo_ps_log_marginal_likelihood_estimate(self._h)

    @property
    def traces(self):
        out = np.empty((self.D, self.n))
        L.mo_ps_read_state(self._h, out.ctypes.data_as(dp))
        return out

    @property
    def log_weights(self):
        out = np.empty(self.n)
        L.mo_ps_read_log_weights(self._h, out.ctypes.data_as(dp))
        return out

    @property
    def parents(self):
        out = np.empty(self.n, dtype=np.int64)
        L.mo_ps_read_parents(self._h, out.ctypes.data_as(i64p))
        return out

    def write_state(self, st):
        a, p = _d(np.asarray(st).reshape(self.D, self.n))
        L.mo_ps_write_state(self._h, p)

    def write_log_weights(self, lw):
        a, p = _d(lw)
        L.mo_ps_write_log_weights(self._h, p)

    def __del__(self):
        if getattr(self, "_h", None):
            L.mo_ps_free(self._h)
            self._h = None


def importance_sampling(model, args, obs, n, seed=0, batch=0):
    aa, ap = _d(args)
    oa, op = _d(obs)
    nl = L.mo_is_num_latents(model.encode())
    lat = np.empty((nl, n))
    lnw = np.empty(n)
    lml = C.c_double()
    rc = L.mo_importance_sampling(model.encode(), ap, aa.size, op, oa.size, n, seed, batch, lat.ctypes.data_as(dp), lnw.ctypes.data_as(dp), C.byref(lml))
    assert rc == 0, rc
    return lat, lnw, lml.value


def importance_resampling_indices(lnw, n_ret, seed=0, batch=0):
    a, p = _d(lnw)
    idx = np.empty(n_ret, dtype=np.int64)
    assert L.mo_importance_resampling_indices(p, a.size, n_ret, seed, batch, idx.ctypes.data_as(i64p)) == 0
    return idx


class OracleChains:
    def __init__(self, model, args, obs, n, seed=0, offset=0):
        aa, ap = _d(args)
        oa, op = _d(obs)
        self.n = n
        self._h = L.mo_chains_new(model.encode(), ap, aa.size, op, oa.size, n, seed, offset)
        assert self._h
        self.slots = L.mo_chains_num_slots(self._h)

    def move(self, move, parg=1.0, mask=0, n_steps=1):
        acc = C.c_uint64()
        rc = L.mo_chains_move(self._h, move, parg, mask, n_steps, C.byref(acc))
        assert rc == 0, rc
        return acc.value

    def read(self):
        out = np.empty((self.slots, self.n))
        L.mo_chains_read(self._h, out.ctypes.data_as(dp))
        return out

    def write(self, st):
        a, p = _d(np.asarray(st).reshape(self.slots, self.n))
        L.mo_chains_write(self._h, p)

    def __del__(self):
        if getattr(self, "_h", None):
            L.mo_chains_free(self._h)
            self._h = None


def hier_logjp(xs, ys, st):
    xa, xp = _d(xs)
    ya, yp = _d(ys)
    sa, sp = _d(st)
    return L.mo_hier_logjp(xp, yp, xa.size, sp)


def hier_mh_alpha(xs, ys, cur, prop, move, parg):
    xa, xp = _d(xs)
    ya, yp = _d(ys)
    ca, cp = _d(cur)
    pa, pp = _d(prop)
    out = np.empty(3)
    alpha = L.mo_hier_mh_alpha(xp, yp, xa.size, cp, pp, move, parg, out.ctypes.data_as(dp))
    return alpha, out


def hmm_forward(prior, emission, transition, obs):
    """emission[s][o], transition[from][to] (row-stochastic, as written in tests/particle_filter.rs:41-50)"""
    pr, prp = _d(prior)
    em = np.asarray(emission, float)
    tr = np.asarray(transition, float)
    K, M = em.shape
    ea, ep = _d(em.T.ravel())
    ta, tp = _d(tr.T.ravel())
    o = (C.c_int * len(obs))(*obs)
    return L.mo_hmm_forward(prp, ep, tp, K, M, o, len(obs))


def kalman_lml_lgssm4(q, r, x0, ys):
    a, p = _d(np.asarray(ys).reshape(-1, 2))
    return L.mo_kalman_lml_lgssm4(q, r, x0, p, a.shape[0])


def line_model_lml(xs, ys):
    xa, xp = _d(xs)
    ya, yp = _d(ys)
    return L.mo_line_model_lml(xp, yp, xa.size)


def hier_model_lml(xs, ys):
    xa, xp = _d(xs)
    ya, yp = _d(ys)
    pl = C.c_double()
    v = L.mo_hier_model_lml(xp, yp, xa.size, C.byref(pl))
    return v, pl.value
