"""Model registry front-end: the restricted vectorisable models the engine has device functors for.

Each helper mirrors a fixture of the reference (citations relative to /root/reference/modppl/):
  spiral_model        tests/dyngenfns/unfold.rs:14-33     (DynUnfold kernel, config 1)
  hmm                 tests/hmm/model.rs:24-81            (hand-coded GenFn, the particle-filter accuracy test)
  line_model          tests/dyngenfns/simple.rs:10-23
  hierarchical_model  tests/dyngenfns/hierarchical.rs:32-46
  pointed_model       tests/pointed_model/model.rs / tests/dyngenfns/simple.rs:27-31
lgssm4 and stochastic_volatility are the config-4/5 workloads of BASELINE.json (not in the reference).
"""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import lib, check_handle


class Model:
    def __init__(self, name, params):
        self.name = name
        self.params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
        self._h = check_handle(lib.mpl_model_create(name.encode(), self.params.ctypes.data_as(_lib.c_double_p), self.params.size))

    @property
    def state_dim(self):
        return lib.mpl_model_state_dim(self._h)

    @property
    def obs_dim(self):
        return lib.mpl_model_obs_dim(self._h)

    @property
    def num_latents(self):
        return lib.mpl_model_num_latents(self._h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.mpl_model_destroy(h)


class CompiledModel(Model):
    """An Unfold model written as a declarative spec (dict or JSON string; include/modppl_b200.h: mpl_model_compile) -- the
    stand-in for authoring with `dyngen!` (modppl-macros/src/lib.rs:20-114).  NVRTC compiles it against the library's own kernel
    headers on first use; `compile(dtype)` does it right away and returns the compiler's log."""

    def __init__(self, spec):
        import json
        self.spec = spec if isinstance(spec, str) else json.dumps(spec)
        self.name = "compiled"
        self.params = np.zeros(0)
        self._h = check_handle(lib.mpl_model_compile(self.spec.encode()))

    def compile(self, dtype="f32"):
        log = C.create_string_buffer(1 << 16)
        rc = lib.mpl_model_jit_compile(self._h, 0 if dtype in ("f32", 0) else 1, log, len(log))
        if rc != 0:
            raise _lib.MplError(rc, _lib.last_error())
        return log.value.decode()

    def source(self, dtype="f32"):
        return lib.mpl_model_jit_source(self._h, 0 if dtype in ("f32", 0) else 1).decode()


def compile_model(spec):
    return CompiledModel(spec)


def lgssm4_spec(q_std=0.1, r_std=0.5, x0_std=1.0):
    """config 4's model written in the spec language; compiles to the arithmetic of the built-in functor"""
    return {"name": "lgssm4_spec", "state_dim": 4, "obs_dim": 2, "params": {"q": q_std, "r": r_std, "x0": x0_std},
            "init": [{"dist": "normal", "args": ["0", "x0"]}] * 4,
            "step": [{"dist": "normal", "args": ["x[0] + x[2]", "q"]}, {"dist": "normal", "args": ["x[1] + x[3]", "q"]},
                     {"dist": "normal", "args": ["x[2]", "q"]}, {"dist": "normal", "args": ["x[3]", "q"]}],
            "observe": [{"dist": "normal", "value": "y[0]", "args": ["x[0]", "r"]}, {"dist": "normal", "value": "y[1]", "args": ["x[1]", "r"]}]}


def spiral_spec(dr_std=0.1, dtheta_mean=0.4, dtheta_std=0.2, obs_var=0.001):
    """tests/dyngenfns/unfold.rs:14-33 written in the spec language"""
    return {"name": "spiral_spec", "state_dim": 2, "obs_dim": 2,
            "params": {"dr_std": dr_std, "dth_mean": dtheta_mean, "dth_std": dtheta_std, "ov": obs_var, "two_pi": 2.0 * np.pi},
            "init": [{"dist": "uniform", "args": ["0", "1"]}, {"dist": "uniform", "args": ["0", "two_pi"]}],
            "step": [{"dist": "normal", "args": ["0", "dr_std"], "add_to": "x[0]"}, {"dist": "normal", "args": ["dth_mean", "dth_std"], "add_to": "x[1]"}],
            "observe": [{"dist": "mvnormal2", "value": ["y[0]", "y[1]"], "mean": ["x[0] * cos(x[1])", "x[0] * sin(x[1])"], "cov": ["ov", "0", "0", "ov"]}]}


def lgssm4(q_std=0.1, r_std=0.5, x0_std=1.0):
    return Model("lgssm4", [q_std, r_std, x0_std])


def spiral_model(dr_std=0.1, dtheta_mean=0.4, dtheta_std=0.2, obs_var=0.001):
    return Model("spiral", [dr_std, dtheta_mean, dtheta_std, obs_var])


def stochastic_volatility(mu=-1.024, phi=0.9702, sigma=0.178):
    return Model("sv", [mu, phi, sigma])


def hmm(prior, emission, transition):
    """prior[K]; emission[s][o] = P(o | s); transition[from][to] -- the row-stochastic matrices the reference test writes
    down before transposing them (tests/particle_filter.rs:41-50)."""
    prior = np.asarray(prior, float)
    em = np.asarray(emission, float)
    tr = np.asarray(transition, float)
    K, M = em.shape
    params = np.concatenate([[K, M], prior, em.T.ravel(), tr.T.ravel()])   # emission[o*K+s], transition[to*K+from]
    return Model("hmm", params)


def line_model(xs):
    return Model("line", xs)


def hierarchical_model(xs):
    return Model("hierarchical", xs)


def pointed_model(bounds, cov):
    """bounds = (xmin, xmax, ymin, ymax); cov 2x2."""
    return Model("pointed", np.concatenate([np.asarray(bounds, float).ravel(), np.asarray(cov, float).ravel()]))
