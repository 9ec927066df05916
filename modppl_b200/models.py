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
