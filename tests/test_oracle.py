"""CPU checks of the oracle's own internal consistency (the restatement has to agree with itself and with exact
ground truths before it is allowed to judge the CUDA path)."""
import math

import numpy as np
import pytest

import oracle_lib as O


def test_logsumexp_matches_definition_and_guards():     # lib.rs:34-45
    rng = np.random.default_rng(1)
    x = rng.normal(size=1000) * 30
    ref = float(np.log(np.sum(np.exp(x - x.max()))) + x.max())
    assert abs(O.logsumexp(x) - ref) < 1e-12
    assert O.logsumexp([-math.inf, -math.inf]) == -math.inf
    assert O.logsumexp([0.0] * 8) == math.log(8.0)


def test_cumsum_is_numpy_sequential():
    rng = np.random.default_rng(2)
    p = rng.random(200000)
    p /= p.sum()
    assert np.array_equal(O.cumsum_sequential(p), np.cumsum(p))   # numpy's 1-D float64 cumsum is the same sequential loop


@pytest.mark.parametrize("scheme", [0, 1])
def test_faithful_equals_fast(scheme):
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 1000):
        p = rng.random(n) ** 4
        p[rng.random(n) < 0.2] = 0.0
        if p.sum() == 0:
            p[0] = 1.0
        p /= p.sum()
        u = rng.random(n if scheme == 0 else 1)
        a = O.resample_indices(p, u, n_draws=n, scheme=scheme, faithful=True)
        b = O.resample_indices(p, u, n_draws=n, scheme=scheme, faithful=False)
        assert np.array_equal(a, b)
        assert a.min() >= 0 and a.max() < n


def test_resample_edge_cases():
    p = np.array([0.0, 0.0, 1.0, 0.0])
    assert list(O.resample_indices(p, [0.0, 0.5, 1.0, 0.9999999])) == [0, 2, 2, 2]   # u == 0 clamps to 0 (quirk Q2)
    assert list(O.resample_indices(p, [1.5])) == [3]                                  # past the total: clamp to n-1


def test_philox_known_answers():                   # Random123 kat_vectors, philox4x32-10
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_fixed_weight_is_exp():
    for d in [0.0, -0.1, -1.0, -5.5, -20.0, -27.0]:
        q = O.L.mo_fixed_weight(d, 38)
        assert abs(q / 2.0**38 - math.exp(d)) <= 3e-6 * math.exp(d) + 2.0**-38
    assert O.L.mo_fixed_weight(0.0, 38) == 1 << 38
    assert O.L.mo_fixed_weight(-math.inf, 38) == 0
    assert O.L.mo_fixed_weight(float("nan"), 38) == 0
    assert O.L.mo_fixed_kbits(1 << 24) == 38 and O.L.mo_fixed_kbits(1000) == 40 and O.L.mo_fixed_kbits(1 << 26) == 36


def test_fixed_systematic_properties():
    rng = np.random.default_rng(4)
    n = 5000
    lw = (rng.normal(size=n) * 3).astype(np.float32)
    anc, lse, W = O.fixed_systematic(lw, 0x123456789ABCDEF0)
    assert np.all(np.diff(anc) >= 0) and anc.min() >= 0 and anc.max() < n
    w = np.exp(lw.astype(np.float64) - lw.max())
    expect = n * w / w.sum()
    counts = np.bincount(anc, minlength=n)
    assert np.all(np.abs(counts - expect) < 1.0 + 1e-3 * expect)      # systematic: offspring within 1 of N w_i
    assert abs(lse - O.logsumexp(lw.astype(np.float64))) < 1e-5
    # degenerate: one particle holds everything
    lw2 = np.full(n, -1000.0, dtype=np.float32)
    lw2[1234] = 0.0
    anc2, _, _ = O.fixed_systematic(lw2, 42)
    assert np.all(anc2 == 1234)


def test_fixed_multinomial_distribution():
    rng = np.random.default_rng(5)
    n = 20000
    lw = np.log(np.array([0.1, 0.3, 0.2, 0.1, 0.05, 0.25]))[rng.integers(0, 6, size=n)].astype(np.float32)
    anc, lse, W = O.fixed_multinomial(lw, 7, 3)
    w = np.exp(lw.astype(np.float64))
    # offspring mass of each weight class within 5 sigma
    for v in np.unique(lw):
        sel = lw == v
        p = w[sel].sum() / w.sum()
        got = np.isin(anc, np.nonzero(sel)[0]).mean()
        assert abs(got - p) < 5 * math.sqrt(p * (1 - p) / n)


def lgssm_data(T, seed=4, q=0.1, r=0.5, x0=1.0):
    rng = np.random.default_rng(seed)
    A = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], float)
    x = rng.normal(size=4) * x0
    ys = []
    for t in range(T):
        if t > 0:
            x = A @ x + q * rng.normal(size=4)
        ys.append(x[:2] + r * rng.normal(size=2))
    return np.array(ys)


def kalman_numpy(ys, q=0.1, r=0.5, x0=1.0):
    A = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], float)
    H = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], float)
    m, P, lml = np.zeros(4), np.eye(4) * x0**2, 0.0
    for t, y in enumerate(ys):
        if t > 0:
            m, P = A @ m, A @ P @ A.T + q * q * np.eye(4)
        S = H @ P @ H.T + r * r * np.eye(2)
        v = y - H @ m
        lml += -0.5 * (2 * math.log(2 * math.pi) + math.log(np.linalg.det(S)) + v @ np.linalg.solve(S, v))
        K = P @ H.T @ np.linalg.inv(S)
        m, P = m + K @ v, P - K @ H @ P
    return lml


def test_kalman_truth_matches_numpy():
    ys = lgssm_data(50)
    assert abs(O.kalman_lml_lgssm4(0.1, 0.5, 1.0, ys) - kalman_numpy(ys)) < 1e-9


@pytest.mark.parametrize("dtype,scheme", [("f64", 0), ("f64", 1), ("f32", 2), ("f32", 3)])
def test_oracle_pf_vs_kalman(dtype, scheme):
    # bootstrap filter on the config-4 model: the log-ML estimate has std ~0.15-0.25 at N = 5e4, T = 10 (measured over
    # seeds; the velocity prior is vague).  Six seeds: each within 4 sigma, the mean within 3 sigma of the mean.
    T, N = 10, 50000
    ys = lgssm_data(T)
    truth = O.kalman_lml_lgssm4(0.1, 0.5, 1.0, ys)
    errs = []
    for seed in range(6):
        ps = O.OraclePS("lgssm4", [0.1, 0.5, 1.0], N, dtype=dtype, seed=100 + seed)
        ps.init_step(ys[0])
        ps.resample(scheme)
        for y in ys[1:]:
            ps.step(y)
            ps.resample(scheme)
        errs.append(ps.log_marginal_likelihood_estimate() - truth)
    assert max(abs(e) for e in errs) < 1.0
    assert abs(np.mean(errs)) < 0.3


def test_stale_ess_quirk():                       # particle_filter.rs:31,98-100 (quirk Q1)
    ps = O.OraclePS("sv", [-1.024, 0.9702, 0.178], 1000, seed=5)
    ps.init_step([0.3])
    assert abs(ps.effective_sample_size(True) - 1.0 / 1000) < 1e-15   # all-zero stale buffer: exp(-ln N)
    fresh = ps.effective_sample_size(False)
    ps.resample(0)
    assert abs(ps.effective_sample_size(True) - fresh) < 1e-9        # stale value = ESS at the last normalisation
    assert abs(ps.effective_sample_size(False) - 1000) < 1e-9        # weights were reset to zero


def test_importance_sampling_lml_vs_closed_form():
    xs = np.arange(-5.0, 6.0)
    rng = np.random.default_rng(6)
    ys = 0.5 * xs - 1.0 + 0.1 * rng.normal(size=11)
    truth = O.line_model_lml(xs, ys)
    lat, lnw, lml = O.importance_sampling("line", xs, ys, 400000, seed=1)
    assert abs(np.exp(lnw).sum() - 1.0) < 1e-9
    assert abs(lml - truth) < 0.5          # prior-as-proposal is a poor estimator here (tests/importance.rs:90-92 says as much)
    post_slope = np.sum(np.exp(lnw) * lat[0])
    assert abs(post_slope - 0.5) < 0.05


def test_hier_mh_alpha_pieces():
    xs = np.arange(-5.0, 6.0)
    ys = 0.3 + 0.4 * xs + 0.5 * xs * xs
    cur = [0.0, 0.2, 0.4, 0.45]        # quadratic
    prop = [1.0, 0.21, 0.41, 0.0]      # add/remove proposes linear
    alpha, (w, fwd, bwd) = O.hier_mh_alpha(xs, ys, cur, prop, 1, 0.025)
    n = O.normal_logpdf
    assert abs(w - (O.hier_logjp(xs, ys, prop) - O.hier_logjp(xs, ys, cur))) < 1e-9
    assert abs(fwd - (n(0.21, 0.2, 0.025) + n(0.41, 0.4, 0.025) + math.log(0.5))) < 1e-12
    # backward move re-proposes c from prev_c' = 0 because the new trace has no c  (hierarchical.rs:54-58)
    assert abs(bwd - (n(0.2, 0.21, 0.025) + n(0.4, 0.41, 0.025) + math.log(0.5) + n(0.45, 0.0, 0.025))) < 1e-12
    assert abs(alpha - (w - fwd + bwd)) < 1e-12
    # drift: symmetric proposal => forward == backward bit for bit
    alpha, (w, fwd, bwd) = O.hier_mh_alpha(xs, ys, cur, [0.0, 0.25, 0.38, 0.5], 0, 0.1)
    assert fwd == bwd


def test_mh_posterior_moves_toward_truth():
    xs = np.arange(-5.0, 6.0)
    rng = np.random.default_rng(7)
    ys = 0.3 + 0.4 * xs + 0.5 * xs * xs + 0.1 * rng.normal(size=11)
    ch = O.OracleChains("hierarchical", xs, ys, 64, seed=2)
    for _ in range(60):
        ch.move(1, 0.025, n_steps=1)
        ch.move(0, 0.1, n_steps=3)
        ch.move(0, 0.01, n_steps=10)
    st = ch.read()
    # cached logjp stays consistent with the state
    for i in range(0, 64, 7):
        assert abs(st[4, i] - O.hier_logjp(xs, ys, st[:4, i])) < 1e-9
    assert st[4].mean() > -2000        # far above the prior draw's typical logjp (~ -1e5)


def test_oracle_particle_filter_is_independent_of_host_threads():
    # bench.py times the port on one and on all host threads: same numbers either way (counter-based RNG per particle,
    # every integer sum taken by one thread per chunk)
    ys = np.random.default_rng(0).normal(size=(4, 2))
    out = []
    for th in (1, 3):
        O.L.mo_set_threads(th)
        ps = O.OraclePS("lgssm4", [0.1, 0.5, 1.0], 40000, dtype="f32", seed=3)
        ps.init_step(ys[0]); ps.resample(4)
        for t in range(1, 4):
            ps.step(ys[t]); ps.resample(4 if t % 2 else 2)
        out.append((ps.log_marginal_likelihood_estimate(), ps.traces.copy(), ps.parents.copy()))
    O.L.mo_set_threads(1)
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


@pytest.mark.parametrize("n,reps", [(1000, 3000), (2 * (1 << 17) + 5000, 250)])
def test_nested_systematic_is_unbiased(n, reps):
    # three nested exact systematic levels: E[#offspring of i] = N w_i / sum w (up to the 2^-22 quantisation), checked by
    # averaging the oracle over many random offset words (the larger case spans three sections on different scales)
    rng = np.random.default_rng(7)
    lw = (rng.normal(size=n) * 2.0).astype(np.float32)
    lw[::7] -= 30.0                                   # a few particles far below their chunk's maximum
    lw -= (np.arange(n) // (1 << 17)).astype(np.float32) * np.float32(3.3)
    w = np.exp(lw.astype(np.float64)); expect = n * w / w.sum()
    mean = np.zeros(n)
    for r in range(reps):
        anc, _, _ = O.nested_systematic(lw, int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2)))
        mean += np.bincount(anc, minlength=n)
    mean /= reps
    # systematic counts differ from their mean by less than 1 per level: variance of the average <= 3 / reps
    assert np.max(np.abs(mean - expect)) < 6.0 * math.sqrt(3.0 / reps)
    assert abs(mean.sum() - n) < 1e-9
