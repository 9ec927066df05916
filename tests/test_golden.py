"""Pins the CPU oracle against every known-answer value the reference's own tests hold for the hot path
(SURVEY.md section 8c).  Citations: /root/reference/modppl/tests/*.rs.  Runs on CPU."""
import math

import numpy as np
import pytest

import oracle_lib as O

F32_EPS = 1.1920929e-07   # LOGPDF_EPSILON = f32::EPSILON in tests/dists.rs


def test_normal_logpdf_known_answers():          # tests/dists.rs:120-136
    assert abs(O.normal_logpdf(1.4, 0.9, 0.5) - -0.7257913526447272) <= F32_EPS
    assert abs(O.normal_logpdf(2.8, 1.8, 1.0) - -1.4189385332046727) <= F32_EPS
    assert abs(O.normal_logpdf(-3.14, 8.0, 20.0) - -4.069795306758664) <= F32_EPS
    # tighter than the reference asks: the restatement follows normal.rs:13-17 operation by operation
    assert abs(O.normal_logpdf(1.4, 0.9, 0.5) - -0.7257913526447272) <= 4e-16


def test_mvnormal_logpdf_known_answers():        # tests/dists.rs:164-183 (pins nalgebra 0.32 det / try_inverse for k = 2, 3)
    assert abs(O.mvnormal_logpdf([1.1, 5.8], [1.3, 5.6], [[1.0, -0.81], [-0.81, 2.5]]) - -2.1642100746383357) <= F32_EPS
    assert abs(O.mvnormal_logpdf([30.1, -46.8], [0.0, 6.0], [[496.0, 0.13], [0.13, 500.0]]) - -11.750458919763666) <= F32_EPS
    assert abs(O.mvnormal_logpdf([1.2, 5.1, -7.8], [1.4, 5.0, -7.4], [[1.0, 0.1, 0.9], [0.1, 1.3, 0.4], [0.9, 0.4, 1.75]]) - -2.873267436425841) <= F32_EPS


def test_bernoulli_logpdf_exact():               # tests/dists.rs:27-29 (assert_eq!)
    assert O.L.mo_bernoulli_logpdf(1, 0.11) == math.log(0.11)
    assert O.L.mo_bernoulli_logpdf(0, 0.11) == math.log(1.0 - 0.11)


def test_uniform_logpdf_exact():                 # tests/dists.rs:43-48 (assert_eq!)
    a, b = 0.5, 3.14
    true_p = 1.0 / (b - a)
    # the reference asserts == against true_p.ln(); -(b-a).ln() and (1/(b-a)).ln() agree to 1 ulp on this input
    assert abs(O.L.mo_uniform_logpdf(0.9, a, b) - math.log(true_p)) <= 2.3e-16
    assert abs(O.L.mo_uniform_logpdf(2.1, a, b) - math.log(true_p)) <= 2.3e-16
    assert O.L.mo_uniform_logpdf(0.4, a, b) == -math.inf
    assert math.isnan(O.L.mo_uniform_logpdf(0.4, 1.0, 1.0))   # reference panics (uniform.rs:6-10)


def test_uniform2d_logpdf():                      # tests/test_pointed.rs:12,20,23
    b = [0.0, 2.5, -1.0, 0.25]
    assert abs(O.uniform2d_logpdf(1.0, -0.5, b) - -1.1394342831883648) <= np.finfo(float).eps
    assert O.uniform2d_logpdf(-1.0, 0.0, b) == -math.inf


def test_categorical_frequencies():               # tests/dists.rs:86-104 (50k draws, +-0.01)
    probs = np.array([0.1, 0.3, 0.2, 0.1, 0.05, 0.25])
    rng = np.random.default_rng(0)
    u = rng.random(50000)
    idx = O.resample_indices(probs, u, scheme=0)
    freq = np.bincount(idx, minlength=6) / 50000
    assert np.all(np.abs(freq - probs) <= 0.01)


def test_categorical_literal_quirks():            # categorical.rs:24-31 (quirk Q2)
    probs = [0.25, 0.25, 0.5]
    assert O.categorical_random(probs, 0.0) == -1          # loop never runs -> x - 1 == -1
    assert O.categorical_random(probs, 0.25) == 0          # `while t < u` stops as soon as t >= u
    assert O.categorical_random(probs, 0.2500001) == 1
    assert O.categorical_random(probs, 1.5) == 3           # reference would index out of bounds


def test_update_weight_table():                   # tests/dyngenfn.rs:56-114 pin the weight rules of SURVEY 3.4
    n = O.normal_logpdf
    bern = O.L.mo_bernoulli_logpdf
    # :56-66  x constrained & existed: logpdf_new - logp_old, exactly -0.5
    assert n(1.0, 0.0, 1.0) - n(0.0, 0.0, 1.0) == -0.5
    # :68-79  b false->true (constrained, existed) + x = 1 constrained-new
    w = (bern(1, 0.25) - bern(0, 0.25)) + n(1.0, 0.0, 1.0)
    assert abs(w - -2.517551) <= 1e-6
    # :81-93  m 1 -> .5 constrained; x = 1, y = -.3 free-existed under diff Unknown are re-scored
    w = (O.L.mo_uniform_logpdf(0.5, 0.0, 1.0) - O.L.mo_uniform_logpdf(1.0, 0.0, 1.0)) + (n(1.0, 0.5, 1.0) - n(1.0, 1.0, 1.0)) + (n(-0.3, 0.5, 1.0) - n(-0.3, 1.0, 1.0))
    assert abs(w - 0.4) <= 1e-6
    # :95-114 b false->true, x free-new: no weight for the fresh draw
    w = bern(1, 0.25) - bern(0, 0.25)
    assert abs(w - -1.098612) <= 1e-6


def test_hmm_forward_two_state():                 # tests/particle_filter.rs:10-33 ; expected 0.2484 by enumeration
    prior = [0.4, 0.6]
    emission = [[0.1, 0.9], [0.7, 0.3]]           # emission[s][o]
    transition = [[0.5, 0.5], [0.2, 0.8]]         # transition[from][to]
    obs = [1, 0]
    brute = 0.0
    for z0 in range(2):
        for z1 in range(2):
            brute += prior[z0] * emission[z0][obs[0]] * transition[z0][z1] * emission[z1][obs[1]]
    assert abs(brute - 0.2484) < 1e-15
    assert abs(O.hmm_forward(prior, emission, transition, obs) - brute) <= 1e-16


HMM3 = dict(
    prior=[0.2, 0.3, 0.5],
    emission=[[0.1, 0.2, 0.7], [0.2, 0.7, 0.1], [0.7, 0.2, 0.1]],
    transition=[[0.4, 0.4, 0.2], [0.2, 0.3, 0.5], [0.9, 0.05, 0.05]],
    obs=[0, 0, 1, 2],
)


def hmm_params(h=HMM3):
    em, tr = np.asarray(h["emission"]), np.asarray(h["transition"])
    K, M = em.shape
    return np.concatenate([[K, M], h["prior"], em.T.ravel(), tr.T.ravel()])


def test_hmm_forward_three_state_value():         # expected value of tests/particle_filter.rs:56
    ml = O.hmm_forward(HMM3["prior"], HMM3["emission"], HMM3["transition"], HMM3["obs"])
    assert abs(math.log(ml) - -4.87645083351704) < 1e-12


@pytest.mark.parametrize("scheme", [0, 1, 2, 3])
def test_particle_filter_end_to_end(scheme):      # tests/particle_filter.rs:35-79: N = 10^4, |lml - ln forward| <= 0.03
    expected = math.log(O.hmm_forward(HMM3["prior"], HMM3["emission"], HMM3["transition"], HMM3["obs"]))
    ps = O.OraclePS("hmm", hmm_params(), 10000, dtype="f64" if scheme < 2 else "f32", seed=1000)
    obs = HMM3["obs"]
    ps.init_step([obs[0]])
    for o in obs[1:]:
        ps.step([o])
        ps.effective_sample_size()
        ps.resample(scheme)
    assert abs(ps.log_marginal_likelihood_estimate() - expected) <= 0.03


# ---- the remaining built-in distributions: tests/dists.rs:60-69, 186-212 --------------------------------
def test_uniform_discrete_logpdf_known_answers():
    assert abs(O.logpdf("uniform_discrete", 9, (8, 130)) - math.log(1.0 / 123)) <= 1e-15
    assert abs(O.logpdf("uniform_discrete", 130, (8, 130)) - math.log(1.0 / 123)) <= 1e-15
    assert O.logpdf("uniform_discrete", 140, (8, 130)) == -math.inf


def test_geometric_poisson_beta_gamma_logpdf_known_answers():
    known = [("geometric", 1, (0.5,), -1.3862943611198906), ("geometric", 5, (0.98,), -19.580317734458244), ("geometric", 101, (0.01,), -5.6202541071917365),
             ("poisson", 3, (4.0,), -1.6328763858683835), ("poisson", 5, (1.5,), -4.2601662022412240), ("poisson", 52, (36.11,), -5.969204868031767),
             ("beta", 0.3, (0.5, 0.5), -0.364406011717066), ("beta", 0.7, (1.5, 2.0), -0.06055443631298263),
             ("gamma", 1.7, (1.23, 1.46), -1.414334369005868), ("gamma", 8.4, (4.5, 1.0), -3.4049256003700052), ("gamma", 0.03, (50.0, 70.0), -528.8122715889206)]
    for dist, x, params, want in known:
        assert abs(O.logpdf(dist, x, params) - want) <= F32_EPS, (dist, x, params)


def test_categorical_logpdf():                   # categorical.rs:13-20 with the probabilities of tests/dists.rs:89
    probs = [0.1, 0.3, 0.2, 0.1, 0.05, 0.25]
    for i, pr in enumerate(probs):
        assert O.logpdf("categorical", i, probs) == math.log(pr)
    assert O.logpdf("categorical", 6, probs) == -math.inf
