"""GPU parity for batched importance sampling (reference inference/importance.rs) and many-chain MH (inference/mh.rs).
The engine and the oracle draw from the same Philox streams, so sample paths agree to fp64 round-off (device libm vs
glibc); tolerances: 1e-9 relative as BASELINE.json's north_star states for fp64."""
import math

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
m = pytest.importorskip("modppl_b200")

XS = np.arange(-5.0, 6.0)
# the oracle numbers its moves; the engine selects a proposal by the reference fixture's name (mh.rs:9-14 is generic over it)
PROPOSAL = {0: m.HIER_DRIFT, 1: m.HIER_ADD_REMOVE, 3: m.POINTED_DRIFT}


def line_data(seed=6):
    rng = np.random.default_rng(seed)
    return 0.5 * XS - 1.0 + 0.1 * rng.normal(size=11)           # tests/importance.rs:61-69


def hier_data(seed=7):
    rng = np.random.default_rng(seed)
    return 0.3 + 0.4 * XS + 0.5 * XS * XS + 0.1 * rng.normal(size=11)   # tests/importance.rs:98-106


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


@pytest.mark.parametrize("name", ["line", "hierarchical", "pointed"])
def test_importance_sampling_matches_oracle(name):
    n = 20000
    if name == "line":
        model, args, obs = m.line_model(XS), XS, line_data()
    elif name == "hierarchical":
        model, args, obs = m.hierarchical_model(XS), XS, hier_data()
    else:
        bounds, cov = [-5.0, 5.0, -5.0, 5.0], [1.0, -0.6, -0.6, 2.0]     # tests/importance.rs:26-28
        model, args, obs = m.pointed_model(bounds, cov), bounds + cov, [0.0, 0.0]
    lat, lnw, lml = m.importance_sampling(model, obs, n, seed=3, batch=2)
    rlat, rlnw, rlml = O.importance_sampling(name, args, obs, n, seed=3, batch=2)
    assert rel(lat, rlat) <= 1e-9
    assert rel(lnw, rlnw) <= 1e-9
    assert abs(lml - rlml) <= 1e-9 * max(1.0, abs(rlml))
    assert abs(np.exp(lnw).sum() - 1.0) < 1e-9                        # importance.rs:23-25: normalised


def test_importance_resampling_indices_match_oracle():            # importance.rs:37-51
    n, n_ret = 50000, 1000
    obs = line_data()
    lat, idx, lml = m.importance_resampling(m.line_model(XS), obs, n, n_ret, seed=4, batch=1)
    rlat, rlnw, rlml = O.importance_sampling("line", XS, obs, n, seed=4, batch=1)
    ridx = O.importance_resampling_indices(rlnw, n_ret, seed=4, batch=1)
    assert idx.min() >= 0 and idx.max() < n
    # same uniforms, weights equal to ~1e-13: the indices agree except where a uniform lands within round-off of a boundary
    assert np.mean(idx == ridx) >= 0.999
    assert abs(lml - rlml) <= 1e-9 * abs(rlml)
    # returns ALL traces plus indices (quirk Q10)
    assert lat.shape == (2, n)
    post_slope = lat[0, idx].mean()
    assert abs(post_slope - 0.5) < 0.1


def test_importance_sampling_lml_converges_to_closed_form():      # config 2 shape: 2^20 proposals per batch
    ys = line_data()
    truth = O.line_model_lml(XS, ys)
    est = []
    for b in range(4):
        _, lnw, lml = m.importance_sampling(m.line_model(XS), ys, 1 << 20, seed=9, batch=b)
        est.append(lml)
    # independent batches are independent streams
    assert len(set(est)) == 4
    lme = math.log(np.mean(np.exp(np.array(est) - max(est)))) + max(est)
    assert abs(lme - truth) < 0.25
    yh = hier_data()
    htruth, p_lin = O.hier_model_lml(XS, yh)
    lat, lnw, lml = m.importance_sampling(m.hierarchical_model(XS), yh, 1 << 20, seed=10)
    w = np.exp(lnw)
    assert np.sum(w * lat[0]) < 0.05 and p_lin < 0.05               # the quadratic branch explains these data
    assert abs(lml - htruth) < 3.0                                   # prior proposals: a high-variance estimator (tests/importance.rs:90-92)


@pytest.mark.parametrize("move,parg,mask", [(0, 0.1, 0), (0, 0.01, 0), (1, 0.025, 0), (2, 1.0, 1), (2, 1.0, 2), (2, 1.0, 4), (2, 1.0, 7), (2, 1.0, 8)])
def test_hierarchical_mh_matches_oracle(move, parg, mask):
    n, steps = 2048, 12
    ys = hier_data()
    ch = m.Chains(m.hierarchical_model(XS), ys, n, seed=5, chain_offset=100)
    ref = O.OracleChains("hierarchical", XS, ys, n, seed=5, offset=100)
    assert rel(ch.read(), ref.read()) <= 1e-9                       # chains start from generate(args, observations)
    # warm up with a few accepted moves so that both linear and quadratic states occur
    for mv, pa in [(1, 0.025), (0, 0.1), (1, 0.025)]:
        a = m.mh(ch, PROPOSAL[mv], pa, 3); b = ref.move(mv, pa, n_steps=3)
        assert a == b
    ch.write(ref.read())                                            # identical inputs for the move under test
    acc = m.regen_mh(ch, mask, steps) if move == 2 else m.mh(ch, PROPOSAL[move], parg, steps)
    racc = ref.move(move, parg, mask, steps)
    st, rst = ch.read(), ref.read()
    same = np.all(np.abs(st - rst) <= 1e-9 * np.maximum(1.0, np.abs(rst)), axis=0)
    # an accept test decided within round-off of log(u) == alpha may flip one chain; everything else is identical
    assert same.mean() >= 0.999
    assert abs(acc - racc) <= max(2, 0.001 * racc)
    for i in range(0, n, 97):                                       # cached logjp is the model's log joint (dyngenfn.rs:512)
        assert abs(st[4, i] - O.hier_logjp(XS, ys, st[:4, i])) <= 1e-9 * abs(st[4, i])


def test_regen_mask_c_on_linear_trace_is_noop():                  # nothing visited under the mask => w = 0, accepted
    ys = hier_data()
    n = 512
    ch = m.Chains(m.hierarchical_model(XS), ys, n, seed=11)
    st = ch.read()
    st[0] = 1.0; st[3] = 0.0                                        # force linear traces
    for i in range(n):
        st[4, i] = O.hier_logjp(XS, ys, st[:4, i])
    ch.write(st)
    acc = m.regen_mh(ch, m.MASK_C, 5)
    assert acc == 5 * n
    got = ch.read()
    assert np.array_equal(got[:4], st[:4])                          # choices untouched
    assert np.max(np.abs(got[4] - st[4]) / np.abs(st[4])) < 1e-12    # logjp re-derived on the device


def test_pointed_mh_matches_oracle_and_rejects_out_of_bounds():   # tests/mh.rs:21-46
    n = 4096
    bounds, cov = [-5.0, 5.0, -5.0, 5.0], [1.0, -0.6, -0.6, 2.0]
    ch = m.Chains(m.pointed_model(bounds, cov), [0.0, 0.0], n, seed=2)
    ref = O.OracleChains("pointed", bounds + cov, [0.0, 0.0], n, seed=2)
    assert rel(ch.read(), ref.read()) <= 1e-9
    acc = m.mh(ch, m.POINTED_DRIFT, 0.5, 40)                                      # drift cov 0.25 I (tests/mh.rs:28)
    racc = ref.move(3, 0.5, n_steps=40)
    st, rst = ch.read(), ref.read()
    same = np.all(np.abs(st - rst) <= 1e-9 * np.maximum(1.0, np.abs(rst)), axis=0)
    assert same.mean() >= 0.999 and abs(acc - racc) <= 4
    assert np.all(np.abs(st[:2]) <= 5.0)                            # uniform_2d = -inf outside => always rejected
    # posterior of latent | obs = (0, 0) is N(0, cov) truncated to the box
    many = m.Chains(m.pointed_model(bounds, cov), [0.0, 0.0], 1 << 16, seed=3)
    m.mh(many, m.POINTED_DRIFT, 0.5, 400)
    s = many.read()
    assert abs(np.mean(s[0])) < 0.03 and abs(np.mean(s[1])) < 0.04
    assert abs(np.var(s[0]) - 1.0) < 0.05 and abs(np.var(s[1]) - 2.0) < 0.08 and abs(np.mean(s[0] * s[1]) + 0.6) < 0.05


def test_hierarchical_sweeps_posterior():                          # config 3 shape (fewer chains/steps): tests/mh.rs:93-106 schedule
    ys = hier_data()
    n = 1 << 14
    ch = m.Chains(m.hierarchical_model(XS), ys, n, seed=8)
    acc, ms = m.hierarchical_sweeps(ch, 200, timed=True)
    st = ch.read()
    assert 0 < acc < 14 * 200 * n
    # the fused sweep equals the same moves issued one by one
    ch2 = m.Chains(m.hierarchical_model(XS), ys, 256, seed=8)
    ch3 = m.Chains(m.hierarchical_model(XS), ys, 256, seed=8)
    a2 = m.hierarchical_sweeps(ch2, 3)
    a3 = 0
    for _ in range(3):
        a3 += m.mh(ch3, m.HIER_ADD_REMOVE, 0.025, 1) + m.mh(ch3, m.HIER_DRIFT, 0.1, 3) + m.mh(ch3, m.HIER_DRIFT, 0.01, 10)
    assert a2 == a3 and np.array_equal(ch2.read(), ch3.read())
    # chains that reached the quadratic mode sit at the least-squares coefficients
    quad = st[0] == 0.0
    assert quad.mean() > 0.5
    A = np.stack([np.ones(11), XS, XS * XS], 1)
    ls = np.linalg.lstsq(A, ys, rcond=None)[0]
    good = quad & (st[4] > np.percentile(st[4], 60))
    assert np.all(np.abs(np.median(st[1:4, good], axis=1) - ls) < 0.1)


def test_proposals_are_registered_by_name_and_unknown_names_fail():
    hm = m.hierarchical_model(XS)
    assert m.proposals(hm) == ["hierarchical_drift_proposal", "add_or_remove_param_proposal"]       # tests/dyngenfns/hierarchical.rs:48-71
    assert m.proposals(m.pointed_model([-5.0, 5.0, -5.0, 5.0], [1.0, -0.6, -0.6, 2.0])) == ["pointed_2d_drift_proposal"]
    assert m.proposals(m.line_model(XS)) == []
    ch = m.Chains(hm, hier_data(), 64, seed=1)
    with pytest.raises(m.MplError):
        m.mh(ch, "pointed_2d_drift_proposal", 0.5, 1)                # registered for another model
    with pytest.raises(m.MplError):
        m.mh(ch, m.HIER_DRIFT, -1.0, 1)
    pc = m.Chains(m.pointed_model([-5.0, 5.0, -5.0, 5.0], [1.0, -0.6, -0.6, 2.0]), [0.0, 0.0], 64, seed=1)
    with pytest.raises(m.MplError):
        m.regen_mh(pc, 1, 1)                                         # no regenerate() registered for this model


def test_full_sweep_schedule_equals_the_moves_issued_one_by_one():  # config 3: 14 proposal moves + 4 regen_mh per sweep, one launch
    ys = hier_data()
    a = m.Chains(m.hierarchical_model(XS), ys, 512, seed=4)
    b = m.Chains(m.hierarchical_model(XS), ys, 512, seed=4)
    acc_a = m.hierarchical_full_sweeps(a, 4)
    acc_b = 0
    for _ in range(4):
        acc_b += m.mh(b, m.HIER_ADD_REMOVE, 0.025, 1) + m.mh(b, m.HIER_DRIFT, 0.1, 3) + m.mh(b, m.HIER_DRIFT, 0.01, 10)
        for mask in (m.MASK_IS_LINEAR, m.MASK_A, m.MASK_B, m.MASK_C):
            acc_b += m.regen_mh(b, mask, 1)
    assert acc_a == acc_b and np.array_equal(a.read(), b.read())


def test_importance_resampling_long_input_uses_the_parallel_exact_cumsum():
    # 2^20 proposals: the running sum goes through the parallel emulation of the sequential f64 sum (bit-exact, cumsum_exact.cuh)
    n, n_ret = 1 << 20, 1 << 10
    obs = line_data()
    lat, idx, lml = m.importance_resampling(m.line_model(XS), obs, n, n_ret, seed=4, batch=3)
    _, lnw, lml2 = m.importance_sampling(m.line_model(XS), obs, n, seed=4, batch=3)
    assert lml == lml2
    ridx = O.importance_resampling_indices(lnw, n_ret, seed=4, batch=3)     # the device's own weights through the oracle's sequential sum
    assert np.sum(idx != ridx) <= 1          # (device exp vs glibc exp differ in the last bit of some weights: a draw within 1e-13 of a boundary may flip)
