"""The reference's MCMC test (modppl/tests/mh.rs:60-112) on the engine, many chains at once: the hierarchical regression model
(tests/dyngenfns/hierarchical.rs:17-46), one sweep = add_or_remove_param_proposal(.025), 3 x hierarchical_drift_proposal(.1),
10 x hierarchical_drift_proposal(.01) -- the body of the reference's loop -- plus a regen_mh on each coefficient.

    python examples/mh_regression.py [log2_chains] [sweeps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m


def main():
    log2c = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    xs = np.arange(-5.0, 6.0)
    rng = np.random.default_rng(2)
    ys = 0.3 + 0.4 * xs + 0.5 * xs * xs + 0.1 * rng.normal(size=xs.size)       # generating coefficients (a, b, c) = (.3, .4, .5)
    model = m.hierarchical_model(xs)
    chains = m.Chains(model, ys, 1 << log2c, seed=2)
    # one call per move, as the reference's loop is written ...
    accepted = m.mh(chains, "add_or_remove_param_proposal", 0.025, 1)
    accepted += m.mh(chains, "hierarchical_drift_proposal", 0.1, 3)
    accepted += m.mh(chains, "hierarchical_drift_proposal", 0.01, 10)
    accepted += m.regen_mh(chains, m.MASK_A, 1) + m.regen_mh(chains, m.MASK_B, 1) + m.regen_mh(chains, m.MASK_C, 1)
    # ... or the whole loop body handed over as a schedule: every sweep of every chain in one launch, state in registers
    acc, ms = m.hierarchical_full_sweeps(chains, sweeps, timed=True)
    st = chains.read()                                                          # [slots][chains]: is_linear, a, b, c, log joint
    quad = st[0] < 0.5
    print(f"{1 << log2c} chains x {sweeps} sweeps of 18 moves in {ms:.1f} ms ({(1 << log2c) * sweeps * 18 / (ms * 1e-3):.3g} chain-steps/s), "
          f"acceptance {acc / ((1 << log2c) * sweeps * 18):.3f}; quadratic in {quad.mean():.3f} of the chains; "
          f"posterior mean (a, b, c) = ({st[1][quad].mean():.3f}, {st[2][quad].mean():.3f}, {st[3][quad].mean():.3f})")
    chains.close()


if __name__ == "__main__":
    main()
