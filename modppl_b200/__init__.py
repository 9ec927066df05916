"""modppl_b200: B200-native SMC/MCMC engine behind modppl's inference vocabulary.

    ParticleSystem, importance_sampling, importance_resampling, metropolis_hastings (mh),
    regenerative_metropolis_hastings (regen_mh)   -- reference modppl/src/inference/mod.rs:8-10

All compute happens in hand-written sm_100a kernels inside lib/libmodppl_b200.so (C ABI: include/modppl_b200.h).
"""
from ._lib import MplError, device_count, LIB_PATH
from .models import (Model, CompiledModel, compile_model, lgssm4_spec, spiral_spec, lgssm4, spiral_model, stochastic_volatility, hmm, line_model, hierarchical_model,
                     pointed_model)
from .particle_filter import ParticleSystem, F32, F64, MULTINOMIAL, SYSTEMATIC, SYSTEMATIC_FIXED, MULTINOMIAL_FIXED, SYSTEMATIC_NESTED
from .importance import importance_sampling, importance_resampling
from .mh import (Chains, metropolis_hastings, mh, regenerative_metropolis_hastings, regen_mh, hierarchical_sweeps, hierarchical_full_sweeps,
                 run_schedule, proposals, HIER_DRIFT, HIER_ADD_REMOVE, POINTED_DRIFT, MASK_A, MASK_B, MASK_C, MASK_IS_LINEAR)
from . import parity
