/*
 * modppl_b200.h -- C ABI of the B200-native SMC/MCMC engine (libmodppl_b200.so).
 *
 * This is the drop-in boundary for modppl's inference hot path.  modppl itself has no FFI: its seam is the Rust
 * trait `GenFn` (modppl/src/gfi.rs:49-92) as consumed by `ParticleSystem` (src/inference/particle_filter.rs),
 * `importance_sampling/_resampling` (src/inference/importance.rs) and `mh/regen_mh` (src/inference/mh.rs).
 * A Rust shim crate binds the entry points below with `extern "C"` and re-exposes them under
 * the reference's names and signatures; the Python package modppl_b200/ does the same over ctypes.
 *
 * Conventions
 *   - every function returns an int status (MPL_OK == 0, negative == error) unless it returns a handle (NULL ==
 *     error); `mpl_last_error()` gives the thread-local message.  The reference panics instead (gfi.rs:72,
 *     dyngenfn.rs:526-529); a shim turns a non-zero status into a panic.
 *   - plain pointers and sizes only.  Host pointers unless a parameter is named `dev_*`.
 *   - models are the "restricted vectorisable form": a registered device functor (name + parameter vector of
 *     doubles) with fixed-shape SoA state and built-in normal/mvnormal/bernoulli/uniform/categorical log-densities.
 *   - `ThreadRng` (unseedable) is replaced by a counter-based Philox4x32-10 keyed by `seed`; the counter is
 *     (global particle/chain id, step, purpose, draw block) so results do not depend on how particles are
 *     sharded across GPUs.
 *   - one handle is used by one host thread at a time.  Work is queued on the handle's CUDA stream; calls that
 *     return a value to the host synchronise that stream.
 *   - there is NO CPU fallback: if no CUDA device is usable every compute entry point fails with MPL_ERR_CUDA.
 */
#ifndef MODPPL_B200_H
#define MODPPL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPL_OK 0
#define MPL_ERR_INVALID (-1)      /* bad argument / unknown model                                   */
#define MPL_ERR_CUDA (-2)         /* CUDA runtime error, or no device                                */
#define MPL_ERR_DEGENERATE (-3)   /* all weights are -inf (reference: logsumexp -> -inf, then NaN)   */
#define MPL_ERR_UNSUPPORTED (-4)

#define MPL_F32 0
#define MPL_F64 1

/* resampling schemes.  0 is the reference's (particle_filter.rs:37-41 -> categorical.rs:22-32): normalised f64
 * weights, SEQUENTIAL f64 running sum, parent = min{k : S_k >= u}.  1 keeps that cumsum convention with positions
 * (u+i)/N.  2 and 3 quantise weights to integers (exact, order-independent prefix sums; identical ancestors for
 * any sharding) -- 2 is the throughput path. */
#define MPL_RESAMPLE_MULTINOMIAL 0
#define MPL_RESAMPLE_SYSTEMATIC 1
#define MPL_RESAMPLE_SYSTEMATIC_FIXED 2
#define MPL_RESAMPLE_MULTINOMIAL_FIXED 3
#define MPL_RESAMPLE_SYSTEMATIC_NESTED 4   /* integer weights quantised per 128-particle chunk; sections (2^17 particles), chunks, then particles
                                            * resampled systematically, each level exactly; the throughput scheme */

#define MPL_READ_STATE 0          /* double[D * N], SoA: state[d * N + i]  (`traces[i].retv.last()`, dynunfold.rs:76) */
#define MPL_READ_LOG_WEIGHTS 1    /* double[N]                              (`log_weights`, particle_filter.rs:15)      */
#define MPL_READ_PARENTS 2        /* int64[N]                               (`parents`, particle_filter.rs:20)          */

typedef struct mpl_model mpl_model;
typedef struct mpl_ps mpl_ps;
typedef struct mpl_chains mpl_chains;

const char* mpl_last_error(void);
const char* mpl_version(void);
int mpl_device_count(int* count);

/* ---- model registry (replaces dyngen!/DynUnfold authoring, modppl-macros/src/lib.rs:20-114) --------------------
 * Unfold models (particle filter):  "lgssm4" {q_std, r_std, x0_std}; "spiral" {dr_std, dtheta_mean, dtheta_std,
 *   obs_var} (tests/dyngenfns/unfold.rs:14-33); "sv" {mu, phi, sigma}; "hmm" {K, M, prior[K], emission[M*K],
 *   transition[K*K]} (tests/hmm/model.rs:24-81).
 * Static models (importance sampling / MH): "line" {xs...} (tests/dyngenfns/simple.rs:10-23); "hierarchical"
 *   {xs...} (tests/dyngenfns/hierarchical.rs:32-46); "pointed" {xmin,xmax,ymin,ymax, cov[4]}
 *   (tests/pointed_model/model.rs). */
mpl_model* mpl_model_create(const char* name, const double* params, size_t n_params);
/* Model front-end (stands in for the `dyngen!` macro + DynUnfold, modppl-macros/src/lib.rs:20-114, dynunfold.rs:41-100): an Unfold
 * model in the restricted vectorisable form written as a JSON spec --
 *   {"name", "state_dim", "obs_dim", "params": {name: value, ...},
 *    "init":    [one sample statement per state component, drawn at t = 0],
 *    "step":    [one per component, drawn at t > 0 from the previous state],
 *    "observe": [observation terms scored on the new state]}
 * sample statement: {"dist": "normal" | "uniform" | "delta", "args": [..], "add_to": expr (optional: an increment)};
 * observation term: {"dist": "normal" | "uniform" | "bernoulli", "value": expr, "args": [..]}, {"dist": "mvnormal2", "value": [2],
 * "mean": [2], "cov": [4]}, {"dist": "expr", "value": log-density}.  Expressions are C++ expressions over x[i] (state), y[i]
 * (observation), t and the parameter names.  The spec becomes a device functor compiled by NVRTC against the library's own kernel
 * headers on first use (or by mpl_model_jit_compile): the model then runs through the same kernels as the built-in ones, without
 * rebuilding the library.  NULL on a malformed spec (mpl_last_error()). */
mpl_model* mpl_model_compile(const char* spec_json);
int mpl_model_jit_compile(mpl_model*, int dtype, char* log, size_t log_bytes);   /* compile now (no device needed); log: compiler messages */
const char* mpl_model_jit_source(const mpl_model*, int dtype);                   /* the generated translation unit, for inspection */
void mpl_model_destroy(mpl_model*);
int mpl_model_state_dim(const mpl_model*);
int mpl_model_obs_dim(const mpl_model*);

/* ---- ParticleSystem (src/inference/particle_filter.rs:8-121) ------------------------------------------------- */
typedef struct mpl_pf_config {
    int dtype;             /* MPL_F32 | MPL_F64: storage + arithmetic type of state and log-weights            */
    int device;            /* CUDA device ordinal, -1 = current                                                 */
    uint64_t seed;         /* replaces the ThreadRng argument of ParticleSystem::new (particle_filter.rs:44)   */
    uint64_t gid_offset;   /* global id of this shard's first particle (0 on a single GPU)                     */
    uint64_t n_global;     /* total particles over all shards (0 = num_particles)                              */
} mpl_pf_config;

mpl_ps* mpl_particle_system_new(const mpl_model*, uint64_t num_particles, const mpl_pf_config*);   /* ::new       :44-57  */
void mpl_ps_destroy(mpl_ps*);
int mpl_ps_init_step(mpl_ps*, const double* obs, size_t n_obs);                                     /* init_step   :60-70  */
int mpl_ps_step(mpl_ps*, const double* obs, size_t n_obs);                                          /* step        :73-95  */
int mpl_ps_effective_sample_size(mpl_ps*, int stale_like_reference, double* out);                   /* ESS         :98-100 */
int mpl_ps_resample(mpl_ps*, int scheme, double* log_total_weight);                                 /* resample    :103-116; log_total_weight may be NULL (no host sync) */
/* step() then resample() -- the body of the reference's filtering loop (tests/smc.rs:78-81) -- as ONE call: same results as
 * the two calls; with MPL_RESAMPLE_SYSTEMATIC_NESTED on fp32 the extend kernel quantises the weights in its epilogue */
int mpl_ps_step_resample(mpl_ps*, const double* obs, size_t n_obs, int scheme, double* log_total_weight);
int mpl_ps_log_marginal_likelihood_estimate(mpl_ps*, double* out);                                  /* lml         :119-121 */
/* checkpoint / resume (single GPU): the filter's resumable state in its native precision.  A restored filter continues
 * exactly as the original would have (same seed required); the trajectory log is not part of a checkpoint. */
int mpl_ps_checkpoint_size(mpl_ps*, uint64_t* bytes);
int mpl_ps_checkpoint(mpl_ps*, void* dst, uint64_t bytes);
int mpl_ps_restore(mpl_ps*, const void* src, uint64_t bytes);
int mpl_ps_read(mpl_ps*, int what, void* host_dst, size_t bytes);                                   /* `pub traces` :13     */
int mpl_ps_write(mpl_ps*, int what, const void* host_src, size_t bytes);                            /* parity hook: inject state / log-weights */
int mpl_ps_num_particles(const mpl_ps*, uint64_t* out);
/* Trajectories: the reference keeps every trace's whole history (`retv: Vec<State>`, dynunfold.rs:91-92) and clones it on
 * each resample.  Here an optional log of per-step states and ancestors (max_steps x (D+1) x N elements, so for small N)
 * is back-traced on demand: out[k][t][d] = state at step t of the lineage of particle ids[k]. */
int mpl_ps_history_enable(mpl_ps*, uint64_t max_steps);
int mpl_ps_trajectories(mpl_ps*, const int64_t* ids, uint64_t n_ids, double* out, size_t bytes, uint64_t* n_steps);
int mpl_ps_sync(mpl_ps*);

/* Device-resident filter loop: `n_steps` x (step; [ESS test]; resample) with all observations already in HBM.
 * ess_threshold <= 0: resample after every step (tests/smc.rs:80-85); otherwise resample when fresh ESS <
 * ess_threshold * N.  Nothing is copied to the host inside the loop.  elapsed_ms (nullable) is the CUDA-event time of
 * the loop on the handle's stream. */
int mpl_ps_upload_observations(mpl_ps*, const double* obs, size_t n_steps, size_t n_obs);
int mpl_ps_run(mpl_ps*, size_t first_step, size_t n_steps, int scheme, double ess_threshold, float* elapsed_ms);
int mpl_ps_num_resamples(mpl_ps*, uint64_t* out);   /* resamples performed so far (integer schemes) */

/* per-kernel CUDA-event timing (bench.py's roofline leg).  names: "extend", "fixed_reduce", "fixed_scan", ... */
int mpl_ps_profile_enable(mpl_ps*, int on);
int mpl_ps_profile_get(mpl_ps*, const char* kernel, double* total_ms, uint64_t* launches);
int mpl_ps_launch_count(mpl_ps*, uint64_t* out);

/* ---- importance sampling (src/inference/importance.rs:12-51) -------------------------------------------------- */
/* latents: double[L * n] SoA (may be NULL), log_norm_weights: double[n] (may be NULL), lml: log-ML estimate.
 * `batch` selects an independent RNG stream (config 2 runs 64 batches). */
int mpl_importance_sampling(const mpl_model*, const double* obs, size_t n_obs, uint32_t num_samples, uint64_t seed,
                            uint64_t batch, double* latents, double* log_norm_weights, double* lml);
int mpl_importance_resampling(const mpl_model*, const double* obs, size_t n_obs, uint32_t num_samples,
                              uint32_t num_ret_samples, uint64_t seed, uint64_t batch, double* latents,
                              int64_t* resampled_indices, double* lml);
int mpl_model_num_latents(const mpl_model*);

/* ---- Metropolis-Hastings over many independent chains (src/inference/mh.rs:9-76) -----------------------------
 * `mh(model, trace, proposal, proposal_args)` is generic over the proposal GenFn (mh.rs:9-14).  Here a static model registers
 * its proposals as device functors under the names of the reference's fixtures and `mpl_mh` selects one BY NAME:
 *   "hierarchical": "hierarchical_drift_proposal" (std; hierarchical.rs:63-71), "add_or_remove_param_proposal" (std; :48-61)
 *   "pointed":      "pointed_2d_drift_proposal"   (s: covariance s^2 I; tests/pointed_model/proposal.rs)
 * A new proposal is a struct with make / propose / assess next to its model in csrc/is_mh.cu, listed in Model::Proposals. */
int mpl_model_num_proposals(const mpl_model*);
const char* mpl_model_proposal_name(const mpl_model*, int index);       /* NULL when out of range */
int mpl_model_proposal_index(const mpl_model*, const char* name);       /* >= 0, or MPL_ERR_INVALID */
mpl_chains* mpl_chains_new(const mpl_model*, const double* obs, size_t n_obs, uint64_t n_chains, uint64_t seed,
                           uint64_t chain_offset, int device);                     /* trace = model.generate(args, obs).0 per chain */
void mpl_chains_destroy(mpl_chains*);
int mpl_mh(mpl_chains*, const char* proposal, double proposal_arg, uint32_t n_steps, uint64_t* n_accepted);   /* mh       :9-50  */
/* regen_mh(model, trace, mask): mask bits over the model's latent slots; "hierarchical": 1 coeffs/a, 2 coeffs/b, 4 coeffs/c,
 * 8 is_linear; 0 = everything (dyngenfn.rs:571) */
int mpl_regen_mh(mpl_chains*, uint32_t mask_bits, uint32_t n_steps, uint64_t* n_accepted);                      /* regen_mh :54-76 */
/* A sweep as ONE launch: `moves` is the body of the caller's loop (e.g. tests/mh.rs:93-106: 1 add/remove(.025), 3 drift(.1),
 * 10 drift(.01)), run n_sweeps times per chain with the chain's state in registers throughout. */
#define MPL_MOVE_MH 0
#define MPL_MOVE_REGEN 1
typedef struct mpl_move {
    int32_t kind;        /* MPL_MOVE_MH | MPL_MOVE_REGEN                                   */
    int32_t proposal;    /* MPL_MOVE_MH: mpl_model_proposal_index(model, name)             */
    double arg;          /* MPL_MOVE_MH: proposal_args                                     */
    uint32_t mask;       /* MPL_MOVE_REGEN: mask bits                                      */
    uint32_t repeat;     /* how many times in a row                                        */
} mpl_move;
int mpl_mh_schedule(mpl_chains*, const mpl_move* moves, uint32_t n_moves, uint32_t n_sweeps, uint64_t* n_accepted, float* elapsed_ms);
int mpl_chains_num_slots(const mpl_chains*);
int mpl_chains_read(mpl_chains*, double* host_dst, size_t bytes);    /* double[slots * n] SoA */
int mpl_chains_write(mpl_chains*, const double* host_src, size_t bytes);

/* ---- parity hooks: injected inputs, no RNG --------------------------------------------------------------------- */
/* categorical.rs:22-32 / particle_filter.rs:37-41 on the device: bit-exact against the sequential f64 routine.
 * scheme: MPL_RESAMPLE_MULTINOMIAL (n_draws uniforms) or MPL_RESAMPLE_SYSTEMATIC (uniforms[0] only). */
int mpl_resample_indices(const double* probs, const double* uniforms, uint64_t n, uint64_t n_draws, int scheme,
                         int64_t* parents);
int mpl_cumsum_sequential(const double* probs, uint64_t n, double* out);       /* the exact running sum itself */
/* lib.rs:34-45 + particle_filter.rs:27-35,98-100 in one pass: lse, ESS = 1/sum(w~^2), max. dtype of lw. */
int mpl_logsumexp_stats(const void* lw, uint64_t n, int dtype, double* lse, double* ess, double* max);
/* integer-weight resamplers on injected f32 log-weights; the systematic offset word and the multinomial per-output
 * words come from Philox(seed, t) exactly as inside a particle system resampling the weights of step t. */
int mpl_fixed_resample(const float* lw, uint64_t n, int scheme, uint64_t seed, uint32_t t, int32_t* anc,
                       double* lse, uint64_t* total_weight);
/* built-in log-densities evaluated on the device (tests/dists.rs known answers): "normal" (mu, std), "bernoulli" (p),
 * "uniform" (a, b), "uniform_2d" (bounds), "mvnormal2" (mu[2], cov[4]; determinant and inverse hoisted, as the models use
 * it), "mvnormal" (x[k]; mu[k], cov[k*k] row-major, 1 <= k <= 8: mvnormal.rs:14-22 literally), "uniform_discrete" (a, b),
 * "geometric" (p), "poisson" (rate), "beta" (a, b), "gamma" (shape, scale), "categorical" (the probabilities, at most 8) */
int mpl_logpdf(const char* dist, const double* x, const double* params, size_t n_params, double* out);

/* ---- multi-GPU (one process per GPU) ---------------------------------------------------------------- */
/* Peer (NVLink) exchange without a host round trip.  Each rank exports a handle blob, the caller all-gathers the
 * blobs with whatever transport it has (torch.distributed in bench.py) and attaches them. */
#define MPL_PEER_BLOB_BYTES 1024
int mpl_ps_peer_export(mpl_ps*, void* blob /* MPL_PEER_BLOB_BYTES */);
int mpl_ps_peer_attach(mpl_ps*, int rank, int world, const void* blobs /* world * MPL_PEER_BLOB_BYTES */);
int mpl_ps_peer_detach(mpl_ps*);
int mpl_ps_peer_barrier(mpl_ps*);           /* queues a device-side rendezvous of all ranks' streams (no host round trip); every rank calls it */
int mpl_ps_peer_error(mpl_ps*, int* out);   /* 1 if a kernel gave up waiting for a peer (bounded spin) */
int mpl_ps_nvlink_bytes(mpl_ps*, uint64_t* out);   /* payload bytes requested from the peers' memory so far (remote parents, weights, records) */
int mpl_ps_trace(mpl_ps*, long long* out16);   /* device time stamps (ns) of the last sharded step's phases; diagnostics */
/* Islands -- the local-resample-then-rebalance variant: each GPU filters N / G particles of its own and resamples
 * locally, so no step waits for another GPU.  The islands' weights are their log-ML increments; the host compares them now and
 * then and, when the island-level ESS has dropped, resamples whole islands: a surviving island is copied over NVLink into the
 * place of one that died out (orchestration: modppl_b200/distributed.py, IslandParticleSystem).  A different estimator from the
 * global scheme above (not the same ancestors), unbiased for the likelihood all the same. */
int mpl_ps_island_export(mpl_ps*, void* blob /* MPL_PEER_BLOB_BYTES */);
int mpl_ps_island_attach(mpl_ps*, int rank, int n_islands, const void* blobs /* n_islands * MPL_PEER_BLOB_BYTES */);
int mpl_ps_live_buffer(mpl_ps*, int* out);                 /* which state buffer is live (a pending resample is applied first) */
int mpl_ps_island_copy_from(mpl_ps*, int src_island, int src_live_buffer);   /* this island := a copy of island src (uniform weights) */
int mpl_ps_copy_state(mpl_ps* dst, mpl_ps* src);           /* the same between two particle systems of one process */
/* Test hook: the nested scheme computes level 1 (the first output slot of every chunk) in a plan pass of its own for large shards
 * and inside the expansion for an unsharded population of at most 2^22 particles (-1, the default); 0 / 1 force one or the other, also
 * for shards.  Same results. */
int mpl_test_set_inline_level1(int mode);
/* Test hook: `world` shards emulated on ONE GPU run the multi-GPU kernels phase by phase (remote loads/stores become
 * local).  init + resample, then steps with a resample after each except the last; outputs the final state
 * double[D * n_global] (SoA), log-weights double[n_global] and the log-ML estimate. */
int mpl_test_virtual_shards(const mpl_model*, uint64_t n_global, int world, int dtype, uint64_t seed, const double* obs,
                            size_t n_steps, size_t n_obs, double* state_out, double* lw_out, double* lml_out, double* loop_ms);
/* same, with the resampling scheme chosen: MPL_RESAMPLE_SYSTEMATIC_FIXED or MPL_RESAMPLE_SYSTEMATIC_NESTED (shards of
 * whole 2^17-particle sections) */
int mpl_test_virtual_shards_scheme(const mpl_model*, uint64_t n_global, int world, int dtype, uint64_t seed, int scheme, const double* obs,
                                   size_t n_steps, size_t n_obs, double* state_out, double* lw_out, double* lml_out, double* loop_ms);

#ifdef __cplusplus
}
#endif
#endif
