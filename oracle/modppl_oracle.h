/*
 * modppl_oracle.h -- CPU restatement of agarret7/modppl's inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the reported CPU baseline.
 * The product (modppl_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED for log-densities, logsumexp, Update/Regenerate weight
 * identities and the particle-filter log-ML (against every known-answer value the
 * reference's own tests hold -- see tests/test_golden.py).  UNPINNED by the
 * reference for: resampled ancestor indices (no reference test injects uniforms;
 * the restatement of categorical.rs:22-32 is the definition), importance-sampling
 * log-ML, MH acceptance, regen_mh and systematic resampling (absent upstream).
 * The reference (Rust) cannot be built here (no cargo/rustc), so there is no
 * oracle/_ref.
 *
 * All file:line citations are relative to /root/reference/modppl/.
 */
#ifndef MODPPL_ORACLE_H
#define MODPPL_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/lib.rs:34-45 ------------------------------------------------------ */
double mo_logsumexp(const double* xs, size_t n);

/* ---- src/modeling/dists (one file per distribution) ------------------------------------------------ */
double mo_normal_logpdf(double x, double mu, double std);                 /* normal.rs:13-17   */
double mo_bernoulli_logpdf(int a, double p);                              /* bernoulli.rs:12-14 */
double mo_uniform_logpdf(double x, double a, double b);                   /* uniform.rs:22-26 (NaN where the reference panics) */
double mo_uniform_discrete_logpdf(int64_t x, int64_t a, int64_t b);       /* uniform.rs:43-47 (NaN where the reference panics) */
double mo_geometric_logpdf(int64_t k, double p);                          /* geometric.rs:16-19 */
double mo_poisson_logpdf(int64_t k, double rate);                         /* poisson.rs:16-18 */
double mo_beta_logpdf(double x, double a, double b);                      /* beta.rs:17-21 (Gamma function: crate compute 0.2.3 -> std::tgamma) */
double mo_gamma_logpdf(double x, double shape, double scale);             /* gamma.rs:17-20 */
double mo_uniform2d_logpdf(double x, double y, const double bounds[4]);   /* tests/pointed_model/types_2d.rs:15-21; bounds = xmin,xmax,ymin,ymax */
double mo_mvnormal_logpdf(const double* x, const double* mu, const double* cov, int k); /* mvnormal.rs:14-22, nalgebra 0.32 det/inverse, k<=4, cov row-major */
int64_t mo_categorical_random(const double* probs, size_t n, double u);   /* categorical.rs:22-32 with injected u; literal (may return -1 or run past n: returns n) */
double mo_categorical_logpdf(int64_t x, const double* probs, size_t n);   /* categorical.rs:13-20 */

/* ---- resampling (particle_filter.rs:37-41 + categorical.rs:22-32) ----------- */
#define MO_SCHEME_MULTINOMIAL 0
#define MO_SCHEME_SYSTEMATIC 1
void mo_cumsum_sequential(const double* probs, size_t n, double* out);
/* faithful: literal per-draw linear scan (O(n) per draw).  fast: sequential cumsum + binary search.  Both clamp to [0,n-1]
 * (quirk Q2) and must agree bit-for-bit.  systematic: uniforms[0] only, positions (u+i)/n_draws. */
int mo_resample_indices_faithful(const double* probs, const double* uniforms, size_t n, size_t n_draws, int scheme, int64_t* parents);
int mo_resample_indices(const double* probs, const double* uniforms, size_t n, size_t n_draws, int scheme, int64_t* parents);

/* ---- counter-based RNG the engine uses instead of ThreadRng ----------------- */
void mo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double mo_u01_f64(uint32_t hi, uint32_t lo);      /* [0,1), 53 bit */
float mo_u01_f32(uint32_t x);                     /* [0,1), 24 bit */

/* ---- fixed-point (integer) weights: engine-defined, shard-invariant --------- */
float mo_exp2_poly(float f);                      /* reproducible 2^f on [0,1) */
uint64_t mo_fixed_weight(float d, int kbits);     /* rint(exp(d) * 2^kbits), d = lw - max <= 0 */
int mo_fixed_kbits(uint64_t n_total);
/* host threads used by the per-particle / per-chunk loops of the particle filter (default 1; results do not depend on it) */
void mo_set_threads(int n);
int mo_get_threads(void);
/* lw: float log-weights.  Returns total W; writes ancestors (systematic, offset word `u64rand`). */
uint64_t mo_fixed_systematic(const float* lw, size_t n, uint64_t u64rand, int32_t* anc, double* lse_out);
uint64_t mo_fixed_multinomial(const float* lw, size_t n, uint64_t seed, uint32_t t, int32_t* anc, double* lse_out);
/* nested systematic on integer weights: weights quantised against the max of their own 128-particle chunk (power-of-two
 * reference), chunks resampled systematically by their totals, particles systematically inside each chunk. */
uint64_t mo_nested_systematic(const float* lw, size_t n, uint64_t u64rand, int32_t* anc, double* lse_out);

/* ---- particle filter (inference/particle_filter.rs) -------------------------- */
typedef struct mo_ps mo_ps;
#define MO_F32 0
#define MO_F64 1
#define MO_RESAMPLE_MULTINOMIAL 0        /* reference scheme: sequential f64 cumsum + per-draw search */
#define MO_RESAMPLE_SYSTEMATIC 1         /* same cumsum convention, positions (u+i)/N */
#define MO_RESAMPLE_SYSTEMATIC_FIXED 2   /* integer weights */
#define MO_RESAMPLE_MULTINOMIAL_FIXED 3
#define MO_RESAMPLE_SYSTEMATIC_NESTED 4
mo_ps* mo_ps_new(const char* model, const double* params, size_t n_params, uint64_t num_particles,
                 int dtype, uint64_t seed, uint64_t gid_offset, uint64_t n_global);
void mo_ps_free(mo_ps*);
int mo_ps_init_step(mo_ps*, const double* obs, size_t n_obs);
int mo_ps_step(mo_ps*, const double* obs, size_t n_obs);
double mo_ps_effective_sample_size(mo_ps*, int stale_like_reference);
double mo_ps_resample(mo_ps*, int scheme);
double mo_ps_log_marginal_likelihood_estimate(mo_ps*);
int mo_ps_state_dim(mo_ps*);
void mo_ps_read_state(mo_ps*, double* out /* D x N, SoA */);
void mo_ps_read_log_weights(mo_ps*, double* out);
void mo_ps_read_parents(mo_ps*, int64_t* out);
void mo_ps_write_state(mo_ps*, const double* in);
void mo_ps_write_log_weights(mo_ps*, const double* in);
/* reference-cost variant of resample(): O(N^2) categorical with per-draw clone + sum, as categorical.rs:22-32 is written */
double mo_ps_resample_faithful_cost(mo_ps*);

/* ---- importance sampling (inference/importance.rs) --------------------------- */
/* model: "line" (tests/dyngenfns/simple.rs:10-23), "hierarchical" (hierarchical.rs:32-46), "pointed" (pointed_model/model.rs) */
int mo_importance_sampling(const char* model, const double* args, size_t n_args, const double* obs, size_t n_obs,
                           uint32_t num_samples, uint64_t seed, uint64_t batch,
                           double* latents /* L x n SoA */, double* log_norm_weights, double* lml);
int mo_importance_resampling_indices(const double* log_norm_weights, uint32_t n, uint32_t n_ret, uint64_t seed, uint64_t batch, int64_t* idx);
int mo_is_num_latents(const char* model);

/* ---- Metropolis-Hastings (inference/mh.rs) ----------------------------------- */
typedef struct mo_chains mo_chains;
#define MO_MOVE_HIER_DRIFT 0           /* hierarchical_drift_proposal(std)     hierarchical.rs:63-71 */
#define MO_MOVE_HIER_ADD_REMOVE 1      /* add_or_remove_param_proposal         hierarchical.rs:48-61 */
#define MO_MOVE_HIER_REGEN 2           /* regen_mh, mask bits: 1=a 2=b 4=c 8=is_linear(extension) */
#define MO_MOVE_POINTED_DRIFT 3        /* pointed drift proposal, cov = s^2 I  pointed_model/proposal.rs */
mo_chains* mo_chains_new(const char* model, const double* args, size_t n_args, const double* obs, size_t n_obs,
                         uint64_t n_chains, uint64_t seed, uint64_t chain_offset);
void mo_chains_free(mo_chains*);
int mo_chains_move(mo_chains*, int move, double parg, uint32_t mask, uint32_t n_steps, uint64_t* n_accepted);
int mo_chains_num_slots(mo_chains*);
void mo_chains_read(mo_chains*, double* out /* slots x n SoA; hierarchical: is_linear,a,b,c,logjp */);
void mo_chains_write(mo_chains*, const double* in);
/* single-transition weight pieces for parity tests (hierarchical): returns alpha; out[0..2] = w_update, fwd, bwd */
double mo_hier_mh_alpha(const double* xs, const double* ys, size_t n, const double cur[4], const double prop[4],
                        int move, double parg, double out[3]);
double mo_hier_logjp(const double* xs, const double* ys, size_t n, const double st[4]);

/* ---- ground truths ------------------------------------------------------------ */
double mo_hmm_forward(const double* prior, const double* emission /* n_obs x K row-major: emission[o*K+s] */,
                      const double* transition /* K x K: transition[to*K+from] */, int K, int n_obs_sym,
                      const int* obs, int T);                                    /* tests/hmm/forward.rs:3-23 */
double mo_kalman_lml_lgssm4(double q_std, double r_std, double x0_std, const double* ys /* T x 2 */, int T);
double mo_line_model_lml(const double* xs, const double* ys, int n);            /* closed-form Gaussian marginal */
double mo_hier_model_lml(const double* xs, const double* ys, int n, double* p_linear);

#ifdef __cplusplus
}
#endif
#endif
