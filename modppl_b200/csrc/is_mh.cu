// is_mh.cu -- batched importance sampling (reference modppl/src/inference/importance.rs:12-51) and many-chain
// Metropolis-Hastings (src/inference/mh.rs:9-76).
//
// The reference's entry points are generic over the GenFn trait (gfi.rs:49-92): importance_sampling takes ANY model, mh ANY
// (model, proposal) pair.  Here the same seam is a pair of device-functor concepts, and every kernel is a template over them:
//
//   StaticModel   L                         number of latent slots (fixed-shape choice vector; absent choices hold 0)
//                 generate(g, z) -> w       GenFn::generate(args, constraints): unconstrained choices from the prior, weight =
//                                           sum of the constrained choices' log-densities        (gfi.rs:70-78, dyngenfn.rs:115-141)
//                 log_joint(z)              the trace's cached logjp                              (dyngenfn.rs:512)
//                 update(cur, cur_logjp, next, &next_logjp) -> w
//                                           GenFn::update(trace, args, NoChange, choices) with every proposed choice replaced;
//                                           flattened weight algebra of dyngenfn.rs:143-221
//                 regenerate(g, cur, mask, next) -> w         (kHasRegenerate)   GenFn::regenerate, dyngenfn.rs:223-273
//                 Proposals                 std::tuple of the proposal functors registered for this model
//   Proposal      name()                    the reference fixture's name (what mpl_mh's `proposal` argument selects)
//                 make(arg)                 proposal_args -> functor (constants hoisted once per launch)
//                 propose(model, g, cur, next) -> log q(next | cur)       GenFn::propose   (gfi.rs:80-84)
//                 assess(model, from, to) -> log q(to | from)             GenFn::assess    (gfi.rs:86-92)
//
// RNG convention of one MH move (stream = (seed, chain id, move number, P_MH)): the proposal draws from Philox blocks 0..2, the
// accept / reject uniform is block 3 -- so a move's draws do not depend on which proposal ran before it.
// One thread per proposal / chain; all state in registers; fp64 like the reference: these paths are register-resident
// (FP64 pipe / issue bound), not HBM bound.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <tuple>
#include <utility>
#include "engine.h"

namespace mpl {

constexpr int kMaxPoints = 32;
constexpr uint32_t kAcceptBlock = 3;
__device__ __forceinline__ double std_normal_logpdf(double x) { return -(x * x + 1.8378770664093453) / 2. - 0.; }
__device__ __forceinline__ double normal_lp(double x, double mu, double sd, double ln_sd) { double z = (x - mu) / sd; return -(z * z + 1.8378770664093453) / 2. - ln_sd; }

struct RegressionData {
    int n;
    double xs[kMaxPoints], ys[kMaxPoints];
};

// ---- tests/dyngenfns/simple.rs:10-23 --------------------------------------------------------------------------------------
struct LineModel {
    static constexpr int L = 2;                       // slope, intercept
    static constexpr bool kHasRegenerate = false;
    using Proposals = std::tuple<>;
    RegressionData d;
    __device__ __forceinline__ double log_lik(const double (&z)[L]) const {
        double w = 0.;
        const double ln_noise = -2.3025850929940455;   // ln 0.1
        for (int j = 0; j < d.n; ++j) { double r = (d.ys[j] - (z[0] * d.xs[j] + z[1])) / 0.1; w += -(r * r + 1.8378770664093453) / 2. - ln_noise; }
        return w;
    }
    __device__ __forceinline__ double generate(Rng64& g, double (&z)[L]) const {
        double z0, z1;
        g.normal2(z0, z1);
        z[0] = z0 * 1. + 0.; z[1] = z1 * 2. + 0.;      // simple.rs:12-13
        return log_lik(z);
    }
    __device__ __forceinline__ double log_joint(const double (&z)[L]) const { return std_normal_logpdf(z[0]) + normal_lp(z[1], 0., 2., 0.6931471805599453) + log_lik(z); }
    __device__ __forceinline__ double update(const double (&)[L], double cur_logjp, const double (&next)[L], double& next_logjp) const { next_logjp = log_joint(next); return next_logjp - cur_logjp; }
};

// ---- tests/dyngenfns/hierarchical.rs:32-71 ---------------------------------------------------------------------------------
struct HierDrift;
struct HierAddRemove;
struct HierarchicalModel {
    static constexpr int L = 4;                       // is_linear (1 / 0), coeffs/a, coeffs/b, coeffs/c (0 while is_linear)
    static constexpr bool kHasRegenerate = true;
    using Proposals = std::tuple<HierDrift, HierAddRemove>;
    RegressionData d;
    __device__ __forceinline__ double log_lik(const double (&z)[L]) const {
        const bool lin = z[0] != 0.;
        double w = 0.;
        const double ln_noise = -2.3025850929940455;   // ln 0.1
        for (int i = 0; i < d.n; ++i) {
            double x = d.xs[i];
            double mean = lin ? z[1] + z[2] * x : z[1] + z[2] * x + z[3] * x * x;    // hierarchical.rs:36-44
            double r = (d.ys[i] - mean) / 0.1;
            w += -(r * r + 1.8378770664093453) / 2. - ln_noise;                       // normal.rs:13-17
        }
        return w;
    }
    __device__ __forceinline__ double log_joint(const double (&z)[L]) const {
        const bool lin = z[0] != 0.;
        double lp = log(lin ? 0.7 : 1. - 0.7) + std_normal_logpdf(z[1]) + std_normal_logpdf(z[2]);   // bernoulli.rs:12-14
        if (!lin) lp += std_normal_logpdf(z[3]);
        return lp + log_lik(z);
    }
    __device__ __forceinline__ double generate(Rng64& g, double (&z)[L]) const {
        const bool lin = 0.7 > g.uniform();             // bernoulli.rs:16-18
        double a, b;
        g.normal2(a, b);
        double c = g.normal();
        z[0] = lin ? 1. : 0.; z[1] = a; z[2] = b; z[3] = lin ? 0. : c;
        return log_lik(z);
    }
    __device__ __forceinline__ double update(const double (&)[L], double cur_logjp, const double (&next)[L], double& next_logjp) const {
        next_logjp = log_joint(next);                   // every latent is (re)proposed or kept: weight = delta logjp
        return next_logjp - cur_logjp;
    }
    // dyngenfn.rs:223-266: masked choices are redrawn from the prior; a choice that appears because the branch changed is drawn
    // fresh too; weight = likelihood ratio.  mask bits: 1 coeffs/a, 2 coeffs/b, 4 coeffs/c, 8 is_linear.
    __device__ __forceinline__ double regenerate(Rng64& g, const double (&cur)[L], uint32_t mask, double (&next)[L]) const {
        double za, zb;
        g.normal2(za, zb);
        const double zc = g.normal(), uf = g.uniform();
        const bool cur_lin = cur[0] != 0.;
        const bool lin = (mask & 8u) ? (0.7 > uf) : cur_lin;
        next[0] = lin ? 1. : 0.;
        next[1] = (mask & 1u) ? za * 1. + 0. : cur[1];
        next[2] = (mask & 2u) ? zb * 1. + 0. : cur[2];
        next[3] = cur[3];
        if (lin) next[3] = 0.;
        else if (cur_lin) next[3] = zc * 1. + 0.;
        else if (mask & 4u) next[3] = zc * 1. + 0.;
        return log_lik(next) - log_lik(cur);
    }
};
struct HierDrift {                                     // hierarchical_drift_proposal, hierarchical.rs:63-71
    static const char* name() { return "hierarchical_drift_proposal"; }
    double sd, ln_sd;
    __host__ __device__ static HierDrift make(double arg) { return HierDrift{arg, log(arg)}; }
    __device__ __forceinline__ double propose(const HierarchicalModel&, Rng64& g, const double (&cur)[4], double (&next)[4]) const {
        double za, zb;
        g.normal2(za, zb);
        const double zc = g.normal();
        const bool lin = cur[0] != 0.;
        next[0] = cur[0]; next[1] = za * sd + cur[1]; next[2] = zb * sd + cur[2];
        next[3] = lin ? cur[3] : zc * sd + cur[3];
        return assess_pair(cur, next, lin);
    }
    __device__ __forceinline__ double assess(const HierarchicalModel&, const double (&from)[4], const double (&to)[4]) const { return assess_pair(from, to, from[0] != 0.); }
    __device__ __forceinline__ double assess_pair(const double (&from)[4], const double (&to)[4], bool lin) const {
        double w = normal_lp(to[1], from[1], sd, ln_sd) + normal_lp(to[2], from[2], sd, ln_sd);
        if (!lin) w += normal_lp(to[3], from[3], sd, ln_sd);
        return w;
    }
};
struct HierAddRemove {                                 // add_or_remove_param_proposal, hierarchical.rs:48-61
    static const char* name() { return "add_or_remove_param_proposal"; }
    double sd, ln_sd;
    __host__ __device__ static HierAddRemove make(double arg) { return HierAddRemove{arg, log(arg)}; }
    __device__ __forceinline__ double propose(const HierarchicalModel& m, Rng64& g, const double (&cur)[4], double (&next)[4]) const {
        double za, zb;
        g.normal2(za, zb);
        const double zc = g.normal(), uf = g.uniform();
        const bool lin = 0.5 > uf;
        const double prev_c = cur[0] != 0. ? 0. : cur[3];
        next[0] = lin ? 1. : 0.; next[1] = za * sd + cur[1]; next[2] = zb * sd + cur[2];
        next[3] = lin ? 0. : zc * sd + prev_c;
        return assess(m, cur, next);
    }
    // log q(to | from): a, b drift, the branch flips a fair coin, c is proposed around the previous c (0 if there was none)
    __device__ __forceinline__ double assess(const HierarchicalModel&, const double (&from)[4], const double (&to)[4]) const {
        const double ln_half = -0.6931471805599453;
        double w = normal_lp(to[1], from[1], sd, ln_sd) + normal_lp(to[2], from[2], sd, ln_sd) + ln_half;
        const double prev_c = from[0] != 0. ? 0. : from[3];
        if (to[0] == 0.) w += normal_lp(to[3], prev_c, sd, ln_sd);
        return w;
    }
};

// ---- tests/pointed_model/{model,proposal}.rs (hand-coded twins of tests/dyngenfns/simple.rs:27-39) ---------------------------
struct PointedDrift;
struct PointedModel {
    static constexpr int L = 2;                       // latent (x, y)
    static constexpr bool kHasRegenerate = false;
    using Proposals = std::tuple<PointedDrift>;
    double bounds[4], prec[4], log_norm;               // uniform_2d bounds; obs covariance as precision, k ln2pi + ln det (hoisted, quirk Q8)
    double obs[2];
    __device__ __forceinline__ double obs_lp(double lx, double ly) const { return mvnormal2_logpdf(obs[0], obs[1], lx, ly, prec, log_norm); }
    __device__ __forceinline__ double generate(Rng64& g, double (&z)[L]) const {
        double u0, u1;
        g.uniform2(u0, u1);
        z[0] = u0 * (bounds[1] - bounds[0]) + bounds[0]; z[1] = u1 * (bounds[3] - bounds[2]) + bounds[2];   // types_2d.rs:23-30
        return obs_lp(z[0], z[1]);
    }
    __device__ __forceinline__ double log_joint(const double (&z)[L]) const { return uniform2d_logpdf(z[0], z[1], bounds) + obs_lp(z[0], z[1]); }
    __device__ __forceinline__ double update(const double (&cur)[L], double cur_logjp, const double (&next)[L], double& next_logjp) const {
        double v = cur_logjp;                                                 // model.rs:76-102: incremental, in this order
        v -= uniform2d_logpdf(cur[0], cur[1], bounds);
        v += uniform2d_logpdf(next[0], next[1], bounds);
        v -= obs_lp(cur[0], cur[1]);
        v += obs_lp(next[0], next[1]);
        next_logjp = v;
        return v - cur_logjp;
    }
};
struct PointedDrift {                                  // pointed_2d_drift_proposal: mvnormal(latent, s^2 I)   proposal.rs
    static const char* name() { return "pointed_2d_drift_proposal"; }
    double s_std, dprec[4], dnorm;
    __host__ __device__ static PointedDrift make(double arg) {
        PointedDrift q;
        q.s_std = arg;
        const double var = arg * arg, det = var * var - 0. * 0.;
        q.dprec[0] = var / det; q.dprec[1] = -0. / det; q.dprec[2] = -0. / det; q.dprec[3] = var / det;
        q.dnorm = 2. * 1.8378770664093453 + log(det);
        return q;
    }
    __device__ __forceinline__ double propose(const PointedModel& m, Rng64& g, const double (&cur)[2], double (&next)[2]) const {
        double z0, z1;
        g.normal2(z0, z1);
        next[0] = s_std * z0 + cur[0]; next[1] = s_std * z1 + cur[1];          // mvnormal.rs:36 (L z + mu, L = s I)
        return assess(m, cur, next);                                           // proposal.rs:24
    }
    __device__ __forceinline__ double assess(const PointedModel&, const double (&from)[2], const double (&to)[2]) const {
        return mvnormal2_logpdf(to[0], to[1], from[0], from[1], dprec, dnorm);  // proposal.rs:39
    }
};

// ---- importance sampling: n x generate(args, constraints), importance.rs:18-20 ---------------------------------------------------
template <class Model>
__global__ void __launch_bounds__(256) is_kernel(Model model, uint32_t n, uint64_t seed, uint32_t batch, double* __restrict__ latents, double* __restrict__ w) {
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        Rng64 g(seed, i, batch, P_IS);
        double z[Model::L];
        w[i] = model.generate(g, z);
        if (latents) {   // (null: the caller wants the log-ML estimate only -- the traces never leave the registers)
#pragma unroll
            for (int k = 0; k < Model::L; ++k) latents[(size_t)k * n + i] = z[k];
        }
    }
}

// importance.rs:23-25 (log-normalised weights) and :46 (probs = exp)
__global__ void __launch_bounds__(256) is_normalize_kernel(const double* __restrict__ w, uint32_t n, const DeviceStats* st, double* __restrict__ lnw, double* __restrict__ probs) {
    const double lse = (st->max == -INFINITY) ? -INFINITY : st->max + log(st->sumexp);
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        double v = w[i] - lse;
        lnw[i] = v;
        if (probs) probs[i] = exp(v);
    }
}

__global__ void __launch_bounds__(256) is_resample_search_kernel(const double* __restrict__ S, uint32_t n, uint32_t n_ret, uint64_t seed, uint32_t batch, long long* __restrict__ idx) {
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n_ret; i += gridDim.x * 256) {
        Rng64 g(seed, i, batch, P_IS_RESAMPLE);
        idx[i] = search_cumsum(S, n, g.uniform());
    }
}

// ---- Metropolis-Hastings ------------------------------------------------------------------------------------------------
struct ChainArgs {
    double* st;          // (L + 1) x n SoA: the latents, then the cached logjp
    uint64_t n, seed, offset;
    uint32_t step_base;  // move counter of every chain (RNG tag)
    unsigned long long* accepted;
};
template <int L> struct ChainState { double z[L], logjp; };
template <int L> __device__ __forceinline__ ChainState<L> chain_load(const ChainArgs& a, uint64_t i) {
    ChainState<L> s;
#pragma unroll
    for (int k = 0; k < L; ++k) s.z[k] = a.st[(uint64_t)k * a.n + i];
    s.logjp = a.st[(uint64_t)L * a.n + i];
    return s;
}
template <int L> __device__ __forceinline__ void chain_store(const ChainArgs& a, uint64_t i, const ChainState<L>& s) {
#pragma unroll
    for (int k = 0; k < L; ++k) a.st[(uint64_t)k * a.n + i] = s.z[k];
    a.st[(uint64_t)L * a.n + i] = s.logjp;
}

template <class Model>
__global__ void __launch_bounds__(256) chains_init_kernel(Model model, ChainArgs a) {
    // tests/mh.rs:34,61,91 : trace = model.generate(args, observations).0
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * 256) {
        Rng64 g(a.seed, a.offset + i, 0, P_MH_INIT);
        ChainState<Model::L> s;
        model.generate(g, s.z);
        s.logjp = model.log_joint(s.z);
        chain_store<Model::L>(a, i, s);
    }
}

__device__ __forceinline__ bool mh_accept(double alpha, uint64_t seed, uint64_t id, uint32_t step) {
    Rng64 g(seed, id, step, P_MH);
    g.skip(kAcceptBlock);
    return log(g.uniform()) < alpha;                       // mh.rs:35 / :62
}

// mh.rs:9-40: propose -> update -> assess the discard under the new trace -> accept with probability min(1, exp(alpha))
template <class Model, class Proposal>
__device__ __forceinline__ bool mh_transition(const Model& m, const Proposal& q, ChainState<Model::L>& s, uint64_t seed, uint64_t id, uint32_t step) {
    Rng64 g(seed, id, step, P_MH);
    ChainState<Model::L> p;
    const double fwd = q.propose(m, g, s.z, p.z);          // mh.rs:17-22  (proposal_choices, fwd_weight)
    const double w = m.update(s.z, s.logjp, p.z, p.logjp); // mh.rs:23-24  (new_trace, discard, weight)
    const double bwd = q.assess(m, p.z, s.z);              // mh.rs:25-33  bwd_weight: the discard scored under the new trace
    const double alpha = w - fwd + bwd;                    // mh.rs:34
    if (mh_accept(alpha, seed, id, step)) { s = p; return true; }
    return false;
}
// mh.rs:54-67
template <class Model>
__device__ __forceinline__ bool regen_transition(const Model& m, uint32_t mask, ChainState<Model::L>& s, uint64_t seed, uint64_t id, uint32_t step) {
    Rng64 g(seed, id, step, P_MH);
    ChainState<Model::L> p;
    const double alpha = m.regenerate(g, s.z, mask, p.z);  // mh.rs:60-61  (new_trace, weight)
    p.logjp = m.log_joint(p.z);
    if (mh_accept(alpha, seed, id, step)) { s = p; return true; }
    return false;
}

__device__ __forceinline__ void count_accepts(unsigned long long acc, unsigned long long* out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// a sweep = a short list of moves, each repeated `repeat` times (the loop body of tests/mh.rs:93-106 is one such list)
constexpr int kMaxScheduleMoves = 16;
struct Schedule {
    int n_moves;
    mpl_move mv[kMaxScheduleMoves];
};

// move `mv` repeated mv.repeat times: mh with the proposal the model registered at index mv.proposal, or regen_mh
template <class Model, size_t I>
__device__ __forceinline__ unsigned long long run_mh_repeat(const Model& m, const mpl_move& mv, ChainState<Model::L>& s, uint64_t seed, uint64_t id, uint32_t& step) {
    const auto q = std::tuple_element_t<I, typename Model::Proposals>::make(mv.arg);
    unsigned long long acc = 0;
    for (uint32_t k = 0; k < mv.repeat; ++k) acc += mh_transition(m, q, s, seed, id, step++) ? 1 : 0;
    return acc;
}
template <class Model, size_t... I>
__device__ __forceinline__ unsigned long long run_move(const Model& m, const mpl_move& mv, ChainState<Model::L>& s, uint64_t seed, uint64_t id, uint32_t& step, std::index_sequence<I...>) {
    unsigned long long acc = 0;
    if (mv.kind == MPL_MOVE_REGEN) {
        if constexpr (Model::kHasRegenerate)
            for (uint32_t k = 0; k < mv.repeat; ++k) acc += regen_transition(m, mv.mask, s, seed, id, step++) ? 1 : 0;
        return acc;
    }
    ((mv.proposal == (int)I ? (void)(acc += run_mh_repeat<Model, I>(m, mv, s, seed, id, step)) : (void)0), ...);
    return acc;
}

template <class Model>
__global__ void __launch_bounds__(128) mh_schedule_kernel(Model model, ChainArgs a, Schedule sched, uint32_t n_sweeps) {
    const uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long acc = 0;
    if (i < a.n) {
        ChainState<Model::L> s = chain_load<Model::L>(a, i);
        uint32_t step = a.step_base;
        for (uint32_t sw = 0; sw < n_sweeps; ++sw)
            for (int k = 0; k < sched.n_moves; ++k)
                acc += run_move(model, sched.mv[k], s, a.seed, a.offset + i, step, std::make_index_sequence<std::tuple_size<typename Model::Proposals>::value>());
        chain_store<Model::L>(a, i, s);
    }
    count_accepts(acc, a.accepted);
}

// ---- host side: the registry of static models (name -> functor built from the parameter vector and the observations) --------
static int build_regression(const mpl_model* m, const double* obs, size_t n_obs, RegressionData& d) {
    std::memset(&d, 0, sizeof d);
    const size_t n = m->params.size();
    if (n == 0 || n > (size_t)kMaxPoints) return fail(MPL_ERR_INVALID, "static regression models take 1..32 points");
    if (n_obs != n || !obs) return fail(MPL_ERR_INVALID, "one observation per regressor is required");
    d.n = (int)n;
    for (size_t i = 0; i < n; ++i) { d.xs[i] = m->params[i]; d.ys[i] = obs[i]; }
    return MPL_OK;
}
static int build_pointed(const mpl_model* m, const double* obs, size_t n_obs, PointedModel& d) {
    std::memset(&d, 0, sizeof d);
    if (n_obs != 2 || !obs) return fail(MPL_ERR_INVALID, "pointed model observes a 2-vector");
    const double* p = m->params.data();
    for (int k = 0; k < 4; ++k) d.bounds[k] = p[k];
    const double m11 = p[4], m12 = p[5], m21 = p[6], m22 = p[7];
    const double det = m11 * m22 - m21 * m12;                  // nalgebra 2x2 closed forms, hoisted (mvnormal.rs:17-18)
    if (det == 0.) return fail(MPL_ERR_INVALID, "singular observation covariance");
    d.prec[0] = m22 / det; d.prec[1] = -m12 / det; d.prec[2] = -m21 / det; d.prec[3] = m11 / det;
    d.log_norm = 2. * std::log(2. * kPi) + std::log(det);
    d.obs[0] = obs[0]; d.obs[1] = obs[1];
    return MPL_OK;
}
// calls f(functor) with the device functor of a static model; f is a generic lambda (one instantiation per registered model)
template <class F>
static int with_static_model(const mpl_model* m, const double* obs, size_t n_obs, F&& f) {
    int rc;
    switch (m->kind) {
        case M_LINE: { LineModel md; if ((rc = build_regression(m, obs, n_obs, md.d))) return rc; return f(md); }
        case M_HIER: { HierarchicalModel md; if ((rc = build_regression(m, obs, n_obs, md.d))) return rc; return f(md); }
        case M_POINTED: { PointedModel md; if ((rc = build_pointed(m, obs, n_obs, md))) return rc; return f(md); }
    }
    return fail(MPL_ERR_INVALID, "not a static (importance / MH) model");
}
template <class Model, size_t... I>
static int proposal_index_of(const char* name, std::index_sequence<I...>) {
    int found = -1;
    ((std::strcmp(name, std::tuple_element_t<I, typename Model::Proposals>::name()) == 0 ? (void)(found = (int)I) : (void)0), ...);
    return found;
}
template <class Model, size_t... I>
static const char* proposal_name_of(int idx, std::index_sequence<I...>) {
    const char* out = nullptr;
    ((idx == (int)I ? (void)(out = std::tuple_element_t<I, typename Model::Proposals>::name()) : (void)0), ...);
    return out;
}
template <class Model> using ProposalSeq = std::make_index_sequence<std::tuple_size<typename Model::Proposals>::value>;

static int require_device() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(MPL_ERR_CUDA, "no CUDA device: modppl_b200 has no CPU fallback"); }
    return MPL_OK;
}

// device workspace of the importance entry points, kept per host thread and device and only ever grown: a batch of 2^20
// proposals runs in tens of microseconds, several cudaMalloc/cudaFree pairs per call would cost more than the kernels.
// Everything is queued on the workspace's own stream; scalars come back through pinned memory.
struct IsWorkspace {
    int device = -1;
    size_t cap_n = 0, cap_lat = 0, cap_ret = 0, cap_grid = 0;
    double *lat = nullptr, *w = nullptr, *lnw = nullptr, *probs = nullptr, *cum = nullptr;
    long long* idx = nullptr;
    DeviceStats* st = nullptr;
    DeviceStats* st_host = nullptr;   // pinned
    Lse3<double>* part = nullptr;
    cudaStream_t stream = nullptr;
};

static int importance_impl(const mpl_model* m, const double* obs, size_t n_obs, uint32_t n, uint32_t n_ret, uint64_t seed, uint64_t batch,
                           double* latents, double* lnw_out, long long* idx_out, double* lml) {
    if (!m || n == 0) return fail(MPL_ERR_INVALID, "bad argument");
    int rc = require_device();
    if (rc) return rc;
    const int L = m->num_latents;
    const int grid = (int)std::min<size_t>(((size_t)n + 255) / 256, (size_t)kNumSMs * 8);
    static thread_local IsWorkspace ws;
    int dev = 0;
    MPL_CUDA_OK(cudaGetDevice(&dev));
    if (ws.device != dev) { ws = IsWorkspace(); ws.device = dev; }   // (buffers of another device are left to the driver's teardown)
    if (!ws.stream) MPL_CUDA_OK(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
    if ((size_t)L * n > ws.cap_lat) { cudaFree(ws.lat); ws.lat = nullptr; ws.cap_lat = 0; MPL_CUDA_OK(cudaMalloc(&ws.lat, (size_t)L * n * 8)); ws.cap_lat = (size_t)L * n; }
    if ((size_t)n > ws.cap_n) {
        cudaFree(ws.w); cudaFree(ws.lnw); cudaFree(ws.probs); cudaFree(ws.cum); ws.w = ws.lnw = ws.probs = ws.cum = nullptr; ws.cap_n = 0;
        MPL_CUDA_OK(cudaMalloc(&ws.w, (size_t)n * 8)); MPL_CUDA_OK(cudaMalloc(&ws.lnw, (size_t)n * 8));
        MPL_CUDA_OK(cudaMalloc(&ws.probs, (size_t)n * 8)); MPL_CUDA_OK(cudaMalloc(&ws.cum, (size_t)n * 8));
        ws.cap_n = n;
    }
    if ((size_t)n_ret > ws.cap_ret) { cudaFree(ws.idx); ws.idx = nullptr; ws.cap_ret = 0; MPL_CUDA_OK(cudaMalloc(&ws.idx, (size_t)n_ret * 8)); ws.cap_ret = n_ret; }
    if (!ws.st) { MPL_CUDA_OK(cudaMalloc(&ws.st, sizeof(DeviceStats))); MPL_CUDA_OK(cudaMallocHost(&ws.st_host, sizeof(DeviceStats))); }
    if ((size_t)grid > ws.cap_grid) { cudaFree(ws.part); ws.part = nullptr; ws.cap_grid = 0; MPL_CUDA_OK(cudaMalloc(&ws.part, (size_t)grid * sizeof(Lse3<double>))); ws.cap_grid = grid; }
    double* dprobs = n_ret ? ws.probs : nullptr;
    cudaStream_t s = ws.stream;
    MPL_CUDA_OK(cudaMemsetAsync(ws.st, 0, sizeof(DeviceStats), s));
    rc = with_static_model(m, obs, n_obs, [&](auto model) {
        is_kernel<<<grid, 256, 0, s>>>(model, n, seed, (uint32_t)batch, latents ? ws.lat : nullptr, ws.w);
        return MPL_OK;
    });
    if (rc) return rc;
    weight_reduce_kernel<double><<<grid, 256, 0, s>>>(ws.w, n, ws.st, ws.part);
    if (lnw_out || n_ret) is_normalize_kernel<<<grid, 256, 0, s>>>(ws.w, n, ws.st, ws.lnw, dprobs);   // (log-ML only: the two scalars of the reduction suffice)
    if (n_ret) {   // importance.rs:46-48: n_ret categorical draws over exp(lnw) -- the sequential-f64 running sum, reproduced in parallel
        if ((rc = launch_cumsum_exact(nullptr, dprobs, n, ws.cum, s))) return rc;
        is_resample_search_kernel<<<(n_ret + 255) / 256, 256, 0, s>>>(ws.cum, n, n_ret, seed, (uint32_t)batch, ws.idx);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws.st_host, ws.st, sizeof(DeviceStats), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && latents) e = cudaMemcpyAsync(latents, ws.lat, (size_t)L * n * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && lnw_out) e = cudaMemcpyAsync(lnw_out, ws.lnw, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && idx_out && n_ret) e = cudaMemcpyAsync(idx_out, ws.idx, (size_t)n_ret * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    const DeviceStats& h = *ws.st_host;
    if (lml) *lml = ((h.max == -INFINITY) ? -INFINITY : h.max + std::log(h.sumexp)) - std::log((double)n);   // importance.rs:21-22
    return MPL_OK;
}

}  // namespace mpl

using namespace mpl;

struct mpl_chains {
    mpl_model model;
    std::vector<double> obs;
    uint64_t n, seed, offset;
    int slots, device;
    uint32_t step;
    double* st;
    unsigned long long* accepted;
    unsigned long long* accepted_host;   // pinned
    cudaStream_t stream;
};

extern "C" int mpl_importance_sampling(const mpl_model* m, const double* obs, size_t n_obs, uint32_t num_samples, uint64_t seed, uint64_t batch,
                                       double* latents, double* log_norm_weights, double* lml) {
    return importance_impl(m, obs, n_obs, num_samples, 0, seed, batch, latents, log_norm_weights, nullptr, lml);
}
extern "C" int mpl_importance_resampling(const mpl_model* m, const double* obs, size_t n_obs, uint32_t num_samples, uint32_t num_ret_samples, uint64_t seed,
                                         uint64_t batch, double* latents, int64_t* resampled_indices, double* lml) {
    if (num_ret_samples == 0 || !resampled_indices) return fail(MPL_ERR_INVALID, "num_ret_samples must be positive");
    return importance_impl(m, obs, n_obs, num_samples, num_ret_samples, seed, batch, latents, nullptr, (long long*)resampled_indices, lml);
}

extern "C" int mpl_model_num_proposals(const mpl_model* m) {
    if (!m) return MPL_ERR_INVALID;
    switch (m->kind) {
        case M_LINE: return (int)std::tuple_size<LineModel::Proposals>::value;
        case M_HIER: return (int)std::tuple_size<HierarchicalModel::Proposals>::value;
        case M_POINTED: return (int)std::tuple_size<PointedModel::Proposals>::value;
    }
    return 0;
}
extern "C" const char* mpl_model_proposal_name(const mpl_model* m, int index) {
    if (!m) return nullptr;
    switch (m->kind) {
        case M_HIER: return proposal_name_of<HierarchicalModel>(index, ProposalSeq<HierarchicalModel>());
        case M_POINTED: return proposal_name_of<PointedModel>(index, ProposalSeq<PointedModel>());
    }
    return nullptr;
}
extern "C" int mpl_model_proposal_index(const mpl_model* m, const char* name) {
    if (!m || !name) return fail(MPL_ERR_INVALID, "null argument");
    int idx = -1;
    switch (m->kind) {
        case M_HIER: idx = proposal_index_of<HierarchicalModel>(name, ProposalSeq<HierarchicalModel>()); break;
        case M_POINTED: idx = proposal_index_of<PointedModel>(name, ProposalSeq<PointedModel>()); break;
    }
    if (idx < 0) return fail(MPL_ERR_INVALID, std::string("no proposal named '") + name + "' is registered for model '" + m->name + "'");
    return idx;
}

extern "C" mpl_chains* mpl_chains_new(const mpl_model* m, const double* obs, size_t n_obs, uint64_t n_chains, uint64_t seed, uint64_t chain_offset, int device) {
    if (!m || n_chains == 0) { fail(MPL_ERR_INVALID, "bad argument"); return nullptr; }
    if (m->num_latents <= 0) { fail(MPL_ERR_INVALID, "MH chains need a static model"); return nullptr; }
    if (require_device()) return nullptr;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { fail(MPL_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    auto* c = new mpl_chains();
    c->model = *m;
    if (obs && n_obs) c->obs.assign(obs, obs + n_obs);
    c->n = n_chains; c->seed = seed; c->offset = chain_offset; c->step = 0;
    c->slots = m->num_latents + 1;
    cudaGetDevice(&c->device);
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&c->st, (size_t)c->slots * n_chains * 8) == cudaSuccess;
    ok = ok && cudaMalloc(&c->accepted, 8) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->accepted_host, 8) == cudaSuccess;
    if (ok) {
        ChainArgs a{c->st, c->n, c->seed, c->offset, 0, c->accepted};
        const int grid = (int)std::min<uint64_t>((n_chains + 255) / 256, (uint64_t)kNumSMs * 16);
        int rc = with_static_model(&c->model, c->obs.data(), c->obs.size(), [&](auto model) {
            chains_init_kernel<<<grid, 256, 0, c->stream>>>(model, a);
            return MPL_OK;
        });
        if (rc) { mpl_chains_destroy(c); return nullptr; }
        ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess;
    }
    if (!ok) { fail(MPL_ERR_CUDA, std::string("chains allocation/init failed: ") + cudaGetErrorString(cudaGetLastError())); mpl_chains_destroy(c); return nullptr; }
    return c;
}
extern "C" void mpl_chains_destroy(mpl_chains* c) {
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->st); cudaFree(c->accepted); cudaFreeHost(c->accepted_host);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" int mpl_chains_num_slots(const mpl_chains* c) { return c ? c->slots : MPL_ERR_INVALID; }
extern "C" int mpl_chains_read(mpl_chains* c, double* host_dst, size_t bytes) {
    if (!c || !host_dst || bytes != (size_t)c->slots * c->n * 8) return fail(MPL_ERR_INVALID, "chains buffer must be double[slots*n]");
    MPL_CUDA_OK(cudaMemcpyAsync(host_dst, c->st, bytes, cudaMemcpyDeviceToHost, c->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(c->stream));
    return MPL_OK;
}
extern "C" int mpl_chains_write(mpl_chains* c, const double* host_src, size_t bytes) {
    if (!c || !host_src || bytes != (size_t)c->slots * c->n * 8) return fail(MPL_ERR_INVALID, "chains buffer must be double[slots*n]");
    MPL_CUDA_OK(cudaMemcpyAsync(c->st, host_src, bytes, cudaMemcpyHostToDevice, c->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(c->stream));
    return MPL_OK;
}

extern "C" int mpl_mh_schedule(mpl_chains* c, const mpl_move* moves, uint32_t n_moves, uint32_t n_sweeps, uint64_t* n_accepted, float* elapsed_ms) {
    if (!c || !moves) return fail(MPL_ERR_INVALID, "null argument");
    if (n_moves == 0 || n_moves > (uint32_t)kMaxScheduleMoves) return fail(MPL_ERR_INVALID, "a schedule holds 1..16 moves");
    Schedule sched;
    std::memset(&sched, 0, sizeof sched);
    sched.n_moves = (int)n_moves;
    const int n_prop = mpl_model_num_proposals(&c->model);
    uint64_t per_sweep = 0;
    for (uint32_t k = 0; k < n_moves; ++k) {
        mpl_move mv = moves[k];
        if (mv.kind == MPL_MOVE_MH) {
            if (mv.proposal < 0 || mv.proposal >= n_prop) return fail(MPL_ERR_INVALID, "schedule: proposal index out of range for this model (mpl_model_proposal_index)");
            if (!(mv.arg > 0.)) return fail(MPL_ERR_INVALID, "proposal std must be positive");
        } else if (mv.kind == MPL_MOVE_REGEN) {
            if (c->model.kind != M_HIER) return fail(MPL_ERR_UNSUPPORTED, "regen_mh: this model registers no regenerate()");
            if (mv.mask == 0) mv.mask = 15u;   // empty top-level mask regenerates everything (dyngenfn.rs:571)
        } else return fail(MPL_ERR_INVALID, "schedule: move kind must be MPL_MOVE_MH or MPL_MOVE_REGEN");
        sched.mv[k] = mv;
        per_sweep += mv.repeat;
    }
    if (per_sweep * n_sweeps > 0xffffffffull - c->step) return fail(MPL_ERR_INVALID, "move counter would overflow 32 bits");
    MPL_CUDA_OK(cudaSetDevice(c->device));
    MPL_CUDA_OK(cudaMemsetAsync(c->accepted, 0, 8, c->stream));
    ChainArgs a{c->st, c->n, c->seed, c->offset, c->step, c->accepted};
    const unsigned int grid = (unsigned int)((c->n + 127) / 128);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (elapsed_ms) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
    int rc = with_static_model(&c->model, c->obs.data(), c->obs.size(), [&](auto model) {
        mh_schedule_kernel<<<grid, 128, 0, c->stream>>>(model, a, sched, n_sweeps);
        return MPL_OK;
    });
    if (rc) return rc;
    c->step += (uint32_t)(per_sweep * n_sweeps);
    if (elapsed_ms) cudaEventRecord(e1, c->stream);
    MPL_CUDA_OK(cudaGetLastError());
    MPL_CUDA_OK(cudaMemcpyAsync(c->accepted_host, c->accepted, 8, cudaMemcpyDeviceToHost, c->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (elapsed_ms) { cudaEventElapsedTime(elapsed_ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    if (n_accepted) *n_accepted = *c->accepted_host;
    return MPL_OK;
}

extern "C" int mpl_mh(mpl_chains* c, const char* proposal, double proposal_arg, uint32_t n_steps, uint64_t* n_accepted) {
    if (!c || !proposal) return fail(MPL_ERR_INVALID, "null argument");
    const int idx = mpl_model_proposal_index(&c->model, proposal);
    if (idx < 0) return idx;
    mpl_move mv;
    mv.kind = MPL_MOVE_MH; mv.proposal = idx; mv.arg = proposal_arg; mv.mask = 0; mv.repeat = n_steps;
    return mpl_mh_schedule(c, &mv, 1, 1, n_accepted, nullptr);
}
extern "C" int mpl_regen_mh(mpl_chains* c, uint32_t mask_bits, uint32_t n_steps, uint64_t* n_accepted) {
    if (!c) return fail(MPL_ERR_INVALID, "null handle");
    mpl_move mv;
    mv.kind = MPL_MOVE_REGEN; mv.proposal = -1; mv.arg = 1.; mv.mask = mask_bits; mv.repeat = n_steps;
    return mpl_mh_schedule(c, &mv, 1, 1, n_accepted, nullptr);
}

// built-in log-densities on the device (tests/dists.rs known answers)
__global__ void logpdf_kernel(int which, const double* x, const double* p, double* out) {
    if (which == 0) *out = normal_logpdf<double>(x[0], p[0], p[1]);
    else if (which == 1) *out = bernoulli_logpdf(x[0] != 0., p[0]);
    else if (which == 2) *out = (p[0] >= p[1]) ? NAN : uniform_logpdf(x[0], p[0], p[1]);
    else if (which == 3) *out = uniform2d_logpdf(x[0], x[1], p);
    else if (which == 4) {   // mvnormal k=2 with the determinant / inverse hoisted as the models do: p = mu[2], cov[4]
        double det = p[2] * p[5] - p[4] * p[3];
        double prec[4] = {p[5] / det, -p[3] / det, -p[4] / det, p[2] / det};
        *out = mvnormal2_logpdf(x[0], x[1], p[0], p[1], prec, 2. * 1.8378770664093453 + log(det));
    }
    else if (which == 5) *out = (p[0] > p[1]) ? NAN : uniform_discrete_logpdf((long long)x[0], (long long)p[0], (long long)p[1]);
    else if (which == 6) *out = geometric_logpdf((long long)x[0], p[0]);
    else if (which == 7) *out = poisson_logpdf((long long)x[0], p[0]);
    else if (which == 8) *out = beta_logpdf(x[0], p[0], p[1]);
    else if (which == 9) *out = gamma_logpdf(x[0], p[0], p[1]);
    else if (which >= 32) {  // mvnormal.rs:14-22 for k = which - 32: p = mu[k], cov[k*k] (row-major)
        const int k = which - 32;
        *out = mvnormal_logpdf_k(x, p, p + k, k);
    }
    else if (which >= 16) {   // categorical.rs:13-20 with K = which - 16 probabilities
        const long long k = (long long)x[0];
        *out = k < 0 ? NAN : (k < (long long)(which - 16) ? log(p[k]) : -INFINITY);
    }
}
extern "C" int mpl_logpdf(const char* dist, const double* x, const double* params, size_t n_params, double* out) {
    if (!dist || !x || !params || !out) return fail(MPL_ERR_INVALID, "null argument");
    std::string s(dist);
    int which = s == "normal" ? 0 : s == "bernoulli" ? 1 : s == "uniform" ? 2 : s == "uniform_2d" ? 3 : s == "mvnormal2" ? 4
              : s == "uniform_discrete" ? 5 : s == "geometric" ? 6 : s == "poisson" ? 7 : s == "beta" ? 8 : s == "gamma" ? 9 : s == "categorical" ? 16
              : s == "mvnormal" ? 32 : -1;
    static const size_t kNeed[10] = {2, 1, 2, 4, 6, 2, 1, 1, 2, 2};
    size_t nx = (which == 3 || which == 4) ? 2 : 1;
    if (which == 16) {   // the probability vector is the parameter list (at most 8 categories)
        if (n_params < 1 || n_params > 8) return fail(MPL_ERR_INVALID, "categorical: 1..8 probabilities");
        which = 16 + (int)n_params;
    } else if (which == 32) {   // x has k entries, params = mu[k] followed by cov[k*k]
        int k = 0;
        for (int c = 1; c <= kMvnMaxK; ++c) if ((size_t)(c + c * c) == n_params) k = c;
        if (!k) return fail(MPL_ERR_INVALID, "mvnormal: params = mu[k], cov[k*k] with 1 <= k <= 8");
        which = 32 + k; nx = (size_t)k;
    } else if (which < 0 || n_params != kNeed[which]) return fail(MPL_ERR_INVALID, "unknown distribution or wrong parameter count");
    int rc = require_device();
    if (rc) return rc;
    constexpr size_t kX = kMvnMaxK, kP = kMvnMaxK + kMvnMaxK * kMvnMaxK;
    double* d = nullptr;
    MPL_CUDA_OK(cudaMalloc(&d, (kX + kP + 1) * 8));
    double h[kX + kP + 1] = {0};
    for (size_t i = 0; i < nx; ++i) h[i] = x[i];
    for (size_t i = 0; i < n_params; ++i) h[kX + i] = params[i];
    MPL_CUDA_OK(cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
    logpdf_kernel<<<1, 1>>>(which, d, d + kX, d + kX + kP);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d + kX + kP, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    return MPL_OK;
}

// multi-GPU peer attach: see multi_gpu.cu
