// is_mh.cu -- batched importance sampling (reference modppl/src/inference/importance.rs:12-51) and many-chain
// Metropolis-Hastings (src/inference/mh.rs:9-76) for the flattened fixture models of
// tests/dyngenfns/{simple,hierarchical}.rs and tests/pointed_model/*.rs.  One thread per proposal / chain; all state in
// registers; fp64 like the reference.  These paths are register-resident (FP64/SFU bound), not HBM bound.
#include <cmath>
#include <cstring>
#include "engine.h"

namespace mpl {

constexpr int kMaxPoints = 32;
struct StaticData {
    int kind, n;
    double xs[kMaxPoints], ys[kMaxPoints];   // line / hierarchical: regressors and observations
    double bounds[4], prec[4], log_norm;     // pointed: uniform_2d bounds; obs covariance as precision + k ln2pi + ln det
    double obs[2];
};

__device__ __forceinline__ double hier_loglik(const StaticData& d, bool L, double a, double b, double c) {
    double w = 0.;
    const double ln_noise = -2.3025850929940455;   // ln 0.1
    for (int i = 0; i < d.n; ++i) {
        double x = d.xs[i];
        double mean = L ? a + b * x : a + b * x + c * x * x;              // hierarchical.rs:36-44
        double z = (d.ys[i] - mean) / 0.1;
        w += -(z * z + 1.8378770664093453) / 2. - ln_noise;              // normal.rs:13-17
    }
    return w;
}
__device__ __forceinline__ double std_normal_logpdf(double x) { return -(x * x + 1.8378770664093453) / 2. - 0.; }
__device__ __forceinline__ double normal_lp(double x, double mu, double sd, double ln_sd) { double z = (x - mu) / sd; return -(z * z + 1.8378770664093453) / 2. - ln_sd; }
__device__ __forceinline__ double hier_logjp(const StaticData& d, bool L, double a, double b, double c) {
    double lp = log(L ? 0.7 : 1. - 0.7) + std_normal_logpdf(a) + std_normal_logpdf(b);   // bernoulli.rs:12-14
    if (!L) lp += std_normal_logpdf(c);
    return lp + hier_loglik(d, L, a, b, c);
}
__device__ __forceinline__ double pointed_obs_lp(const StaticData& d, double lx, double ly) {
    return mvnormal2_logpdf(d.obs[0], d.obs[1], lx, ly, d.prec, d.log_norm);
}

// ---- importance sampling: n x generate(args, constraints) ------------------------------------------------------------
__global__ void __launch_bounds__(256) is_kernel(StaticData d, uint32_t n, uint64_t seed, uint32_t batch, double* __restrict__ latents, double* __restrict__ w) {
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        Rng64 g(seed, i, batch, P_IS);
        if (d.kind == M_LINE) {
            double z0, z1;
            g.normal2(z0, z1);
            double slope = z0 * 1. + 0., intercept = z1 * 2. + 0.;   // simple.rs:12-13
            double ww = 0.;
            const double ln_noise = -2.3025850929940455;
            for (int j = 0; j < d.n; ++j) { double z = (d.ys[j] - (slope * d.xs[j] + intercept)) / 0.1; ww += -(z * z + 1.8378770664093453) / 2. - ln_noise; }
            latents[i] = slope; latents[(size_t)n + i] = intercept; w[i] = ww;
        } else if (d.kind == M_HIER) {
            bool L = 0.7 > g.uniform();   // bernoulli.rs:16-18
            double a, b;
            g.normal2(a, b);
            double c = g.normal();
            if (L) c = 0.;
            latents[i] = L ? 1. : 0.; latents[(size_t)n + i] = a; latents[2 * (size_t)n + i] = b; latents[3 * (size_t)n + i] = c;
            w[i] = hier_loglik(d, L, a, b, c);
        } else {
            double u0, u1;
            g.uniform2(u0, u1);
            double lx = u0 * (d.bounds[1] - d.bounds[0]) + d.bounds[0], ly = u1 * (d.bounds[3] - d.bounds[2]) + d.bounds[2];   // types_2d.rs:23-30
            latents[i] = lx; latents[(size_t)n + i] = ly;
            w[i] = pointed_obs_lp(d, lx, ly);
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
    double* st;          // slots x n SoA
    uint64_t n, seed, offset;
    uint32_t step_base;  // move counter of every chain (RNG tag)
    unsigned long long* accepted;
};

__global__ void __launch_bounds__(256) chains_init_kernel(StaticData d, ChainArgs a) {
    // tests/mh.rs:34,61,91 : trace = model.generate(args, observations).0
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * 256) {
        Rng64 g(a.seed, a.offset + i, 0, P_MH_INIT);
        if (d.kind == M_HIER) {
            bool L = 0.7 > g.uniform();
            double aa, bb;
            g.normal2(aa, bb);
            double cc = g.normal();
            if (L) cc = 0.;
            a.st[i] = L ? 1. : 0.; a.st[a.n + i] = aa; a.st[2 * a.n + i] = bb; a.st[3 * a.n + i] = cc;
            a.st[4 * a.n + i] = hier_logjp(d, L, aa, bb, cc);
        } else {
            double u0, u1;
            g.uniform2(u0, u1);
            double lx = u0 * (d.bounds[1] - d.bounds[0]) + d.bounds[0], ly = u1 * (d.bounds[3] - d.bounds[2]) + d.bounds[2];
            a.st[i] = lx; a.st[a.n + i] = ly;
            a.st[2 * a.n + i] = uniform2d_logpdf(lx, ly, d.bounds) + pointed_obs_lp(d, lx, ly);
        }
    }
}

struct HierState {
    bool L;
    double a, b, c, logjp;
};

// One transition (mh.rs:9-40 / :54-67) of the flattened hierarchical model; weight algebra per SURVEY.md section 3.4.
__device__ __forceinline__ bool hier_transition(const StaticData& d, HierState& s, int move, double parg, double ln_parg, uint32_t mask, uint64_t seed, uint64_t id, uint32_t step) {
    Rng64 g(seed, id, step, P_MH);
    double za, zb;
    g.normal2(za, zb);           // block 0
    double zc = g.normal();      // block 1
    double uf = g.uniform();     // block 2
    double ua = g.uniform();     // block 3
    HierState p = s;
    double alpha;
    if (move == MPL_MOVE_HIER_DRIFT) {                     // hierarchical.rs:63-71
        p.a = za * parg + s.a; p.b = zb * parg + s.b;
        if (!s.L) p.c = zc * parg + s.c;
        p.logjp = hier_logjp(d, p.L, p.a, p.b, p.c);
        double w = p.logjp - s.logjp;                      // update weight = delta logjp
        double fwd = normal_lp(p.a, s.a, parg, ln_parg) + normal_lp(p.b, s.b, parg, ln_parg);
        double bwd = normal_lp(s.a, p.a, parg, ln_parg) + normal_lp(s.b, p.b, parg, ln_parg);
        if (!s.L) { fwd += normal_lp(p.c, s.c, parg, ln_parg); bwd += normal_lp(s.c, p.c, parg, ln_parg); }
        alpha = w - fwd + bwd;                             // mh.rs:34
    } else if (move == MPL_MOVE_HIER_ADD_REMOVE) {         // hierarchical.rs:48-61
        p.a = za * parg + s.a; p.b = zb * parg + s.b;
        p.L = 0.5 > uf;
        double prev_c = s.L ? 0. : s.c;
        p.c = p.L ? 0. : zc * parg + prev_c;
        p.logjp = hier_logjp(d, p.L, p.a, p.b, p.c);
        double w = p.logjp - s.logjp;
        const double ln_half = -0.6931471805599453;
        double fwd = normal_lp(p.a, s.a, parg, ln_parg) + normal_lp(p.b, s.b, parg, ln_parg) + ln_half;
        if (!p.L) fwd += normal_lp(p.c, prev_c, parg, ln_parg);
        double prev_c_bwd = p.L ? 0. : p.c;
        double bwd = normal_lp(s.a, p.a, parg, ln_parg) + normal_lp(s.b, p.b, parg, ln_parg) + ln_half;
        if (!s.L) bwd += normal_lp(s.c, prev_c_bwd, parg, ln_parg);
        alpha = w - fwd + bwd;
    } else {                                               // regen_mh: dyngenfn.rs:223-266
        p.L = (mask & 8u) ? (0.7 > uf) : s.L;
        if (mask & 1u) p.a = za * 1. + 0.;
        if (mask & 2u) p.b = zb * 1. + 0.;
        if (p.L) p.c = 0.;
        else if (s.L) p.c = zc * 1. + 0.;
        else if (mask & 4u) p.c = zc * 1. + 0.;
        alpha = hier_loglik(d, p.L, p.a, p.b, p.c) - hier_loglik(d, s.L, s.a, s.b, s.c);
        p.logjp = hier_logjp(d, p.L, p.a, p.b, p.c);
    }
    if (log(ua) < alpha) { s = p; return true; }           // mh.rs:35 / :62
    return false;
}

__device__ __forceinline__ void count_accepts(unsigned long long acc, unsigned long long* out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(128) mh_hier_kernel(StaticData d, ChainArgs a, int move, double parg, uint32_t mask, uint32_t n_steps) {
    const uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long acc = 0;
    if (i < a.n) {
        HierState s{a.st[i] != 0., a.st[a.n + i], a.st[2 * a.n + i], a.st[3 * a.n + i], a.st[4 * a.n + i]};
        const double ln_parg = log(parg);
        for (uint32_t k = 0; k < n_steps; ++k) acc += hier_transition(d, s, move, parg, ln_parg, mask, a.seed, a.offset + i, a.step_base + k) ? 1 : 0;
        a.st[i] = s.L ? 1. : 0.; a.st[a.n + i] = s.a; a.st[2 * a.n + i] = s.b; a.st[3 * a.n + i] = s.c; a.st[4 * a.n + i] = s.logjp;
    }
    count_accepts(acc, a.accepted);
}

// tests/mh.rs:93-106: per sweep 1 add/remove(.025) + 3 drift(.1) + 10 drift(.01)
__global__ void __launch_bounds__(128) mh_hier_sweep_kernel(StaticData d, ChainArgs a, uint32_t n_sweeps) {
    const uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long acc = 0;
    if (i < a.n) {
        HierState s{a.st[i] != 0., a.st[a.n + i], a.st[2 * a.n + i], a.st[3 * a.n + i], a.st[4 * a.n + i]};
        const double l025 = log(0.025), l1 = log(0.1), l01 = log(0.01);
        uint32_t step = a.step_base;
        for (uint32_t sw = 0; sw < n_sweeps; ++sw) {
            acc += hier_transition(d, s, MPL_MOVE_HIER_ADD_REMOVE, 0.025, l025, 0, a.seed, a.offset + i, step++) ? 1 : 0;
            for (int k = 0; k < 3; ++k) acc += hier_transition(d, s, MPL_MOVE_HIER_DRIFT, 0.1, l1, 0, a.seed, a.offset + i, step++) ? 1 : 0;
            for (int k = 0; k < 10; ++k) acc += hier_transition(d, s, MPL_MOVE_HIER_DRIFT, 0.01, l01, 0, a.seed, a.offset + i, step++) ? 1 : 0;
        }
        a.st[i] = s.L ? 1. : 0.; a.st[a.n + i] = s.a; a.st[2 * a.n + i] = s.b; a.st[3 * a.n + i] = s.c; a.st[4 * a.n + i] = s.logjp;
    }
    count_accepts(acc, a.accepted);
}

__global__ void __launch_bounds__(128) mh_pointed_kernel(StaticData d, ChainArgs a, double s_std, uint32_t n_steps) {
    // mh.rs:9-40 with tests/pointed_model/{model,proposal}.rs; drift covariance s^2 I
    const uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long acc = 0;
    if (i < a.n) {
        double lx = a.st[i], ly = a.st[a.n + i], logjp = a.st[2 * a.n + i];
        const double var = s_std * s_std;
        const double det = var * var - 0. * 0.;
        const double dprec[4] = {var / det, -0. / det, -0. / det, var / det};
        const double dnorm = 2. * 1.8378770664093453 + log(det);
        for (uint32_t k = 0; k < n_steps; ++k) {
            Rng64 g(a.seed, a.offset + i, a.step_base + k, P_MH);
            double z0, z1;
            g.normal2(z0, z1);
            double nx = s_std * z0 + lx, ny = s_std * z1 + ly;                    // mvnormal.rs:36
            double fwd = mvnormal2_logpdf(nx, ny, lx, ly, dprec, dnorm);           // proposal.rs:24
            double new_logjp = logjp;                                              // model.rs:76-102
            new_logjp -= uniform2d_logpdf(lx, ly, d.bounds);
            new_logjp += uniform2d_logpdf(nx, ny, d.bounds);
            new_logjp -= pointed_obs_lp(d, lx, ly);
            new_logjp += pointed_obs_lp(d, nx, ny);
            double w = new_logjp - logjp;
            double bwd = mvnormal2_logpdf(lx, ly, nx, ny, dprec, dnorm);           // proposal.rs:39
            double alpha = w - fwd + bwd;
            g.skip(2);
            double u = g.uniform();
            if (log(u) < alpha) { lx = nx; ly = ny; logjp = new_logjp; acc++; }
        }
        a.st[i] = lx; a.st[a.n + i] = ly; a.st[2 * a.n + i] = logjp;
    }
    count_accepts(acc, a.accepted);
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
static int build_static(const mpl_model* m, const double* obs, size_t n_obs, StaticData& d) {
    std::memset(&d, 0, sizeof d);
    d.kind = m->kind;
    if (m->kind == M_LINE || m->kind == M_HIER) {
        size_t n = m->params.size();
        if (n == 0 || n > (size_t)kMaxPoints) return fail(MPL_ERR_INVALID, "static regression models take 1..32 points");
        if (n_obs != n || !obs) return fail(MPL_ERR_INVALID, "one observation per regressor is required");
        d.n = (int)n;
        for (size_t i = 0; i < n; ++i) { d.xs[i] = m->params[i]; d.ys[i] = obs[i]; }
    } else if (m->kind == M_POINTED) {
        if (n_obs != 2 || !obs) return fail(MPL_ERR_INVALID, "pointed model observes a 2-vector");
        const double* p = m->params.data();
        for (int k = 0; k < 4; ++k) d.bounds[k] = p[k];
        double m11 = p[4], m12 = p[5], m21 = p[6], m22 = p[7];
        double det = m11 * m22 - m21 * m12;                  // nalgebra 2x2 closed forms, hoisted (mvnormal.rs:17-18)
        if (det == 0.) return fail(MPL_ERR_INVALID, "singular observation covariance");
        d.prec[0] = m22 / det; d.prec[1] = -m12 / det; d.prec[2] = -m21 / det; d.prec[3] = m11 / det;
        d.log_norm = 2. * std::log(2. * kPi) + std::log(det);
        d.obs[0] = obs[0]; d.obs[1] = obs[1];
    } else return fail(MPL_ERR_INVALID, "not a static (importance / MH) model");
    return MPL_OK;
}

static int require_device() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(MPL_ERR_CUDA, "no CUDA device: modppl_b200 has no CPU fallback"); }
    return MPL_OK;
}

static int importance_impl(const mpl_model* m, const double* obs, size_t n_obs, uint32_t n, uint32_t n_ret, uint64_t seed, uint64_t batch,
                           double* latents, double* lnw_out, long long* idx_out, double* lml) {
    if (!m || n == 0) return fail(MPL_ERR_INVALID, "bad argument");
    int rc = require_device();
    if (rc) return rc;
    StaticData d;
    if ((rc = build_static(m, obs, n_obs, d))) return rc;
    const int L = m->num_latents;
    const int grid = (int)std::min<size_t>(((size_t)n + 255) / 256, (size_t)kNumSMs * 8);
    // device workspace, kept per host thread and device and only ever grown: a batch of 2^20 proposals runs in tens of
    // microseconds, several cudaMalloc/cudaFree pairs per call would cost more than the kernels
    struct Workspace { int device = -1; size_t cap_n = 0, cap_lat = 0, cap_ret = 0, cap_grid = 0;
                       double *lat = nullptr, *w = nullptr, *lnw = nullptr, *probs = nullptr, *cum = nullptr; long long* idx = nullptr;
                       DeviceStats* st = nullptr; Lse3<double>* part = nullptr; };
    static thread_local Workspace ws;
    int dev = 0;
    MPL_CUDA_OK(cudaGetDevice(&dev));
    if (ws.device != dev) { ws = Workspace(); ws.device = dev; }   // (buffers of another device are left to the driver's teardown)
    if ((size_t)L * n > ws.cap_lat) { cudaFree(ws.lat); ws.lat = nullptr; ws.cap_lat = 0; MPL_CUDA_OK(cudaMalloc(&ws.lat, (size_t)L * n * 8)); ws.cap_lat = (size_t)L * n; }
    if ((size_t)n > ws.cap_n) {
        cudaFree(ws.w); cudaFree(ws.lnw); cudaFree(ws.probs); cudaFree(ws.cum); ws.w = ws.lnw = ws.probs = ws.cum = nullptr; ws.cap_n = 0;
        MPL_CUDA_OK(cudaMalloc(&ws.w, (size_t)n * 8)); MPL_CUDA_OK(cudaMalloc(&ws.lnw, (size_t)n * 8));
        MPL_CUDA_OK(cudaMalloc(&ws.probs, (size_t)n * 8)); MPL_CUDA_OK(cudaMalloc(&ws.cum, (size_t)n * 8));
        ws.cap_n = n;
    }
    if ((size_t)n_ret > ws.cap_ret) { cudaFree(ws.idx); ws.idx = nullptr; ws.cap_ret = 0; MPL_CUDA_OK(cudaMalloc(&ws.idx, (size_t)n_ret * 8)); ws.cap_ret = n_ret; }
    if (!ws.st) MPL_CUDA_OK(cudaMalloc(&ws.st, sizeof(DeviceStats)));
    if ((size_t)grid > ws.cap_grid) { cudaFree(ws.part); ws.part = nullptr; ws.cap_grid = 0; MPL_CUDA_OK(cudaMalloc(&ws.part, (size_t)grid * sizeof(Lse3<double>))); ws.cap_grid = grid; }
    double *dlat = ws.lat, *dw = ws.w, *dlnw = ws.lnw, *dprobs = n_ret ? ws.probs : nullptr, *dcum = ws.cum;
    long long* didx = ws.idx;
    DeviceStats* st = ws.st; Lse3<double>* part = ws.part;
    MPL_CUDA_OK(cudaMemsetAsync(st, 0, sizeof(DeviceStats)));
    is_kernel<<<grid, 256>>>(d, n, seed, (uint32_t)batch, dlat, dw);
    weight_reduce_kernel<double><<<grid, 256>>>(dw, n, st, part);
    is_normalize_kernel<<<grid, 256>>>(dw, n, st, dlnw, dprobs);
    if (n_ret) {
        cumsum_seq_kernel<<<1, 256>>>(dprobs, n, dcum);
        is_resample_search_kernel<<<(n_ret + 255) / 256, 256>>>(dcum, n, n_ret, seed, (uint32_t)batch, didx);
    }
    cudaError_t e = cudaGetLastError();
    DeviceStats h;
    if (e == cudaSuccess) e = cudaMemcpy(&h, st, sizeof h, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && latents) e = cudaMemcpy(latents, dlat, (size_t)L * n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && lnw_out) e = cudaMemcpy(lnw_out, dlnw, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && idx_out && n_ret) e = cudaMemcpy(idx_out, didx, (size_t)n_ret * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    if (lml) *lml = ((h.max == -INFINITY) ? -INFINITY : h.max + std::log(h.sumexp)) - std::log((double)n);   // importance.rs:21-22
    return MPL_OK;
}

}  // namespace mpl

using namespace mpl;

struct mpl_chains {
    mpl_model model;
    StaticData data;
    uint64_t n, seed, offset;
    int slots, device;
    uint32_t step;
    double* st;
    unsigned long long* accepted;
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

extern "C" mpl_chains* mpl_chains_new(const mpl_model* m, const double* obs, size_t n_obs, uint64_t n_chains, uint64_t seed, uint64_t chain_offset, int device) {
    if (!m || n_chains == 0) { fail(MPL_ERR_INVALID, "bad argument"); return nullptr; }
    if (m->kind != M_HIER && m->kind != M_POINTED) { fail(MPL_ERR_INVALID, "MH chains: hierarchical or pointed model"); return nullptr; }
    if (require_device()) return nullptr;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { fail(MPL_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    auto* c = new mpl_chains();
    c->model = *m;
    if (build_static(m, obs, n_obs, c->data)) { delete c; return nullptr; }
    c->n = n_chains; c->seed = seed; c->offset = chain_offset; c->step = 0;
    c->slots = m->kind == M_HIER ? 5 : 3;
    cudaGetDevice(&c->device);
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&c->st, (size_t)c->slots * n_chains * 8) == cudaSuccess;
    ok = ok && cudaMalloc(&c->accepted, 8) == cudaSuccess;
    if (ok) {
        ChainArgs a{c->st, c->n, c->seed, c->offset, 0, c->accepted};
        int grid = (int)std::min<uint64_t>((n_chains + 255) / 256, (uint64_t)kNumSMs * 16);
        chains_init_kernel<<<grid, 256, 0, c->stream>>>(c->data, a);
        ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess;
    }
    if (!ok) { fail(MPL_ERR_CUDA, std::string("chains allocation/init failed: ") + cudaGetErrorString(cudaGetLastError())); mpl_chains_destroy(c); return nullptr; }
    return c;
}
extern "C" void mpl_chains_destroy(mpl_chains* c) {
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->st); cudaFree(c->accepted);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" int mpl_chains_num_slots(const mpl_chains* c) { return c ? c->slots : MPL_ERR_INVALID; }
extern "C" int mpl_chains_read(mpl_chains* c, double* host_dst, size_t bytes) {
    if (!c || !host_dst || bytes != (size_t)c->slots * c->n * 8) return fail(MPL_ERR_INVALID, "chains buffer must be double[slots*n]");
    MPL_CUDA_OK(cudaStreamSynchronize(c->stream));
    MPL_CUDA_OK(cudaMemcpy(host_dst, c->st, bytes, cudaMemcpyDeviceToHost));
    return MPL_OK;
}
extern "C" int mpl_chains_write(mpl_chains* c, const double* host_src, size_t bytes) {
    if (!c || !host_src || bytes != (size_t)c->slots * c->n * 8) return fail(MPL_ERR_INVALID, "chains buffer must be double[slots*n]");
    MPL_CUDA_OK(cudaStreamSynchronize(c->stream));
    MPL_CUDA_OK(cudaMemcpy(c->st, host_src, bytes, cudaMemcpyHostToDevice));
    return MPL_OK;
}

static int chains_launch(mpl_chains* c, int move, double parg, uint32_t mask, uint32_t n_steps, uint32_t n_sweeps, uint64_t* n_accepted, float* elapsed_ms) {
    MPL_CUDA_OK(cudaSetDevice(c->device));
    MPL_CUDA_OK(cudaMemsetAsync(c->accepted, 0, 8, c->stream));
    ChainArgs a{c->st, c->n, c->seed, c->offset, c->step, c->accepted};
    const unsigned int grid = (unsigned int)((c->n + 127) / 128);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (elapsed_ms) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
    if (n_sweeps) { mh_hier_sweep_kernel<<<grid, 128, 0, c->stream>>>(c->data, a, n_sweeps); c->step += 14 * n_sweeps; }
    else if (move == MPL_MOVE_POINTED_DRIFT) { mh_pointed_kernel<<<grid, 128, 0, c->stream>>>(c->data, a, parg, n_steps); c->step += n_steps; }
    else { mh_hier_kernel<<<grid, 128, 0, c->stream>>>(c->data, a, move, parg, mask, n_steps); c->step += n_steps; }
    if (elapsed_ms) cudaEventRecord(e1, c->stream);
    MPL_CUDA_OK(cudaGetLastError());
    unsigned long long acc = 0;
    MPL_CUDA_OK(cudaMemcpyAsync(&acc, c->accepted, 8, cudaMemcpyDeviceToHost, c->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (elapsed_ms) { cudaEventElapsedTime(elapsed_ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    if (n_accepted) *n_accepted = acc;
    return MPL_OK;
}

extern "C" int mpl_mh(mpl_chains* c, int move, double proposal_arg, uint32_t n_steps, uint64_t* n_accepted) {
    if (!c) return fail(MPL_ERR_INVALID, "null handle");
    if (move == MPL_MOVE_POINTED_DRIFT) { if (c->model.kind != M_POINTED) return fail(MPL_ERR_INVALID, "pointed drift needs the pointed model"); }
    else if (move == MPL_MOVE_HIER_DRIFT || move == MPL_MOVE_HIER_ADD_REMOVE) { if (c->model.kind != M_HIER) return fail(MPL_ERR_INVALID, "hierarchical proposal needs the hierarchical model"); }
    else return fail(MPL_ERR_INVALID, "unknown proposal");
    if (!(proposal_arg > 0.)) return fail(MPL_ERR_INVALID, "proposal std must be positive");
    return chains_launch(c, move, proposal_arg, 0, n_steps, 0, n_accepted, nullptr);
}
extern "C" int mpl_regen_mh(mpl_chains* c, uint32_t mask_bits, uint32_t n_steps, uint64_t* n_accepted) {
    if (!c) return fail(MPL_ERR_INVALID, "null handle");
    if (c->model.kind != M_HIER) return fail(MPL_ERR_UNSUPPORTED, "regen_mh: hierarchical model only");
    if (mask_bits == 0) mask_bits = 15u;   // empty top-level mask regenerates everything (dyngenfn.rs:571)
    return chains_launch(c, MPL_MOVE_HIER_REGEN, 1., mask_bits, n_steps, 0, n_accepted, nullptr);
}
extern "C" int mpl_mh_hier_sweeps(mpl_chains* c, uint32_t n_sweeps, uint64_t* n_accepted, float* elapsed_ms) {
    if (!c || c->model.kind != M_HIER || n_sweeps == 0) return fail(MPL_ERR_INVALID, "bad argument");
    return chains_launch(c, 0, 0., 0, 0, n_sweeps, n_accepted, elapsed_ms);
}

// built-in log-densities on the device (tests/dists.rs known answers)
__global__ void logpdf_kernel(int which, const double* x, const double* p, double* out) {
    if (which == 0) *out = normal_logpdf<double>(x[0], p[0], p[1]);
    else if (which == 1) *out = bernoulli_logpdf(x[0] != 0., p[0]);
    else if (which == 2) *out = (p[0] >= p[1]) ? NAN : uniform_logpdf(x[0], p[0], p[1]);
    else if (which == 3) *out = uniform2d_logpdf(x[0], x[1], p);
    else if (which == 4) {   // mvnormal k=2: p = mu[2], cov[4]
        double det = p[2] * p[5] - p[4] * p[3];
        double prec[4] = {p[5] / det, -p[3] / det, -p[4] / det, p[2] / det};
        *out = mvnormal2_logpdf(x[0], x[1], p[0], p[1], prec, 2. * 1.8378770664093453 + log(det));
    }
    else if (which == 5) *out = (p[0] > p[1]) ? NAN : uniform_discrete_logpdf((long long)x[0], (long long)p[0], (long long)p[1]);
    else if (which == 6) *out = geometric_logpdf((long long)x[0], p[0]);
    else if (which == 7) *out = poisson_logpdf((long long)x[0], p[0]);
    else if (which == 8) *out = beta_logpdf(x[0], p[0], p[1]);
    else if (which == 9) *out = gamma_logpdf(x[0], p[0], p[1]);
    else if (which >= 16) {   // categorical.rs:13-20 with K = which - 16 probabilities
        const long long k = (long long)x[0];
        *out = k < 0 ? NAN : (k < (long long)(which - 16) ? log(p[k]) : -INFINITY);
    }
}
extern "C" int mpl_logpdf(const char* dist, const double* x, const double* params, size_t n_params, double* out) {
    if (!dist || !x || !params || !out) return fail(MPL_ERR_INVALID, "null argument");
    std::string s(dist);
    int which = s == "normal" ? 0 : s == "bernoulli" ? 1 : s == "uniform" ? 2 : s == "uniform_2d" ? 3 : s == "mvnormal2" ? 4
              : s == "uniform_discrete" ? 5 : s == "geometric" ? 6 : s == "poisson" ? 7 : s == "beta" ? 8 : s == "gamma" ? 9 : s == "categorical" ? 16 : -1;
    static const size_t kNeed[10] = {2, 1, 2, 4, 6, 2, 1, 1, 2, 2};
    if (which == 16) {   // the probability vector is the parameter list (at most 8 categories)
        if (n_params < 1 || n_params > 8) return fail(MPL_ERR_INVALID, "categorical: 1..8 probabilities");
        which = 16 + (int)n_params;
    } else if (which < 0 || n_params != kNeed[which]) return fail(MPL_ERR_INVALID, "unknown distribution or wrong parameter count");
    int rc = require_device();
    if (rc) return rc;
    double* d = nullptr;
    MPL_CUDA_OK(cudaMalloc(&d, (2 + 8 + 1) * 8));
    double h[11] = {0};
    h[0] = x[0]; h[1] = (which == 3 || which == 4) ? x[1] : 0.;
    for (size_t i = 0; i < n_params; ++i) h[2 + i] = params[i];
    MPL_CUDA_OK(cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
    logpdf_kernel<<<1, 1>>>(which, d, d + 2, d + 10);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d + 10, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    return MPL_OK;
}

// multi-GPU peer attach: see multi_gpu.cu
