// modppl_oracle.cpp -- CPU restatement of modppl's inference hot path (see modppl_oracle.h).
// TEST INFRASTRUCTURE ONLY: never linked or imported by the product (modppl_b200/).
// Single-threaded fp64 unless stated; compile with -ffp-contract=off so the sequential
// roundings are the ones the Rust reference performs.
//
// Citations are relative to /root/reference/modppl/.
#include "modppl_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <algorithm>
#include <functional>
#include <thread>

// Host threads for the embarrassingly parallel loops (per-particle propagation, per-chunk quantisation, gather).  Results do
// not depend on it: the RNG is counter-based per particle and every integer sum is taken by one thread per chunk.
static int g_threads = 1;
extern "C" void mo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
extern "C" int mo_get_threads() { return g_threads; }
static void parallel_for(size_t n, const std::function<void(size_t, size_t)>& body) {
    const size_t nt = std::min<size_t>((size_t)g_threads, std::max<size_t>(1, n / 4096));
    if (nt <= 1) { body(0, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (size_t k = 0; k < nt; ++k) {
        const size_t lo = k * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([&body, lo, hi] { body(lo, hi); });
    }
    for (auto& t : th) t.join();
}

namespace {
const double kPi = 3.14159265358979323846;
const double kNegInf = -std::numeric_limits<double>::infinity();
typedef unsigned __int128 u128;
}  // namespace

// ============================================================================================
// src/lib.rs:34-45  logsumexp
// ============================================================================================
extern "C" double mo_logsumexp(const double* xs, size_t n) {
    double mx = kNegInf;                                  // fold(-inf, f64::max): NaN operands are ignored by f64::max
    for (size_t i = 0; i < n; ++i) mx = std::fmax(mx, xs[i]);
    if (mx == kNegInf) return kNegInf;
    double sum_exp = 0.;
    for (size_t i = 0; i < n; ++i) sum_exp += std::exp(xs[i] - mx);
    return mx + std::log(sum_exp);
}

// ============================================================================================
// src/modeling/dists
// ============================================================================================
extern "C" double mo_normal_logpdf(double x, double mu, double sd) {
    // normal.rs:13-17   -(z.abs().powf(2.) + (2.*PI).ln())/2. - std.ln()
    double z = (x - mu) / sd;
    return -(std::pow(std::fabs(z), 2.) + std::log(2. * kPi)) / 2. - std::log(sd);
}

extern "C" double mo_bernoulli_logpdf(int a, double p) {
    // bernoulli.rs:12-14
    return std::log(a ? p : 1. - p);
}

extern "C" double mo_uniform_logpdf(double x, double a, double b) {
    // uniform.rs:22-26 ; check_bounds panics when a >= b -> NaN here
    if (a >= b) return std::numeric_limits<double>::quiet_NaN();
    return (a <= x && x <= b) ? -std::log(b - a) : kNegInf;
}

extern "C" double mo_uniform_discrete_logpdf(int64_t x, int64_t a, int64_t b) {
    // uniform.rs:43-47 ; check_bounds panics when a > b -> NaN here
    if (a > b) return std::numeric_limits<double>::quiet_NaN();
    return (a <= x && x <= b) ? -std::log((double)(b - a + 1)) : kNegInf;
}

extern "C" double mo_geometric_logpdf(int64_t k, double p) {
    // geometric.rs:16-19, literally: ((1-p)^k * p).ln()  (underflows to -inf for very large k, like the reference)
    return std::log(std::pow(1. - p, (double)k) * p);
}

extern "C" double mo_poisson_logpdf(int64_t k, double rate) {
    // poisson.rs:16-18: k ln(rate) - rate - sum_{v=1..k} ln v   (the sum in the iterator's order)
    double s = 0.;
    for (int64_t v = 1; v <= k; ++v) s += std::log((double)v);
    return (double)k * std::log(rate) - rate - s;
}

// beta.rs / gamma.rs call `compute::functions::gamma` (crate compute 0.2.3, not vendored in the reference tree): the
// Gamma function.  std::tgamma stands in for it; parity is anchored on the reference's known answers
// (tests/dists.rs:200-212, epsilon = f32::EPSILON), which both satisfy.
extern "C" double mo_beta_logpdf(double x, double a, double b) {
    // beta.rs:17-21, literally: ln( Gamma(a+b)/(Gamma(a)Gamma(b)) * x^(a-1) * (1-x)^(b-1) )
    const double beta_f = std::tgamma(a + b) / (std::tgamma(a) * std::tgamma(b));
    return std::log(beta_f * std::pow(x, a - 1.) * std::pow(1. - x, b - 1.));
}

extern "C" double mo_gamma_logpdf(double x, double a, double b) {
    // gamma.rs:17-20 (shape a, scale b): (a-1) ln x - x/b - ln Gamma(a) - a ln b
    return (a - 1.) * std::log(x) - x / b - std::log(std::tgamma(a)) - a * std::log(b);
}

extern "C" double mo_uniform2d_logpdf(double x, double y, const double bd[4]) {
    // tests/pointed_model/types_2d.rs:15-21
    if (bd[0] <= x && x <= bd[1] && bd[2] <= y && y <= bd[3])
        return -std::log((bd[1] - bd[0]) * (bd[3] - bd[2]));
    return kNegInf;
}

namespace {
// nalgebra 0.32.2 (pinned by modppl/Cargo.toml:17, not vendored): Matrix::determinant() and try_inverse()
// special-case dims 1..3 with closed forms (cofactor expansion); larger dims go through LU.  All reference call
// sites use k = 2 (one test uses k = 3), pinned by tests/dists.rs:164-183.  k = 4 here uses plain Gauss-Jordan.
double det_small(const double* m, int k) {
    if (k == 1) return m[0];
    if (k == 2) return m[0] * m[3] - m[2] * m[1];  // m11*m22 - m21*m12
    if (k == 3) {
        double m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
        double minor_m12_m23 = m22 * m33 - m32 * m23;
        double minor_m11_m23 = m21 * m33 - m31 * m23;
        double minor_m11_m22 = m21 * m32 - m31 * m22;
        return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    }
    // generic LU with partial pivoting
    std::vector<double> a(m, m + k * k);
    double det = 1.;
    for (int c = 0; c < k; ++c) {
        int p = c;
        for (int r = c + 1; r < k; ++r) if (std::fabs(a[r * k + c]) > std::fabs(a[p * k + c])) p = r;
        if (a[p * k + c] == 0.) return 0.;
        if (p != c) { for (int j = 0; j < k; ++j) std::swap(a[p * k + j], a[c * k + j]); det = -det; }
        det *= a[c * k + c];
        for (int r = c + 1; r < k; ++r) {
            double f = a[r * k + c] / a[c * k + c];
            for (int j = c; j < k; ++j) a[r * k + j] -= f * a[c * k + j];
        }
    }
    return det;
}

bool inv_small(const double* m, int k, double* out) {
    if (k == 1) { if (m[0] == 0.) return false; out[0] = 1. / m[0]; return true; }
    if (k == 2) {
        double m11 = m[0], m12 = m[1], m21 = m[2], m22 = m[3];
        double det = m11 * m22 - m21 * m12;
        if (det == 0.) return false;
        out[0] = m22 / det; out[1] = -m12 / det; out[2] = -m21 / det; out[3] = m11 / det;
        return true;
    }
    if (k == 3) {
        double m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
        double minor_m12_m23 = m22 * m33 - m32 * m23;
        double minor_m11_m23 = m21 * m33 - m31 * m23;
        double minor_m11_m22 = m21 * m32 - m31 * m22;
        double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
        if (det == 0.) return false;
        out[0] = minor_m12_m23 / det;
        out[1] = (m13 * m32 - m33 * m12) / det;
        out[2] = (m12 * m23 - m22 * m13) / det;
        out[3] = -minor_m11_m23 / det;
        out[4] = (m11 * m33 - m31 * m13) / det;
        out[5] = (m13 * m21 - m23 * m11) / det;
        out[6] = minor_m11_m22 / det;
        out[7] = (m12 * m31 - m32 * m11) / det;
        out[8] = (m11 * m22 - m21 * m12) / det;
        return true;
    }
    std::vector<double> a(k * 2 * k, 0.);
    for (int r = 0; r < k; ++r) { for (int c = 0; c < k; ++c) a[r * 2 * k + c] = m[r * k + c]; a[r * 2 * k + k + r] = 1.; }
    for (int c = 0; c < k; ++c) {
        int p = c;
        for (int r = c + 1; r < k; ++r) if (std::fabs(a[r * 2 * k + c]) > std::fabs(a[p * 2 * k + c])) p = r;
        if (a[p * 2 * k + c] == 0.) return false;
        if (p != c) for (int j = 0; j < 2 * k; ++j) std::swap(a[p * 2 * k + j], a[c * 2 * k + j]);
        double d = a[c * 2 * k + c];
        for (int j = 0; j < 2 * k; ++j) a[c * 2 * k + j] /= d;
        for (int r = 0; r < k; ++r) if (r != c) {
            double f = a[r * 2 * k + c];
            for (int j = 0; j < 2 * k; ++j) a[r * 2 * k + j] -= f * a[c * 2 * k + j];
        }
    }
    for (int r = 0; r < k; ++r) for (int c = 0; c < k; ++c) out[r * k + c] = a[r * 2 * k + k + c];
    return true;
}
}  // namespace

extern "C" double mo_mvnormal_logpdf(const double* x, const double* mu, const double* cov, int k) {
    // mvnormal.rs:14-22:  det and inverse recomputed per call;  -(k ln 2pi + ln det + mahal^2)/2
    if (k < 1 || k > 8) return std::numeric_limits<double>::quiet_NaN();
    double inv[64], c[8], tmp[8];
    double cov_det = det_small(cov, k);
    if (!inv_small(cov, k, inv)) return std::numeric_limits<double>::quiet_NaN();
    for (int i = 0; i < k; ++i) c[i] = x[i] - mu[i];
    // (centered^T * cov_inv) * centered
    for (int j = 0; j < k; ++j) { double s = 0.; for (int i = 0; i < k; ++i) s += c[i] * inv[i * k + j]; tmp[j] = s; }
    double mahal = 0.;
    for (int j = 0; j < k; ++j) mahal += tmp[j] * c[j];
    return -((double)k * std::log(2. * kPi) + std::log(cov_det) + mahal) / 2.;
}

extern "C" int64_t mo_categorical_random(const double* probs, size_t n, double u) {
    // categorical.rs:22-32 with the u01 draw injected.  Literal: may return -1 (u == 0) and would index past the
    // end (panic) when u exceeds the running total; here that case returns n.
    double t = 0.;
    int64_t x = 0;
    while (t < u) {
        if ((size_t)x >= n) return (int64_t)n;
        t += probs[x];
        x += 1;
    }
    return x - 1;
}

extern "C" double mo_categorical_logpdf(int64_t x, const double* probs, size_t n) {
    // categorical.rs:13-20
    return (x < (int64_t)n) ? std::log(probs[x]) : kNegInf;
}

// ============================================================================================
// resampling: particle_filter.rs:37-41 -> categorical.rs:22-32
// ============================================================================================
extern "C" void mo_cumsum_sequential(const double* p, size_t n, double* out) {
    double t = 0.;
    for (size_t i = 0; i < n; ++i) { t += p[i]; out[i] = t; }
}

namespace {
inline int64_t clamp_idx(int64_t x, size_t n) { return x < 0 ? 0 : (x >= (int64_t)n ? (int64_t)n - 1 : x); }
inline int64_t search_cumsum(const double* S, size_t n, double u) {
    // min{k : S_k >= u}, n if none
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (S[mid] >= u) hi = mid; else lo = mid + 1; }
    return (int64_t)lo;
}
}  // namespace

extern "C" int mo_resample_indices_faithful(const double* probs, const double* uniforms, size_t n, size_t n_draws,
                                            int scheme, int64_t* parents) {
    if (n == 0) return -1;
    for (size_t i = 0; i < n_draws; ++i) {
        double u = (scheme == MO_SCHEME_MULTINOMIAL) ? uniforms[i] : (uniforms[0] + (double)i) / (double)n_draws;
        parents[i] = clamp_idx(mo_categorical_random(probs, n, u), n);
    }
    return 0;
}

extern "C" int mo_resample_indices(const double* probs, const double* uniforms, size_t n, size_t n_draws, int scheme,
                                   int64_t* parents) {
    if (n == 0) return -1;
    std::vector<double> S(n);
    mo_cumsum_sequential(probs, n, S.data());
    for (size_t i = 0; i < n_draws; ++i) {
        double u = (scheme == MO_SCHEME_MULTINOMIAL) ? uniforms[i] : (uniforms[0] + (double)i) / (double)n_draws;
        parents[i] = clamp_idx(search_cumsum(S.data(), n, u), n);
    }
    return 0;
}

// ============================================================================================
// Philox4x32-10 (Salmon et al., SC'11) -- replaces the unseedable ThreadRng (distribution.rs:5-7)
// ============================================================================================
extern "C" void mo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
extern "C" double mo_u01_f64(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; }
extern "C" float mo_u01_f32(uint32_t x) { return (float)(x >> 8) * 0x1.0p-24f; }

namespace {
enum Purpose : uint32_t { P_MODEL = 0, P_RESAMPLE_U = 1, P_RESAMPLE_OFFSET = 2, P_IS = 3, P_MH = 4, P_IS_RESAMPLE = 5, P_MH_INIT = 6, P_MODEL_GROUP = 7 };

struct Stream {  // ctr = {id_lo, id_hi, t, purpose<<24 | blk}, key = seed
    uint32_t key[2]; uint32_t id_lo, id_hi, t, purpose;
    Stream(uint64_t seed, uint64_t id, uint32_t t_, uint32_t purpose_)
        : id_lo((uint32_t)id), id_hi((uint32_t)(id >> 32)), t(t_), purpose(purpose_) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); }
    void block(uint32_t blk, uint32_t out[4]) const {
        uint32_t c[4] = {id_lo, id_hi, t, (purpose << 24) | blk};
        mo_philox4x32_10(c, key, out);
    }
};
inline double u01_oc64(uint32_t hi, uint32_t lo) { return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * 0x1.0p-53; }  // (0,1]
inline float u01_oc32(uint32_t x) { return (float)((x >> 8) + 1) * 0x1.0p-24f; }

template <typename Real> struct Draw;
template <> struct Draw<double> {
    // one Philox block -> one Box-Muller pair
    static void normals(const Stream& s, uint32_t first_blk, int count, double* z) {
        for (int i = 0; i < count; i += 2) {
            uint32_t x[4]; s.block(first_blk + i / 2, x);
            double u1 = u01_oc64(x[0], x[1]), u2 = mo_u01_f64(x[2], x[3]);
            double r = std::sqrt(-2. * std::log(u1));
            z[i] = r * std::cos(2. * kPi * u2);
            if (i + 1 < count) z[i + 1] = r * std::sin(2. * kPi * u2);
        }
    }
    static void uniforms(const Stream& s, uint32_t first_blk, int count, double* u) {
        for (int i = 0; i < count; i += 2) {
            uint32_t x[4]; s.block(first_blk + i / 2, x);
            u[i] = mo_u01_f64(x[0], x[1]);
            if (i + 1 < count) u[i + 1] = mo_u01_f64(x[2], x[3]);
        }
    }
};
template <> struct Draw<float> {
    // one Philox block -> two Box-Muller pairs
    static void normals(const Stream& s, uint32_t first_blk, int count, float* z) {
        for (int i = 0; i < count; i += 4) {
            uint32_t x[4]; s.block(first_blk + i / 4, x);
            for (int h = 0; h < 2; ++h) {
                if (i + 2 * h >= count) break;
                float u1 = u01_oc32(x[2 * h]), u2 = mo_u01_f32(x[2 * h + 1]);
                float r = std::sqrt(-2.f * std::log(u1));
                z[i + 2 * h] = r * (float)std::cos(2. * kPi * (double)u2);
                if (i + 2 * h + 1 < count) z[i + 2 * h + 1] = r * (float)std::sin(2. * kPi * (double)u2);
            }
        }
    }
    static void uniforms(const Stream& s, uint32_t first_blk, int count, float* u) {
        for (int i = 0; i < count; i += 4) {
            uint32_t x[4]; s.block(first_blk + i / 4, x);
            for (int h = 0; h < 4 && i + h < count; ++h) u[i + h] = mo_u01_f32(x[h]);
        }
    }
};
}  // namespace

// ============================================================================================
// Fixed-point weights (engine-defined; integer arithmetic => order- and shard-invariant)
// ============================================================================================
extern "C" float mo_exp2_poly(float f) {
    // 2^f on [-0.5, 0.5], degree-7 Taylor in f*ln2, Horner with correctly-rounded fmaf only.
    float p = 1.5252733804059841e-05f;
    p = std::fmaf(p, f, 0.00015403530393381608f);
    p = std::fmaf(p, f, 0.0013333558146428443f);
    p = std::fmaf(p, f, 0.009618129107628477f);
    p = std::fmaf(p, f, 0.05550410866482158f);
    p = std::fmaf(p, f, 0.2402265069591007f);
    p = std::fmaf(p, f, 0.6931471805599453f);
    p = std::fmaf(p, f, 1.0f);
    return p;
}

extern "C" uint64_t mo_fixed_weight(float d, int kbits) {
    if (!(d > -100.0f)) d = -100.0f;       // also NaN, -inf (they come out as 0 below)
    if (d > 0.f) d = 0.f;
    float y = d * 1.44269504088896341f;    // single rounding
    float n = std::rint(y);                // ties-to-even
    float f = y - n;                       // exact
    float p = mo_exp2_poly(f);
    float scale = std::ldexp(1.0f, kbits + (int)n);      // 2^(kbits + n), a normal float (n >= -145)
    float v = p * scale;                                  // exact
    return (uint64_t)std::llrint(v);                      // ties-to-even
}

extern "C" int mo_fixed_kbits(uint64_t n_total) {
    int lg = 0;
    while (((uint64_t)1 << lg) < n_total) ++lg;
    int k = 62 - lg;
    return k > 40 ? 40 : k;
}

namespace {
inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) >> 64); }

struct FixedStats { float mx; uint64_t W; int kbits; double lse; };

FixedStats fixed_quantize(const float* lw, size_t n, size_t n_total, std::vector<uint64_t>& q) {
    FixedStats st;
    float mx = -std::numeric_limits<float>::infinity();
    for (size_t i = 0; i < n; ++i) if (lw[i] > mx) mx = lw[i];   // NaN never wins
    st.mx = mx; st.kbits = mo_fixed_kbits(n_total);
    q.resize(n);
    uint64_t W = 0;
    for (size_t i = 0; i < n; ++i) { q[i] = std::isfinite(mx) ? mo_fixed_weight(lw[i] - mx, st.kbits) : 0; W += q[i]; }
    st.W = W;
    st.lse = (W > 0) ? (double)mx + std::log((double)W) - (double)st.kbits * std::log(2.) : kNegInf;
    return st;
}
}  // namespace

extern "C" uint64_t mo_fixed_systematic(const float* lw, size_t n, uint64_t u64rand, int32_t* anc, double* lse_out) {
    std::vector<uint64_t> q;
    FixedStats st = fixed_quantize(lw, n, n, q);
    if (lse_out) *lse_out = st.lse;
    if (st.W == 0) return 0;
    uint64_t U = mulhi64(u64rand, st.W);
    // ancestor(j) = min{k : C_k * n > j*W + U}
    uint64_t C = 0; size_t j = 0;
    for (size_t k = 0; k < n; ++k) {
        C += q[k];
        u128 lhs = (u128)C * n;
        while (j < n && (u128)j * st.W + U < lhs) anc[j++] = (int32_t)k;
    }
    return st.W;
}

namespace {
inline uint64_t splitmix64_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// #{ j in [0, n_out) : j*W + U < C*n_out }
inline uint64_t slots_below(uint64_t C, uint64_t W, uint64_t U, uint64_t n_out) {
    u128 lhs = (u128)C * n_out;
    if (lhs <= (u128)U) return 0;
    return (uint64_t)((lhs - U - 1) / W) + 1;
}
}  // namespace

// Nested systematic resampling on integer weights, three levels (section of 2^17 particles > chunk of 128 > particle);
// restates modppl_b200/csrc/nested.cuh.  Every level is an exact systematic scheme on integers:
//   chunk c:   e_c = ceil(max y), q_i = rint(2^(y_i - e_c + k)), S_c = sum q_i
//   section s: E_s = max e_c, G_c = S_c >> (E_s - e_c), T_s = sum G_c
//   top:       E = max E_s,  M_s = T_s >> (E - E_s),  W = sum M_s
extern "C" uint64_t mo_nested_systematic(const float* lw, size_t n, uint64_t u64rand, int32_t* anc, double* lse_out) {
    const size_t CH = 128, SEC = (size_t)1 << 17, CPS = SEC / CH;
    const int kbits = 22;   // kNestedBits: a chunk's integer weights are at most 2^22, its sum fits 32 bits
    const size_t nch = (n + CH - 1) / CH, nsec = (n + SEC - 1) / SEC;
    std::vector<uint64_t> q(n, 0), S(nch, 0);
    std::vector<int> e(nch, 0);
    std::vector<char> empty(nch, 1);
    parallel_for(nch, [&](size_t c_lo, size_t c_hi) {
    for (size_t c = c_lo; c < c_hi; ++c) {
        float Y = -std::numeric_limits<float>::infinity();
        for (size_t i = c * CH; i < std::min(n, (c + 1) * CH); ++i) {
            if (lw[i] == lw[i]) { float y = lw[i] * 1.44269504088896341f; if (y > Y) Y = y; }
        }
        if (!(Y > -std::numeric_limits<float>::infinity())) continue;
        empty[c] = 0;
        e[c] = (int)std::ceil(Y);
        for (size_t i = c * CH; i < std::min(n, (c + 1) * CH); ++i) {
            float y = lw[i] * 1.44269504088896341f;
            float z = y - (float)e[c];
            if (!(z > -126.0f)) z = -126.0f;              // also NaN, -inf
            float nn = std::rint(z);
            float f = z - nn;
            float p = mo_exp2_poly(f);
            float v = p * std::ldexp(1.0f, kbits + (int)nn);
            q[i] = (uint64_t)std::llrint(v);
            S[c] += q[i];
        }
    }
    });
    std::vector<uint64_t> G(nch, 0), T(nsec, 0), M(nsec, 0);
    std::vector<int> Es(nsec, 0);
    std::vector<char> sec_empty(nsec, 1);
    bool any = false;
    int E = 0;
    for (size_t s = 0; s < nsec; ++s) {
        const size_t c0 = s * CPS, c1 = std::min(nch, (s + 1) * CPS);
        for (size_t c = c0; c < c1; ++c) if (!empty[c] && (sec_empty[s] || e[c] > Es[s])) { Es[s] = e[c]; sec_empty[s] = 0; }
        if (sec_empty[s]) continue;
        for (size_t c = c0; c < c1; ++c) { if (!empty[c] && Es[s] - e[c] < 64) G[c] = S[c] >> (Es[s] - e[c]); T[s] += G[c]; }
        if (!any || Es[s] > E) { E = Es[s]; any = true; }
    }
    if (!any) { if (lse_out) *lse_out = kNegInf; return 0; }
    uint64_t W = 0;
    for (size_t s = 0; s < nsec; ++s) { if (!sec_empty[s] && E - Es[s] < 64) M[s] = T[s] >> (E - Es[s]); W += M[s]; }
    if (lse_out) *lse_out = W ? (double)E * 0.6931471805599453 + std::log((double)W) - (double)kbits * 0.6931471805599453 : kNegInf;
    if (W == 0) return 0;
    const uint64_t U = mulhi64(u64rand, W);
    uint64_t Mu = 0;
    for (size_t s = 0; s < nsec; ++s) {
        const uint64_t a_s = slots_below(Mu, W, U, n);
        Mu += M[s];
        const uint64_t n_s = slots_below(Mu, W, U, n) - a_s;
        if (n_s == 0) continue;
        const uint64_t Us = mulhi64(splitmix64_mix((u64rand ^ 0x5851F42D4C957F2Dull) + (uint64_t)(s + 1) * 0xD1B54A32D192ED03ull), T[s]);
        uint64_t Gam = 0;
        for (size_t c = s * CPS; c < std::min(nch, (s + 1) * CPS); ++c) {
            const uint64_t a = slots_below(Gam, T[s], Us, n_s);
            Gam += G[c];
            const uint64_t nc = slots_below(Gam, T[s], Us, n_s) - a;
            if (nc == 0) continue;
            const uint64_t Uc = mulhi64(splitmix64_mix(u64rand + (uint64_t)(c + 1) * 0x9E3779B97F4A7C15ull), S[c]);
            uint64_t C = 0, prev = 0;
            for (size_t i = c * CH; i < std::min(n, (c + 1) * CH); ++i) {
                C += q[i];
                const uint64_t cnt = slots_below(C, S[c], Uc, nc);
                for (uint64_t l = prev; l < cnt; ++l) anc[a_s + a + l] = (int32_t)i;
                prev = cnt;
            }
        }
    }
    return W;
}

extern "C" uint64_t mo_fixed_multinomial(const float* lw, size_t n, uint64_t seed, uint32_t t, int32_t* anc, double* lse_out) {
    std::vector<uint64_t> q;
    FixedStats st = fixed_quantize(lw, n, n, q);
    if (lse_out) *lse_out = st.lse;
    if (st.W == 0) return 0;
    std::vector<uint64_t> C(n);
    uint64_t c = 0;
    for (size_t k = 0; k < n; ++k) { c += q[k]; C[k] = c; }
    for (size_t j = 0; j < n; ++j) {
        Stream s(seed, j, t, P_RESAMPLE_U);
        uint32_t x[4]; s.block(0, x);
        uint64_t r = ((uint64_t)x[0] << 32) | x[1];
        uint64_t T = mulhi64(r, st.W);
        size_t lo = 0, hi = n;                       // min{k : C_k > T}
        while (lo < hi) { size_t mid = (lo + hi) >> 1; if (C[mid] > T) hi = mid; else lo = mid + 1; }
        anc[j] = (int32_t)lo;
    }
    return st.W;
}

// ============================================================================================
// Unfold-style models (restricted vectorisable form of modeling/dynunfold.rs:41-100:
//  kernel sees t = 0 at init_step, t = 1 at the first step (quirk Q5); state = last retv;
//  weight = sum of constrained logpdfs -- dyngenfn.rs:121-131)
// ============================================================================================
namespace {
template <typename Real> struct ModelBase {
    virtual ~ModelBase() {}
    virtual int dim() const = 0;
    virtual int n_obs() const = 0;
    // t == 0: sample from the prior; t > 0: transition from x (in/out).  Returns the observation log-likelihood.
    virtual Real kernel(int64_t t, const Stream& s, Real* x, const double* obs) const = 0;
    // models that need ONE normal deviate per particle and step share a Philox block between 4 (fp32) / 2 (fp64) consecutive
    // global ids (stream id = gid / group, purpose P_MODEL_GROUP, deviate number gid % group): kernel_z gets the deviate
    virtual bool group_draws() const { return false; }
    virtual Real kernel_z(int64_t, Real, Real*, const double*) const { return 0; }
};

// -- 4-D constant-velocity linear-Gaussian tracker (config 4; not in the reference) ------------
template <typename Real> struct Lgssm4 : ModelBase<Real> {
    Real q, r, x0;
    Lgssm4(const double* p, size_t n) { q = n > 0 ? p[0] : 0.1; r = n > 1 ? p[1] : 0.5; x0 = n > 2 ? p[2] : 1.0; }
    int dim() const override { return 4; }
    int n_obs() const override { return 2; }
    Real kernel(int64_t t, const Stream& s, Real* x, const double* obs) const override {
        Real z[4]; Draw<Real>::normals(s, 0, 4, z);
        if (t == 0) { for (int d = 0; d < 4; ++d) x[d] = z[d] * x0; }
        else { x[0] = x[0] + x[2] + z[0] * q; x[1] = x[1] + x[3] + z[1] * q; x[2] = x[2] + z[2] * q; x[3] = x[3] + z[3] * q; }
        Real lw = 0;
        for (int d = 0; d < 2; ++d) {   // two independent `normal` observations (normal.rs:13-17)
            Real zz = ((Real)obs[d] - x[d]) / r;
            lw += -(zz * zz + (Real)std::log(2. * kPi)) / 2 - (Real)std::log((double)r);
        }
        return lw;
    }
};

// -- spiral model, tests/dyngenfns/unfold.rs:14-33 (config 1) ---------------------------------
template <typename Real> struct Spiral : ModelBase<Real> {
    double dr_std, dth_mean, dth_std, obs_var;
    Spiral(const double* p, size_t n) { dr_std = n > 0 ? p[0] : 0.1; dth_mean = n > 1 ? p[1] : 0.4; dth_std = n > 2 ? p[2] : 0.2; obs_var = n > 3 ? p[3] : 0.001; }
    int dim() const override { return 2; }
    int n_obs() const override { return 2; }
    Real kernel(int64_t t, const Stream& s, Real* x, const double* obs) const override {
        if (t == 0) {
            Real u[2]; Draw<Real>::uniforms(s, 0, 2, u);
            x[0] = u[0] * (Real)(1. - 0.) + (Real)0.;            // uniform.rs:28-32  u*(b-a)+a
            x[1] = u[1] * (Real)(2. * kPi - 0.) + (Real)0.;
        } else {
            Real z[2]; Draw<Real>::normals(s, 0, 2, z);
            x[0] = x[0] + (z[0] * (Real)dr_std + (Real)0.);     // normal.rs:26  u*c*std + mu
            x[1] = x[1] + (z[1] * (Real)dth_std + (Real)dth_mean);
        }
        if (sizeof(Real) == 8) {
            double pos[2] = {(double)x[0] * std::cos((double)x[1]), (double)x[0] * std::sin((double)x[1])};
            double cov[4] = {obs_var, 0., 0., obs_var};
            return (Real)mo_mvnormal_logpdf(obs, pos, cov, 2);   // mvnormal.rs:14-22
        }
        Real px = x[0] * std::cos(x[1]), py = x[0] * std::sin(x[1]);
        Real inv = (Real)(1. / obs_var);
        Real c0 = (Real)obs[0] - px, c1 = (Real)obs[1] - py;
        Real mahal = c0 * c0 * inv + c1 * c1 * inv;
        return -((Real)(2. * std::log(2. * kPi) + std::log(obs_var * obs_var)) + mahal) / 2;
    }
};

// -- stochastic volatility (config 5; not in the reference) -----------------------------------
template <typename Real> struct StochVol : ModelBase<Real> {
    Real mu, phi, sig;
    StochVol(const double* p, size_t n) { mu = n > 0 ? p[0] : -1.024; phi = n > 1 ? p[1] : 0.9702; sig = n > 2 ? p[2] : 0.178; }
    int dim() const override { return 1; }
    int n_obs() const override { return 1; }
    bool group_draws() const override { return true; }
    Real kernel(int64_t t, const Stream& s, Real* x, const double* obs) const override {
        Real z[1]; Draw<Real>::normals(s, 0, 1, z);
        return kernel_z(t, z[0], x, obs);
    }
    Real kernel_z(int64_t t, Real z0, Real* x, const double* obs) const override {
        if (t == 0) x[0] = mu + (sig / std::sqrt(1 - phi * phi)) * z0;
        else x[0] = mu + phi * (x[0] - mu) + sig * z0;
        Real sd = std::exp(x[0] / 2);
        Real zz = (Real)obs[0] / sd;
        return -(zz * zz + (Real)std::log(2. * kPi)) / 2 - x[0] / 2;   // normal.rs:13-17 with ln(std) = x/2
    }
};

// -- K-state HMM, tests/hmm/model.rs:24-81 (the reference's only end-to-end particle-filter check) ----
template <typename Real> struct Hmm : ModelBase<Real> {
    int K, M; std::vector<double> prior, emis, trans;
    Hmm(const double* p, size_t n) {
        K = (int)p[0]; M = (int)p[1];
        prior.assign(p + 2, p + 2 + K);
        emis.assign(p + 2 + K, p + 2 + K + M * K);
        trans.assign(p + 2 + K + M * K, p + 2 + K + M * K + K * K);
        (void)n;
    }
    int dim() const override { return 1; }
    int n_obs() const override { return 1; }
    Real kernel(int64_t t, const Stream& s, Real* x, const double* obs) const override {
        double u[1]; Draw<double>::uniforms(s, 0, 1, u);
        std::vector<double> probs(K);
        if (t == 0) probs = prior;                                        // model.rs:63
        else { int prev = (int)x[0]; for (int k = 0; k < K; ++k) probs[k] = trans[k * K + prev]; }   // model.rs:75-79
        int64_t st = clamp_idx(mo_categorical_random(probs.data(), K, u[0]), K);   // model.rs:41
        x[0] = (Real)st;
        int o = (int)obs[0];
        return (Real)std::log(emis[o * K + st]);                           // model.rs:42-44 categorical.logpdf
    }
};
}  // namespace

// ============================================================================================
// inference/particle_filter.rs
// ============================================================================================
struct mo_ps {
    virtual ~mo_ps() {}
    virtual int init_step(const double* obs, size_t n) = 0;
    virtual int step(const double* obs, size_t n) = 0;
    virtual double ess(int stale) = 0;
    virtual double resample(int scheme) = 0;
    virtual double resample_faithful_cost() = 0;
    virtual double lml() = 0;
    virtual int dim() = 0;
    virtual void read_state(double*) = 0;
    virtual void read_lw(double*) = 0;
    virtual void read_parents(int64_t*) = 0;
    virtual void write_state(const double*) = 0;
    virtual void write_lw(const double*) = 0;
};

namespace {
template <typename Real> struct PS : mo_ps {
    // particle_filter.rs:8-24
    size_t num_particles; uint64_t n_global, gid_offset, seed;
    ModelBase<Real>* model; int D;
    std::vector<Real> state;          // SoA: D x N   ("traces": only the live state, quirk Q11)
    std::vector<Real> log_weights;
    std::vector<double> log_normalized_weights, two_times_log_normalized_weights, normalized_weights;
    std::vector<int64_t> parents;
    double log_ml_estimate; int64_t t; uint32_t resample_count;

    PS(ModelBase<Real>* m, size_t n, uint64_t seed_, uint64_t off, uint64_t ng)
        : num_particles(n), n_global(ng), gid_offset(off), seed(seed_), model(m), D(m->dim()), state((size_t)m->dim() * n, 0),
          log_weights(n, 0), log_normalized_weights(n, 0.), two_times_log_normalized_weights(n, 0.), normalized_weights(n, 0.),
          parents(n, 0), log_ml_estimate(0.), t(0), resample_count(0) {}           // :44-57
    ~PS() override { delete model; }

    std::vector<double> lw_f64() const { return std::vector<double>(log_weights.begin(), log_weights.end()); }

    double normalize_weights() {                                                      // :27-35
        std::vector<double> lw = lw_f64();
        double log_total_weight = mo_logsumexp(lw.data(), num_particles);
        for (size_t i = 0; i < num_particles; ++i) {
            log_normalized_weights[i] = lw[i] - log_total_weight;
            two_times_log_normalized_weights[i] = 2.0 * log_normalized_weights[i];
            normalized_weights[i] = std::exp(log_normalized_weights[i]);
        }
        return log_total_weight;
    }

    // one particle through the model kernel with the engine's stream convention
    Real propagate(uint64_t gid, Real* x, const double* obs) {
        if (model->group_draws()) {
            const uint64_t G = sizeof(Real) == 4 ? 4 : 2;
            Stream sg(seed, gid / G, (uint32_t)t, P_MODEL_GROUP);
            Real z[4]; Draw<Real>::normals(sg, 0, (int)G, z);
            return model->kernel_z(t, z[gid % G], x, obs);
        }
        Stream s(seed, gid, (uint32_t)t, P_MODEL);
        return model->kernel(t, s, x, obs);
    }

    int init_step(const double* obs, size_t n) override {                             // :60-70 (reset instead of push: Q4)
        if ((int)n < model->n_obs()) return -1;
        t = 0;
        parallel_for(num_particles, [&](size_t lo, size_t hi) {
            Real x[8];
            for (size_t i = lo; i < hi; ++i) {
                Real w = propagate(gid_offset + i, x, obs);
                for (int d = 0; d < D; ++d) state[(size_t)d * num_particles + i] = x[d];
                log_weights[i] = w;
            }
        });
        t = 1;
        return 0;
    }

    int step(const double* obs, size_t n) override {                                  // :73-95
        if ((int)n < model->n_obs()) return -1;
        parallel_for(num_particles, [&](size_t lo, size_t hi) {
            Real x[8];
            for (size_t i = lo; i < hi; ++i) {
                for (int d = 0; d < D; ++d) x[d] = state[(size_t)d * num_particles + i];
                Real w = propagate(gid_offset + i, x, obs);
                for (int d = 0; d < D; ++d) state[(size_t)d * num_particles + i] = x[d];
                log_weights[i] = log_weights[i] + w;                                     // :81
            }
        });
        t += 1;
        return 0;
    }

    double ess(int stale) override {                                                  // :98-100 (stale: quirk Q1)
        if (stale) return std::exp(-mo_logsumexp(two_times_log_normalized_weights.data(), num_particles));
        std::vector<double> lw = lw_f64();
        double lse = mo_logsumexp(lw.data(), num_particles);
        for (auto& v : lw) v = 2.0 * (v - lse);
        return std::exp(-mo_logsumexp(lw.data(), num_particles));
    }

    void gather() {                                                                   // :109-113
        std::vector<Real> tmp(state.size());
        parallel_for(num_particles, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
                for (int d = 0; d < D; ++d) tmp[(size_t)d * num_particles + i] = state[(size_t)d * num_particles + parents[i]];
        });
        state.swap(tmp);
        std::fill(log_weights.begin(), log_weights.end(), (Real)0);                   // :114
    }

    double resample(int scheme) override {                                            // :103-116
        uint32_t rt = (uint32_t)(t - 1);    // RNG tag: the step whose weights are being resampled
        double log_total_weight;
        if (scheme == MO_RESAMPLE_MULTINOMIAL || scheme == MO_RESAMPLE_SYSTEMATIC) {
            log_total_weight = normalize_weights();
            log_ml_estimate += log_total_weight - std::log((double)num_particles);    // :105
            std::vector<double> u(scheme == MO_RESAMPLE_MULTINOMIAL ? num_particles : 1);
            if (scheme == MO_RESAMPLE_MULTINOMIAL) {
                for (size_t i = 0; i < num_particles; ++i) {
                    Stream s(seed, gid_offset + i, rt, P_RESAMPLE_U);
                    uint32_t x[4]; s.block(0, x); u[i] = mo_u01_f64(x[0], x[1]);
                }
            } else {
                Stream s(seed, 0, rt, P_RESAMPLE_OFFSET);
                uint32_t x[4]; s.block(0, x); u[0] = mo_u01_f64(x[0], x[1]);
            }
            mo_resample_indices(normalized_weights.data(), u.data(), num_particles, num_particles,
                                scheme == MO_RESAMPLE_MULTINOMIAL ? MO_SCHEME_MULTINOMIAL : MO_SCHEME_SYSTEMATIC, parents.data());
        } else {
            std::vector<float> lwf(log_weights.begin(), log_weights.end());
            std::vector<int32_t> anc(num_particles);
            double lse; uint64_t W;
            if (scheme == MO_RESAMPLE_SYSTEMATIC_FIXED || scheme == MO_RESAMPLE_SYSTEMATIC_NESTED) {
                Stream s(seed, 0, rt, P_RESAMPLE_OFFSET);
                uint32_t x[4]; s.block(0, x);
                const uint64_t word = ((uint64_t)x[0] << 32) | x[1];
                W = scheme == MO_RESAMPLE_SYSTEMATIC_FIXED ? mo_fixed_systematic(lwf.data(), num_particles, word, anc.data(), &lse)
                                                           : mo_nested_systematic(lwf.data(), num_particles, word, anc.data(), &lse);
            } else {
                W = mo_fixed_multinomial(lwf.data(), num_particles, seed, rt, anc.data(), &lse);
            }
            if (W == 0) return kNegInf;
            log_total_weight = lse;
            log_ml_estimate += log_total_weight - std::log((double)num_particles);
            for (size_t i = 0; i < num_particles; ++i) parents[i] = anc[i];
        }
        gather();
        resample_count++;
        return log_total_weight;
    }

    double resample_faithful_cost() override {
        // the reference's cost model: every draw clones the probability vector and re-sums it for the assert
        // (categorical.rs:23), then scans linearly (:25-30).
        double log_total_weight = normalize_weights();
        log_ml_estimate += log_total_weight - std::log((double)num_particles);
        uint32_t rt = (uint32_t)(t - 1);
        volatile double sink = 0.;
        for (size_t i = 0; i < num_particles; ++i) {
            std::vector<double> probs(normalized_weights);                           // .clone()
            double sum = 0.; for (double p : probs) sum += p;                         // assert_abs_diff_eq!(sum, 1.0)
            sink = sink + sum;
            Stream s(seed, gid_offset + i, rt, P_RESAMPLE_U);
            uint32_t x[4]; s.block(0, x);
            parents[i] = clamp_idx(mo_categorical_random(probs.data(), num_particles, mo_u01_f64(x[0], x[1])), num_particles);
        }
        gather();
        return log_total_weight;
    }

    double lml() override {                                                           // :119-121
        std::vector<double> lw = lw_f64();
        return log_ml_estimate + mo_logsumexp(lw.data(), num_particles) - std::log((double)num_particles);
    }
    int dim() override { return D; }
    void read_state(double* o) override { for (size_t i = 0; i < state.size(); ++i) o[i] = (double)state[i]; }
    void read_lw(double* o) override { for (size_t i = 0; i < num_particles; ++i) o[i] = (double)log_weights[i]; }
    void read_parents(int64_t* o) override { std::copy(parents.begin(), parents.end(), o); }
    void write_state(const double* in) override { for (size_t i = 0; i < state.size(); ++i) state[i] = (Real)in[i]; }
    void write_lw(const double* in) override { for (size_t i = 0; i < num_particles; ++i) log_weights[i] = (Real)in[i]; }
};

template <typename Real> ModelBase<Real>* make_model(const std::string& name, const double* p, size_t n) {
    if (name == "lgssm4") return new Lgssm4<Real>(p, n);
    if (name == "spiral") return new Spiral<Real>(p, n);
    if (name == "sv") return new StochVol<Real>(p, n);
    if (name == "hmm") return new Hmm<Real>(p, n);
    return nullptr;
}
}  // namespace

extern "C" mo_ps* mo_ps_new(const char* model, const double* params, size_t n_params, uint64_t num_particles, int dtype,
                            uint64_t seed, uint64_t gid_offset, uint64_t n_global) {
    if (n_global == 0) n_global = num_particles;
    if (dtype == MO_F64) { auto* m = make_model<double>(model, params, n_params); return m ? new PS<double>(m, num_particles, seed, gid_offset, n_global) : nullptr; }
    auto* m = make_model<float>(model, params, n_params);
    return m ? new PS<float>(m, num_particles, seed, gid_offset, n_global) : nullptr;
}
extern "C" void mo_ps_free(mo_ps* p) { delete p; }
extern "C" int mo_ps_init_step(mo_ps* p, const double* obs, size_t n) { return p->init_step(obs, n); }
extern "C" int mo_ps_step(mo_ps* p, const double* obs, size_t n) { return p->step(obs, n); }
extern "C" double mo_ps_effective_sample_size(mo_ps* p, int stale) { return p->ess(stale); }
extern "C" double mo_ps_resample(mo_ps* p, int scheme) { return p->resample(scheme); }
extern "C" double mo_ps_resample_faithful_cost(mo_ps* p) { return p->resample_faithful_cost(); }
extern "C" double mo_ps_log_marginal_likelihood_estimate(mo_ps* p) { return p->lml(); }
extern "C" int mo_ps_state_dim(mo_ps* p) { return p->dim(); }
extern "C" void mo_ps_read_state(mo_ps* p, double* o) { p->read_state(o); }
extern "C" void mo_ps_read_log_weights(mo_ps* p, double* o) { p->read_lw(o); }
extern "C" void mo_ps_read_parents(mo_ps* p, int64_t* o) { p->read_parents(o); }
extern "C" void mo_ps_write_state(mo_ps* p, const double* in) { p->write_state(in); }
extern "C" void mo_ps_write_log_weights(mo_ps* p, const double* in) { p->write_lw(in); }

// ============================================================================================
// Static models for importance sampling / MH (fp64, as the reference)
//   RNG convention for these paths: every draw consumes one whole Philox block:
//     uniform():  u = [0,1) from x0,x1          (second uniform from x2,x3 for 2-D draws)
//     normal():   Box-Muller, cosine branch only  (normal.rs:19-27 also discards half of its pair)
// ============================================================================================
namespace {
struct Rng64 {
    Stream s; uint32_t blk;
    Rng64(uint64_t seed, uint64_t id, uint32_t t, uint32_t purpose) : s(seed, id, t, purpose), blk(0) {}
    double uniform() { uint32_t x[4]; s.block(blk++, x); return mo_u01_f64(x[0], x[1]); }
    void uniform2(double& a, double& b) { uint32_t x[4]; s.block(blk++, x); a = mo_u01_f64(x[0], x[1]); b = mo_u01_f64(x[2], x[3]); }
    void normal2(double& a, double& b) {
        uint32_t x[4]; s.block(blk++, x);
        double u1 = u01_oc64(x[0], x[1]), u2 = mo_u01_f64(x[2], x[3]);
        double r = std::sqrt(-2. * std::log(u1));
        a = r * std::cos(2. * kPi * u2); b = r * std::sin(2. * kPi * u2);
    }
    double normal() { double a, b; normal2(a, b); return a; }
    void skip(uint32_t n) { blk += n; }
};

const double kHierNoise = 0.1;

double hier_loglik(const double* xs, const double* ys, size_t n, int L, double a, double b, double c) {
    double w = 0.;
    for (size_t i = 0; i < n; ++i) {
        double mean = L ? a + b * xs[i] : a + b * xs[i] + c * xs[i] * xs[i];   // hierarchical.rs:36-44
        w += mo_normal_logpdf(ys[i], mean, kHierNoise);
    }
    return w;
}
}  // namespace

extern "C" double mo_hier_logjp(const double* xs, const double* ys, size_t n, const double st[4]) {
    int L = st[0] != 0.;
    double lp = mo_bernoulli_logpdf(L, 0.7) + mo_normal_logpdf(st[1], 0., 1.) + mo_normal_logpdf(st[2], 0., 1.);
    if (!L) lp += mo_normal_logpdf(st[3], 0., 1.);
    return lp + hier_loglik(xs, ys, n, L, st[1], st[2], st[3]);
}

extern "C" int mo_is_num_latents(const char* model) {
    std::string m(model);
    if (m == "line") return 2;
    if (m == "hierarchical") return 4;
    if (m == "pointed") return 2;
    return -1;
}

extern "C" int mo_importance_sampling(const char* model, const double* args, size_t n_args, const double* obs, size_t n_obs,
                                      uint32_t num_samples, uint64_t seed, uint64_t batch, double* latents,
                                      double* log_norm_weights, double* lml) {
    // importance.rs:12-28
    std::string m(model);
    size_t n = num_samples;
    std::vector<double> w(n);
    for (size_t i = 0; i < n; ++i) {
        Rng64 g(seed, i, (uint32_t)batch, P_IS);
        if (m == "line") {                      // tests/dyngenfns/simple.rs:10-23 ; args = xs, obs = ys
            if (n_args != n_obs) return -1;
            double z0, z1; g.normal2(z0, z1);
            double slope = z0 * 1. + 0., intercept = z1 * 2. + 0.;
            double ww = 0.;
            for (size_t j = 0; j < n_obs; ++j) ww += mo_normal_logpdf(obs[j], slope * args[j] + intercept, 0.1);
            latents[0 * n + i] = slope; latents[1 * n + i] = intercept; w[i] = ww;
        } else if (m == "hierarchical") {       // hierarchical.rs:32-46
            if (n_args != n_obs) return -1;
            int L = 0.7 > g.uniform();          // bernoulli.rs:16-18
            double a, b; g.normal2(a, b);
            double c = g.normal();
            if (L) c = 0.;
            latents[0 * n + i] = L; latents[1 * n + i] = a; latents[2 * n + i] = b; latents[3 * n + i] = c;
            w[i] = hier_loglik(args, obs, n_obs, L, a, b, c);
        } else if (m == "pointed") {            // tests/pointed_model/model.rs:25-66 ; args = bounds[4], cov[4] ; obs[2]
            if (n_args != 8 || n_obs != 2) return -1;
            double u0, u1; g.uniform2(u0, u1);
            double lat[2] = {u0 * (args[1] - args[0]) + args[0], u1 * (args[3] - args[2]) + args[2]};   // types_2d.rs:23-30
            latents[0 * n + i] = lat[0]; latents[1 * n + i] = lat[1];
            w[i] = mo_mvnormal_logpdf(obs, lat, args + 4, 2);
        } else return -2;
    }
    double log_total_weight = mo_logsumexp(w.data(), n);                  // :21
    *lml = log_total_weight - std::log((double)num_samples);             // :22
    for (size_t i = 0; i < n; ++i) log_norm_weights[i] = w[i] - log_total_weight;   // :23-25
    return 0;
}

extern "C" int mo_importance_resampling_indices(const double* lnw, uint32_t n, uint32_t n_ret, uint64_t seed, uint64_t batch, int64_t* idx) {
    // importance.rs:44-50
    std::vector<double> probs(n), u(n_ret);
    for (uint32_t i = 0; i < n; ++i) probs[i] = std::exp(lnw[i]);
    for (uint32_t i = 0; i < n_ret; ++i) { Rng64 g(seed, i, (uint32_t)batch, P_IS_RESAMPLE); u[i] = g.uniform(); }
    return mo_resample_indices(probs.data(), u.data(), n, n_ret, MO_SCHEME_MULTINOMIAL, idx);
}

// ============================================================================================
// inference/mh.rs, flattened per SURVEY.md section 3.4 (weight table from modeling/dyngenfn.rs:115-273,454-486)
// ============================================================================================
extern "C" double mo_hier_mh_alpha(const double* xs, const double* ys, size_t n, const double cur[4], const double prop[4],
                                   int move, double parg, double out[3]) {
    int L = cur[0] != 0., Lp = prop[0] != 0.;
    double w = mo_hier_logjp(xs, ys, n, prop) - mo_hier_logjp(xs, ys, n, cur);   // update weight == delta logjp
    double fwd = 0., bwd = 0.;
    if (move == MO_MOVE_HIER_DRIFT) {                  // hierarchical.rs:63-71
        fwd = mo_normal_logpdf(prop[1], cur[1], parg) + mo_normal_logpdf(prop[2], cur[2], parg);
        bwd = mo_normal_logpdf(cur[1], prop[1], parg) + mo_normal_logpdf(cur[2], prop[2], parg);
        if (!L) { fwd += mo_normal_logpdf(prop[3], cur[3], parg); bwd += mo_normal_logpdf(cur[3], prop[3], parg); }
    } else if (move == MO_MOVE_HIER_ADD_REMOVE) {      // hierarchical.rs:48-61
        double prev_c = L ? 0. : cur[3], prev_c_bwd = Lp ? 0. : prop[3];
        fwd = mo_normal_logpdf(prop[1], cur[1], parg) + mo_normal_logpdf(prop[2], cur[2], parg) + mo_bernoulli_logpdf(Lp, 0.5);
        if (!Lp) fwd += mo_normal_logpdf(prop[3], prev_c, parg);
        bwd = mo_normal_logpdf(cur[1], prop[1], parg) + mo_normal_logpdf(cur[2], prop[2], parg) + mo_bernoulli_logpdf(L, 0.5);
        if (!L) bwd += mo_normal_logpdf(cur[3], prev_c_bwd, parg);
    }
    if (out) { out[0] = w; out[1] = fwd; out[2] = bwd; }
    return w - fwd + bwd;                                // mh.rs:34
}

struct mo_chains {
    std::string model; std::vector<double> args, obs;
    uint64_t n, seed, offset; int slots;
    std::vector<double> st;            // slots x n SoA
    std::vector<uint32_t> step;        // per-chain move counter (RNG tag)
};

extern "C" mo_chains* mo_chains_new(const char* model, const double* args, size_t n_args, const double* obs, size_t n_obs,
                                    uint64_t n_chains, uint64_t seed, uint64_t chain_offset) {
    auto* c = new mo_chains;
    c->model = model; c->args.assign(args, args + n_args); c->obs.assign(obs, obs + n_obs);
    c->n = n_chains; c->seed = seed; c->offset = chain_offset;
    if (c->model == "hierarchical") c->slots = 5; else if (c->model == "pointed") c->slots = 3; else { delete c; return nullptr; }
    c->st.assign((size_t)c->slots * n_chains, 0.); c->step.assign(n_chains, 0);
    size_t n = n_chains;
    for (size_t i = 0; i < n; ++i) {       // tests/mh.rs:34,61,91: trace = model.generate(args, observations).0
        Rng64 g(seed, chain_offset + i, 0, P_MH_INIT);
        if (c->model == "hierarchical") {
            int L = 0.7 > g.uniform(); double a, b; g.normal2(a, b); double cc = g.normal(); if (L) cc = 0.;
            double s4[4] = {(double)L, a, b, cc};
            c->st[0 * n + i] = L; c->st[1 * n + i] = a; c->st[2 * n + i] = b; c->st[3 * n + i] = cc;
            c->st[4 * n + i] = mo_hier_logjp(c->args.data(), c->obs.data(), n_obs, s4);
        } else {
            double u0, u1; g.uniform2(u0, u1);
            const double* bd = c->args.data();
            double lat[2] = {u0 * (bd[1] - bd[0]) + bd[0], u1 * (bd[3] - bd[2]) + bd[2]};
            c->st[0 * n + i] = lat[0]; c->st[1 * n + i] = lat[1];
            c->st[2 * n + i] = mo_uniform2d_logpdf(lat[0], lat[1], bd) + mo_mvnormal_logpdf(c->obs.data(), lat, bd + 4, 2);
        }
    }
    return c;
}
extern "C" void mo_chains_free(mo_chains* c) { delete c; }
extern "C" int mo_chains_num_slots(mo_chains* c) { return c->slots; }
extern "C" void mo_chains_read(mo_chains* c, double* out) { std::copy(c->st.begin(), c->st.end(), out); }
extern "C" void mo_chains_write(mo_chains* c, const double* in) { std::copy(in, in + c->st.size(), c->st.begin()); }

extern "C" int mo_chains_move(mo_chains* c, int move, double parg, uint32_t mask, uint32_t n_steps, uint64_t* n_accepted) {
    size_t n = c->n; uint64_t acc = 0;
    const double* xs = c->args.data(); const double* ys = c->obs.data(); size_t m = c->obs.size();
    bool hier = c->model == "hierarchical";
    if (hier != (move != MO_MOVE_POINTED_DRIFT)) return -1;
    for (size_t i = 0; i < n; ++i) {
        for (uint32_t s = 0; s < n_steps; ++s) {
            Rng64 g(c->seed, c->offset + i, c->step[i]++, P_MH);
            if (move == MO_MOVE_POINTED_DRIFT) {
                // mh.rs:9-40 with pointed_model/{model,proposal}.rs
                const double* bd = xs; const double* cov = xs + 4;
                double lat[2] = {c->st[0 * n + i], c->st[1 * n + i]}; double logjp = c->st[2 * n + i];
                double z0, z1; g.normal2(z0, z1);
                double nl[2] = {parg * z0 + lat[0], parg * z1 + lat[1]};              // mvnormal.rs:36  L z + mu, L = s I
                double dcov[4] = {parg * parg, 0., 0., parg * parg};
                double fwd = mo_mvnormal_logpdf(nl, lat, dcov, 2);                    // proposal.rs:24
                double new_logjp = logjp;                                             // model.rs:76-102
                new_logjp -= mo_uniform2d_logpdf(lat[0], lat[1], bd);
                new_logjp += mo_uniform2d_logpdf(nl[0], nl[1], bd);
                new_logjp -= mo_mvnormal_logpdf(ys, lat, cov, 2);
                new_logjp += mo_mvnormal_logpdf(ys, nl, cov, 2);
                double w = new_logjp - logjp;
                double bwd = mo_mvnormal_logpdf(lat, nl, dcov, 2);                    // proposal.rs:39
                double alpha = w - fwd + bwd;
                g.skip(2);
                double u = g.uniform();
                if (std::log(u) < alpha) { c->st[0 * n + i] = nl[0]; c->st[1 * n + i] = nl[1]; c->st[2 * n + i] = new_logjp; acc++; }
                continue;
            }
            double cur[4] = {c->st[0 * n + i], c->st[1 * n + i], c->st[2 * n + i], c->st[3 * n + i]};
            double logjp = c->st[4 * n + i];
            int L = cur[0] != 0.;
            double za, zb; g.normal2(za, zb);       // blk 0
            double zc = g.normal();                 // blk 1
            double uf = g.uniform();                // blk 2
            double ua = g.uniform();                // blk 3
            double prop[4] = {cur[0], cur[1], cur[2], cur[3]};
            double alpha, new_logjp;
            if (move == MO_MOVE_HIER_DRIFT) {
                prop[1] = za * parg + cur[1]; prop[2] = zb * parg + cur[2];
                if (!L) prop[3] = zc * parg + cur[3];
                alpha = mo_hier_mh_alpha(xs, ys, m, cur, prop, move, parg, nullptr);
                new_logjp = mo_hier_logjp(xs, ys, m, prop);
            } else if (move == MO_MOVE_HIER_ADD_REMOVE) {
                prop[1] = za * parg + cur[1]; prop[2] = zb * parg + cur[2];
                int Lp = 0.5 > uf;
                prop[0] = Lp;
                double prev_c = L ? 0. : cur[3];
                prop[3] = Lp ? 0. : zc * parg + prev_c;
                alpha = mo_hier_mh_alpha(xs, ys, m, cur, prop, move, parg, nullptr);
                new_logjp = mo_hier_logjp(xs, ys, m, prop);
            } else if (move == MO_MOVE_HIER_REGEN) {
                // mh.rs:54-67 ; dyngenfn.rs:223-266: masked choices resampled from the prior (weight 0), later
                // choices re-scored => alpha = delta log-likelihood.  bit 8 (is_linear) is an engine extension
                // (the reference panics on quadratic->linear, SURVEY 3.4 "Hazard").
                int Ln = (mask & 8u) ? (0.7 > uf) : L;        // the branch of the *new* trace decides what is visited
                prop[0] = Ln;
                if (mask & 1u) prop[1] = za * 1. + 0.;
                if (mask & 2u) prop[2] = zb * 1. + 0.;
                if (Ln) prop[3] = 0.;                          // c absent (n/a, or dropped: extension)
                else if (L) prop[3] = zc * 1. + 0.;            // c is a new choice: sampled from the prior, no weight (:261-266)
                else if (mask & 4u) prop[3] = zc * 1. + 0.;    // c masked: resampled from the prior (:223-231)
                alpha = hier_loglik(xs, ys, m, Ln, prop[1], prop[2], prop[3]) - hier_loglik(xs, ys, m, L, cur[1], cur[2], cur[3]);
                new_logjp = mo_hier_logjp(xs, ys, m, prop);
            } else return -2;
            (void)logjp;
            if (std::log(ua) < alpha) {                                            // mh.rs:35 / :62
                for (int k = 0; k < 4; ++k) c->st[k * n + i] = prop[k];
                c->st[4 * n + i] = new_logjp; acc++;
            }
        }
    }
    if (n_accepted) *n_accepted = acc;
    return 0;
}

// ============================================================================================
// ground truths
// ============================================================================================
extern "C" double mo_hmm_forward(const double* prior, const double* emission, const double* transition, int K, int M,
                                 const int* obs, int T) {
    // tests/hmm/forward.rs:3-23
    (void)M;
    double ml = 1.0;
    std::vector<double> alpha(prior, prior + K), post(K);
    for (int t = 0; t < T; ++t) {
        double evidence = 0.;
        for (int s = 0; s < K; ++s) { post[s] = alpha[s] * emission[obs[t] * K + s]; evidence += post[s]; }
        for (int s = 0; s < K; ++s) post[s] /= evidence;
        for (int to = 0; to < K; ++to) { double a = 0.; for (int f = 0; f < K; ++f) a += transition[to * K + f] * post[f]; alpha[to] = a; }
        ml *= evidence;
    }
    return ml;
}

namespace {
// small dense helpers for the closed-form truths
bool cholesky(std::vector<double>& a, int n) {   // in place, lower
    for (int j = 0; j < n; ++j) {
        double d = a[j * n + j];
        for (int k = 0; k < j; ++k) d -= a[j * n + k] * a[j * n + k];
        if (d <= 0.) return false;
        d = std::sqrt(d); a[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = a[i * n + j];
            for (int k = 0; k < j; ++k) s -= a[i * n + k] * a[j * n + k];
            a[i * n + j] = s / d;
        }
    }
    return true;
}
double gauss_logpdf_zero_mean(const double* y, std::vector<double> cov, int n) {
    if (!cholesky(cov, n)) return std::numeric_limits<double>::quiet_NaN();
    double logdet = 0., mahal = 0.;
    std::vector<double> v(n);
    for (int i = 0; i < n; ++i) {
        double s = y[i];
        for (int k = 0; k < i; ++k) s -= cov[i * n + k] * v[k];
        v[i] = s / cov[i * n + i];
        mahal += v[i] * v[i];
        logdet += 2. * std::log(cov[i * n + i]);
    }
    return -0.5 * ((double)n * std::log(2. * kPi) + logdet + mahal);
}
}  // namespace

extern "C" double mo_kalman_lml_lgssm4(double q, double r, double x0, const double* ys, int T) {
    // x_0 ~ N(0, x0^2 I); y_t = H x_t + N(0, r^2 I) for t = 0..T-1 ; x_t = A x_{t-1} + N(0, q^2 I)
    double m[4] = {0, 0, 0, 0}, P[16] = {0};
    for (int i = 0; i < 4; ++i) P[i * 4 + i] = x0 * x0;
    const double A[16] = {1, 0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1};
    double lml = 0.;
    for (int t = 0; t < T; ++t) {
        if (t > 0) {   // predict
            double m2[4], AP[16], P2[16];
            for (int i = 0; i < 4; ++i) { m2[i] = 0; for (int j = 0; j < 4; ++j) m2[i] += A[i * 4 + j] * m[j]; }
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * P[k * 4 + j]; AP[i * 4 + j] = s; }
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += AP[i * 4 + k] * A[j * 4 + k]; P2[i * 4 + j] = s + (i == j ? q * q : 0.); }
            std::memcpy(m, m2, sizeof m); std::memcpy(P, P2, sizeof P);
        }
        // update with H = [I2 0]
        double S[4] = {P[0] + r * r, P[1], P[4], P[5] + r * r};
        double v[2] = {ys[2 * t] - m[0], ys[2 * t + 1] - m[1]};
        double det = S[0] * S[3] - S[1] * S[2];
        double Si[4] = {S[3] / det, -S[1] / det, -S[2] / det, S[0] / det};
        double mahal = v[0] * (Si[0] * v[0] + Si[1] * v[1]) + v[1] * (Si[2] * v[0] + Si[3] * v[1]);
        lml += -0.5 * (2. * std::log(2. * kPi) + std::log(det) + mahal);
        double Kg[8];   // 4x2 = P H^T S^-1
        for (int i = 0; i < 4; ++i) { Kg[i * 2 + 0] = P[i * 4 + 0] * Si[0] + P[i * 4 + 1] * Si[2]; Kg[i * 2 + 1] = P[i * 4 + 0] * Si[1] + P[i * 4 + 1] * Si[3]; }
        double P2[16];
        for (int i = 0; i < 4; ++i) m[i] += Kg[i * 2] * v[0] + Kg[i * 2 + 1] * v[1];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) P2[i * 4 + j] = P[i * 4 + j] - (Kg[i * 2] * P[0 * 4 + j] + Kg[i * 2 + 1] * P[1 * 4 + j]);
        std::memcpy(P, P2, sizeof P);
    }
    return lml;
}

extern "C" double mo_line_model_lml(const double* xs, const double* ys, int n) {
    // y = slope*x + intercept + N(0,.1^2), slope~N(0,1), intercept~N(0,2^2)  =>  y ~ N(0, x x^T + 4 11^T + .01 I)
    std::vector<double> cov(n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) cov[i * n + j] = xs[i] * xs[j] + 4. + (i == j ? 0.01 : 0.);
    return gauss_logpdf_zero_mean(ys, cov, n);
}

extern "C" double mo_hier_model_lml(const double* xs, const double* ys, int n, double* p_linear) {
    std::vector<double> cl(n * n), cq(n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
        cl[i * n + j] = 1. + xs[i] * xs[j] + (i == j ? 0.01 : 0.);
        cq[i * n + j] = cl[i * n + j] + xs[i] * xs[i] * xs[j] * xs[j];
    }
    double ll = std::log(0.7) + gauss_logpdf_zero_mean(ys, cl, n), lq = std::log(0.3) + gauss_logpdf_zero_mean(ys, cq, n);
    double v[2] = {ll, lq};
    double tot = mo_logsumexp(v, 2);
    if (p_linear) *p_linear = std::exp(ll - tot);
    return tot;
}
