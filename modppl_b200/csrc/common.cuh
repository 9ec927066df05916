// common.cuh -- shared device helpers: error handling, Philox4x32-10, uniform/normal transforms, reductions.
// sm_100a only.
#pragma once
#ifdef __CUDACC_RTC__
// compiled at run time by NVRTC (csrc/jit.cu: the kernel headers are embedded in the library and a model spec becomes one more
// functor): no host headers there, only what device code needs
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#define INFINITY __int_as_float(0x7f800000)
#define NAN __int_as_float(0x7fffffff)
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#endif

namespace mpl {

#ifndef __CUDACC_RTC__
// ------------------------------------------------------------------------------------------------
// error plumbing (thread-local message behind mpl_last_error())
// ------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
#endif

#define MPL_CUDA_OK(expr)                                                                                     \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return ::mpl::fail(-2, std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                                       std::to_string(__LINE__));                                             \
    } while (0)

constexpr int kNumSMs = 148;   // B200

// Programmatic dependent launch: the next kernel of the step is launched while this one drains and its blocks wait here
// until this grid has completed and its memory is visible.  Hides ~4 us of kernel-boundary latency per launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
constexpr double kPi = 3.14159265358979323846;

// ------------------------------------------------------------------------------------------------
// Philox4x32-10.  ctr = {id_lo, id_hi, t, purpose<<24 | block}, key = seed.  Replaces ThreadRng
// (reference modppl/src/modeling/dists/distribution.rs:5-7): counter-based, so a draw depends only on
// (seed, global id, step, purpose, block) and never on which GPU or thread computes it.
// ------------------------------------------------------------------------------------------------
enum Purpose : uint32_t { P_MODEL = 0, P_RESAMPLE_U = 1, P_RESAMPLE_OFFSET = 2, P_IS = 3, P_MH = 4, P_IS_RESAMPLE = 5, P_MH_INIT = 6, P_MODEL_GROUP = 7 };

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

struct Stream {
    uint2 key;
    uint32_t id_lo, id_hi, t, purpose;
    __host__ __device__ __forceinline__ Stream(uint64_t seed, uint64_t id, uint32_t t_, uint32_t purpose_)
        : key(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))), id_lo((uint32_t)id), id_hi((uint32_t)(id >> 32)), t(t_), purpose(purpose_) {}
    __host__ __device__ __forceinline__ uint4 block(uint32_t blk) const {
        return philox4x32_10(make_uint4(id_lo, id_hi, t, (purpose << 24) | blk), key);
    }
};

__host__ __device__ __forceinline__ double u01_co64(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; }        // [0,1)
__host__ __device__ __forceinline__ double u01_oc64(uint32_t hi, uint32_t lo) { return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * 0x1.0p-53; }  // (0,1]
__host__ __device__ __forceinline__ float u01_co32(uint32_t x) { return (float)(x >> 8) * 0x1.0p-24f; }
__host__ __device__ __forceinline__ float u01_oc32(uint32_t x) { return (float)((x >> 8) + 1) * 0x1.0p-24f; }

// fp32 Box-Muller on the SFU path (the extend kernel is instruction-issue bound, not HBM bound, with libm-grade
// logf/sincospif): -2 ln(u1) from MUFU.LG2 with a 3-term series where u1 is within 2^-6 of 1 (there MUFU.LG2's
// absolute error would dominate the tiny result), sqrt.approx, MUFU.SIN/COS on an argument folded into [-pi, pi).
// Absolute error of a normal deviate ~1e-6, far inside the 1e-4 fp32 parity tolerance.
// k1, k2: the top 24 bits of two Philox words.  u1 = (k1 + 1) 2^-24 in (0, 1], u2 = k2 2^-24 in [0, 1); the 2^-24 scalings
// are folded into the constants of the following FFMAs (lg2(u1) = lg2(k1 + 1) - 24 exactly).
__device__ __forceinline__ float neg2log_fast_k(float kf /* k1 + 1 as float, exact */) {
    float l = fmaf(__log2f(kf), -1.3862943611198906f, 33.27106466687737f);   // -2 ln2 (lg2(kf) - 24)
    float v = fmaf(kf, -0x1.0p-24f, 1.0f);                                     // 1 - u1, exact
    float s = v * fmaf(v, fmaf(v, 0.66666667f, 1.0f), 2.0f);                  // -2 ln(1 - v) = 2v + v^2 + (2/3) v^3 + O(v^4)
    return (v < 0.015625f) ? s : l;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void box_muller_bits(uint32_t x1, uint32_t x2, float& z0, float& z1) {
    float r = -sqrt_approx(neg2log_fast_k((float)((x1 >> 8) + 1u)));
    float a = fmaf((float)(x2 >> 8), 3.7450703e-07f /* 2 pi 2^-24 */, -3.141592653589793f);   // 2 pi u2 - pi in [-pi, pi)
    z0 = r * __cosf(a);                                               // cos(2 pi u2) = -cos(a)
    z1 = r * __sinf(a);
}
__device__ __forceinline__ void box_muller(float u1, float u2, float& z0, float& z1) {   // (generic entry; the hot path uses the _bits form)
    float l = __log2f(u1) * -1.3862943611198906f;
    float v = 1.0f - u1;
    float s = v * fmaf(v, fmaf(v, 0.66666667f, 1.0f), 2.0f);
    float r = -sqrt_approx((v < 0.015625f) ? s : l);
    float a = fmaf(u2, 6.283185307179586f, -3.141592653589793f);
    z0 = r * __cosf(a);
    z1 = r * __sinf(a);
}
__device__ __forceinline__ void box_muller(double u1, double u2, double& z0, double& z1) {
    double r = sqrt(-2. * log(u1));
    double s, c;
    sincospi(2. * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// COUNT standard normals starting at Philox block `first_blk`.
//   float : one block -> two Box-Muller pairs (z[4b..4b+3]);   double: one block -> one pair (z[2b], z[2b+1]).
template <int COUNT>
__device__ __forceinline__ void draw_normals(const Stream& s, uint32_t first_blk, float (&z)[COUNT]) {
#pragma unroll
    for (int i = 0; i < COUNT; i += 4) {
        uint4 x = s.block(first_blk + i / 4);
        float a, b;
        box_muller_bits(x.x, x.y, a, b);
        z[i] = a;
        if (i + 1 < COUNT) z[i + 1] = b;
        if (i + 2 < COUNT) {
            box_muller_bits(x.z, x.w, a, b);
            z[i + 2] = a;
            if (i + 3 < COUNT) z[i + 3] = b;
        }
    }
}
template <int COUNT>
__device__ __forceinline__ void draw_normals(const Stream& s, uint32_t first_blk, double (&z)[COUNT]) {
#pragma unroll
    for (int i = 0; i < COUNT; i += 2) {
        uint4 x = s.block(first_blk + i / 2);
        double a, b;
        box_muller(u01_oc64(x.x, x.y), u01_co64(x.z, x.w), a, b);
        z[i] = a;
        if (i + 1 < COUNT) z[i + 1] = b;
    }
}
template <int COUNT>
__device__ __forceinline__ void draw_uniforms(const Stream& s, uint32_t first_blk, float (&u)[COUNT]) {
#pragma unroll
    for (int i = 0; i < COUNT; i += 4) {
        uint4 x = s.block(first_blk + i / 4);
        u[i] = u01_co32(x.x);
        if (i + 1 < COUNT) u[i + 1] = u01_co32(x.y);
        if (i + 2 < COUNT) u[i + 2] = u01_co32(x.z);
        if (i + 3 < COUNT) u[i + 3] = u01_co32(x.w);
    }
}
template <int COUNT>
__device__ __forceinline__ void draw_uniforms(const Stream& s, uint32_t first_blk, double (&u)[COUNT]) {
#pragma unroll
    for (int i = 0; i < COUNT; i += 2) {
        uint4 x = s.block(first_blk + i / 2);
        u[i] = u01_co64(x.x, x.y);
        if (i + 1 < COUNT) u[i + 1] = u01_co64(x.z, x.w);
    }
}

// fp64 draw helper for the IS / MH paths: every draw consumes one whole Philox block.
struct Rng64 {
    Stream s;
    uint32_t blk;
    __device__ __forceinline__ Rng64(uint64_t seed, uint64_t id, uint32_t t, uint32_t purpose) : s(seed, id, t, purpose), blk(0) {}
    __device__ __forceinline__ double uniform() { uint4 x = s.block(blk++); return u01_co64(x.x, x.y); }
    __device__ __forceinline__ void uniform2(double& a, double& b) { uint4 x = s.block(blk++); a = u01_co64(x.x, x.y); b = u01_co64(x.z, x.w); }
    __device__ __forceinline__ void normal2(double& a, double& b) { uint4 x = s.block(blk++); box_muller(u01_oc64(x.x, x.y), u01_co64(x.z, x.w), a, b); }
    __device__ __forceinline__ double normal() { double a, b; normal2(a, b); return a; }
    __device__ __forceinline__ void skip(uint32_t n) { blk += n; }
};

// ------------------------------------------------------------------------------------------------
// built-in log-densities (reference modppl/src/modeling/dists/*.rs), generic over Real
// ------------------------------------------------------------------------------------------------
template <typename Real>
__host__ __device__ __forceinline__ Real normal_logpdf(Real x, Real mu, Real sd) {
    // normal.rs:13-17   -(|z|^2 + ln 2pi)/2 - ln std
    Real z = (x - mu) / sd;
    return -(z * z + (Real)1.8378770664093453) / 2 - log(sd);
}
// ---- the remaining built-in distributions (reference src/modeling/dists/*.rs; known answers tests/dists.rs:60-69,186-212) ----
__host__ __device__ __forceinline__ double uniform_discrete_logpdf(long long x, long long a, long long b) {           // uniform.rs:43-47
    return (a <= x && x <= b) ? -log((double)(b - a + 1)) : -INFINITY;
}
__host__ __device__ __forceinline__ double geometric_logpdf(long long k, double p) { return log(pow(1. - p, (double)k) * p); }   // geometric.rs:16-19, literally
__host__ __device__ __forceinline__ double poisson_logpdf(long long k, double rate) {                                  // poisson.rs:16-18
    double s = 0.;
    for (long long v = 1; v <= k; ++v) s += log((double)v);
    return (double)k * log(rate) - rate - s;
}
__host__ __device__ __forceinline__ double beta_logpdf(double x, double a, double b) {                                 // beta.rs:17-21, literally
    const double beta_f = tgamma(a + b) / (tgamma(a) * tgamma(b));
    return log(beta_f * pow(x, a - 1.) * pow(1. - x, b - 1.));
}
__host__ __device__ __forceinline__ double gamma_logpdf(double x, double a, double b) {                                // gamma.rs:17-20 (shape a, scale b)
    return (a - 1.) * log(x) - x / b - log(tgamma(a)) - a * log(b);
}
__host__ __device__ __forceinline__ double bernoulli_logpdf(bool a, double p) { return log(a ? p : 1. - p); }   // bernoulli.rs:12-14
__host__ __device__ __forceinline__ double uniform_logpdf(double x, double a, double b) {                        // uniform.rs:22-26
    return (a <= x && x <= b) ? -log(b - a) : -INFINITY;
}
__host__ __device__ __forceinline__ double uniform2d_logpdf(double x, double y, const double* bd) {              // tests/pointed_model/types_2d.rs:15-21
    return (bd[0] <= x && x <= bd[1] && bd[2] <= y && y <= bd[3]) ? -log((bd[1] - bd[0]) * (bd[3] - bd[2])) : -INFINITY;
}
// mvnormal.rs:14-22 for k = 2, with det and inverse hoisted to the host once per model (quirk Q8):
// prec = {inv00, inv01, inv10, inv11}, log_norm = k ln 2pi + ln det
__host__ __device__ __forceinline__ double mvnormal2_logpdf(double x0, double x1, double m0, double m1, const double* prec, double log_norm) {
    double c0 = x0 - m0, c1 = x1 - m1;
    double t0 = c0 * prec[0] + c1 * prec[2], t1 = c0 * prec[1] + c1 * prec[3];
    return -(log_norm + (t0 * c0 + t1 * c1)) / 2.;
}

// mvnormal.rs:14-22 for any k <= 8, literally: determinant and inverse of the covariance recomputed per call (row-major cov).
// nalgebra 0.32.2 (modppl/Cargo.toml:17; not vendored) special-cases k <= 3 with cofactor closed forms -- restated here so that
// the reference's known answers (tests/dists.rs:164-183, k = 2 and k = 3) are reproduced to the last bits -- and uses LU beyond.
constexpr int kMvnMaxK = 8;
__host__ __device__ inline double mvnormal_logpdf_k(const double* x, const double* mu, const double* cov, int k) {
    if (k < 1 || k > kMvnMaxK) return NAN;
    double inv[kMvnMaxK * kMvnMaxK], det;
    if (k == 1) { det = cov[0]; inv[0] = 1. / cov[0]; }
    else if (k == 2) {
        const double m11 = cov[0], m12 = cov[1], m21 = cov[2], m22 = cov[3];
        det = m11 * m22 - m21 * m12;
        inv[0] = m22 / det; inv[1] = -m12 / det; inv[2] = -m21 / det; inv[3] = m11 / det;
    } else if (k == 3) {
        const double m11 = cov[0], m12 = cov[1], m13 = cov[2], m21 = cov[3], m22 = cov[4], m23 = cov[5], m31 = cov[6], m32 = cov[7], m33 = cov[8];
        const double minor_m12_m23 = m22 * m33 - m32 * m23, minor_m11_m23 = m21 * m33 - m31 * m23, minor_m11_m22 = m21 * m32 - m31 * m22;
        det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
        inv[0] = minor_m12_m23 / det; inv[1] = (m13 * m32 - m33 * m12) / det; inv[2] = (m12 * m23 - m22 * m13) / det;
        inv[3] = -minor_m11_m23 / det; inv[4] = (m11 * m33 - m31 * m13) / det; inv[5] = (m13 * m21 - m23 * m11) / det;
        inv[6] = minor_m11_m22 / det; inv[7] = (m12 * m31 - m32 * m11) / det; inv[8] = (m11 * m22 - m21 * m12) / det;
    } else {
        // LU with partial pivoting for the determinant, Gauss-Jordan for the inverse
        double a[kMvnMaxK * kMvnMaxK];
        for (int i = 0; i < k * k; ++i) a[i] = cov[i];
        det = 1.;
        for (int c = 0; c < k; ++c) {
            int p = c;
            for (int r = c + 1; r < k; ++r) if (fabs(a[r * k + c]) > fabs(a[p * k + c])) p = r;
            if (a[p * k + c] == 0.) { det = 0.; break; }
            if (p != c) { for (int j = 0; j < k; ++j) { double t = a[p * k + j]; a[p * k + j] = a[c * k + j]; a[c * k + j] = t; } det = -det; }
            det *= a[c * k + c];
            for (int r = c + 1; r < k; ++r) {
                const double f = a[r * k + c] / a[c * k + c];
                for (int j = c; j < k; ++j) a[r * k + j] -= f * a[c * k + j];
            }
        }
        double g[kMvnMaxK * 2 * kMvnMaxK];
        for (int r = 0; r < k; ++r) for (int c = 0; c < 2 * k; ++c) g[r * 2 * k + c] = c < k ? cov[r * k + c] : (c - k == r ? 1. : 0.);
        for (int c = 0; c < k; ++c) {
            int p = c;
            for (int r = c + 1; r < k; ++r) if (fabs(g[r * 2 * k + c]) > fabs(g[p * 2 * k + c])) p = r;
            if (g[p * 2 * k + c] == 0.) return NAN;
            if (p != c) for (int j = 0; j < 2 * k; ++j) { double t = g[p * 2 * k + j]; g[p * 2 * k + j] = g[c * 2 * k + j]; g[c * 2 * k + j] = t; }
            const double d = g[c * 2 * k + c];
            for (int j = 0; j < 2 * k; ++j) g[c * 2 * k + j] /= d;
            for (int r = 0; r < k; ++r) if (r != c) {
                const double f = g[r * 2 * k + c];
                for (int j = 0; j < 2 * k; ++j) g[r * 2 * k + j] -= f * g[c * 2 * k + j];
            }
        }
        for (int r = 0; r < k; ++r) for (int c = 0; c < k; ++c) inv[r * k + c] = g[r * 2 * k + k + c];
    }
    if (det == 0.) return NAN;
    double c[kMvnMaxK], mahal = 0.;
    for (int i = 0; i < k; ++i) c[i] = x[i] - mu[i];
    for (int j = 0; j < k; ++j) {   // (centered^T * cov_inv) * centered
        double t = 0.;
        for (int i = 0; i < k; ++i) t += c[i] * inv[i * k + j];
        mahal += t * c[j];
    }
    return -((double)k * 1.8378770664093453 + log(det) + mahal) / 2.;
}

// ------------------------------------------------------------------------------------------------
// online log-sum-exp triple: (m, s, s2) represents sum exp(x) = s*exp(m), sum exp(2x) = s2*exp(2m)
// ------------------------------------------------------------------------------------------------
template <typename Acc>
struct Lse3 {
    Acc m, s, s2;
};
template <typename Acc>
__host__ __device__ __forceinline__ Lse3<Acc> lse3_identity() { return Lse3<Acc>{(Acc)-INFINITY, (Acc)0, (Acc)0}; }
// exp for weight statistics: fp32 uses MUFU.EX2 (the sums carry ~1e-6 relative error in fp32 anyway; the integer
// resampler never reads them), fp64 stays libm-grade.
__device__ __forceinline__ float stat_exp(float x) { return __expf(x); }
__device__ __forceinline__ double stat_exp(double x) { return exp(x); }

template <typename Acc>
__device__ __forceinline__ Lse3<Acc> lse3_combine(const Lse3<Acc>& a, const Lse3<Acc>& b) {
    Acc m = fmax(a.m, b.m);   // fmax ignores NaN like f64::max (lib.rs:35)
    if (m == (Acc)-INFINITY) return Lse3<Acc>{m, (Acc)0, (Acc)0};
    Acc ea = stat_exp(a.m - m), eb = stat_exp(b.m - m);
    return Lse3<Acc>{m, a.s * ea + b.s * eb, a.s2 * ea * ea + b.s2 * eb * eb};
}
template <typename Acc>
__device__ __forceinline__ Lse3<Acc> lse3_shfl_xor(const Lse3<Acc>& v, int lane_mask) {
    return Lse3<Acc>{__shfl_xor_sync(0xffffffffu, v.m, lane_mask), __shfl_xor_sync(0xffffffffu, v.s, lane_mask), __shfl_xor_sync(0xffffffffu, v.s2, lane_mask)};
}
template <typename Acc>
__device__ __forceinline__ Lse3<Acc> lse3_warp_reduce(Lse3<Acc> v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = lse3_combine(v, lse3_shfl_xor(v, o));
    return v;
}

// order-preserving float <-> uint mapping for atomicMax on floats
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// ------------------------------------------------------------------------------------------------
// fixed-point weights: q = rint(exp(d) * 2^kbits), built only from correctly-rounded IEEE operations so that a
// CPU restatement reproduces it bit for bit (oracle/modppl_oracle.cpp: mo_fixed_weight).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float exp2_poly(float f) {
    float p = 1.5252733804059841e-05f;
    p = __fmaf_rn(p, f, 0.00015403530393381608f);
    p = __fmaf_rn(p, f, 0.0013333558146428443f);
    p = __fmaf_rn(p, f, 0.009618129107628477f);
    p = __fmaf_rn(p, f, 0.05550410866482158f);
    p = __fmaf_rn(p, f, 0.2402265069591007f);
    p = __fmaf_rn(p, f, 0.6931471805599453f);
    p = __fmaf_rn(p, f, 1.0f);
    return p;
}
__device__ __forceinline__ uint64_t fixed_weight(float d, int kbits, float* qf = nullptr) {
    // branch-free; d = lw - max <= 0.  NaN and anything below -100 (incl. -inf) come out as 0.  *qf: the weight as a float.
    d = fmaxf(d, -100.0f);                                   // fmaxf(NaN, x) = x
    float y = __fmul_rn(d, 1.44269504088896341f);
    float t = __fadd_rn(y, 12582912.0f);                     // 1.5 * 2^23: rounds y to the nearest integer (ties to even)
    float n = __fsub_rn(t, 12582912.0f);
    int ni = __float_as_int(t) - 0x4B400000;                 // the same integer, from the mantissa bits
    float f = __fsub_rn(y, n);                               // exact, in [-0.5, 0.5]
    float p = exp2_poly(f);                                  // in [0.70, 1.42]
    float scale = __int_as_float((127 + kbits + ni) << 23);  // 2^(kbits + n): exponent >= 127 + 36 - 145 > 0, never denormal
    float v = __fmul_rn(p, scale);                           // exact product
    if (qf) *qf = v;
    return __float2ull_rn(v);                                // round-to-nearest-even to an integer
}
__host__ __device__ inline int fixed_kbits(uint64_t n_total) {
    int lg = 0;
    while (((uint64_t)1 << lg) < n_total) ++lg;
    int k = 62 - lg;
    return k > 40 ? 40 : k;
}

}  // namespace mpl
