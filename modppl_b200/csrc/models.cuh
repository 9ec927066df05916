// models.cuh -- device functors: the "restricted vectorisable form" of modppl's Unfold kernels.
//
// A functor mirrors DynUnfold::generate / update(Extend) (reference modppl/src/modeling/dynunfold.rs:41-100) for one
// particle: `kernel(t, stream, x, obs)` sees t = 0 at init_step and t = k at the k-th step (quirk Q5), mutates the
// fixed-shape state x in place (state = last retv, dynunfold.rs:76) and returns the weight = sum of the constrained
// choices' log-densities (Generate-mode semantics, dyngenfn.rs:121-136).
#pragma once
#include "common.cuh"

namespace mpl {

struct Obs {
    double v[4];
};

enum ModelKind : int { M_LGSSM4 = 0, M_SPIRAL = 1, M_SV = 2, M_HMM = 3, M_LINE = 10, M_HIER = 11, M_POINTED = 12,
                       M_JIT = 20 /* an Unfold model compiled from a spec at run time (jit.cu) */ };

// ---- config 4: 4-D constant-velocity linear-Gaussian tracker, two independent `normal` observations -----------
template <typename Real>
struct Lgssm4 {
    static constexpr int D = 4;
    static constexpr int NOBS = 2;
    static constexpr bool kGroupDraws = false;
    Real q, r, x0, ln_r, inv_r, lw_const;   // lw_const = ln 2pi + 2 ln r
    __device__ __forceinline__ Real kernel(int64_t t, const Stream& s, Real (&x)[D], const Obs& obs) const {
        Real z[4];
        draw_normals<4>(s, 0, z);
        if (t == 0) {
#pragma unroll
            for (int d = 0; d < 4; ++d) x[d] = z[d] * x0;
        } else {
            x[0] = x[0] + x[2] + z[0] * q;
            x[1] = x[1] + x[3] + z[1] * q;
            x[2] = x[2] + z[2] * q;
            x[3] = x[3] + z[3] * q;
        }
        // two independent `normal` observations (normal.rs:13-17), summed in closed form:
        //   sum_d -(z_d^2 + ln 2pi)/2 - ln r  =  -(z_0^2 + z_1^2)/2 - (ln 2pi + 2 ln r)
        const Real z0 = ((Real)obs.v[0] - x[0]) * inv_r, z1 = ((Real)obs.v[1] - x[1]) * inv_r;   // 1/std hoisted
        return (Real)-0.5 * (z0 * z0 + z1 * z1) - lw_const;
    }
};

// ---- config 1: spiral model, reference tests/dyngenfns/unfold.rs:14-33 ------------------------------------------
template <typename Real>
struct Spiral {
    static constexpr int D = 2;
    static constexpr int NOBS = 2;
    static constexpr bool kGroupDraws = false;
    Real dr_std, dth_mean, dth_std;
    double prec[4], log_norm;   // mvnormal.rs:14-22 with det/inverse hoisted (quirk Q8)
    __device__ __forceinline__ Real kernel(int64_t t, const Stream& s, Real (&x)[D], const Obs& obs) const {
        if (t == 0) {
            Real u[2];
            draw_uniforms<2>(s, 0, u);
            x[0] = u[0] * (Real)(1. - 0.) + (Real)0.;                     // uniform.rs:28-32
            x[1] = u[1] * (Real)(2. * kPi - 0.) + (Real)0.;
        } else {
            Real z[2];
            draw_normals<2>(s, 0, z);
            x[0] = x[0] + (z[0] * dr_std + (Real)0.);                     // normal.rs:26
            x[1] = x[1] + (z[1] * dth_std + dth_mean);
        }
        // mvnormal.rs:14-22 in double for both precisions (det / inverse hoisted, quirk Q8) -- the form a run-time spec of this model
        // compiles to as well (jit.cu), so the two agree to the last bit
        return (Real)mvnormal2_logpdf((double)(Real)obs.v[0], (double)(Real)obs.v[1], (double)(x[0] * cos(x[1])), (double)(x[0] * sin(x[1])), prec, log_norm);
    }
};

// ---- config 5: stochastic volatility -------------------------------------------------------------------------------
template <typename Real>
struct StochVol {
    static constexpr int D = 1;
    static constexpr int NOBS = 1;
    // One normal deviate per particle and step: a Philox block (two Box-Muller pairs in fp32, one in fp64) serves the 4 (2)
    // consecutive particles one thread handles -- stream id = global id / 4 (2), purpose P_MODEL_GROUP -- instead of one block
    // per particle of which three quarters would be thrown away.
    static constexpr bool kGroupDraws = true;
    Real mu, phi, sig, sd0;
    __device__ __forceinline__ Real kernel_z(int64_t t, Real z, Real (&x)[D], const Obs& obs) const {
        if (t == 0) x[0] = mu + sd0 * z;
        else x[0] = mu + phi * (x[0] - mu) + sig * z;
        Real sd = exp(x[0] / 2);
        Real zz = (Real)obs.v[0] / sd;
        return -(zz * zz + (Real)1.8378770664093453) / 2 - x[0] / 2;
    }
};

// ---- K-state HMM, reference tests/hmm/model.rs:24-81 --------------------------------------------------------------
constexpr int kHmmMaxK = 8;
template <typename Real>
struct Hmm {
    static constexpr int D = 1;
    static constexpr int NOBS = 1;
    static constexpr bool kGroupDraws = false;
    int K, M;
    double prior[kHmmMaxK], log_emis[kHmmMaxK * kHmmMaxK], trans[kHmmMaxK * kHmmMaxK];   // log_emis[o*K+s], trans[to*K+from]
    __device__ __forceinline__ Real kernel(int64_t t, const Stream& s, Real (&x)[D], const Obs& obs) const {
        double u[1];
        draw_uniforms<1>(s, 0, u);
        int prev = (t == 0) ? 0 : (int)x[0];
        // categorical.rs:22-32: x = first index whose running sum reaches u, clamped (quirk Q2)
        double acc = 0.;
        int st = -1;
        for (int k = 0; k < K; ++k) {
            if (!(acc < u[0])) break;
            acc += (t == 0) ? prior[k] : trans[k * K + prev];
            st = k;
        }
        st = st < 0 ? 0 : st;
        x[0] = (Real)st;
        int o = (int)obs.v[0];
        return (Real)log_emis[o * K + st];
    }
};

}  // namespace mpl
