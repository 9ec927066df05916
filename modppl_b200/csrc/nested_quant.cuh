// nested_quant.cuh -- per-chunk quantisation of the nested systematic scheme (see nested.cuh); needed by the extend kernel's
// fused epilogue, hence separate.
#pragma once
#include "common.cuh"

namespace mpl {

constexpr int kChunk = 128;                       // particles per chunk == one warp x 4 particles

constexpr size_t kSection = (size_t)1 << 17;      // particles per section (1024 chunks; one block of the section pass)

constexpr int kNestedBits = 22;                   // a chunk's integer weights: rint(w_i / 2^e_c * 2^22) <= 2^22, so a chunk sum fits 32 bits

struct ChunkRecords {
    int* e;                  // power-of-two reference of the chunk (INT_MIN: no finite weight)
    unsigned int* S;         // sum of the chunk's integer weights (< 2^30)
    float* sq;               // sum of squared integer weights (for the ESS), as a float
};
constexpr int kChunkEmpty = -2147483647 - 1;

// arrays of the levels above the chunks (nested.cuh)
struct NestedPrefixes {
    unsigned long long* tile_pre;   // [local tile] exclusive prefix of the tile's chunk masses inside its section (scale E_s)
    int* sec_E;                     // [GLOBAL section] E_s (kChunkEmpty: no finite weight)
    unsigned long long* sec_T;      // [GLOBAL section] T_s
    double* sec_sq;                 // [GLOBAL section] sum of squared integer weights at scale E_s (ESS)
    unsigned long long* sec_pre;    // [GLOBAL section] exclusive prefix of M_s                       (top-level pass)
    unsigned long long* sec_M;      // [GLOBAL section] M_s
    unsigned long long* sec_a;      // [GLOBAL section] first output slot of the section                (top-level pass)
    unsigned long long* sec_n;      // [GLOBAL section] number of output slots of the section
    unsigned int* P;                // [GLOBAL chunk, + 1 sentinel] first output slot of the chunk (monotone; chunk c owns [P[c], P[c+1]))  (plan pass)
    unsigned int* F;                // [GLOBAL output tile] the chunk that owns the tile's first slot
    unsigned int sec0;              // global number of this shard's first section
    unsigned int n_sec;             // sections of this shard
    unsigned int n_sec_global;
};

__device__ __forceinline__ unsigned long long splitmix64_mix(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// warp-wide maximum of non-NaN floats: one redux on the order-preserving integer image
__device__ __forceinline__ float warp_max_f32(float v) {
    int i = __float_as_int(v);
    i ^= (i >> 31) & 0x7fffffff;
    i = __reduce_max_sync(0xffffffffu, i);
    i ^= (i >> 31) & 0x7fffffff;
    return __int_as_float(i);
}

// integer weight of one particle against the chunk reference 2^e_c; y = lw * log2(e) already multiplied.
// Everything is IEEE-exact (no SFU approximation), so the oracle reproduces it bit for bit.  The final rounding to an
// integer rides on the float adder: v + 2^23 rounds v <= 2^22 to the nearest (even) integer, which is then simply the low
// mantissa bits -- no conversion instruction.  *qf receives the same integer as a float.
__device__ __forceinline__ unsigned int nested_weight(float y, float e_c, float* qf) {
    float z = fmaxf(__fsub_rn(y, e_c), -126.0f);            // fmaxf(NaN, x) = x; below -126 the weight rounds to 0
    float t = __fadd_rn(z, 12582912.0f);
    float n = __fsub_rn(t, 12582912.0f);
    int ni = __float_as_int(t) - 0x4B400000;
    float f = __fsub_rn(z, n);
    float p = exp2_poly(f);
    float scale = __int_as_float((127 + kNestedBits + ni) << 23);  // >= 2^(22 - 126): never denormal
    float v = __fmul_rn(p, scale);
    float r = __fadd_rn(v, 8388608.0f);
    *qf = __fsub_rn(r, 8388608.0f);
    return (unsigned int)__float_as_int(r) - 0x4B000000u;
}

// how an integer weight is kept in the log-weight array until the expansion kernel has consumed it
template <typename Real> __device__ __forceinline__ Real nested_store(unsigned int q) {
    if constexpr (sizeof(Real) == 4) return __uint_as_float(q); else return (Real)q;
}
template <typename Real> __device__ __forceinline__ unsigned int nested_load(Real v) {
    if constexpr (sizeof(Real) == 4) return __float_as_uint(v); else return __double2uint_rz(v);
}

// One warp quantises one chunk: w[4] are the lane's 4 consecutive log-weights (already masked: invalid lanes hold -inf /
// NaN).  Returns the integer weights and leaves the chunk record with every lane.
__device__ __forceinline__ void warp_quantise_chunk(const float (&w)[4], unsigned int (&q)[4], int& e_c, unsigned int& S_c, float& sq_c) {
    float y[4], ymax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) { y[j] = __fmul_rn(w[j], 1.44269504088896341f); ymax = fmaxf(ymax, y[j]); }   // fmaxf skips NaN
    ymax = warp_max_f32(ymax);
    if (!(ymax > -INFINITY)) {   // no finite weight in the chunk
        e_c = kChunkEmpty; S_c = 0u; sq_c = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = 0u;
        return;
    }
    const float ef = ceilf(ymax);
    e_c = (int)ef;
    unsigned int s = 0;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float qf;
        q[j] = nested_weight(y[j], ef, &qf);
        s += q[j];
        sq = fmaf(qf, qf, sq);
    }
    S_c = __reduce_add_sync(0xffffffffu, s);   // exact: at most 128 * 2^22
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    sq_c = sq;
}


}  // namespace mpl
