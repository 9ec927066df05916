// nested_quant.cuh -- per-chunk quantisation of the nested systematic scheme (see nested.cuh); needed by the extend kernel's
// fused epilogue, hence separate.
#pragma once
#include "common.cuh"

namespace mpl {

constexpr int kChunk = 128;                       // particles per chunk == one warp x 4 particles

struct ChunkRecords {
    int* e;                  // power-of-two reference of the chunk (INT_MIN: no finite weight)
    unsigned long long* S;   // sum of the chunk's integer weights
    float* sq;               // sum of squared integer weights (for the ESS), as a float
};
constexpr int kChunkEmpty = -2147483647 - 1;

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

// integer weight of one particle against the chunk reference 2^e_c; y = lw * log2(e) already multiplied
__device__ __forceinline__ unsigned long long nested_weight(float y, float e_c, int kbits, float* qf) {
    float z = fmaxf(__fsub_rn(y, e_c), -126.0f);            // fmaxf(NaN, x) = x; below -126 the weight rounds to 0
    float t = __fadd_rn(z, 12582912.0f);
    float n = __fsub_rn(t, 12582912.0f);
    int ni = __float_as_int(t) - 0x4B400000;
    float f = __fsub_rn(z, n);
    float p = exp2_poly(f);
    float scale = __int_as_float((127 + kbits + ni) << 23);  // >= 2^(36 - 126 + 127) > 0: never denormal
    float v = __fmul_rn(p, scale);
    *qf = v;
    return __float2ull_rn(v);
}

// One warp quantises one chunk: w[4] are the lane's 4 consecutive log-weights (already masked: invalid lanes hold -inf /
// NaN).  Returns the integer weights (as Real-exact floats in qv) and leaves the chunk record with lane 0.
__device__ __forceinline__ void warp_quantise_chunk(const float (&w)[4], int kbits, float (&qv)[4], int& e_c, unsigned long long& S_c, float& sq_c) {
    float y[4], ymax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) { y[j] = __fmul_rn(w[j], 1.44269504088896341f); ymax = fmaxf(ymax, y[j]); }   // fmaxf skips NaN
    ymax = warp_max_f32(ymax);
    if (!(ymax > -INFINITY)) {   // no finite weight in the chunk
        e_c = kChunkEmpty; S_c = 0ull; sq_c = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) qv[j] = 0.f;
        return;
    }
    const float ef = ceilf(ymax);
    e_c = (int)ef;
    unsigned long long s = 0;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float qf;
        unsigned long long q = nested_weight(y[j], ef, kbits, &qf);
        qv[j] = rintf(qf);
        s += q;
        sq = fmaf(qf, qf, sq);
    }
    // the lane's sum is below 2^42 (4 weights of at most 2^40): two integer redux instructions give the exact chunk sum
    S_c = ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned int)(s >> 24)) << 24) + (unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned int)s & 0xFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    sq_c = sq;
}


}  // namespace mpl
