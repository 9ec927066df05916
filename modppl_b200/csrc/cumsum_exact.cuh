// cumsum_exact.cuh -- the reference's SEQUENTIAL f64 running sum, computed in parallel, bit for bit.
//
// modppl draws ancestors with `categorical.random` (reference modppl/src/modeling/dists/categorical.rs:22-32): a running
// sum t += probs[x] in index order, each addition rounded to nearest-even.  Floating-point addition is not associative,
// so a tree or look-back scan gives different low bits -- and at N = 2^24 a different ancestor for some draws.  This file
// reproduces the sequential roundings exactly:
//
//   While the running sum S stays inside one binade [2^e, 2^(e+1)) it lives on the grid g = 2^(e-52): S = M*g with M an
//   integer, and fl(S + p) = (M + I + r)*g where p/g = I + f, r = [f > 1/2] or, on an exact tie f = 1/2, r = (M + I) mod 2
//   (round half to even).  So inside a binade the recurrence is an INTEGER recurrence whose only dependence on the
//   running value is its parity on ties: each element is the pair (increment if M is even, increment if M is odd), and
//   pairs compose associatively.  Integer prefix "sums" of such pairs are exact in any evaluation order.
//
//   Binade crossings (a few dozen for normalised weights) are the only sequential part.  An approximate fp64 scan
//   predicts, per tile of 2048 elements, the binade it lives in; tiles that might contain a crossing are "irregular".
//   A single block then walks the tiles in order carrying the exact S: regular tiles cost one integer update (after an
//   exact check that the prediction holds), irregular tiles are added element by element, exactly.  Finally every
//   regular tile expands its exact prefix values from its exact start.  Correctness never depends on the prediction
//   (only speed does): every use of it is re-verified against the exact running value.
#pragma once
#include "common.cuh"

namespace mpl {

constexpr int kCxThreads = 256;
constexpr int kCxIpt = 8;
constexpr int kCxTile = kCxThreads * kCxIpt;   // 2048

struct CxTile {
    double approx_sum, approx_start;
    unsigned long long inc_even, inc_odd;   // tile function on the integer grid of binade e_pred (valid when regular)
    double s_start;                          // exact running sum before the tile (written by the walk)
    int e_pred;                              // predicted grid exponent (>= -1022)
    int regular;                             // 1: predicted to stay inside one binade; the walk may still demote it (-> 2)
};

struct CxPair {
    unsigned long long e, o;   // increment of M when M is even / odd before the element(s)
};
__device__ __forceinline__ CxPair cx_compose(const CxPair& a, const CxPair& b) {   // a first, then b
    // saturating at kCxHuge (2^54): every operand is <= 2^54, so the sums cannot wrap; a saturated value only says
    // "leaves the binade" and its parity is never used
    CxPair c;
    c.e = min(a.e + ((a.e & 1ull) ? b.o : b.e), 1ull << 54);
    c.o = min(a.o + (((a.o + 1ull) & 1ull) ? b.o : b.e), 1ull << 54);
    return c;
}
constexpr unsigned long long kCxHuge = 1ull << 54;   // increment that certainly leaves the binade

// grid exponent of a non-negative finite double (subnormals share the grid of the first normal binade)
__device__ __forceinline__ int cx_grid_exp(double s) {
    int f = (int)((unsigned long long)__double_as_longlong(s) >> 52) & 0x7ff;
    return (f == 0 ? 1 : f) - 1023;
}
__device__ __forceinline__ unsigned long long cx_mantissa(double s) {   // M with s = M * 2^(e - 52)
    unsigned long long b = (unsigned long long)__double_as_longlong(s);
    unsigned long long m = b & 0xfffffffffffffull;
    return ((b >> 52) & 0x7ff) ? (m | (1ull << 52)) : m;
}
__device__ __forceinline__ double cx_from_grid(unsigned long long M, int e) {   // M < 2^53, e >= -1022
    if (M >= (1ull << 52)) return __longlong_as_double((long long)(((unsigned long long)(e + 1023) << 52) | (M - (1ull << 52))));
    return __longlong_as_double((long long)M);   // subnormal (only possible when e == -1022)
}
// element p on the grid of exponent e: pair of integer increments
__device__ __forceinline__ CxPair cx_element(double p, int e) {
    if (!(p > 0.)) return CxPair{0ull, 0ull};   // zero (negative / NaN inputs are rejected before this path is chosen)
    const int ep = cx_grid_exp(p);
    const unsigned long long mp = cx_mantissa(p);
    const int sh = e - ep;
    if (sh < 0) return CxPair{kCxHuge, kCxHuge};          // p alone is at least 2^e: the sum leaves the binade
    if (sh == 0) return CxPair{mp, mp};
    if (sh > 54) return CxPair{0ull, 0ull};               // p < g/2 (strictly): rounds away
    const unsigned long long I = sh < 64 ? (mp >> sh) : 0ull;
    const unsigned long long rem = mp & ((1ull << sh) - 1ull);
    const unsigned long long half = 1ull << (sh - 1);
    if (rem > half) return CxPair{I + 1ull, I + 1ull};
    if (rem < half) return CxPair{I, I};
    return CxPair{I + (I & 1ull), I + 1ull - (I & 1ull)};  // exact tie: round half to even
}

// ---- pass 1: approximate tile sums (any order) + input validation ----------------------------------------------------------
static __global__ void __launch_bounds__(kCxThreads) cx_tilesum_kernel(const double* __restrict__ p, size_t n, CxTile* tiles, int* bad) {
    const size_t base = (size_t)blockIdx.x * kCxTile;
    double s = 0.;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < kCxIpt; ++i) {
        size_t idx = base + (size_t)i * kCxThreads + threadIdx.x;
        double v = idx < n ? p[idx] : 0.;
        ok = ok && (v >= 0.) && (v < INFINITY);
        s += v;
    }
    __shared__ double ws[kCxThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    if (!ok) *bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int i = 0; i < kCxThreads / 32; ++i) t += ws[i];
        tiles[blockIdx.x].approx_sum = t;
    }
}

// ---- pass 2: approximate exclusive prefix over tiles + classification (one block) -------------------------------------------
static __global__ void __launch_bounds__(1024) cx_classify_kernel(CxTile* tiles, unsigned int num_tiles, const int* bad) {
    __shared__ double part[1024];
    const unsigned int per = (num_tiles + 1023) / 1024;
    const unsigned int lo = threadIdx.x * per, hi = min(lo + per, num_tiles);
    double s = 0.;
    for (unsigned int t = lo; t < hi; ++t) s += tiles[t].approx_sum;
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double run = 0.; for (int i = 0; i < 1024; ++i) { double v = part[i]; part[i] = run; run += v; } }
    __syncthreads();
    double run = part[threadIdx.x];
    const double delta = 1e-7;   // generous: the approximate and the sequential sums differ by < 2e-9 relative at n = 2^24
    for (unsigned int t = lo; t < hi; ++t) {
        const double a = run, b = run + tiles[t].approx_sum;
        tiles[t].approx_start = a;
        int reg = 0, e = 0;
        if (a > 0.) {
            const double al = a * (1. - delta), bh = b * (1. + delta);
            e = cx_grid_exp(al);
            // whole (padded) range strictly inside the binade of exponent e, away from both ends
            const double lo_edge = ldexp(1., e), hi_edge = ldexp(1., e + 1);
            reg = (e > -1022) && (al > lo_edge) && (bh < hi_edge) && (cx_grid_exp(a) == e) && (*bad == 0);   // bad input: everything goes the exact element-by-element way
        }
        tiles[t].e_pred = e;
        tiles[t].regular = reg;
        run = b;
    }
}

// ---- pass 3: tile functions of the regular tiles ---------------------------------------------------------------------------
// thread i owns elements [i*8, i*8+8) of the tile (blocked order, needed for the ordered composition)
static __device__ __forceinline__ CxPair cx_block_inclusive(const CxPair (&el)[kCxIpt], CxPair (&incl)[kCxIpt], CxPair* warp_tot /* smem[8] */) {
    // inclusive composition along the tile; returns the tile's total function
    incl[0] = el[0];
#pragma unroll
    for (int i = 1; i < kCxIpt; ++i) incl[i] = cx_compose(incl[i - 1], el[i]);
    CxPair mine = incl[kCxIpt - 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CxPair sc = mine;   // inclusive over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        CxPair up{__shfl_up_sync(0xffffffffu, sc.e, o), __shfl_up_sync(0xffffffffu, sc.o, o)};
        if (lane >= o) sc = cx_compose(up, sc);
    }
    if (lane == 31) warp_tot[warp] = sc;
    CxPair excl_lane{__shfl_up_sync(0xffffffffu, sc.e, 1), __shfl_up_sync(0xffffffffu, sc.o, 1)};
    if (lane == 0) excl_lane = CxPair{0ull, 0ull};
    __syncthreads();
    CxPair pre{0ull, 0ull}, total{0ull, 0ull};
    for (int w = 0; w < kCxThreads / 32; ++w) { if (w < warp) pre = cx_compose(pre, warp_tot[w]); total = cx_compose(total, warp_tot[w]); }
    pre = cx_compose(pre, excl_lane);
#pragma unroll
    for (int i = 0; i < kCxIpt; ++i) incl[i] = cx_compose(pre, incl[i]);
    return total;
}

static __global__ void __launch_bounds__(kCxThreads) cx_pairs_kernel(const double* __restrict__ p, size_t n, CxTile* tiles) {
    CxTile& T = tiles[blockIdx.x];
    if (!T.regular) return;
    __shared__ CxPair warp_tot[kCxThreads / 32];
    const int e = T.e_pred;
    const size_t base = (size_t)blockIdx.x * kCxTile + (size_t)threadIdx.x * kCxIpt;
    CxPair el[kCxIpt], incl[kCxIpt];
#pragma unroll
    for (int i = 0; i < kCxIpt; ++i) el[i] = cx_element(base + i < n ? p[base + i] : 0., e);
    CxPair total = cx_block_inclusive(el, incl, warp_tot);
    if (threadIdx.x == 0) { T.inc_even = total.e; T.inc_odd = total.o; }
}

// ---- pass 4: the sequential walk over tiles (one block) ----------------------------------------------------------------------
constexpr int kCxMaxRounds = 4;      // binade crossings handled by block scans inside one irregular tile before the element-by-element path takes over
constexpr int kCxWalkChunk = kCxThreads;   // tile records staged in shared memory per round: one per thread
static __global__ void __launch_bounds__(kCxThreads) cx_walk_kernel(const double* __restrict__ p, size_t n, CxTile* tiles, unsigned int num_tiles, double* __restrict__ out, const int* bad) {
    __shared__ double buf[kCxTile];
    __shared__ unsigned long long c_even[kCxWalkChunk], c_odd[kCxWalkChunk];
    __shared__ double c_start[kCxWalkChunk];
    __shared__ int c_epred[kCxWalkChunk];
    __shared__ signed char c_reg[kCxWalkChunk];
    __shared__ double s_run, s_start_tile;
    __shared__ unsigned int run_len, n_fast, cross_at, nz_total, nz_warp[kCxThreads / 32];
    __shared__ double c_sum[kCxWalkChunk], s_prev, s_cross;
    __shared__ CxPair run_warp[kCxThreads / 32];
    __shared__ unsigned short nz_rank[kCxTile];
    const int tid = threadIdx.x;
    const bool clean = *bad == 0;   // negative / NaN / infinite inputs: no integer recurrence, every tile is added element by element
    if (tid == 0) s_run = 0.;
    for (unsigned int chunk0 = 0; chunk0 < num_tiles; chunk0 += kCxWalkChunk) {
        const unsigned int cnt = min((unsigned int)kCxWalkChunk, num_tiles - chunk0);
        __syncthreads();
        for (unsigned int i = tid; i < cnt; i += kCxThreads) {
            const CxTile& T = tiles[chunk0 + i];
            c_even[i] = T.inc_even; c_odd[i] = T.inc_odd; c_epred[i] = T.e_pred; c_reg[i] = (signed char)T.regular; c_sum[i] = T.approx_sum;
        }
        __syncthreads();
        unsigned int k = 0;
        while (k < cnt) {
            // The run of regular tiles from k on that are predicted to live in the binade of tile k, thread i <-> tile k + i: their
            // tile functions compose associatively, so one block scan gives every tile's exact start -- as long as the running sum
            // really is in that binade at k and has not left it by the end of a tile (increments are non-negative: the tiles that
            // pass are a prefix of the run).  The first tile that fails is walked element by element, exactly as a sequential walk
            // over the tiles would decide.
            if (tid == 0) { run_len = kCxThreads; n_fast = kCxThreads; }
            __syncthreads();
            const double S = s_run;
            const int e_run = c_epred[k];
            const unsigned int j = k + tid;
            const bool in_run = j < cnt && c_reg[j] == 1 && c_epred[j] == e_run && c_even[j] < kCxHuge && c_odd[j] < kCxHuge;
            if (!in_run) atomicMin(&run_len, (unsigned int)tid);
            __syncthreads();
            const bool mine = (unsigned int)tid < run_len;
            CxPair rin = mine ? CxPair{c_even[j], c_odd[j]} : CxPair{0ull, 0ull}, rex;
            {
                const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    CxPair up{__shfl_up_sync(0xffffffffu, rin.e, o), __shfl_up_sync(0xffffffffu, rin.o, o)};
                    if (lane >= o) rin = cx_compose(up, rin);
                }
                if (lane == 31) run_warp[warp] = rin;
                rex = CxPair{__shfl_up_sync(0xffffffffu, rin.e, 1), __shfl_up_sync(0xffffffffu, rin.o, 1)};
                if (lane == 0) rex = CxPair{0ull, 0ull};
                __syncthreads();
                CxPair pre{0ull, 0ull};
                for (int w = 0; w < warp; ++w) pre = cx_compose(pre, run_warp[w]);
                rin = cx_compose(pre, rin);
                rex = cx_compose(pre, rex);
            }
            const bool at_home = S > 0. && cx_grid_exp(S) == e_run;
            const unsigned long long M = cx_mantissa(S);
            const unsigned long long m_start = M + ((M & 1ull) ? rex.o : rex.e), m_end = M + ((M & 1ull) ? rin.o : rin.e);
            if (mine && !(at_home && m_end < (1ull << 53))) atomicMin(&n_fast, (unsigned int)tid);
            __syncthreads();
            const unsigned int fast = min(n_fast, run_len);
            if ((unsigned int)tid < fast) {
                c_start[j] = cx_from_grid(m_start, e_run);
                if ((unsigned int)tid == fast - 1) s_run = cx_from_grid(m_end, e_run);
            }
            __syncthreads();
            k += fast;
            if (fast != 0u) continue;   // (the tile after the run: the next round decides -- a new run, or the exact path right below)
            if (tid == 0) { c_reg[k] = 2; c_start[k] = S; }
            // Irregular tile.  thread t owns elements [8 t, 8 t + 8).
            const size_t base = (size_t)(chunk0 + k) * kCxTile;
            double v[kCxIpt];
#pragma unroll
            for (int i = 0; i < kCxIpt; ++i) { const size_t idx = base + (size_t)tid * kCxIpt + i; v[i] = idx < n ? p[idx] : 0.; }
            // (a) A tile that holds one or two binade crossings (the usual case for resampling weights: its predicted sum is
            // comparable to the running sum): between two crossings the integer recurrence holds, so the block scans the pairs
            // from `pos` on, finds the first addition that leaves the binade, writes everything before it, performs that one
            // addition in fp64 -- exactly what the sequential loop does there -- and goes on from the element after it.
            unsigned int pos = 0;
            if (clean && S > 0. && c_sum[k] < 3. * S) {
                double Sr = S;
                for (int round = 0; round < kCxMaxRounds && pos < (unsigned int)kCxTile; ++round) {
                    const int e = cx_grid_exp(Sr);
                    const unsigned long long M = cx_mantissa(Sr);
                    CxPair el[kCxIpt], incl[kCxIpt];
#pragma unroll
                    for (int i = 0; i < kCxIpt; ++i) el[i] = (unsigned int)(tid * kCxIpt + i) >= pos ? cx_element(v[i], e) : CxPair{0ull, 0ull};
                    if (tid == 0) cross_at = kCxTile;
                    cx_block_inclusive(el, incl, run_warp);   // (one block barrier inside: orders the store above, too)
                    unsigned long long m[kCxIpt];
#pragma unroll
                    for (int i = 0; i < kCxIpt; ++i) m[i] = M + ((M & 1ull) ? incl[i].o : incl[i].e);
#pragma unroll
                    for (int i = kCxIpt - 1; i >= 0; --i) if ((unsigned int)(tid * kCxIpt + i) >= pos && m[i] >= (1ull << 53)) atomicMin(&cross_at, (unsigned int)(tid * kCxIpt + i));
                    __syncthreads();
                    const unsigned int cross = cross_at;
#pragma unroll
                    for (int i = 0; i < kCxIpt; ++i) {
                        const unsigned int q = tid * kCxIpt + i;
                        if (q >= pos && q < cross) {
                            const double val = cx_from_grid(m[i], e);
                            if (base + q < n) out[base + q] = val;
                            if (q + 1 == cross || q + 1 == (unsigned int)kCxTile) s_prev = val;   // the running sum before the crossing / at the end of the tile
                        }
                        if (q == cross) s_cross = v[i];
                    }
                    __syncthreads();
                    if (cross >= (unsigned int)kCxTile) { pos = kCxTile; Sr = s_prev; break; }
                    Sr = __dadd_rn(cross > pos ? s_prev : Sr, s_cross);
                    if (tid == 0 && base + cross < n) out[base + cross] = Sr;
                    pos = cross + 1;
                    __syncthreads();
                }
                if (tid == 0) s_run = Sr;
                __syncthreads();
            }
            // (b) Whatever is left (tiles with many crossings -- importance weights spread over hundreds of binades -- and the first
            // tile of all): exact additions in index order by one thread, staged through shared memory.  Adding +0 never changes
            // the running sum, and importance weights are mostly exact zeros (exp underflow): the zeros are compacted away by all
            // threads first (order kept), the one thread adds the non-zero elements only, and every position then reads the sum
            // after the last non-zero element at or before it.
            if (pos < (unsigned int)kCxTile) {
                {
                    unsigned int c = 0;   // inclusive count of the non-zero elements from `pos` on (warp scan + warp totals)
#pragma unroll
                    for (int i = 0; i < kCxIpt; ++i) { if ((unsigned int)(tid * kCxIpt + i) < pos) v[i] = 0.; c += v[i] != 0. ? 1u : 0u; }
                    unsigned int incl = c;
                    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
                    if (lane == 31) nz_warp[warp] = incl;
                    __syncthreads();
                    unsigned int before = incl - c;
                    for (int w = 0; w < warp; ++w) before += nz_warp[w];
#pragma unroll
                    for (int i = 0; i < kCxIpt; ++i) {
                        if (v[i] != 0.) buf[before++] = v[i];                       // compacted, in index order
                        nz_rank[tid * kCxIpt + i] = (unsigned short)before;       // non-zero elements at or before this position
                    }
                    if (tid == kCxThreads - 1) nz_total = before;
                }
                __syncthreads();
                if (tid == 0) {
                    double Sq = s_run;
                    const unsigned int cnt_nz = nz_total;
                    for (unsigned int i = 0; i < cnt_nz; i += 8) {   // (loads hoisted out of the dependent chain of additions)
                        double r[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) r[j] = i + j < cnt_nz ? buf[i + j] : 0.;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { Sq = __dadd_rn(Sq, r[j]); r[j] = Sq; }
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (i + j < cnt_nz) buf[i + j] = r[j];
                    }
                    s_start_tile = s_run;
                    s_run = Sq;
                }
                __syncthreads();
                for (unsigned int i = tid; i < (unsigned int)kCxTile; i += kCxThreads)
                    if (i >= pos && base + i < n) { const unsigned int rk = nz_rank[i]; out[base + i] = rk ? buf[rk - 1] : s_start_tile; }
            }
            k += 1;
            __syncthreads();
            continue;
        }
        __syncthreads();
        for (unsigned int i = tid; i < cnt; i += kCxThreads) { tiles[chunk0 + i].s_start = c_start[i]; tiles[chunk0 + i].regular = c_reg[i]; }
    }
}

// ---- pass 5: exact prefix values inside the regular tiles -------------------------------------------------------------------
static __global__ void __launch_bounds__(kCxThreads) cx_fill_kernel(const double* __restrict__ p, size_t n, const CxTile* tiles, double* __restrict__ out) {
    const CxTile& T = tiles[blockIdx.x];
    if (T.regular != 1) return;   // irregular tiles were written by the walk
    __shared__ CxPair warp_tot[kCxThreads / 32];
    const int e = T.e_pred;
    const size_t base = (size_t)blockIdx.x * kCxTile + (size_t)threadIdx.x * kCxIpt;
    CxPair el[kCxIpt], incl[kCxIpt];
#pragma unroll
    for (int i = 0; i < kCxIpt; ++i) el[i] = cx_element(base + i < n ? p[base + i] : 0., e);
    cx_block_inclusive(el, incl, warp_tot);
    const unsigned long long M = cx_mantissa(T.s_start);
    const bool odd = M & 1ull;
#pragma unroll
    for (int i = 0; i < kCxIpt; ++i)
        if (base + i < n) out[base + i] = cx_from_grid(M + (odd ? incl[i].o : incl[i].e), e);
}

}  // namespace mpl
