// pf_kernels.cuh -- particle-filter kernels for sm_100a.
//
//   pf_extend_kernel     K1/K2 (+K6 fused): ancestor gather -> Philox extend -> observation log-likelihood ->
//                        SoA stores, with an online (max, sum exp, sum exp^2) epilogue (K3) and a last-block
//                        finalisation.  Replaces the per-particle generate/update loops of
//                        reference modppl/src/inference/particle_filter.rs:65-69,76-82.
//   weight_reduce_kernel K3 stand-alone: lib.rs:34-45 + particle_filter.rs:27-35,98-100 in one pass.
//   fixed_reduce_kernel / fixed_scan_kernel / fixed_overflow_kernel
//                        K5: integer-weight systematic resampling -- quantise, two-level exact integer prefix scan,
//                        in-tile expansion (run heads + max-scan in shared memory) that writes ancestors coalesced.
//   normalize / cumsum_seq / search kernels
//                        K4: the reference's multinomial routine (categorical.rs:22-32) with its SEQUENTIAL f64
//                        running sum, bit-exact.
//   gather_kernel        K6 stand-alone (only when the host reads state right after a resample).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "nested_quant.cuh"

namespace mpl {

typedef unsigned __int128 u128;

// device-resident bookkeeping of one ParticleSystem (reference particle_filter.rs:8-24 scalars)
struct DeviceStats {
    double max, sumexp, sumexp2;   // of the current log-weights
    double lse;                    // log total weight at the last normalisation (return value of resample())
    double lml_acc;                // log_ml_estimate                                  (particle_filter.rs:23)
    double ess;                    // fresh ESS of the current weights
    double ess_stale;              // ESS as of the last normalize_weights()            (quirk Q1)
    unsigned long long W;          // total integer weight (fixed schemes)
    unsigned long long rand_word;  // systematic offset word of this resample
    long long t;                   // next kernel time index (device copy, for mpl_ps_run)
    unsigned int blocks_done;      // last-block-done counter of the extend epilogue
    unsigned int ticket;           // last-block-done counter of the expansion's "ancestors written" signal
    unsigned int overflow_count;
    int degenerate;                // all weights -inf seen
    int resampled;                 // 1: ancestors pending (next extend gathers, weights are zero)
    int do_resample;               // ESS trigger decision for the dynamic path
    // sharded runs: block 0 of a kernel waits for the peers (system scope), publishes the global values into this
    // struct and then raises the matching ready word; the kernel's other blocks only watch that local word
    unsigned long long c_offset;   // integer weight of all lower-ranked shards
    int resampled_flag[2];         // [t & 1]: the extend of step t must gather through the ancestors (set by the scan of step t-1)
    unsigned long long n_resamples;   // resamples performed so far
    unsigned long long max_bits[2];   // exact max of the log-weights written by the extend of step t, slot t & 1 (order-preserving bits)
    long long ready_stats, ready_w, ready_done;
    long long trace[16];           // %globaltimer stamps of the last step's phases (sharded runs; mpl_ps_trace)
    int nest_E;                    // nested scheme: the global power-of-two reference of this resample
    unsigned int pad2;
    unsigned long long nvlink_bytes;   // sharded runs: payload bytes this GPU has requested from its peers' memory so far (diagnostics: mpl_ps_nvlink_bytes)
};
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---- multi-GPU: one process per GPU, peers reached through NVLink-mapped pointers (CUDA IPC) ----------------------------
// Every rank owns a Mailbox; rank g writes column g of every peer's mailbox (remote stores), then a flag carrying the
// step number, and kernels on the receiving GPU spin on their LOCAL copy of the flags.  No host round trip, no NCCL call
// on the data path: (max, sum exp, sum exp^2) after the extend, the integer weight total after the reduce pass, and a
// "my ancestors are written" flag after the scan.
constexpr int kMaxPeers = 8;
constexpr int kMaxSections = 2048;   // nested scheme: sections of 2^17 particles, up to 2^28 particles in all
struct Mailbox {
    // self-validating words (the trick of NCCL's LL protocol): every 8-byte word carries 32 payload bits and the 32-bit
    // step number, and an aligned 8-byte store is performed atomically -- so neither side needs a memory fence (a
    // fence.sys costs ~4 us on B200 and there would be three on the critical path of every step)
    unsigned long long stats_ll[kMaxPeers][6];   // (max, sum exp, sum exp^2) of shard h as six tagged halves
    unsigned long long w_ll[kMaxPeers][4];       // integer weight total of shard h, and its sum of squared weights
    long long flag_done[kMaxPeers];              // shard h has written every ancestor it owes for this step
    int error;
    int pad;
    unsigned long long nvlink_polled;   // bytes of section records fetched from peers (one successful poll each; diagnostics)
    unsigned long long barrier_seq;     // mpl_ps_peer_barrier: this rank has reached barrier number ... (polled by the peers)
    // nested scheme: (E_s, T_s, sum q^2) of every section of THIS shard, indexed by [step & 1][GLOBAL section number]; published
    // here by the owner and polled by the peers.  Two parities: a rank may already publish step t + 1 while a slow peer is still
    // collecting step t (it cannot get to t + 2 before that peer has published t + 1, i.e. is done with t)
    unsigned long long sec_ll[2][kMaxSections][6];
};
__device__ __forceinline__ void ll_write64(unsigned long long* dst2, unsigned long long value, unsigned int epoch) {
    *(volatile unsigned long long*)(dst2 + 0) = (value & 0xffffffffull) | ((unsigned long long)epoch << 32);
    *(volatile unsigned long long*)(dst2 + 1) = (value >> 32) | ((unsigned long long)epoch << 32);
}
struct PeerTable {
    int world, rank;
    unsigned int n_loc;   // particles per shard (equal shards)
    int shift;            // log2(n_loc) when it is a power of two, else -1
    const void* state[2][kMaxPeers];   // [buffer][rank]: SoA state of every shard
    int32_t* anc[kMaxPeers];           // ancestor slots of every shard (single-level scheme: ancestors are pushed)
    Mailbox* mail[kMaxPeers];
    // nested scheme: everything is PULLED.  [parity][rank]: the integer weights (in the log-weight array) and the chunk records
    // the extend of a step left; a sharded run alternates two buffers so that a rank which is already extending step t + 1 does
    // not overwrite what a slower peer still reads for the resampling of step t
    const void* lw[2][kMaxPeers];
    const int* rec_e[2][kMaxPeers];
    const unsigned int* rec_S[2][kMaxPeers];
    const unsigned long long* tile_pre[2][kMaxPeers];   // per tile of 32 chunks: exclusive prefix of the chunk masses inside its section (section pass)
};
__device__ __forceinline__ unsigned int peer_owner(const PeerTable& p, unsigned int gid) { return p.shift >= 0 ? gid >> p.shift : gid / p.n_loc; }

// one thread: bounded spins (~20 s, then Mailbox::error is raised and every later wait returns at once)
struct SpinGuard {
    volatile int* err;
    long long t0;
    __device__ __forceinline__ SpinGuard(const PeerTable& p) {
        err = &p.mail[p.rank]->error;
#ifdef __CUDA_ARCH__
        t0 = clock64();
#endif
    }
    __device__ __forceinline__ bool give_up() {
#ifdef __CUDA_ARCH__
        if (*err) return true;
        if (clock64() - t0 > 40000000000ll) { *err = 1; return true; }
        __nanosleep(100);
#endif
        return false;
    }
};
__device__ __forceinline__ unsigned long long ll_read64(const unsigned long long* src2, unsigned int epoch, SpinGuard& g) {
    const volatile unsigned long long* p = src2;
    unsigned long long lo, hi;
    while ((unsigned int)((lo = p[0]) >> 32) != epoch) if (g.give_up()) return 0ull;
    while ((unsigned int)((hi = p[1]) >> 32) != epoch) if (g.give_up()) return 0ull;
    return (lo & 0xffffffffull) | (hi << 32);
}
__device__ __forceinline__ void peer_wait_done(const PeerTable& p, long long epoch) {
    SpinGuard g(p);
    const volatile long long* f = p.mail[p.rank]->flag_done;
    for (int h = 0; h < p.world; ++h) while (f[h] < epoch) if (g.give_up()) return;
    __threadfence();
}
__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ordered_bits(double x) {
    long long b = __double_as_longlong(x);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long u) {
    if (u == 0ull) return -INFINITY;   // slot never touched: no finite weight
    return __longlong_as_double((long long)((u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u));
}

template <int V> struct AncVecOf;
template <> struct AncVecOf<4> { typedef int4 type; };
template <> struct AncVecOf<2> { typedef int2 type; };
template <typename Real> struct VecOf;
template <> struct VecOf<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct VecOf<double> { typedef double2 type; static constexpr int N = 2; };

template <typename Real> __device__ __forceinline__ void vec_load(const Real* p, Real (&v)[VecOf<Real>::N]);
template <> __device__ __forceinline__ void vec_load<float>(const float* p, float (&v)[4]) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
template <> __device__ __forceinline__ void vec_load<double>(const double* p, double (&v)[2]) { double2 t = *reinterpret_cast<const double2*>(p); v[0] = t.x; v[1] = t.y; }
template <typename Real> __device__ __forceinline__ void vec_store(Real* p, const Real (&v)[VecOf<Real>::N]);
template <> __device__ __forceinline__ void vec_store<float>(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void vec_store<double>(double* p, const double (&v)[2]) { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
// streaming variants (ld/st.global.cs: evict-first).  The state is touched once per step and must not displace the arrays
// that one kernel of the step hands to the next through the 126 MB L2: log-weights -> integer weights -> ancestors.
template <typename Real> __device__ __forceinline__ void vec_store_stream(Real* p, const Real (&v)[VecOf<Real>::N]);
template <> __device__ __forceinline__ void vec_store_stream<float>(float* p, const float (&v)[4]) { __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3])); }
template <> __device__ __forceinline__ void vec_store_stream<double>(double* p, const double (&v)[2]) { __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1])); }

// ---- particle-major state: the D components of particle i are contiguous (state[i * D + d]).  A parent is fetched with ONE
// vector load (16 bytes for D = 4 in fp32) instead of D scalar loads from D arrays -- the ancestor-indirect gather is the
// latency-critical access of the extend kernel -- and a thread's V consecutive particles are V * D contiguous elements.
template <typename Real, int D>
__device__ __forceinline__ void load_particle_ro(const Real* __restrict__ state, size_t i, Real (&x)[D]) {   // read-only path (LDG.CONSTANT)
    constexpr int BYTES = D * (int)sizeof(Real);
    const Real* p = state + i * D;
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int k = 0; k < BYTES / 16; ++k) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + k);
            memcpy(reinterpret_cast<char*>(x) + 16 * k, &v, 16);
        }
    } else if constexpr (BYTES == 8) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        memcpy(x, &v, 8);
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __ldg(p + d);
    }
}
template <typename Real, int D>
__device__ __forceinline__ void load_particle(const Real* state, size_t i, Real (&x)[D]) {   // plain loads (a peer's memory over NVLink)
    constexpr int BYTES = D * (int)sizeof(Real);
    const Real* p = state + i * D;
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int k = 0; k < BYTES / 16; ++k) {
            const uint4 v = *(reinterpret_cast<const uint4*>(p) + k);
            memcpy(reinterpret_cast<char*>(x) + 16 * k, &v, 16);
        }
    } else if constexpr (BYTES == 8) {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        memcpy(x, &v, 8);
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = p[d];
    }
}
// a thread's V consecutive particles (V * D contiguous elements, a multiple of 16 bytes): 128-bit accesses
template <typename Real, int V, int D>
__device__ __forceinline__ void load_particles(const Real* __restrict__ state, size_t i0, Real (&x)[V][D]) {
    static_assert((V * D * sizeof(Real)) % 16 == 0, "V consecutive particles are whole 16-byte words");
    const uint4* p = reinterpret_cast<const uint4*>(state + i0 * D);
#pragma unroll
    for (int k = 0; k < (int)(V * D * sizeof(Real)) / 16; ++k) {
        const uint4 v = p[k];
        memcpy(reinterpret_cast<char*>(x) + 16 * k, &v, 16);
    }
}
template <typename Real, int V, int D>
__device__ __forceinline__ void store_particles_stream(Real* __restrict__ state, size_t i0, const Real (&x)[V][D]) {   // st.global.cs: touched once per step
    static_assert((V * D * sizeof(Real)) % 16 == 0, "V consecutive particles are whole 16-byte words");
    uint4* p = reinterpret_cast<uint4*>(state + i0 * D);
#pragma unroll
    for (int k = 0; k < (int)(V * D * sizeof(Real)) / 16; ++k) {
        uint4 v;
        memcpy(&v, reinterpret_cast<const char*>(x) + 16 * k, 16);
        __stcs(p + k, v);
    }
}

// ================================================================================================
// K1/K2/K6/K3: extend
// ================================================================================================
enum ExtendMode : int { EXT_INIT = 0, EXT_ACCUM = 1, EXT_GATHER = 2, EXT_DYNAMIC = 3 };

template <typename Real>
struct ExtendArgs {
    const Real* state_in;    // ld x D, particle-major
    Real* state_out;         // ld x D
    const Real* lw_in;       // ld: log-weights to accumulate onto (EXT_ACCUM)
    Real* lw;                // ld: log-weights written (== lw_in on one GPU; the other buffer of the pair when sharded)
    const int32_t* anc;      // ld (parent: global particle id)
    size_t n, ld;
    uint64_t seed, gid_offset;
    long long t;             // kernel time index; < 0: read stats->t (device-resident loop)
    Obs obs;
    const double* obs_dev;   // if non-null: observations of step t at obs_dev[t * nobs ...]
    int nobs;
    DeviceStats* stats;
    Lse3<double>* partials;  // gridDim.x
    PeerTable peer;          // world == 1: single GPU
    int cur;                 // which state buffer is the input (index into peer.state)
    int wait_done;           // sharded: the ancestors were PUSHED by the peers (single-level scheme): wait for their "done" flags
    ChunkRecords rec;        // NESTED: chunk records written by the fused quantisation epilogue
    int kbits;
};

constexpr int kExtendThreads = 256;

// NESTED (fp32 only): every warp iteration covers one aligned 128-particle chunk, whose log-weights are quantised on the
// spot against the chunk's own maximum (nested.cuh) -- the reduce pass over the particles disappears.  NESTED == 1: the
// integer weights replace the log-weights in HBM (a resample follows for sure); NESTED == 2: only the chunk records are
// kept and the log-weights stay (ESS-triggered loop: the resample may be skipped, the expansion re-quantises if not).
template <class Model, typename Real, int MODE, bool SHARDED = false, int NESTED = 0>
__global__ void __launch_bounds__(kExtendThreads, 4) pf_extend_kernel(ExtendArgs<Real> a, Model model) {
    constexpr int D = Model::D;
    constexpr int V = VecOf<Real>::N;
    const int tid = threadIdx.x;
    pdl_wait();

    long long t = a.t;
    if (t < 0) t = a.stats->t;
    Obs obs = a.obs;
    if (a.obs_dev != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) obs.v[k] = (k < a.nobs) ? a.obs_dev[(size_t)t * a.nobs + k] : 0.;
    }
    bool gather = (MODE == EXT_GATHER);
    bool accum = (MODE == EXT_ACCUM);
    if (MODE == EXT_DYNAMIC) {
        gather = a.stats->resampled_flag[t & 1] != 0;
        accum = !gather;
    }
    constexpr bool sharded = SHARDED;
    if (sharded && gather && a.wait_done) gate_done(a.peer, a.stats, t);   // ancestors written everywhere; old buffer no longer read
    // (nested scheme: nothing is pushed -- this rank computed its own ancestors after the step's only gate, the section records)

    Real run_max = (Real)-INFINITY;
    if (blockIdx.x == 0 && tid == 0) { a.stats->max_bits[(t + 1) & 1] = 0ull; a.stats->trace[13] = global_ns(); }   // slot of the next step (its last reader finished before this launch)
    const size_t stride = (size_t)gridDim.x * kExtendThreads * V;
    typedef typename AncVecOf<V>::type AncVec;
    size_t base = ((size_t)blockIdx.x * kExtendThreads + tid) * V;
    AncVec anc_next = AncVec();
    if (gather && base < a.ld) anc_next = __ldcs(reinterpret_cast<const AncVec*>(a.anc + base));
    for (; base - (size_t)(tid & 31) * V < a.n; base += stride) {   // warp-uniform trip count (arrays are padded to ld)
        Real x[V][D];
        Real w[V];
        int32_t par[V];
        const bool full = base + V <= a.n;
        if (gather) {
            // software pipeline: this iteration's ancestors were loaded one iteration ago; fetch the next ones now and
            // pull the parents' cache lines towards L2 so the dependent gather of the next iteration starts warm
            par[0] = anc_next.x; par[1] = anc_next.y;
            if constexpr (V == 4) { par[2] = anc_next.z; par[3] = anc_next.w; }
            if (base + stride < a.ld) anc_next = __ldcs(reinterpret_cast<const AncVec*>(a.anc + base + stride));   // last use of these ancestors
            bool local = true;   // all parents in this shard (always, on one GPU; nearly always when sharded)
            if (sharded) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    par[v] = (full || base + v < a.n) ? par[v] : (int32_t)a.gid_offset;
                    local = local && ((unsigned int)par[v] - (unsigned int)a.gid_offset < a.peer.n_loc);
                }
            }
            if (local) {
                const unsigned int off = sharded ? (unsigned int)a.gid_offset : 0u;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const size_t src = (sharded || full || base + v < a.n) ? (size_t)((unsigned int)par[v] - off) : 0;
                    load_particle_ro<Real, D>(a.state_in, src, x[v]);
                }
            } else {   // parents are global ids: read them where they live (a peer's HBM over NVLink)
                unsigned int n_remote = 0;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const unsigned int g = (unsigned int)par[v];
                    const unsigned int r = peer_owner(a.peer, g);
                    n_remote += r != (unsigned int)a.peer.rank ? 1u : 0u;
                    load_particle<Real, D>(reinterpret_cast<const Real*>(a.peer.state[a.cur][r]), (size_t)(g - r * a.peer.n_loc), x[v]);
                }
                atomicAdd(&a.stats->nvlink_bytes, (unsigned long long)(n_remote * D * sizeof(Real)));   // (rare branch: shard edges only)
            }
        } else if (MODE != EXT_INIT) {
            load_particles<Real, V, D>(a.state_in, base, x);
        }
        if (accum) vec_load<Real>(a.lw_in + base, w);
        else {
#pragma unroll
            for (int v = 0; v < V; ++v) w[v] = 0;
        }
        if constexpr (Model::kGroupDraws) {   // one Philox block for the thread's V particles (shards start at multiples of 4)
            Real zs[V];
            Stream sg(a.seed, (a.gid_offset + base) / V, (uint32_t)t, P_MODEL_GROUP);
            draw_normals<V>(sg, 0, zs);
#pragma unroll
            for (int v = 0; v < V; ++v) w[v] += model.kernel_z(t, zs[v], x[v], obs);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                Stream s(a.seed, a.gid_offset + base + v, (uint32_t)t, P_MODEL);
                w[v] += model.kernel(t, s, x[v], obs);
            }
        }
        store_particles_stream<Real, V, D>(a.state_out, base, x);
        if constexpr (NESTED) {
            float wm[4], sqc;
            unsigned int qv[4];
#pragma unroll
            for (int v = 0; v < V; ++v) wm[v] = (full || base + v < a.n) ? (float)w[v] : -INFINITY;
            int e_c;
            unsigned int S_c;
            warp_quantise_chunk(wm, qv, e_c, S_c, sqc);
            if constexpr (NESTED == 2) vec_store<Real>(a.lw + base, w);
            else {
                Real qr[V];   // fp32: the integer's bit pattern takes the log-weight's slot; fp64: its value
#pragma unroll
                for (int v = 0; v < V; ++v) qr[v] = nested_store<Real>(qv[v]);
                vec_store<Real>(a.lw + base, qr);
            }
            if ((tid & 31) == 0) { const size_t chunk = base / kChunk; a.rec.e[chunk] = e_c; a.rec.S[chunk] = S_c; a.rec.sq[chunk] = sqc; }
        } else {
            vec_store<Real>(a.lw + base, w);
        }

        if (gather && base + stride < a.n && (unsigned int)anc_next.x - (unsigned int)a.gid_offset < (unsigned int)a.n) {
            const size_t nsrc = (size_t)((unsigned int)anc_next.x - (unsigned int)a.gid_offset);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.state_in + nsrc * D));
        }

        // exact running max of the log-weights (NaN and padding lanes count as -inf, quirk Q9).  Sum statistics are not
        // needed on this path: the integer resampler derives the log total weight and the ESS from its own sums, and
        // ESS / log-ML queries run weight_reduce_kernel on demand.
        if constexpr (!NESTED) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                bool ok = (full || base + v < a.n) && (w[v] == w[v]);
                run_max = fmax(run_max, ok ? w[v] : (Real)-INFINITY);
            }
        }
    }

    pdl_trigger();
    if (blockIdx.x == gridDim.x - 1 && tid == 0) a.stats->trace[14] = global_ns();
    if constexpr (NESTED) return;   // the nested resampler needs no global maximum (and, sharded, no exchange here)
    __shared__ Real warp_max_s[kExtendThreads / 32];
    __shared__ bool is_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) run_max = fmax(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
    if ((tid & 31) == 0) warp_max_s[tid >> 5] = run_max;
    __syncthreads();
    if (tid == 0) {
        Real bm = warp_max_s[0];
#pragma unroll
        for (int i = 1; i < kExtendThreads / 32; ++i) bm = fmax(bm, warp_max_s[i]);
        if (bm > (Real)-INFINITY) atomicMax(&a.stats->max_bits[t & 1], ordered_bits((double)bm));
        is_last = false;
        if (a.peer.world > 1) {   // sharded: the last block posts this shard's max to every rank
            __threadfence();
            is_last = (atomicAdd(&a.stats->blocks_done, 1u) == gridDim.x - 1);
            if (is_last) {
                a.stats->blocks_done = 0;
                a.stats->trace[2] = global_ns();
                __threadfence();
                const double m = from_ordered_bits(*(volatile unsigned long long*)&a.stats->max_bits[t & 1]);
                for (int h = 0; h < a.peer.world; ++h) {
                    unsigned long long* dst = a.peer.mail[h]->stats_ll[a.peer.rank];
                    ll_write64(dst + 0, (unsigned long long)__double_as_longlong(m), (unsigned int)(t + 1));
                }
            }
        }
    }
    (void)is_last;
}

// ================================================================================================
// K3 stand-alone weight reduction (also the parity hook mpl_logsumexp_stats)
// ================================================================================================
// Rounds of kReduceIpt elements per thread, held in registers: the warp's maximum first (shuffles only), then ONE exp per
// element against it and two shuffle sums -- the running (max, sum, sum of squares) is rescaled once per round and warp, not
// once per element and shuffle step (an fp64 exp is ~100 instructions; with one online update per element and a combine per
// shuffle step the pass spent three of them per weight).
constexpr int kReduceIpt = 8;
template <typename Real>
__global__ void __launch_bounds__(256) weight_reduce_kernel(const Real* __restrict__ lw, size_t n, DeviceStats* stats, Lse3<double>* partials) {
    typedef Real Acc;
    const int tid = threadIdx.x;
    Lse3<Acc> run = lse3_identity<Acc>();   // of this warp (every lane holds the same value)
    for (size_t base = (size_t)blockIdx.x * (256 * kReduceIpt); base < n; base += (size_t)gridDim.x * (256 * kReduceIpt)) {
        Real w[kReduceIpt];
        Acc m = (Acc)-INFINITY;
#pragma unroll
        for (int k = 0; k < kReduceIpt; ++k) {
            const size_t i = base + (size_t)k * 256 + tid;
            w[k] = i < n ? lw[i] : (Real)-INFINITY;
            if (!(w[k] == w[k])) w[k] = (Real)-INFINITY;   // NaN counts as -inf (quirk Q9)
            m = fmax(m, (Acc)w[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (m == (Acc)-INFINITY) continue;   // (warp-uniform) nothing finite in this round
        Acc s = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < kReduceIpt; ++k) { const Acc e = stat_exp((Acc)w[k] - m); s += e; s2 += e * e; }   // exp(-inf) = 0
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        run = lse3_combine(run, Lse3<Acc>{m, s, s2});
    }
    __shared__ Lse3<double> warp_part[8];
    __shared__ bool is_last;
    if ((tid & 31) == 0) warp_part[tid >> 5] = Lse3<double>{(double)run.m, (double)run.s, (double)run.s2};
    __syncthreads();
    if (tid == 0) {
        Lse3<double> b = warp_part[0];
        for (int i = 1; i < 8; ++i) b = lse3_combine(b, warp_part[i]);
        partials[blockIdx.x] = b;
        __threadfence();
        is_last = (atomicAdd(&stats->blocks_done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        Lse3<double> acc = lse3_identity<double>();
        for (unsigned int i = tid; i < gridDim.x; i += 256) acc = lse3_combine(acc, partials[i]);
        acc = lse3_warp_reduce(acc);
        if ((tid & 31) == 0) warp_part[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            Lse3<double> b = warp_part[0];
            for (int i = 1; i < 8; ++i) b = lse3_combine(b, warp_part[i]);
            stats->max = b.m; stats->sumexp = b.s; stats->sumexp2 = b.s2;
            stats->ess = (b.s2 > 0.) ? (b.s * b.s) / b.s2 : 0.;
            stats->degenerate = (b.m == -INFINITY) ? 1 : 0;
            stats->blocks_done = 0;
        }
    }
}

// ================================================================================================
// K5: integer-weight systematic resampling
// ================================================================================================

template <typename Real>
struct FixedArgs {
    Real* lw;                 // in: log-weights.  The reduce pass overwrites them IN PLACE with the integer weights (exactly
                              // representable in Real): after a resample the log-weights are zero by definition
                              // (particle_filter.rs:114), and the scan then reads 4 bytes per particle without re-evaluating exp
    size_t n;                 // local particles
    int kbits;
    unsigned long long n_out; // global number of offspring (N_global)
    unsigned long long c_offset;   // integer weight of all lower-ranked shards
    unsigned long long out_base;   // global index of this shard's first output slot
    unsigned long long n_out_local;
    double log_n_global;
    int32_t* anc;             // local output slots [out_base, out_base + n_out_local)
    int32_t src_base;         // value written for local particle 0 (global id of it, or 0)
    unsigned long long* desc;
    DeviceStats* stats;
    unsigned long long* partials;   // gridDim.x of the reduce kernel
    uint64_t seed;
    long long rt;             // RNG tag of this resample (step whose weights are resampled); < 0: stats->t - 1
    int accumulate_lml;
    int dynamic;              // 1: ESS-triggered (device-resident loop): the reduce pass decides, the scan skips unless stats->do_resample
    double ess_threshold;     // dynamic: resample when ESS < ess_threshold (absolute, in particles)
    PeerTable peer;           // world == 1: single GPU
    long long epoch;          // step number the mailbox flags must have reached; < 0: stats->t
    int max_slot;             // >= 0: the max lives in stats->max_bits[max_slot] (left by the extend); < 0: in stats->max
    int overflow_follows;     // 1: a fixed_overflow_kernel launch follows the scan (heavy tiles are queued for it)
    int* overflow_seen_host;  // mapped host word, raised when any tile exceeds the heavy cap; the three 8-byte words at byte 16 of the
                              // same buffer carry the log total weight of this resample to the host (tagged with host_seq)
    unsigned int host_seq;    // tag of this resample call (nested scheme: the host polls the mapped words instead of synchronising)
    double* sq_partials;      // per tile: sum of (q * 2^-k)^2, for ESS = W^2 / sum q^2
    int inline_level1;        // nested scheme, small shards: no plan pass -- every warp of the expansion derives its chunks' slot starts itself
};
template <typename Real>
__device__ __forceinline__ float fixed_max(const FixedArgs<Real>& a) {
    return (float)(a.max_slot >= 0 ? from_ordered_bits(a.stats->max_bits[a.max_slot]) : a.stats->max);
}

// ---- sharded gates: called by every thread of a block at the top of a kernel -----------------------------------------------
__device__ __forceinline__ void local_ready_wait(const long long* word, long long epoch, const PeerTable& p) {
    const volatile long long* w = word;
    volatile int* err = &p.mail[p.rank]->error;
    while (*w < epoch) { if (*err) break; __nanosleep(50); }
    __threadfence();
}
__device__ __forceinline__ void local_ready_set(long long* word, long long epoch) {
    __threadfence();
    *(volatile long long*)word = epoch;
}
// every shard's max has arrived: the global max is exact and order-free
__device__ __forceinline__ void gate_stats(const PeerTable& p, DeviceStats* st, long long epoch) {
    if (p.world <= 1) return;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            st->trace[3] = global_ns();
            Mailbox* mb = p.mail[p.rank];
            SpinGuard g(p);
            double m = -INFINITY;
            for (int h = 0; h < p.world; ++h) m = fmax(m, __longlong_as_double((long long)ll_read64(&mb->stats_ll[h][0], (unsigned int)epoch, g)));
            st->max = m;
            st->trace[4] = global_ns();
            local_ready_set(&st->ready_stats, epoch);
        } else local_ready_wait(&st->ready_stats, epoch, p);
    }
    __syncthreads();
}
// every shard's integer weight total has arrived: global W and this shard's prefix
__device__ __forceinline__ void gate_weights(const PeerTable& p, DeviceStats* st, long long epoch, int dynamic = 0, double ess_threshold = 0.) {
    if (p.world <= 1) return;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            st->trace[6] = global_ns();
            Mailbox* mb = p.mail[p.rank];
            SpinGuard g(p);
            unsigned long long W = 0, c = 0;
            double sq = 0.;
            for (int h = 0; h < p.world; ++h) {
                unsigned long long w = ll_read64(&mb->w_ll[h][0], (unsigned int)epoch, g);
                sq += __longlong_as_double((long long)ll_read64(&mb->w_ll[h][2], (unsigned int)epoch, g));
                if (h < p.rank) c += w;
                W += w;
            }
            st->W = W; st->c_offset = c;
            st->sumexp2 = sq;
            st->ess = sq > 0. ? ((double)W * (double)W) / sq : 0.;
            if (dynamic) st->do_resample = (st->ess < ess_threshold) ? 1 : 0;
            st->trace[7] = global_ns();
            local_ready_set(&st->ready_w, epoch);
        } else local_ready_wait(&st->ready_w, epoch, p);
    }
    __syncthreads();
}
// every shard has written all the ancestors it owes (and is done reading the state buffer about to be overwritten)
__device__ __forceinline__ void gate_done(const PeerTable& p, DeviceStats* st, long long epoch) {
    if (p.world <= 1) return;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) { st->trace[0] = global_ns(); peer_wait_done(p, epoch); st->trace[1] = global_ns(); local_ready_set(&st->ready_done, epoch); }
        else local_ready_wait(&st->ready_done, epoch, p);
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long resample_rand_word(uint64_t seed, long long rt, const DeviceStats* st) {
    if (rt < 0) rt = st->t - 1;
    Stream s(seed, 0, (uint32_t)rt, P_RESAMPLE_OFFSET);
    uint4 x = s.block(0);
    return ((unsigned long long)x.x << 32) | x.y;
}

// 4 consecutive log-weights -> 4 integer weights; out-of-range lanes give 0 (FULL: the whole tile is in range).
// WRITEBACK: store the integer weights over the log-weights (as Real; every q is an integer below 2^41 with at most 24
// significant bits, hence exact in fp32).
template <typename Real, bool FULL, bool WRITEBACK = false>
__device__ __forceinline__ void load_q4(Real* lw, size_t idx, size_t n, float mx, int kbits, unsigned long long (&q)[4], float* sq = nullptr) {
    float w[4];
    if constexpr (sizeof(Real) == 4) {
        float4 v = *reinterpret_cast<const float4*>(lw + idx);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        double2 a = *reinterpret_cast<const double2*>(lw + idx), b = *reinterpret_cast<const double2*>(lw + idx + 2);
        w[0] = (float)a.x; w[1] = (float)a.y; w[2] = (float)b.x; w[3] = (float)b.y;
    }
    float qr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float qf;
        unsigned long long v = fixed_weight(__fsub_rn(w[j], mx), kbits, &qf);
        const bool in = FULL || idx + j < n;
        q[j] = in ? v : 0ull;
        qr[j] = in ? rintf(qf) : 0.f;
        if (sq && in) *sq = fmaf(qf, qf, *sq);
    }
    if constexpr (WRITEBACK) {
        if constexpr (sizeof(Real) == 4) *reinterpret_cast<float4*>(lw + idx) = make_float4(qr[0], qr[1], qr[2], qr[3]);
        else {
            *reinterpret_cast<double2*>(lw + idx) = make_double2((double)qr[0], (double)qr[1]);
            *reinterpret_cast<double2*>(lw + idx + 2) = make_double2((double)qr[2], (double)qr[3]);
        }
    }
}
// 4 consecutive STORED integer weights (left by the reduce pass)
template <typename Real, bool FULL>
__device__ __forceinline__ void load_stored_q4(const Real* lw, size_t idx, size_t n, unsigned long long (&q)[4]) {
    float w[4];
    if constexpr (sizeof(Real) == 4) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(lw + idx));   // last use
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        double2 a = *reinterpret_cast<const double2*>(lw + idx), b = *reinterpret_cast<const double2*>(lw + idx + 2);
        w[0] = (float)a.x; w[1] = (float)a.y; w[2] = (float)b.x; w[3] = (float)b.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = (FULL || idx + j < n) ? __float2ull_rn(w[j]) : 0ull;
}

// R1: per-tile integer weight sums; the last block to finish turns them into exclusive tile prefixes (in place, in
// `desc`) and the grand total W.  Integer addition is associative, so W and every prefix are exact and independent of
// the order in which blocks run (and of how particles are sharded).  One tile (kScanTile particles) per block.
constexpr int kScanThreads = 256;
constexpr int kScanRounds = 4;
constexpr int kScanTile = kScanThreads * 4 * kScanRounds;   // 4096 particles per tile
constexpr unsigned int kHeavyCap = 32u * kScanTile;        // tiles with more offspring than this go to the overflow pass

template <typename Real, bool WRITEBACK>
__global__ void __launch_bounds__(kScanThreads) fixed_reduce_kernel(FixedArgs<Real> a, unsigned int num_tiles) {
    // one tile per block; warps add their partial sums straight into desc[tile] (zeroed by the previous scan pass), so
    // the only block-wide barrier is the one in front of the last-block test.  WRITEBACK = false (ESS-triggered runs): the
    // log-weights must survive when the decision is "do not resample".
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned long long ws[kScanThreads / 32];
    __shared__ double wsq[kScanThreads / 32];
    __shared__ unsigned long long carry_s;
    __shared__ bool is_last;
    pdl_wait();
    gate_stats(a.peer, a.stats, a.epoch < 0 ? a.stats->t : a.epoch);
    const float mx = fixed_max<Real>(a);
    {
        const unsigned int tile = blockIdx.x;
        unsigned long long sum = 0;
        float sqf = 0.f;   // sum of squared weights of this thread's 16 particles (for the ESS)
        const size_t tile_base = (size_t)tile * kScanTile;
        if (tile_base + kScanTile <= a.n) {
#pragma unroll
            for (int r = 0; r < kScanRounds; ++r) {
                unsigned long long q[4];
                load_q4<Real, true, WRITEBACK>(a.lw, tile_base + (size_t)r * (kScanThreads * 4) + (size_t)tid * 4, a.n, mx, a.kbits, q, &sqf);
                sum += q[0] + q[1] + q[2] + q[3];
            }
        } else {
#pragma unroll
            for (int r = 0; r < kScanRounds; ++r) {
                size_t idx = tile_base + (size_t)r * (kScanThreads * 4) + (size_t)tid * 4;
                if (idx < a.n) {
                    unsigned long long q[4];
                    load_q4<Real, false, WRITEBACK>(a.lw, idx, a.n, mx, a.kbits, q, &sqf);
                    sum += q[0] + q[1] + q[2] + q[3];
                }
            }
        }
        double sq = (double)sqf;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
        if (lane == 0 && sum) atomicAdd(&a.desc[tile], sum);
        if (lane == 0) wsq[warp] = sq;
    }
    pdl_trigger();
    __syncthreads();
    if (tid == 0) {   // fixed summation order: the ESS is reproducible
        double b = 0.;
#pragma unroll
        for (int i = 0; i < kScanThreads / 32; ++i) b += wsq[i];
        a.sq_partials[blockIdx.x] = b;
    }
    if (tid == 0) {
        __threadfence();
        is_last = (atomicAdd(&a.stats->blocks_done, 1u) == gridDim.x - 1);
        carry_s = 0ull;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // sharded: the shard total first -- every rank is waiting for it -- then the prefix scan overlaps the NVLink latency
    if (a.peer.world > 1) {
        unsigned long long tot = 0;
        double sqt0 = 0.;
        for (unsigned int i = tid; i < num_tiles; i += kScanThreads) { tot += a.desc[i]; sqt0 += a.sq_partials[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, o); sqt0 += __shfl_xor_sync(0xffffffffu, sqt0, o); }
        __syncthreads();
        if (lane == 0) { ws[warp] = tot; wsq[warp] = sqt0; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long b = 0;
            double sq = 0.;
#pragma unroll
            for (int i = 0; i < kScanThreads / 32; ++i) { b += ws[i]; sq += wsq[i]; }
            DeviceStats* st = a.stats;
            const long long epoch = a.epoch < 0 ? st->t : a.epoch;
            st->trace[5] = global_ns();
            for (int h = 0; h < a.peer.world; ++h) {
                ll_write64(a.peer.mail[h]->w_ll[a.peer.rank], b, (unsigned int)epoch);
                ll_write64(a.peer.mail[h]->w_ll[a.peer.rank] + 2, (unsigned long long)__double_as_longlong(sq), (unsigned int)epoch);
            }
        }
        __syncthreads();
    }
    // exclusive scan of the tile sums, 4 consecutive tiles per thread per pass
    for (unsigned int base = 0; base < num_tiles; base += kScanThreads * 4) {
        unsigned long long v[4], tot = 0;
        const unsigned int first = base + tid * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = (first + i < num_tiles) ? a.desc[first + i] : 0ull; tot += v[i]; }
        unsigned long long incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        __syncthreads();
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        unsigned long long pre = carry_s + incl - tot, all = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) { unsigned long long x = ws[w]; if (w < warp) pre += x; all += x; }
#pragma unroll
        for (int i = 0; i < 4; ++i) { if (first + i < num_tiles) a.desc[first + i] = pre; pre += v[i]; }
        __syncthreads();
        if (tid == 0) carry_s += all;
        __syncthreads();
    }
    // sum of squared weights over the tiles, in tile order
    double sqt = 0.;
    for (unsigned int i = tid; i < num_tiles; i += kScanThreads) sqt += a.sq_partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sqt += __shfl_xor_sync(0xffffffffu, sqt, o);
    __syncthreads();
    if (lane == 0) wsq[warp] = sqt;
    __syncthreads();
    if (tid == 0) {
        DeviceStats* st = a.stats;
        st->overflow_count = 0; st->blocks_done = 0;
        double sq = 0.;
#pragma unroll
        for (int i = 0; i < kScanThreads / 32; ++i) sq += wsq[i];
        if (a.peer.world <= 1) {
            st->W = carry_s; st->c_offset = 0;
            st->sumexp2 = sq;
            st->ess = sq > 0. ? ((double)carry_s * (double)carry_s) / sq : 0.;   // 1 / sum(w~^2), particle_filter.rs:98-100
            if (a.dynamic) st->do_resample = (st->ess < a.ess_threshold) ? 1 : 0;
        }   // (sharded: already posted above)
    }
}

// Offspring arithmetic.  Output slot j (0 <= j < n_out) sits at integer position j*W + U (U in [0, W)); a particle whose
// inclusive integer prefix is C owns every slot below C*n_out.  #slots below C*n_out = floor(X / W) + 1 with
// X = C*n_out - U - 1 (0 when X < 0).  Per tile the quotient/remainder of the tile's exclusive prefix are found once,
// exactly, in 128-bit arithmetic: X_t = Q*W + R with Q = n_start - 1 (Q = -1, R = X_t + W when X_t < 0).
struct TileBase {
    unsigned long long n_start;   // slots owned by everything before the tile
    unsigned long long rem;       // R
};
__device__ __forceinline__ TileBase tile_base_exact(unsigned long long C, unsigned long long W, unsigned long long U, unsigned long long n_out, double inv_w) {
    u128 lhs = (u128)C * n_out;
    if (lhs <= (u128)U) return TileBase{0ull, (unsigned long long)(lhs + W - U - 1)};
    u128 X = lhs - U - 1;
    double ed = ((double)C * (double)n_out - (double)U) * inv_w;
    unsigned long long e = ed <= 0. ? 0ull : (ed >= (double)n_out ? n_out : (unsigned long long)ed);
    u128 p = (u128)e * W;
    while (p > X) { --e; p -= W; }
    while (p + W <= X) { ++e; p += W; }
    return TileBase{e + 1, (unsigned long long)(X - p)};
}
// Slots owned by the tile's particles up to local inclusive prefix c: floor((R + c*n_out) / W).  fp64 gets this right
// unless the quotient lands within 2^-16 of an integer (the estimate's error is < 2^-19 even for 2^31 offspring); only
// then is the exact 128-bit form evaluated (a ~3e-5 fraction of particles).
__device__ __forceinline__ unsigned int local_count(unsigned long long c, unsigned long long rem, double rem_d, unsigned long long W, double n_out_d,
                                                    unsigned long long n_out, double inv_w) {
    double fd = fma((double)c, n_out_d, rem_d) * inv_w;
    double fl = floor(fd);
    double frac = fd - fl;
    unsigned long long f = (unsigned long long)fl;
    if (frac < 0x1.0p-16 || frac > 1. - 0x1.0p-16) {
        u128 y = (u128)c * n_out + rem;
        u128 p = (u128)f * W;
        while (p > y) { --f; p -= W; }
        while (p + W <= y) { ++f; p += W; }
    }
    return (unsigned int)f;
}

struct __align__(16) ScanShared {
    unsigned long long warp_tot[kScanRounds][kScanThreads / 32];
    unsigned long long tile_excl;
    unsigned long long rand_word;
    TileBase base;
};

// Loads a tile, quantises, and leaves in excl[r] the tile-local exclusive prefix of this thread's round-r chunk.
// Returns the tile aggregate.  Element order inside the tile: e = r*1024 + 4*tid + j.
template <typename Real>
__device__ __forceinline__ unsigned long long tile_local_scan(const FixedArgs<Real>& a, ScanShared& sh, unsigned int tile, float mx,
                                                              unsigned long long (&q)[kScanRounds][4], unsigned long long (&excl)[kScanRounds]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t tile_base = (size_t)tile * kScanTile;
    unsigned long long incl[kScanRounds];
    if (tile_base + kScanTile <= a.n) {
#pragma unroll
        for (int r = 0; r < kScanRounds; ++r) {
            load_stored_q4<Real, true>(a.lw, tile_base + (size_t)r * (kScanThreads * 4) + (size_t)tid * 4, a.n, q[r]);
            incl[r] = q[r][0] + q[r][1] + q[r][2] + q[r][3];
        }
    } else {
#pragma unroll
        for (int r = 0; r < kScanRounds; ++r) {
            size_t idx = tile_base + (size_t)r * (kScanThreads * 4) + (size_t)tid * 4;
            if (idx < a.n) load_stored_q4<Real, false>(a.lw, idx, a.n, q[r]);
            else { q[r][0] = q[r][1] = q[r][2] = q[r][3] = 0ull; }
            incl[r] = q[r][0] + q[r][1] + q[r][2] + q[r][3];
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int r = 0; r < kScanRounds; ++r) {
            unsigned long long up = __shfl_up_sync(0xffffffffu, incl[r], o);
            if (lane >= o) incl[r] += up;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int r = 0; r < kScanRounds; ++r) sh.warp_tot[r][warp] = incl[r];
    }
    __syncthreads();
    unsigned long long carry = 0, aggregate = 0;
#pragma unroll
    for (int r = 0; r < kScanRounds; ++r) {
        unsigned long long before = 0, round_tot = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) {
            unsigned long long v = sh.warp_tot[r][w];
            if (w < warp) before += v;
            round_tot += v;
        }
        unsigned long long own = q[r][0] + q[r][1] + q[r][2] + q[r][3];
        excl[r] = carry + before + incl[r] - own;
        carry += round_tot;
    }
    aggregate = carry;
    return aggregate;
}

// ================================================================================================
// K4: the reference's multinomial routine, bit-exact
// ================================================================================================
// normalize_weights (particle_filter.rs:27-35): probs = exp(lw - lse); also the scalar bookkeeping of resample()
template <typename Real>
__global__ void __launch_bounds__(256) normalize_kernel(const Real* __restrict__ lw, size_t n, double* __restrict__ probs, DeviceStats* st,
                                                        double log_n, int accumulate_lml) {
    const double lse = st->max + log(st->sumexp);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        double lnw = (double)lw[i] - lse;
        probs[i] = exp(lnw);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->lse = lse;
        st->ess_stale = st->ess;
        if (accumulate_lml) st->lml_acc += lse - log_n;
        st->resampled = 1;
    }
}

// Sequential f64 running sum S_k = fl(S_{k-1} + p_k) (categorical.rs:25-30).  One block: all threads stage chunks
// through shared memory (coalesced), thread 0 performs the dependent adds in index order.
constexpr int kSeqChunk = 2048;
static __global__ void __launch_bounds__(256) cumsum_seq_kernel(const double* __restrict__ p, size_t n, double* __restrict__ out) {
    __shared__ double buf[2][kSeqChunk];
    __shared__ double carry;
    const int tid = threadIdx.x;
    if (tid == 0) carry = 0.;
    size_t nchunks = (n + kSeqChunk - 1) / kSeqChunk;
    for (int i = tid; i < kSeqChunk; i += 256) buf[0][i] = ((size_t)i < n) ? p[i] : 0.;
    __syncthreads();
    for (size_t c = 0; c < nchunks; ++c) {
        int cur = c & 1;
        size_t base = c * kSeqChunk;
        if (tid == 0) {
            double t = carry;
            double* b = buf[cur];
            size_t cnt = min((size_t)kSeqChunk, n - base);
            for (size_t i = 0; i < cnt; ++i) { t = __dadd_rn(t, b[i]); b[i] = t; }
            carry = t;
        } else if (c + 1 < nchunks) {   // the other threads prefetch the next chunk meanwhile
            size_t nb = base + kSeqChunk;
            for (int i = tid - 1; i < kSeqChunk; i += 255) buf[cur ^ 1][i] = (nb + i < n) ? p[nb + i] : 0.;
        }
        __syncthreads();
        for (int i = tid; i < kSeqChunk; i += 256) if (base + i < n) out[base + i] = buf[cur][i];
        __syncthreads();
    }
}

// parent = min{k : S_k >= u}, clamped to [0, n-1] (quirk Q2)
__device__ __forceinline__ long long search_cumsum(const double* __restrict__ S, size_t n, double u) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (S[mid] >= u) hi = mid; else lo = mid + 1; }
    return (long long)(lo >= n ? n - 1 : lo);
}

// mode 0: injected uniforms; 1: systematic from injected u0; 2: Philox multinomial; 3: Philox systematic
template <typename Out>
__global__ void __launch_bounds__(256) search_kernel(const double* __restrict__ S, size_t n, size_t n_draws, int mode, const double* __restrict__ uniforms,
                                                     uint64_t seed, uint64_t gid_offset, uint32_t t, Out* __restrict__ parents) {
    double u0 = 0.;
    if (mode == 1) u0 = uniforms[0];
    if (mode == 3) { Stream s(seed, 0, t, P_RESAMPLE_OFFSET); uint4 x = s.block(0); u0 = u01_co64(x.x, x.y); }
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_draws; i += (size_t)gridDim.x * 256) {
        double u;
        if (mode == 0) u = uniforms[i];
        else if (mode == 2) { Stream s(seed, gid_offset + i, t, P_RESAMPLE_U); uint4 x = s.block(0); u = u01_co64(x.x, x.y); }
        else u = (u0 + (double)i) / (double)n_draws;
        parents[i] = (Out)search_cumsum(S, n, u);
    }
}

// integer multinomial: materialised integer cumsum + per-output search
template <typename Real>
__global__ void __launch_bounds__(kScanThreads) fixed_cumsum_kernel(FixedArgs<Real> a, unsigned long long* __restrict__ C) {
    // inclusive integer prefix sums: tile prefix from the reduce pass + tile-local scan
    __shared__ ScanShared sh;
    const int tid = threadIdx.x;
    const unsigned int tile = blockIdx.x;
    const float mx = fixed_max<Real>(a);
    unsigned long long q[kScanRounds][4], excl[kScanRounds];
    tile_local_scan<Real>(a, sh, tile, mx, q, excl);
    const unsigned long long tile_excl = a.desc[tile];
    __syncthreads();
    if (tid == 0) a.desc[tile] = 0ull;   // ready for the next reduce pass
#pragma unroll
    for (int r = 0; r < kScanRounds; ++r) {
        size_t idx = (size_t)tile * kScanTile + (size_t)r * (kScanThreads * 4) + (size_t)tid * 4;
        unsigned long long c = tile_excl + excl[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) { c += q[r][j]; if (idx + j < a.n) C[idx + j] = c; }
    }
    if (tile == 0 && tid == 0) {
        DeviceStats* st = a.stats;
        double lse = (double)mx + log((double)st->W) - (double)a.kbits * 0.6931471805599453;
        st->lse = lse; st->ess_stale = st->ess;
        if (a.accumulate_lml) st->lml_acc += lse - a.log_n_global;
        st->resampled = 1;
    }
}

static __global__ void __launch_bounds__(256) fixed_multinomial_search_kernel(const unsigned long long* __restrict__ C, size_t n, size_t n_draws, const DeviceStats* st,
                                                                       uint64_t seed, uint64_t gid_offset, uint32_t t, int32_t* __restrict__ anc) {
    const unsigned long long W = st->W;
    for (size_t j = (size_t)blockIdx.x * 256 + threadIdx.x; j < n_draws; j += (size_t)gridDim.x * 256) {
        if (W == 0ull) { anc[j] = (int32_t)j; continue; }
        Stream s(seed, gid_offset + j, t, P_RESAMPLE_U);
        uint4 x = s.block(0);
        unsigned long long T = __umul64hi(((unsigned long long)x.x << 32) | x.y, W);
        size_t lo = 0, hi = n;   // min{k : C_k > T}
        while (lo < hi) { size_t mid = (lo + hi) >> 1; if (C[mid] > T) hi = mid; else lo = mid + 1; }
        anc[j] = (int32_t)(lo >= n ? n - 1 : lo);
    }
}

// ================================================================================================
// K6 stand-alone gather (trace clone loop, particle_filter.rs:109-113) -- only the live state (quirk Q11)
// ================================================================================================
template <typename Real>
__global__ void __launch_bounds__(256) gather_kernel(const Real* __restrict__ in, Real* __restrict__ out, const int32_t* __restrict__ anc, size_t n, size_t ld, int D) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        size_t src = (size_t)anc[i];
        for (int d = 0; d < D; ++d) out[i * D + d] = __ldg(in + src * D + d);
    }
}

// trajectories by back-tracing the ancestor log: one thread per requested particle, newest step first
template <typename Real>
__global__ void __launch_bounds__(128) backtrace_kernel(const Real* __restrict__ hist_state, const int32_t* __restrict__ hist_anc, const int* __restrict__ resampled,
                                                        size_t ld, int D, int T, int start_after_resample, const long long* __restrict__ ids, size_t n_ids,
                                                        double* __restrict__ out /* [n_ids][T][D] */) {
    for (size_t k = (size_t)blockIdx.x * 128 + threadIdx.x; k < n_ids; k += (size_t)gridDim.x * 128) {
        size_t cur = (size_t)ids[k];
        if (start_after_resample) cur = (size_t)hist_anc[(size_t)(T - 1) * ld + cur];   // ids name post-resample particles
        for (int t = T - 1; t >= 0; --t) {
            for (int d = 0; d < D; ++d) out[(k * T + t) * D + d] = (double)hist_state[((size_t)t * ld + cur) * D + d];
            if (t > 0 && resampled[t - 1]) cur = (size_t)hist_anc[(size_t)(t - 1) * ld + cur];
        }
    }
}

template <typename Real>
__global__ void __launch_bounds__(256) fill_kernel(Real* p, size_t n, Real v) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = v;
}

template <typename Real>
__global__ void __launch_bounds__(256) to_f64_kernel(const Real* __restrict__ in, double* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = (double)in[i];
}
template <typename Real>
__global__ void __launch_bounds__(256) from_f64_kernel(const double* __restrict__ in, Real* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = (Real)in[i];
}
// host-facing reads / writes of the state are double[D][N] (`state[d * N + i]`); on the device it is particle-major in Real
template <typename Real>
__global__ void __launch_bounds__(256) state_to_f64_kernel(const Real* __restrict__ in, double* __restrict__ out, size_t n, size_t ld, int D) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        for (int d = 0; d < D; ++d) out[(size_t)d * ld + i] = (double)in[i * D + d];
}
template <typename Real>
__global__ void __launch_bounds__(256) state_from_f64_kernel(const double* __restrict__ in, Real* __restrict__ out, size_t n, size_t ld, int D) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        for (int d = 0; d < D; ++d) out[i * D + d] = (Real)in[(size_t)d * ld + i];
}
static __global__ void __launch_bounds__(256) i32_to_i64_kernel(const int32_t* __restrict__ in, long long* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = (long long)in[i];
}

}  // namespace mpl

#include "scan2.cuh"
#include "nested.cuh"
