// nested.cuh -- nested systematic resampling on integer weights (MPL_RESAMPLE_SYSTEMATIC_NESTED).
//
// The single-level integer scheme quantises every weight against the GLOBAL maximum, so a whole pass over the
// log-weights (the reduce pass) has to wait for that maximum -- on several GPUs, for an exchange -- before it can even
// start.  Here nothing is ever scaled against a global quantity until only a handful of numbers is left:
//   chunk c   (128 particles, one warp iteration of the extend kernel):
//       e_c = ceil(max_i lw_i * log2 e),   q_i = rint(2^(lw_i log2 e - e_c) * 2^22) (32-bit),   S_c = sum q_i
//     -> needs nothing outside the warp, fuses into the extend kernel's epilogue
//   section s (2^17 particles = 1024 chunks, one block of the section pass):
//       E_s = max e_c,   G_c = S_c >> (E_s - e_c),   T_s = sum G_c
//   top       (at most 2048 sections; across GPUs: the ONLY exchange, one record per section):
//       E = max E_s,   M_s = T_s >> (E - E_s),   W = sum M_s
// Resampling runs the exact integer systematic scheme three times: the N output slots over the sections by M_s (slot j
// at j*W + U), a section's n_s slots over its chunks by G_c (local slot l at l*T_s + U_s), a chunk's n_c slots over its
// particles by q_i (l*S_c + U_c); U_s and U_c are hashed from the step's random word and the section / chunk number.
// All of it is integer arithmetic: unbiased up to the floors (relative 2^-22), independent of thread order and of how
// the particles are sharded (sections are aligned groups of global ids).
// Kernels per resample: [quantise, unless the extend kernel's epilogue did it] -> section pass (+ top level) -> plan pass
// (level 1: first output slot of every chunk) -> expansion (level 2: offspring counts of every particle, expanded into the
// ancestor slots) [-> heavy-tile pass once a heavy warp tile has been seen].
//
// Levels 1 and 2 are organised by the slots a GPU has to FILL, not by the particles it owns: a block of the plan pass takes
// one section whose slots fall into this GPU's slot range -- wherever that section's particles live -- and the expansion
// walks the chunks that own those slots.  Across GPUs everything is therefore a remote LOAD (chunk records, integer
// weights) issued after the step's only gate, the exchange of section records; nothing is ever stored into another GPU's
// memory (a kernel that has stored into a peer's memory cannot complete before those stores are acknowledged, ~4 us at
// every kernel boundary), and no rank waits for another's ancestors.
#pragma once

namespace mpl {

constexpr int kChunksPerTile = kScanTile / kChunk;   // 32
constexpr int kTilesPerSection = 32;                 // one block of the section pass (one warp scan over its tile sums)
static_assert(kSection == (size_t)kTilesPerSection * kScanTile, "a section is 32 tiles");


// warp-wide sum of per-lane values below 2^48: two integer redux instructions instead of ten 32-bit shuffles
__device__ __forceinline__ unsigned long long warp_sum_u48(unsigned long long v) {
    const unsigned int lo = (unsigned int)v & 0xFFFFFFu, hi = (unsigned int)(v >> 24);
    return ((unsigned long long)__reduce_add_sync(0xffffffffu, hi) << 24) + (unsigned long long)__reduce_add_sync(0xffffffffu, lo);
}

// ---- stand-alone quantisation pass (call-per-step API; the device-resident loop fuses this into the extend kernel) -----
// KEEP: only the chunk records are written, the log-weights stay (ESS-triggered loop: the resample may be skipped)
template <typename Real, bool KEEP>
__global__ void __launch_bounds__(kScanThreads) nested_quantise_kernel(FixedArgs<Real> a, ChunkRecords rec) {
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t wt_base = (size_t)blockIdx.x * kScanTile + (size_t)warp * kWarpTile;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const size_t idx = wt_base + (size_t)r * kChunk + (size_t)lane * 4;
        const size_t chunk = idx / kChunk;   // (same for the whole warp)
        if (wt_base + (size_t)r * kChunk >= a.n) break;
        float w[4];
        if constexpr (sizeof(Real) == 4) { float4 v = *reinterpret_cast<const float4*>(a.lw + idx); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
        else { double2 u = *reinterpret_cast<const double2*>(a.lw + idx), v = *reinterpret_cast<const double2*>(a.lw + idx + 2); w[0] = (float)u.x; w[1] = (float)u.y; w[2] = (float)v.x; w[3] = (float)v.y; }
#pragma unroll
        for (int j = 0; j < 4; ++j) if (idx + j >= a.n) w[j] = -INFINITY;
        unsigned int qv[4], S_c;
        float sq;
        int e_c;
        warp_quantise_chunk(w, qv, e_c, S_c, sq);
        if constexpr (!KEEP) {
            if constexpr (sizeof(Real) == 4) *reinterpret_cast<float4*>(a.lw + idx) = make_float4(nested_store<float>(qv[0]), nested_store<float>(qv[1]), nested_store<float>(qv[2]), nested_store<float>(qv[3]));
            else { *reinterpret_cast<double2*>(a.lw + idx) = make_double2((double)qv[0], (double)qv[1]); *reinterpret_cast<double2*>(a.lw + idx + 2) = make_double2((double)qv[2], (double)qv[3]); }
        }
        if (lane == 0) { rec.e[chunk] = e_c; rec.S[chunk] = S_c; rec.sq[chunk] = sq; }
    }
    pdl_trigger();
}

__device__ __forceinline__ unsigned long long nested_shift(unsigned long long S, int e, int E) {   // S >> (E - e), e <= E
    if (e == kChunkEmpty) return 0ull;
    const int sft = E - e;
    return sft < 64 ? (S >> sft) : 0ull;
}
__device__ __forceinline__ unsigned long long nested_section_offset(unsigned long long word, unsigned long long section, unsigned long long T_s) {
    return __umul64hi(splitmix64_mix((word ^ 0x5851F42D4C957F2Dull) + (section + 1ull) * 0xD1B54A32D192ED03ull), T_s);
}
__device__ __forceinline__ unsigned long long nested_chunk_offset(unsigned long long word, unsigned long long chunk, unsigned long long S_c) {
    return __umul64hi(splitmix64_mix(word + (chunk + 1ull) * 0x9E3779B97F4A7C15ull), S_c);
}

// 2^(-2d), exactly, for 0 <= d < 500 (rescales a sum of squares between two power-of-two references)
__device__ __forceinline__ double pow2_neg2(int d) { return __longlong_as_double((long long)(1023 - 2 * d) << 52); }

// Top level, by the section pass' last block (every thread of it): E, M_s, W, the ESS, and level 0 of the resampling (slot j sits
// at j*W + U; section s owns the slots [a_s, a_s + n_s)) from the section records.  One section per thread, 256 per round: up to
// 2^25 particles the prefixes never leave the registers, and the exact 128-bit slot base is one evaluation deep.
__device__ __forceinline__ void nested_top_level_block(const NestedPrefixes& nb, DeviceStats* st, uint64_t seed, long long rt, unsigned long long n_out, int dynamic,
                                                       double ess_threshold) {
    __shared__ unsigned long long s_tot[kScanThreads / 32];
    __shared__ double s_sq[kScanThreads / 32];
    __shared__ int s_E[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int n_sec = nb.n_sec_global;
    int E = kChunkEmpty;
    for (unsigned int i = tid; i < n_sec; i += kScanThreads) E = max(E, __ldcg(nb.sec_E + i));
    E = __reduce_max_sync(0xffffffffu, E);
    if (lane == 0) s_E[warp] = E;
    const unsigned long long word = resample_rand_word(seed, rt, st);   // (independent of the loads)
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) E = max(E, s_E[w]);
    unsigned long long carry = 0, v = 0, pre = 0;
    double sqt = 0.;
    for (unsigned int base = 0; base < n_sec; base += kScanThreads) {
        const unsigned int sc = base + tid;
        const bool ok = sc < n_sec;
        const int e_s = ok ? __ldcg(nb.sec_E + sc) : kChunkEmpty;
        const unsigned long long T = ok ? __ldcg(nb.sec_T + sc) : 0ull;
        const double sq = ok ? __ldcg(nb.sec_sq + sc) : 0.;
        v = nested_shift(T, e_s, E);
        if (e_s != kChunkEmpty && E - e_s < 500) sqt += sq * pow2_neg2(E - e_s);
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) s_tot[warp] = incl;
        __syncthreads();
        unsigned long long before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) { const unsigned long long t = s_tot[w]; if (w < warp) before += t; all += t; }
        pre = carry + before + incl - v;
        if (ok && n_sec > (unsigned int)kScanThreads) { nb.sec_pre[sc] = pre; nb.sec_M[sc] = v; }   // (several rounds: parked until W is known)
        carry += all;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sqt += __shfl_xor_sync(0xffffffffu, sqt, o);
    if (lane == 0) s_sq[warp] = sqt;
    const unsigned long long W = carry;
    if (W != 0ull) {
        const unsigned long long U = __umul64hi(word, W);
        const double inv_w = 1. / (double)W;
        for (unsigned int base = 0; base < n_sec; base += kScanThreads) {
            const unsigned int sc = base + tid;
            if (sc >= n_sec) break;
            if (n_sec > (unsigned int)kScanThreads) { pre = nb.sec_pre[sc]; v = nb.sec_M[sc]; }   // (what this thread wrote itself)
            const TileBase sb = tile_base_exact(pre, W, U, n_out, inv_w);
            nb.sec_a[sc] = sb.n_start;
            nb.sec_n[sc] = local_count(v, sb.rem, (double)sb.rem, W, (double)n_out, n_out, inv_w);
        }
    }
    __syncthreads();
    if (tid == 0) {
        sqt = 0.;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) sqt += s_sq[w];
        st->W = W; st->c_offset = 0; st->nest_E = E; st->rand_word = word;
        st->sumexp2 = sqt;
        const double ess = sqt > 0. ? ((double)W * (double)W) / sqt : 0.;
        st->ess = ess;
        if (dynamic) st->do_resample = (ess < ess_threshold) ? 1 : 0;   // ESS trigger, decided where the numbers are
    }
}

// several GPUs: collect every other shard's section records (the caller then runs the top level), the whole block: one record
// per thread, so that all of them are in flight together -- ONE NVLink round trip.  The records are PULLED: every rank publishes
// its own in its own mailbox and the readers poll them over NVLink -- a kernel that stores into another GPU's memory cannot
// complete before those stores are acknowledged (~4 us at every kernel boundary), a kernel that only loads has nothing to wait for.
__device__ __forceinline__ void collect_section_records(const PeerTable& p, long long epoch, const NestedPrefixes& nb) {
    SpinGuard g(p);
    for (unsigned int sg = threadIdx.x; sg < nb.n_sec_global; sg += blockDim.x) {
        if (sg - nb.sec0 < nb.n_sec) continue;   // own sections: already in place
        const volatile unsigned long long* src = p.mail[sg / nb.n_sec]->sec_ll[epoch & 1][sg];   // (equal shards of whole sections: the owner of global section sg)
        unsigned long long w[6];
        for (;;) {   // all six tagged words in flight together: one NVLink round trip per poll
#pragma unroll
            for (int k = 0; k < 6; ++k) w[k] = src[k];
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 6; ++k) ok = ok && (unsigned int)(w[k] >> 32) == (unsigned int)epoch;
            if (ok || g.give_up()) break;
        }
        atomicAdd(&p.mail[p.rank]->nvlink_polled, 48ull);
        nb.sec_E[sg] = (int)(unsigned int)w[0];
        nb.sec_T[sg] = (w[2] & 0xffffffffull) | (w[3] << 32);
        nb.sec_sq[sg] = __longlong_as_double((long long)((w[4] & 0xffffffffull) | (w[5] << 32)));
    }
    __threadfence();
    __syncthreads();
}

// the tail of the section level, by the block that completed this shard's LAST section: (several GPUs: the other shards' records,)
// the top level
__device__ __forceinline__ void nested_after_sections(const PeerTable& peer, const NestedPrefixes& nb, DeviceStats* st, long long epoch, uint64_t seed, long long rt,
                                                      unsigned long long n_out, int dynamic, double ess_threshold) {
    if (peer.world > 1) { if (threadIdx.x == 0) st->trace[6] = global_ns(); collect_section_records(peer, epoch, nb); }
    nested_top_level_block(nb, st, seed, rt, n_out, dynamic, ess_threshold);
    if (threadIdx.x == 0) st->trace[7] = global_ns();
}

// scalar bookkeeping of resample(): particle_filter.rs:104-105,114 (one thread of the expansion kernel)
// The log total weight goes straight to the host as well: three self-validating words in mapped pinned memory, so the caller
// of resample() has its return value as soon as it exists -- before the expansion has even started -- and can queue the next
// step behind it without ever draining the stream.
template <typename Real>
__device__ __forceinline__ void nested_post_to_host(const FixedArgs<Real>& a, double lse, int degenerate) {
    if (!a.overflow_seen_host || a.host_seq == 0u) return;   // (nobody polls inside the device-resident loop: no store to host memory there)
    volatile unsigned long long* hm = reinterpret_cast<volatile unsigned long long*>(a.overflow_seen_host) + 2;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(lse), tag = (unsigned long long)a.host_seq << 32;
    hm[0] = (bits & 0xffffffffull) | tag;
    hm[1] = (bits >> 32) | tag;
    hm[2] = (unsigned long long)(unsigned int)degenerate | tag;
}

// scalar bookkeeping of resample(): particle_filter.rs:104-105,114 (one thread of the plan pass)
template <typename Real>
__device__ __forceinline__ void nested_bookkeeping(const FixedArgs<Real>& a, DeviceStats* st, long long epoch) {
    const unsigned long long W = st->W;
    if (W == 0ull) {
        st->degenerate = 1; st->lse = -INFINITY; st->resampled = 1; st->resampled_flag[epoch & 1] = 1;
        nested_post_to_host(a, -INFINITY, 1);
        return;
    }
    const double lse = (double)st->nest_E * 0.6931471805599453 + log((double)W) - (double)kNestedBits * 0.6931471805599453;
    nested_post_to_host(a, lse, 0);
    st->lse = lse;
    st->ess_stale = st->ess;
    if (a.accumulate_lml) st->lml_acc += lse - a.log_n_global;
    st->resampled = 1;
    st->resampled_flag[epoch & 1] = 1;
    st->n_resamples += 1;
    st->degenerate = 0;
}

// ---- section pass: one small kernel between the extend and the expansion.
//   phase A (a block per section): section scale, chunk masses, tile sums and their prefixes inside the section; the
//            section record (sent to every other GPU right away)
//   top      the last block to finish phase A collects the other shards' records (several GPUs) and runs the top level
// PHASES = 3: both.  1: phase A only, 2: (collect +) top level only, one block -- separate launches for the test hook that
// emulates shards on one GPU, where a kernel must never wait for one that has not been launched.
template <typename Real, int PHASES>
__global__ void __launch_bounds__(kScanThreads) nested_sections_kernel(FixedArgs<Real> a, ChunkRecords rec, NestedPrefixes nb, unsigned int num_tiles,
                                                                       unsigned int num_chunks) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned long long ts[kTilesPerSection];
    __shared__ double tsq[kTilesPerSection];
    __shared__ int wmax[kScanThreads / 32];
    __shared__ bool is_last;
    DeviceStats* st = a.stats;
    pdl_wait();
    pdl_trigger();   // the expansion kernel may become resident now: it loads its weights while this grid runs
    constexpr int kTilesPerWarp = kTilesPerSection / (kScanThreads / 32);
    const long long epoch = a.epoch < 0 ? st->t : a.epoch;
    if (blockIdx.x == 0 && tid == 0) st->trace[9] = global_ns();
    if constexpr (PHASES == 2) {
        nested_after_sections(a.peer, nb, st, epoch, a.seed, a.rt, a.n_out, a.dynamic, a.ess_threshold);
        return;
    }
    const unsigned int sec = blockIdx.x;
    const unsigned int tile0 = sec * kTilesPerSection + warp * kTilesPerWarp;
    int e[kTilesPerWarp];
    unsigned int S[kTilesPerWarp];
    float sqf[kTilesPerWarp];
    int emax = kChunkEmpty;
#pragma unroll
    for (int i = 0; i < kTilesPerWarp; ++i) {
        const unsigned int c = (tile0 + i) * kChunksPerTile + lane;
        const bool valid = c < num_chunks;
        e[i] = valid ? rec.e[c] : kChunkEmpty;
        S[i] = valid ? rec.S[c] : 0u;
        sqf[i] = valid ? rec.sq[c] : 0.f;
        emax = max(emax, e[i]);
    }
    emax = __reduce_max_sync(0xffffffffu, emax);
    if (lane == 0) wmax[warp] = emax;
    __syncthreads();
    int E_s = kChunkEmpty;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) E_s = max(E_s, wmax[w]);
#pragma unroll
    for (int i = 0; i < kTilesPerWarp; ++i) {
        unsigned long long g = nested_shift(S[i], e[i], E_s);
        double sq = 0.;
        if (e[i] != kChunkEmpty && E_s - e[i] < 500) sq = (double)sqf[i] * pow2_neg2(E_s - e[i]);
        g = warp_sum_u48(g);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) { ts[warp * kTilesPerWarp + i] = g; tsq[warp * kTilesPerWarp + i] = sq; }
    }
    __syncthreads();
    const unsigned int sg = nb.sec0 + sec;   // global section number
    if (warp == 0) {
        const unsigned long long v = ts[lane];
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        double sq = tsq[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const unsigned int tile = sec * kTilesPerSection + lane;
        if (tile < num_tiles) nb.tile_pre[tile] = incl - v;
        const unsigned long long T_s = __shfl_sync(0xffffffffu, incl, 31);
        if (a.peer.world > 1 && lane == 0) {   // published in this GPU's own mailbox (tagged words); the other ranks read it from there
            unsigned long long* dst = a.peer.mail[a.peer.rank]->sec_ll[epoch & 1][sg];
            ll_write64(dst, (unsigned long long)(unsigned int)E_s, (unsigned int)epoch);
            ll_write64(dst + 2, T_s, (unsigned int)epoch);
            ll_write64(dst + 4, (unsigned long long)__double_as_longlong(sq), (unsigned int)epoch);
        }
        if (lane == 31) {
            nb.sec_E[sg] = E_s; nb.sec_T[sg] = T_s; nb.sec_sq[sg] = sq;
            __threadfence();
            is_last = (atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) { st->overflow_count = 0; st->blocks_done = 0; st->trace[5] = global_ns(); }
    if constexpr (PHASES == 3) {
        __syncthreads();
        nested_after_sections(a.peer, nb, st, epoch, a.seed, a.rt, a.n_out, a.dynamic, a.ess_threshold);
    }
}

// ---- plan pass (level 1): one block per GLOBAL section; only the sections whose slots touch this shard's slot range do
// anything.  For such a section: chunk masses G_c = S_c >> (E_s - e_c) from the chunk records (read from the GPU that owns the
// particles), their exclusive prefixes, and from those -- exactly, 128-bit once per thread -- the first output slot of every
// chunk; then per chunk the level-2 constants, and for every output tile that starts inside the section the chunk owning
// its first slot.
constexpr int kChunksPerSection = (int)(kSection / kChunk);   // 1024

template <typename Real>
__global__ void __launch_bounds__(kScanThreads) nested_plan_kernel(FixedArgs<Real> a, NestedPrefixes nb, int par, unsigned int n_chunks_global) {
    __shared__ unsigned long long wtot[kScanThreads / 32];
    __shared__ unsigned int Ps[kChunksPerSection + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DeviceStats* st = a.stats;
    pdl_wait();
    pdl_trigger();   // the expansion may become resident
    const long long epoch = a.epoch < 0 ? st->t : a.epoch;
    if (a.dynamic && !st->do_resample) {   // ESS above the threshold: keep the population, weights keep accumulating
        if (blockIdx.x == 0 && tid == 0) st->resampled_flag[epoch & 1] = 0;
        return;
    }
    if (blockIdx.x == 0 && tid == 0) { st->trace[10] = global_ns(); nested_bookkeeping(a, st, epoch); }
    const unsigned long long W = st->W;
    if (W == 0ull) return;   // degenerate: the expansion writes identity ancestors
    const unsigned int sg = blockIdx.x;
    const unsigned long long n_s = nb.sec_n[sg], a_s = nb.sec_a[sg], T_s = nb.sec_T[sg];
    const int E_s = nb.sec_E[sg];
    // this shard fills the slots [w0, w_end); the tile after its last one starts at w_end and its first chunk is needed too
    const unsigned long long w0 = a.out_base, w_end = a.out_base + a.n_out_local, w1 = w_end < a.n_out ? w_end : a.n_out - 1;
    if (a_s > w_end || a_s + n_s < w0) return;   // (closed comparison: empty sections sitting on an edge of the range are kept: P stays monotone)
    const unsigned int cg0 = sg * (unsigned int)kChunksPerSection;
    const unsigned int cnt = min((unsigned int)kChunksPerSection, n_chunks_global - cg0);   // chunks of this section
    if (n_s == 0ull || T_s == 0ull) {   // no slot: every chunk "starts" at the section's first slot (keeps P monotone)
        for (unsigned int c = tid; c <= cnt; c += kScanThreads) nb.P[cg0 + c] = (unsigned int)a_s;
        return;
    }
    // the section's chunk records, from the GPU that owns its particles
    const unsigned int owner = a.peer.world > 1 ? peer_owner(a.peer, cg0 * (unsigned int)kChunk) : 0u;
    const unsigned int c_loc0 = cg0 - owner * (a.peer.n_loc / (unsigned int)kChunk);
    const int* rec_e = a.peer.rec_e[par][owner] + c_loc0;
    const unsigned int* rec_S = a.peer.rec_S[par][owner] + c_loc0;
    if (tid == 0 && owner != (unsigned int)a.peer.rank) atomicAdd(&st->nvlink_bytes, (unsigned long long)cnt * 8ull);
    unsigned int S[4];
    unsigned long long G[4], tot = 0;
    {
        int4 ev = make_int4(kChunkEmpty, kChunkEmpty, kChunkEmpty, kChunkEmpty);
        uint4 sv = make_uint4(0u, 0u, 0u, 0u);
        if (4u * tid + 3u < cnt) { ev = *reinterpret_cast<const int4*>(rec_e + 4 * tid); sv = *reinterpret_cast<const uint4*>(rec_S + 4 * tid); }
        else {
            int e4[4] = {kChunkEmpty, kChunkEmpty, kChunkEmpty, kChunkEmpty};
            unsigned int s4[4] = {0u, 0u, 0u, 0u};
            for (int i = 0; i < 4; ++i) if (4u * tid + i < cnt) { e4[i] = rec_e[4 * tid + i]; s4[i] = rec_S[4 * tid + i]; }
            ev = make_int4(e4[0], e4[1], e4[2], e4[3]); sv = make_uint4(s4[0], s4[1], s4[2], s4[3]);
        }
        const int e4[4] = {ev.x, ev.y, ev.z, ev.w};
        S[0] = sv.x; S[1] = sv.y; S[2] = sv.z; S[3] = sv.w;
#pragma unroll
        for (int i = 0; i < 4; ++i) { G[i] = nested_shift(S[i], e4[i], E_s); tot += G[i]; }
    }
    unsigned long long incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    unsigned long long pre = incl - tot;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) if (w < warp) pre += wtot[w];
    {   // slots below the thread's first chunk (exact), then its next three relative to that
        const double inv_t = 1. / (double)T_s;
        const TileBase base = tile_base_exact(pre, T_s, nested_section_offset(st->rand_word, sg, T_s), n_s, inv_t);
        unsigned long long g = 0;
        unsigned int below = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (4u * tid + i < cnt) Ps[4 * tid + i] = (unsigned int)(a_s + base.n_start) + below;
            g += G[i];
            below = local_count(g, base.rem, (double)base.rem, T_s, (double)n_s, n_s, inv_t);
        }
    }
    if (tid == 0) Ps[cnt] = (unsigned int)(a_s + n_s);   // every slot of the section is owned by one of its chunks
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const unsigned int c = 4u * tid + i;
        if (c >= cnt) break;
        nb.P[cg0 + c] = Ps[c];
    }
    if (tid == 0) nb.P[cg0 + cnt] = Ps[cnt];   // == first slot of the next section (or the sentinel N after the last one)
    if (a_s + n_s == a.n_out) {   // the section that owns the last slot: F's entry after the last tile names the last chunk that owns a slot
        const unsigned int j = (unsigned int)(a.n_out - 1);
        if (tid == 0) {
            unsigned int c = 0;
#pragma unroll
            for (int step = kChunksPerSection / 2; step > 0; step >>= 1) if (c + step < cnt && Ps[c + step] <= j) c += step;
            nb.F[(a.n_out + kScanTile - 1) / kScanTile] = cg0 + c;
        }
    }
    // output tiles that start inside this section (and belong to this shard, or directly follow its last one)
    const unsigned long long lo = a_s > w0 ? a_s : w0, hi = a_s + n_s - 1 < w1 ? a_s + n_s - 1 : w1;   // first slots of interest, inclusive
    if (lo <= hi) {
        for (unsigned long long T = (lo + kScanTile - 1) / kScanTile + tid; T * kScanTile <= hi; T += kScanThreads) {
            const unsigned int j = (unsigned int)(T * kScanTile);
            unsigned int c = 0;   // max{c < cnt : Ps[c] <= j}
#pragma unroll
            for (int step = kChunksPerSection / 2; step > 0; step >>= 1) if (c + step < cnt && Ps[c + step] <= j) c += step;
            nb.F[T] = cg0 + c;
        }
    }
}

// ---- expansion (level 2) ---------------------------------------------------------------------------------------------------
// exclusive prefix (over the warp) of each lane's 4 particles inside chunk r, for the 4 chunks of a warp tile; lane r < 4 gets
// the total S_c of chunk r in S_w
__device__ __forceinline__ void chunk_exclusive_prefixes(const unsigned int (&q)[4][4], unsigned int (&excl)[4], unsigned int& S_w) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned int own = q[r][0] + q[r][1] + q[r][2] + q[r][3];
        unsigned int inc = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int up = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += up; }
        excl[r] = inc - own;
        const unsigned int t = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == r) S_w = t;
    }
}

// Level 2 count: floor((C * n_c + rem_c) / S_c) + 1 for C <= S_c < 2^30, n_c <= kLevel2FloatMax.  The quotient plus one half is
// estimated in fp32 as x = C (n_c / S_c) + (rem_c / S_c + 1/2), carried from particle to particle (x += q slope: at most eight
// roundings of 2^-24 relative, so x is off by less than (n_c + 2) 2^-21), and rounded to the nearest integer k by the adder (no conversion instruction).
// The quotient lies within 1/2 - |x - k| of the middle between k - 1 and k: unless that leaves less than `eps` = (n_c + 2) 2^-19
// to an integer -- where the floor could go either way -- its floor is k - 1; otherwise the exact 64-bit form is evaluated.
// The result is exact either way.
constexpr unsigned int kLevel2FloatMax = 2048u;
// slots expanded per pass by one warp: a warp tile (512 particles) owns 512 slots on average, so 512 would need a second
// pass half of the time; 768 almost never does
constexpr int kNestedWarpSlots = 768;
__device__ __forceinline__ unsigned int level2_count_exact(unsigned int k, unsigned long long y, unsigned int S_c) {   // floor(y / S_c) + 1 from the guess k
    unsigned long long p = (unsigned long long)k * S_c;
    while (p > y) { --k; p -= S_c; }
    while (p + S_c <= y) { ++k; p += S_c; }
    return k + 1u;
}

// The lane's 16 integer weights of GLOBAL warp tile `wt` (chunks 4 wt .. 4 wt + 3; round r == chunk 4 wt + r), read from the GPU
// that owns those particles, and their exclusive prefixes / chunk totals.
// RECOMPUTE: the array still holds the log-weights (ESS-triggered loop); the integer weights are re-derived from them and the
// chunk's reference e_c exactly as the quantisation did.
template <typename Real, bool RECOMPUTE, bool PULL>
__device__ __forceinline__ void nested_load_warp_tile(const FixedArgs<Real>& a, const int* rec_e_own, int par, unsigned int wt, unsigned int n_chunks_global, unsigned int (&q)[4][4]) {
    const int lane = threadIdx.x & 31;
    const unsigned int k0 = 4u * wt;
    if (k0 >= n_chunks_global) {
#pragma unroll
        for (int r = 0; r < 4; ++r) q[r][0] = q[r][1] = q[r][2] = q[r][3] = 0u;
        return;
    }
    // (a warp tile never straddles two shards: shards are whole sections)
    unsigned int k_loc0 = k0;
    size_t n_own = a.n;
    const Real* lw = a.lw;
    const int* rec_e = rec_e_own;
    if constexpr (PULL) {
        const unsigned int owner = peer_owner(a.peer, k0 * (unsigned int)kChunk);
        k_loc0 = k0 - owner * (a.peer.n_loc / (unsigned int)kChunk);
        n_own = (size_t)a.peer.n_loc;
        lw = reinterpret_cast<const Real*>(a.peer.lw[par][owner]);
        rec_e = a.peer.rec_e[par][owner];
        if (lane == 0 && owner != (unsigned int)a.peer.rank) atomicAdd(&a.stats->nvlink_bytes, (unsigned long long)(kWarpTile * sizeof(Real)));
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const size_t idx = (size_t)(k_loc0 + r) * kChunk + (size_t)lane * 4;
        if constexpr (RECOMPUTE) {
            const int e_c = k0 + r < n_chunks_global ? rec_e[k_loc0 + r] : kChunkEmpty;
            float w[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (idx < n_own) {
                if constexpr (sizeof(Real) == 4) { const float4 v = *reinterpret_cast<const float4*>(lw + idx); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
                else { const double2 u = *reinterpret_cast<const double2*>(lw + idx), v = *reinterpret_cast<const double2*>(lw + idx + 2); w[0] = (float)u.x; w[1] = (float)u.y; w[2] = (float)v.x; w[3] = (float)v.y; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float qf;
                const bool live = e_c != kChunkEmpty && idx + j < n_own;
                q[r][j] = live ? nested_weight(__fmul_rn(w[j], 1.44269504088896341f), (float)e_c, &qf) : 0u;
            }
            continue;
        }
        if (idx < n_own) {   // (chunks are quantised whole: entries past n inside the last chunk hold 0)
            if constexpr (sizeof(Real) == 4) {
                // the integers' bit patterns; last use of an array that is read once: evict-first, so that it does not push the
                // ancestors this kernel writes (the next extend reads them) out of L2.  (A peer's memory: a plain load.)
                uint4 v;
                if constexpr (PULL) v = *reinterpret_cast<const uint4*>(lw + idx);
                else v = __ldcs(reinterpret_cast<const uint4*>(lw + idx));
                q[r][0] = v.x; q[r][1] = v.y; q[r][2] = v.z; q[r][3] = v.w;
            } else {
                const double2 u = *reinterpret_cast<const double2*>(lw + idx), v = *reinterpret_cast<const double2*>(lw + idx + 2);
                q[r][0] = nested_load<double>(u.x); q[r][1] = nested_load<double>(u.y); q[r][2] = nested_load<double>(v.x); q[r][3] = nested_load<double>(v.y);
            }
        } else { q[r][0] = q[r][1] = q[r][2] = q[r][3] = 0u; }
    }
}

// Level 2 for one warp tile: inclusive offspring counts n[r][j] of the lane's 16 particles, counted from the warp tile's
// first output slot `ws` (global); returns the number of slots the warp tile owns.  P_w: lanes 0..4 hold P[4 wt .. 4 wt + 4].
__device__ __forceinline__ unsigned int nested_warp_tile_counts(unsigned int wt, unsigned int P_w, unsigned int S_w, unsigned long long word, const unsigned int (&q)[4][4],
                                                               const unsigned int (&excl)[4], unsigned int (&n)[4][4], unsigned int& ws) {
    const int lane = threadIdx.x & 31;
    ws = __shfl_sync(0xffffffffu, P_w, 0);
    // the per-chunk scalars are computed once, by lane r for chunk r (the offset hash alone is ~35 instructions), and broadcast
    const unsigned int n_own = __shfl_down_sync(0xffffffffu, P_w, 1) - P_w;
    unsigned int rem_own = 0u;
    float inv_own = 0.f;
    if (lane < 4 && S_w != 0u) {
        rem_own = S_w - (unsigned int)nested_chunk_offset(word, 4ull * wt + (unsigned int)lane, (unsigned long long)S_w) - 1u;
        inv_own = __frcp_rn((float)S_w);
    }
    unsigned int we = ws;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned int n_c = __shfl_sync(0xffffffffu, n_own, r);
        const unsigned int cb = we - ws;   // chunks own consecutive slot ranges
        const unsigned int S_c = __shfl_sync(0xffffffffu, S_w, r);
        // local slot l sits at l*S_c + U_c; particle with inclusive chunk prefix C owns the slots below C*n_c
        const unsigned int rem_c = __shfl_sync(0xffffffffu, rem_own, r);
        const float inv_sf = __shfl_sync(0xffffffffu, inv_own, r);
        if (n_c == 0u || S_c == 0u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) n[r][j] = cb;
            continue;
        }
        we += n_c;
        unsigned int C = excl[r];
        if (n_c <= kLevel2FloatMax) {   // (warp-uniform)
            const float n_cf = (float)n_c, slope = __fmul_rn(n_cf, inv_sf), icpt = fmaf((float)rem_c, inv_sf, 0.5f), safe = 0.5f - (n_cf + 2.f) * 0x1.0p-19f;
            const unsigned int cbm1 = cb - 1u;
            float x = fmaf((float)C, slope, icpt);   // (the running estimate itself is carried: one conversion and one fma per particle)
            unsigned int k[4];
            bool close = false;   // some estimate of the lane's four is too close to an integer to be floored safely (rare)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                x = fmaf((float)q[r][j], slope, x);
                const float t = __fadd_rn(x, 8388608.0f);
                const float d = __fsub_rn(x, __fsub_rn(t, 8388608.0f));   // x - k, in [-0.5, 0.5]
                k[j] = (unsigned int)__float_as_int(t) & 0x7fffffu;
                close |= !(fabsf(d) < safe);
            }
            if (close) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    C += q[r][j];
                    k[j] = level2_count_exact(k[j], (unsigned long long)C * n_c + rem_c, S_c);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) n[r][j] = cbm1 + k[j];
        } else {   // a chunk that owns thousands of slots: fp64 estimate with its own exact fallback
            const double inv_s = 1. / (double)S_c, rem_cd = (double)rem_c, n_cd = (double)n_c;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                C += q[r][j];
                n[r][j] = cb + local_count((unsigned long long)C, (unsigned long long)rem_c, rem_cd, (unsigned long long)S_c, n_cd, (unsigned long long)n_c, inv_s);
            }
        }
    }
    return we - ws;
}

// first / last GLOBAL tile of chunks (32 chunks each) that own slots of this shard's range
__device__ __forceinline__ void nested_tile_range(const NestedPrefixes& nb, unsigned long long out_base, unsigned int n_tiles_local, unsigned int n_tiles_global,
                                                  unsigned int& tg_lo, unsigned int& tg_hi) {
    const unsigned int T0 = (unsigned int)(out_base / kScanTile), T1 = T0 + n_tiles_local;
    tg_lo = nb.F[T0] / (unsigned int)kChunksPerTile;
    tg_hi = (T1 < n_tiles_global ? nb.F[T1] : nb.F[n_tiles_global]) / (unsigned int)kChunksPerTile;
}

// passes of kNestedWarpSlots slots of a warp tile's range [ws, ws + total) that can touch this shard's slots
__device__ __forceinline__ void nested_pass_range(unsigned long long out_base, unsigned long long n_out_local, unsigned int ws, unsigned int total, unsigned int& lo, unsigned int& hi) {
    const unsigned long long b = ws, e = (unsigned long long)ws + total, w0 = out_base, w1 = out_base + n_out_local;
    if (e <= w0 || b >= w1) { lo = hi = 0u; return; }
    lo = b >= w0 ? 0u : (unsigned int)((w0 - b) / kNestedWarpSlots) * kNestedWarpSlots;
    hi = e <= w1 ? total : (unsigned int)(w1 - b);
}

// One block per tile of 32 chunks, one warp per 4 chunks (512 particles); warps never meet.  The grid walks the tiles of chunks
// that own slots of THIS shard's slot range -- on one GPU simply all of them -- wherever their particles live (remote loads of the
// integer weights); every warp computes the offspring counts of its 512 particles and expands them into the ancestor slots
// that fall into the shard's range.  PULL: several GPUs.
struct NestedHeavyEntry { unsigned int wt; };
// Level 1 without a plan pass (small shards, where a kernel boundary costs more than the arithmetic): the warp derives the slot
// starts of its own 4 chunks from the section pass' tile prefix -- lane <-> chunk of the 32-chunk tile, exactly the arithmetic of
// the plan pass, hence the same numbers.  Lanes 0..4 return P[4 wt .. 4 wt + 4].
// one warp, lane <-> chunk of the 32-chunk tile `tile_g` (global): returns the tile's first slot in `start` and, per lane, the
// number of the tile's slots up to and including that lane's chunk.  false: the tile's section owns no slot (start = its a_s).
template <typename Real, bool PULL>
__device__ __forceinline__ bool nested_tile_slot_ends(const FixedArgs<Real>& a, const NestedPrefixes& nb, const ChunkRecords& rec_own, int par, unsigned int tile_g,
                                                      unsigned int n_chunks_global, unsigned long long word, unsigned int& start, unsigned int& slot_end) {
    const int lane = threadIdx.x & 31;
    const unsigned int sg = tile_g / kTilesPerSection;
    slot_end = 0u;
    if (sg >= nb.n_sec_global) { start = (unsigned int)a.n_out; return false; }
    const unsigned long long n_s = nb.sec_n[sg], a_s = nb.sec_a[sg], T_s = nb.sec_T[sg];
    start = (unsigned int)a_s;
    if (n_s == 0ull || T_s == 0ull) return false;   // no slot in this section: every chunk "starts" at the section's first slot
    const int E_s = nb.sec_E[sg];
    const unsigned int c = tile_g * kChunksPerTile + lane;   // global chunk of this lane
    unsigned int c_loc = c, t_loc = tile_g;
    const int* rec_e = rec_own.e;
    const unsigned int* rec_S = rec_own.S;
    const unsigned long long* tile_pre = nb.tile_pre;
    if constexpr (PULL) {
        const unsigned int owner = peer_owner(a.peer, tile_g * (unsigned int)kScanTile);
        c_loc = c - owner * (a.peer.n_loc / (unsigned int)kChunk);
        t_loc = tile_g - owner * (a.peer.n_loc / (unsigned int)kScanTile);
        rec_e = a.peer.rec_e[par][owner]; rec_S = a.peer.rec_S[par][owner]; tile_pre = a.peer.tile_pre[par][owner];
        if (lane == 0 && owner != (unsigned int)a.peer.rank) atomicAdd(&a.stats->nvlink_bytes, 8ull * kChunksPerTile + 8ull);
    }
    const bool valid = c < n_chunks_global;
    unsigned long long g = nested_shift(valid ? rec_S[c_loc] : 0u, valid ? rec_e[c_loc] : kChunkEmpty, E_s);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, g, o); if (lane >= o) g += up; }
    const double inv_t = 1. / (double)T_s;
    const TileBase base = tile_base_exact(tile_pre[t_loc], T_s, nested_section_offset(word, sg, T_s), n_s, inv_t);
    slot_end = local_count(g, base.rem, (double)base.rem, T_s, (double)n_s, n_s, inv_t);
    start = (unsigned int)(a_s + base.n_start);
    return true;
}

// Level 1 without a plan pass (small populations, where a kernel boundary costs more than the arithmetic): the warp derives the
// slot starts of its own 4 chunks from the section pass' tile prefix -- exactly the arithmetic of the plan pass, hence the same
// numbers.  Lanes 0..4 return P[4 wt .. 4 wt + 4].
template <typename Real, bool PULL>
__device__ __forceinline__ unsigned int nested_inline_level1(const FixedArgs<Real>& a, const NestedPrefixes& nb, const ChunkRecords& rec_own, int par, unsigned int wt,
                                                            unsigned int n_chunks_global, unsigned long long word) {
    const int lane = threadIdx.x & 31;
    unsigned int start, slot_end;
    nested_tile_slot_ends<Real, PULL>(a, nb, rec_own, par, wt / (kScanThreads / 32), n_chunks_global, word, start, slot_end);
    const unsigned int j = (wt % (kScanThreads / 32)) * 4u + (unsigned int)min(lane, 4);   // lanes 0..4: chunk j of the tile (32: one past its end)
    const unsigned int prev = __shfl_sync(0xffffffffu, slot_end, (int)max(j, 1u) - 1);
    return start + (j == 0u ? 0u : prev);
}

// ---- level-1 pass of a single GPU: one warp per tile of 32 chunks writes their slot starts (the plan pass above does the same
// for the sections a SHARD needs, wherever their particles live, and also finds the chunk range that owns the shard's slots)
template <typename Real>
__global__ void __launch_bounds__(kScanThreads) nested_level1_kernel(FixedArgs<Real> a, NestedPrefixes nb, ChunkRecords rec, unsigned int num_tiles, unsigned int n_chunks_global) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DeviceStats* st = a.stats;
    pdl_wait();
    pdl_trigger();   // the expansion may become resident and load its weights
    const long long epoch = a.epoch < 0 ? st->t : a.epoch;
    if (a.dynamic && !st->do_resample) {   // ESS above the threshold: keep the population, weights keep accumulating
        if (blockIdx.x == 0 && tid == 0) st->resampled_flag[epoch & 1] = 0;
        return;
    }
    if (blockIdx.x == 0 && tid == 0) { st->trace[10] = global_ns(); nested_bookkeeping(a, st, epoch); }
    const unsigned int tile = blockIdx.x * (kScanThreads / 32) + warp;
    if (tile >= num_tiles || st->W == 0ull) return;
    unsigned int start, slot_end;
    nested_tile_slot_ends<Real, false>(a, nb, rec, 0, tile, n_chunks_global, st->rand_word, start, slot_end);
    const unsigned int c = tile * kChunksPerTile + lane;
    const unsigned int prev = __shfl_up_sync(0xffffffffu, slot_end, 1);
    if (c <= n_chunks_global) nb.P[c] = start + (lane == 0 ? 0u : prev);
    if (lane == 31 && c + 1 <= n_chunks_global) nb.P[c + 1] = start + slot_end;   // (== the next tile's first entry; the sentinel after the last chunk)
}

// several GPUs, no plan pass: the tiles of chunks (whole sections) that own slots of this shard's range
__device__ __forceinline__ void nested_tile_range_sections(const NestedPrefixes& nb, unsigned long long w0, unsigned long long w_end, unsigned int& tg_lo, unsigned int& tg_hi) {
    const int lane = threadIdx.x & 31;
    unsigned int lo = 0xffffffffu, hi = 0u;
    for (unsigned int s = lane; s < nb.n_sec_global; s += 32) {
        const unsigned long long a_s = nb.sec_a[s], n_s = nb.sec_n[s];
        if (n_s != 0ull && a_s < w_end && a_s + n_s > w0) { lo = min(lo, s); hi = max(hi, s); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lo == 0xffffffffu) { tg_lo = 1u; tg_hi = 0u; return; }   // (no slot at all: cannot happen while W > 0)
    tg_lo = lo * kTilesPerSection; tg_hi = hi * kTilesPerSection + (kTilesPerSection - 1);
}

template <typename Real, bool PULL>
__device__ __forceinline__ void nested_expand_warp_tile(const FixedArgs<Real>& a, DeviceStats* st, unsigned short* head, unsigned int wt, unsigned int P_w /* lanes 0..4: slot starts */,
                                                        const unsigned int (&q)[4][4], const unsigned int (&excl)[4], unsigned int S_w, NestedHeavyEntry* heavy) {
    const int lane = threadIdx.x & 31;
    unsigned int n[4][4], ws;
    const unsigned int total = nested_warp_tile_counts(wt, P_w, S_w, st->rand_word, q, excl, n, ws);
    unsigned int lo = 0, hi = total;
    if constexpr (PULL) nested_pass_range(a.out_base, a.n_out_local, ws, total, lo, hi);
    if (hi <= lo) return;
    if (total > kWarpHeavyCap && a.overflow_follows) {   // a few particles own a large share of the offspring: the whole grid expands this warp tile
        if (lane == 0) heavy[atomicAdd(&st->overflow_count, 1u)] = NestedHeavyEntry{wt};
        return;
    }
    // no heavy-tile pass was launched for this step (none had been needed so far): the warp expands alone, and the raised
    // host word makes every later step launch the pass
    if (total > kWarpHeavyCap && lane == 0 && a.overflow_seen_host) *(volatile int*)a.overflow_seen_host = 1;
    const int32_t src0 = (int32_t)(wt * (unsigned int)kWarpTile) - 1;   // global id of the warp tile's first particle, minus 1
    // n[][] counts from 0 at the warp tile's first slot
    if (!PULL && total + 3u <= (unsigned int)kNestedWarpSlots) warp_expand_chunk<Real, kNestedWarpSlots, true, false>(a, head, n, 0u, total, 0u, (unsigned long long)ws, src0);
    else
        for (unsigned int chunk_lo = lo; chunk_lo < hi; chunk_lo += kNestedWarpSlots)
            warp_expand_chunk<Real, kNestedWarpSlots, false, PULL>(a, head, n, 0u, total, chunk_lo, (unsigned long long)ws, src0);
}

// lanes 0..4: the slot starts of warp tile wt's 4 chunks (and one past) -- from the plan pass' array, or derived on the spot
template <typename Real, bool PULL>
__device__ __forceinline__ unsigned int nested_slot_starts(const FixedArgs<Real>& a, const NestedPrefixes& nb, const ChunkRecords& rec_own, int par, unsigned int wt,
                                                          unsigned int n_chunks_global, unsigned long long word) {
    if (a.inline_level1) return nested_inline_level1<Real, PULL>(a, nb, rec_own, par, wt, n_chunks_global, word);
    const unsigned int c_w = 4u * wt + (unsigned int)min((int)(threadIdx.x & 31), 4);
    return c_w <= n_chunks_global ? nb.P[c_w] : (unsigned int)a.n_out;
}

template <typename Real, bool RECOMPUTE, bool PULL>
__global__ void __launch_bounds__(kScanThreads, PULL ? 3 : 4) nested_expand_kernel(FixedArgs<Real> a, NestedPrefixes nb, ChunkRecords rec_own, int par, unsigned int n_tiles_local,
                                                                                   unsigned int n_tiles_global, unsigned int n_chunks_global, NestedHeavyEntry* heavy) {
    __shared__ __align__(16) unsigned short head[kScanThreads / 32][kNestedWarpSlots];
    const int tid = threadIdx.x, warp = tid >> 5;
    DeviceStats* st = a.stats;
    unsigned int q[4][4], excl[4], S_w = 0u;
    // One GPU: tile b of chunks belongs to block b, and the integer weights come from the kernel before the section pass -- this
    // grid is only released once every block of the two small passes in between is past its own dependency wait -- so they are
    // complete and visible already: load them (and do the warp-local scans) before waiting for the plan pass' slot ranges.
    if constexpr (!PULL) {
        nested_load_warp_tile<Real, RECOMPUTE, false>(a, rec_own.e, par, blockIdx.x * (kScanThreads / 32) + warp, n_chunks_global, q);
        chunk_exclusive_prefixes(q, excl, S_w);
    }
    pdl_wait();
    pdl_trigger();
    if (blockIdx.x == 0 && tid == 0) st->trace[11] = global_ns();
    const long long epoch = a.epoch < 0 ? st->t : a.epoch;
    if (a.dynamic && !st->do_resample) {   // ESS above the threshold: keep the population, weights keep accumulating
        if (a.inline_level1 && blockIdx.x == 0 && tid == 0) st->resampled_flag[epoch & 1] = 0;   // (else the plan pass did)
        return;
    }
    if (a.inline_level1 && blockIdx.x == 0 && tid == 0) nested_bookkeeping(a, st, epoch);   // (else the plan pass did)
    if (st->W == 0ull) {   // degenerate: every weight is -inf (or NaN): identity ancestors (flagged by the bookkeeping)
        for (size_t i = (size_t)blockIdx.x * kScanThreads + tid; i < a.n; i += (size_t)gridDim.x * kScanThreads) a.anc[i] = a.src_base + (int32_t)i;
        return;
    }
    const unsigned long long word = st->rand_word;
    if constexpr (!PULL) {
        const unsigned int wt = blockIdx.x * (kScanThreads / 32) + warp;
        unsigned int P_w;
        if (a.inline_level1) {   // no level-1 pass was launched: warp 0 derives the slot starts of the block's 32 chunks, the others pick theirs up
            __shared__ unsigned int s_P[kChunksPerTile + 1];
            if (warp == 0) {
                unsigned int start, slot_end;
                nested_tile_slot_ends<Real, false>(a, nb, rec_own, par, blockIdx.x, n_chunks_global, word, start, slot_end);
                s_P[(tid & 31) + 1] = start + slot_end;
                if (tid == 0) s_P[0] = start;
            }
            __syncthreads();
            P_w = s_P[warp * 4 + min(tid & 31, 4)];
        } else P_w = nested_slot_starts<Real, false>(a, nb, rec_own, par, wt, n_chunks_global, word);
        nested_expand_warp_tile<Real, false>(a, st, head[warp], wt, P_w, q, excl, S_w, heavy);
    } else {
        unsigned int tg_lo, tg_hi;
        if (a.inline_level1) nested_tile_range_sections(nb, a.out_base, a.out_base + a.n_out_local, tg_lo, tg_hi);
        else nested_tile_range(nb, a.out_base, n_tiles_local, n_tiles_global, tg_lo, tg_hi);
        for (unsigned int tg = tg_lo + blockIdx.x; tg <= tg_hi; tg += gridDim.x) {
            const unsigned int wt = tg * (kScanThreads / 32) + warp;
            nested_load_warp_tile<Real, RECOMPUTE, true>(a, rec_own.e, par, wt, n_chunks_global, q);   // (in flight while the slot starts are worked out)
            const unsigned int P_w = nested_slot_starts<Real, true>(a, nb, rec_own, par, wt, n_chunks_global, word);
            // (whole sections are walked when there is no plan pass: most tiles of an edge section own none of this shard's slots)
            const unsigned int p_first = __shfl_sync(0xffffffffu, P_w, 0), p_last = __shfl_sync(0xffffffffu, P_w, 4);
            if (p_last <= a.out_base || p_first >= a.out_base + a.n_out_local || p_last == p_first) continue;
            chunk_exclusive_prefixes(q, excl, S_w);
            nested_expand_warp_tile<Real, true>(a, st, head[warp], wt, P_w, q, excl, S_w, heavy);
        }
    }
    if (tid == 0 && atomicAdd(&st->ticket, 1u) == gridDim.x - 1) { st->ticket = 0; st->trace[8] = global_ns(); }   // (time stamp of the last block; diagnostics)
}

// ---- heavy warp tiles: every warp of the grid recomputes the tile's counts (512 particles) and expands its share of the passes.
// Launched only once a heavy tile has been seen (host-mapped flag), like the single-level scheme's overflow pass.
template <typename Real, bool RECOMPUTE, bool PULL>
__global__ void __launch_bounds__(kScanThreads) nested_heavy_kernel(FixedArgs<Real> a, NestedPrefixes nb, ChunkRecords rec_own, int par, unsigned int n_chunks_global,
                                                                    const NestedHeavyEntry* heavy) {
    __shared__ __align__(16) unsigned short head[kScanThreads / 32][kNestedWarpSlots];
    DeviceStats* st = a.stats;
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.dynamic && !st->do_resample) return;
    const unsigned int count = st->overflow_count;
    if (count == 0u || st->W == 0ull) return;
    const unsigned int gw = blockIdx.x * (kScanThreads / 32) + warp, nw = gridDim.x * (kScanThreads / 32);
    for (unsigned int k = 0; k < count; ++k) {
        const unsigned int wt = heavy[k].wt;
        unsigned int q[4][4], excl[4], S_w = 0u, n[4][4], ws;
        nested_load_warp_tile<Real, RECOMPUTE, PULL>(a, rec_own.e, par, wt, n_chunks_global, q);
        chunk_exclusive_prefixes(q, excl, S_w);
        const unsigned int P_w = nested_slot_starts<Real, PULL>(a, nb, rec_own, par, wt, n_chunks_global, st->rand_word);
        const unsigned int total = nested_warp_tile_counts(wt, P_w, S_w, st->rand_word, q, excl, n, ws);
        unsigned int lo = 0, hi = total;
        if constexpr (PULL) nested_pass_range(a.out_base, a.n_out_local, ws, total, lo, hi);
        const int32_t src0 = (int32_t)(wt * (unsigned int)kWarpTile) - 1;
        for (unsigned long long chunk_lo = (unsigned long long)lo + (unsigned long long)gw * kNestedWarpSlots; chunk_lo < hi; chunk_lo += (unsigned long long)nw * kNestedWarpSlots)
            warp_expand_chunk<Real, kNestedWarpSlots, false, PULL>(a, head[warp], n, 0u, total, (unsigned int)chunk_lo, (unsigned long long)ws, src0);
    }
}

}  // namespace mpl
