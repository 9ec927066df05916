// nested.cuh -- nested systematic resampling on integer weights (MPL_RESAMPLE_SYSTEMATIC_NESTED).
//
// The single-level integer scheme quantises every weight against the GLOBAL maximum, so a whole pass over the
// log-weights (the reduce pass) has to wait for that maximum before it can even start.  Here every 128-particle chunk
// (one warp iteration of the extend kernel) is quantised against a power of two just above its OWN maximum:
//     e_c = ceil(max_i lw_i * log2 e),   q_i = rint(2^(lw_i log2 e - e_c) * 2^k),   S_c = sum q_i          (chunk c)
// which needs nothing global and therefore fuses into the extend kernel's epilogue.  Globally only the chunk records
// (e_c, S_c) are reduced -- 1/128 of the data:
//     E = max e_c,   G_c = S_c >> (E - e_c),   W = sum G_c
// Level 1 resamples chunks systematically by G_c (slot j at j*W + U): chunk c gets n_c consecutive slots.  Level 2
// places those n_c slots systematically inside the chunk by q_i (local slot l at l*S_c + U_c).  Both levels are exact
// integer arithmetic, unbiased (E[#offspring of i] = N q_i 2^(e_c - E) / W up to the floor in G_c), and independent of
// thread order and of how the particles are sharded (chunks are aligned groups of global ids).
#pragma once

namespace mpl {

constexpr int kChunksPerTile = kScanTile / kChunk;   // 32
constexpr int kTilesPerChunkBlock = 32;              // tiles summarised by one block of the chunk pass (one warp scan)

// level-1 prefixes: the exclusive prefix of a tile's chunk masses is blk[tile / 32] + tile_pre[tile]
struct NestedPrefixes {
    unsigned long long* tile_pre;   // exclusive prefix of the tile inside its 32-tile block
    unsigned long long* blk;        // block totals, turned into exclusive prefixes by the last block of the chunk pass
    double* blk_sq;                 // per-block sums of squared weights at the global scale (ESS)
};

// warp-wide sum of per-lane values below 2^48: two integer redux instructions instead of ten 32-bit shuffles
__device__ __forceinline__ unsigned long long warp_sum_u48(unsigned long long v) {
    const unsigned int lo = (unsigned int)v & 0xFFFFFFu, hi = (unsigned int)(v >> 24);
    return ((unsigned long long)__reduce_add_sync(0xffffffffu, hi) << 24) + (unsigned long long)__reduce_add_sync(0xffffffffu, lo);
}

// ---- stand-alone quantisation pass (call-per-step API; the device-resident loop fuses this into the extend kernel) -----
template <typename Real>
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
        float qv[4], sq;
        int e_c;
        unsigned long long S_c;
        warp_quantise_chunk(w, a.kbits, qv, e_c, S_c, sq);
        if constexpr (sizeof(Real) == 4) *reinterpret_cast<float4*>(a.lw + idx) = make_float4(qv[0], qv[1], qv[2], qv[3]);
        else { *reinterpret_cast<double2*>(a.lw + idx) = make_double2(qv[0], qv[1]); *reinterpret_cast<double2*>(a.lw + idx + 2) = make_double2(qv[2], qv[3]); }
        if (lane == 0) { rec.e[chunk] = e_c; rec.S[chunk] = S_c; rec.sq[chunk] = sq; }
    }
    pdl_trigger();
}

__device__ __forceinline__ int nested_global_exp(float mx) { return (int)ceilf(__fmul_rn(mx, 1.44269504088896341f)); }
__device__ __forceinline__ unsigned long long nested_chunk_mass(int e_c, unsigned long long S_c, int E) {
    if (e_c == kChunkEmpty) return 0ull;
    const int s = E - e_c;
    return s < 64 ? (S_c >> s) : 0ull;
}

// ---- chunk pass: chunk masses at the global scale -> tile sums -> (per block) tile prefixes + block total; the last block
// turns the block totals into prefixes and publishes W.  One warp per 4 tiles (32 chunks of a tile <-> the 32 lanes).
template <typename Real>
__global__ void __launch_bounds__(kScanThreads) nested_chunk_kernel(FixedArgs<Real> a, ChunkRecords rec, NestedPrefixes nb, unsigned int num_tiles,
                                                                    unsigned int num_chunks) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned long long ts[kTilesPerChunkBlock];
    __shared__ double tsq[kTilesPerChunkBlock];
    __shared__ unsigned long long ws[kScanThreads / 32];
    __shared__ double wsq[kScanThreads / 32];
    __shared__ unsigned long long carry_s;
    __shared__ bool is_last;
    pdl_wait();
    pdl_trigger();   // the expansion kernel may become resident now: it loads its weights while this grid runs
    constexpr int kTilesPerWarp = kTilesPerChunkBlock / (kScanThreads / 32);
    const unsigned int tile0 = blockIdx.x * kTilesPerChunkBlock + warp * kTilesPerWarp;
    int e[kTilesPerWarp];
    unsigned long long S[kTilesPerWarp];
    float sqf[kTilesPerWarp];
#pragma unroll
    for (int i = 0; i < kTilesPerWarp; ++i) {
        const unsigned int c = (tile0 + i) * kChunksPerTile + lane;
        const bool valid = c < num_chunks;
        e[i] = valid ? rec.e[c] : kChunkEmpty;
        S[i] = valid ? rec.S[c] : 0ull;
        sqf[i] = valid ? rec.sq[c] : 0.f;
    }
    gate_stats(a.peer, a.stats, a.epoch < 0 ? a.stats->t : a.epoch);
    const int E = nested_global_exp(fixed_max<Real>(a));
#pragma unroll
    for (int i = 0; i < kTilesPerWarp; ++i) {
        unsigned long long g = nested_chunk_mass(e[i], S[i], E);
        double sq = 0.;
        if (e[i] != kChunkEmpty && E - e[i] < 500) sq = (double)sqf[i] * exp2(-2. * (double)(E - e[i]));
        g = warp_sum_u48(g);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) { ts[warp * kTilesPerWarp + i] = g; tsq[warp * kTilesPerWarp + i] = sq; }
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned long long v = ts[lane];
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        double sq = tsq[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const unsigned int tile = blockIdx.x * kTilesPerChunkBlock + lane;
        if (tile < num_tiles) nb.tile_pre[tile] = incl - v;
        if (lane == 31) {
            nb.blk[blockIdx.x] = incl; nb.blk_sq[blockIdx.x] = sq;
            __threadfence();
            is_last = (atomicAdd(&a.stats->blocks_done, 1u) == gridDim.x - 1);
            carry_s = 0ull;
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const unsigned int nblk = gridDim.x;
    double sqt = 0.;
    for (unsigned int base = 0; base < nblk; base += kScanThreads * 4) {   // exclusive scan of the block totals, in place
        unsigned long long v[4], tot = 0;
        const unsigned int first = base + tid * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = (first + i < nblk) ? __ldcg(nb.blk + first + i) : 0ull; tot += v[i];
            if (first + i < nblk) sqt += __ldcg(nb.blk_sq + first + i);
        }
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
        for (int i = 0; i < 4; ++i) { if (first + i < nblk) nb.blk[first + i] = pre; pre += v[i]; }
        __syncthreads();
        if (tid == 0) carry_s += all;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sqt += __shfl_xor_sync(0xffffffffu, sqt, o);
    if (lane == 0) wsq[warp] = sqt;
    __syncthreads();
    if (tid == 0) {
        DeviceStats* st = a.stats;
        st->overflow_count = 0; st->blocks_done = 0;
        double sq = 0.;
        for (int i = 0; i < kScanThreads / 32; ++i) sq += wsq[i];
        const unsigned long long tot = carry_s;
        if (a.peer.world <= 1) {
            st->W = tot; st->c_offset = 0;
            st->sumexp2 = sq;
            st->ess = sq > 0. ? ((double)tot * (double)tot) / sq : 0.;
        } else {   // the shard's total mass and squared sum go to every rank (the expansion kernel's gate adds them up)
            const long long epoch = a.epoch < 0 ? st->t : a.epoch;
            st->trace[5] = global_ns();
            for (int h = 0; h < a.peer.world; ++h) {
                ll_write64(a.peer.mail[h]->w_ll[a.peer.rank], tot, (unsigned int)epoch);
                ll_write64(a.peer.mail[h]->w_ll[a.peer.rank] + 2, (unsigned long long)__double_as_longlong(sq), (unsigned int)epoch);
            }
        }
    }
}

// ---- expansion: one block per tile, one warp per 4 chunks; no block barrier ---------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(kScanThreads, 4) nested_scan_kernel(FixedArgs<Real> a, ChunkRecords rec, NestedPrefixes nb, unsigned int num_tiles,
                                                                      unsigned int num_chunks) {
    __shared__ __align__(16) unsigned short head[kScanThreads / 32][kWarpChunk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DeviceStats* st = a.stats;
    const unsigned int tile = blockIdx.x;
    // The integer weights and chunk records come from the kernel BEFORE the chunk pass, and this grid is only released once
    // every block of the chunk pass is past its own dependency wait -- so they are complete and visible already: load them
    // (and do the warp-local scans) before waiting for the chunk pass' prefixes.
    const size_t wt_base = (size_t)tile * kScanTile + (size_t)warp * kWarpTile;
    // integer weights stay in their float form (exact: at most 24 significant bits) to keep registers free
    float qf[4][4];
    unsigned long long excl[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const size_t idx = wt_base + (size_t)r * kChunk + (size_t)lane * 4;
        if (idx < a.n) {   // (chunks are quantised whole: entries past n inside the last chunk hold 0)
            if constexpr (sizeof(Real) == 4) {
                float4 v = __ldcs(reinterpret_cast<const float4*>(a.lw + idx));   // last use
                qf[r][0] = v.x; qf[r][1] = v.y; qf[r][2] = v.z; qf[r][3] = v.w;
            } else {
                double2 u = *reinterpret_cast<const double2*>(a.lw + idx), v = *reinterpret_cast<const double2*>(a.lw + idx + 2);
                qf[r][0] = (float)u.x; qf[r][1] = (float)u.y; qf[r][2] = (float)v.x; qf[r][3] = (float)v.y;
            }
        } else { qf[r][0] = qf[r][1] = qf[r][2] = qf[r][3] = 0.f; }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {   // exclusive prefix of the lane's 4 particles inside chunk r
        const unsigned long long own = __float2ull_rz(qf[r][0]) + __float2ull_rz(qf[r][1]) + __float2ull_rz(qf[r][2]) + __float2ull_rz(qf[r][3]);
        unsigned long long inc = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += up; }
        excl[r] = inc - own;
    }
    const unsigned int c_l = tile * kChunksPerTile + lane;   // level 1: lane l <-> chunk l of the tile
    unsigned long long S_l = 0;
    int e_l = kChunkEmpty;
    if (c_l < num_chunks) { S_l = rec.S[c_l]; e_l = rec.e[c_l]; }
    pdl_wait();
    const long long epoch = a.epoch < 0 ? st->t : a.epoch;
    gate_weights(a.peer, st, epoch, 0, 0.);
    const unsigned long long W = st->W;
    if (W == 0ull) {   // degenerate: every weight is -inf (or NaN): identity ancestors, flagged
        for (size_t i = (size_t)tile * kScanTile + tid; i < min((size_t)(tile + 1) * kScanTile, a.n); i += kScanThreads) a.anc[i] = a.src_base + (int32_t)i;
        if (tile == 0 && tid == 0) { st->degenerate = 1; st->lse = -INFINITY; st->resampled = 1; st->resampled_flag[epoch & 1] = 1; }
        return;
    }
    const int E = nested_global_exp(fixed_max<Real>(a));
    const double inv_w = 1. / (double)W;
    const unsigned long long word = resample_rand_word(a.seed, a.rt, st);
    // every warp derives the tile's exact slot base itself
    const unsigned long long U = __umul64hi(word, W);
    const TileBase base = tile_base_exact(st->c_offset + nb.blk[tile / kTilesPerChunkBlock] + nb.tile_pre[tile], W, U, a.n_out, inv_w);
    // masses of the tile's 32 chunks, inclusive prefix, slot offsets of the chunk boundaries
    unsigned long long gi = nested_chunk_mass(e_l, S_l, E);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long up = __shfl_up_sync(0xffffffffu, gi, o); if (lane >= o) gi += up; }
    const double rem_d = (double)base.rem, n_out_d = (double)a.n_out;
    const unsigned int slot_end_l = local_count(gi, base.rem, rem_d, W, n_out_d, a.n_out, inv_w);        // slots of the tile up to and including chunk l
    unsigned int slot_beg_l = __shfl_up_sync(0xffffffffu, slot_end_l, 1);
    if (lane == 0) slot_beg_l = 0;
    if (tile == 0 && tid == 0) {   // scalar bookkeeping of resample(): particle_filter.rs:104-105,114
        double lse = (double)E * 0.6931471805599453 + log((double)W) - (double)a.kbits * 0.6931471805599453;
        st->lse = lse;
        st->ess_stale = st->ess;
        if (a.accumulate_lml) st->lml_acc += lse - a.log_n_global;
        st->resampled = 1;
        st->resampled_flag[epoch & 1] = 1;
        st->n_resamples += 1;
        st->degenerate = 0;
    }
    pdl_trigger();
    // this warp's 4 chunks: 4*warp .. 4*warp + 3  (round r of the lane's 16 particles == chunk 4*warp + r)
    unsigned int n[4][4];
    const unsigned int ws = __shfl_sync(0xffffffffu, slot_beg_l, 4 * warp);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned int cb = __shfl_sync(0xffffffffu, slot_beg_l, 4 * warp + r), ce = __shfl_sync(0xffffffffu, slot_end_l, 4 * warp + r);
        const unsigned long long S_c = __shfl_sync(0xffffffffu, S_l, 4 * warp + r);
        const unsigned int n_c = ce - cb;
        if (n_c == 0u || S_c == 0ull) {
#pragma unroll
            for (int j = 0; j < 4; ++j) n[r][j] = cb;
            continue;
        }
        // level 2: local slot l sits at l*S_c + U_c; particle with inclusive chunk prefix C owns the slots below C*n_c
        const unsigned long long chunk_gid = ((unsigned long long)a.out_base / kChunk) + (unsigned long long)tile * kChunksPerTile + 4 * warp + r;
        const unsigned long long U_c = __umul64hi(splitmix64_mix(word + (chunk_gid + 1ull) * 0x9E3779B97F4A7C15ull), S_c);
        const unsigned long long rem_c = S_c - U_c - 1ull;
        const double inv_s = 1. / (double)S_c, rem_cd = (double)rem_c, n_cd = (double)n_c;
        unsigned long long C = excl[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            C += __float2ull_rz(qf[r][j]);
            n[r][j] = cb + local_count(C, rem_c, rem_cd, S_c, n_cd, (unsigned long long)n_c, inv_s);
        }
    }
    const unsigned int we = __shfl_sync(0xffffffffu, n[3][3], 31);
    const unsigned int total = we - ws;
    if (total == 0u) return;
    if (total > kWarpHeavyCap && lane == 0 && a.overflow_seen_host) *(volatile int*)a.overflow_seen_host = 1;
    // (heavy warp tiles are expanded by the owning warp alone in this scheme)
    const int32_t src0 = a.src_base + (int32_t)(tile * (unsigned int)kScanTile + warp * kWarpTile) - 1;
    for (unsigned int chunk_lo = 0; chunk_lo < total; chunk_lo += kWarpChunk)
        warp_expand_chunk<Real>(a, head[warp], n, ws, total, chunk_lo, base.n_start + ws, src0);
}

}  // namespace mpl
