// scan2.cuh -- integer-weight systematic resampling, warp-granular version (R2).
//
// One block per 4096-particle tile (the granularity of the reduce pass' prefixes), but every warp owns a contiguous
// 512-particle sub-tile and does everything after the single block barrier on its own: offspring counts in registers,
// expansion of its own slot range through a warp-private shared-memory buffer with __syncwarp only.  Compared with the
// block-cooperative version this removes four of the five block barriers per tile and the binary search for the carry.
#pragma once

namespace mpl {

constexpr int kWarpTile = 512;       // particles per warp (16 per lane: 4 rounds of 4 consecutive)
constexpr int kWarpChunk = 512;      // slots expanded per pass by one warp
constexpr unsigned int kWarpHeavyCap = 64u * kWarpChunk;   // warp tiles owning more offspring go to the whole-grid pass

struct OverflowEntry2 {
    unsigned long long rem, n_start, wp;   // tile base (exact) and the warp tile's prefix inside the tile
    unsigned int tile, warp, ws, total;
};

struct Scan2Shared {
    unsigned long long warp_tot[kScanThreads / 32];
    TileBase base;
    __align__(16) unsigned short head[kScanThreads / 32][kWarpChunk];
};

// Loads the lane's 16 integer weights of a warp tile (STORED: left in place of the log-weights by the reduce pass; else
// re-quantised from the log-weights, for ESS-triggered runs where the log-weights must survive a skipped resample); returns per-round lane sums' inclusive scan over lanes
// (incl[r]), the lane's own round sums (own[r]) and the warp-wide round totals (tot[r]).
template <typename Real, bool STORED>
__device__ __forceinline__ void warp_tile_load_scan(const FixedArgs<Real>& a, size_t wt_base, float mx, unsigned long long (&q)[4][4],
                                                    unsigned long long (&incl)[4], unsigned long long (&own)[4], unsigned long long (&tot)[4]) {
    const int lane = threadIdx.x & 31;
    if (wt_base + kWarpTile <= a.n) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if constexpr (STORED) load_stored_q4<Real, true>(a.lw, wt_base + (size_t)r * 128 + (size_t)lane * 4, a.n, q[r]);
            else load_q4<Real, true>(a.lw, wt_base + (size_t)r * 128 + (size_t)lane * 4, a.n, mx, a.kbits, q[r]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            size_t idx = wt_base + (size_t)r * 128 + (size_t)lane * 4;
            if (idx < a.n) { if constexpr (STORED) load_stored_q4<Real, false>(a.lw, idx, a.n, q[r]); else load_q4<Real, false>(a.lw, idx, a.n, mx, a.kbits, q[r]); }
            else { q[r][0] = q[r][1] = q[r][2] = q[r][3] = 0ull; }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) { own[r] = q[r][0] + q[r][1] + q[r][2] + q[r][3]; incl[r] = own[r]; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            unsigned long long up = __shfl_up_sync(0xffffffffu, incl[r], o);
            if (lane >= o) incl[r] += up;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) tot[r] = __shfl_sync(0xffffffffu, incl[r], 31);
}

// Offspring counts of the lane's 16 particles (inclusive, relative to the tile's first slot) from the tile-local prefixes.
__device__ __forceinline__ void warp_tile_counts(unsigned long long wp, const unsigned long long (&q)[4][4], const unsigned long long (&incl)[4],
                                                 const unsigned long long (&own)[4], const unsigned long long (&tot)[4], const TileBase& base,
                                                 unsigned long long W, double inv_w, unsigned long long n_out, unsigned int (&n)[4][4]) {
    const double rem_d = (double)base.rem, n_out_d = (double)n_out;
    unsigned long long round_base = wp;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        unsigned long long c = round_base + incl[r] - own[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c += q[r][j];
            n[r][j] = local_count(c, base.rem, rem_d, W, n_out_d, n_out, inv_w);
        }
        round_base += tot[r];
    }
}

// One warp expands slots [chunk_lo, chunk_lo + CAP) of its own range [0, total) (relative to ws).  CAP = 32 * (slots per
// lane), a multiple of 256: `head` holds CAP 16-bit entries.
// SINGLE: one GPU; the caller guarantees chunk_lo == 0 and total + 3 <= CAP (the whole range in one pass): the window tests drop
// out, and the window is laid out from the 16-byte boundary at or below the first slot, so that the ancestors leave as 128-bit
// stores (4 slots per lane and instruction).
// CLIP: pull organisation (nested scheme on several GPUs) -- this shard fills only its OWN slot range, wherever the parents live:
// slots outside [out_base, out_base + n_out_local) are some other GPU's to fill and are dropped here.
template <typename Real, int CAP = kWarpChunk, bool SINGLE = false, bool CLIP = false>
__device__ __forceinline__ bool warp_expand_chunk(const FixedArgs<Real>& a, unsigned short* head, const unsigned int (&n)[4][4], unsigned int ws,
                                                  unsigned int total, unsigned int chunk_lo, unsigned long long slot_base /* global slot of ws */,
                                                  int32_t src0 /* value for local element 0, minus 1 */) {
    constexpr int PER_LANE = CAP / 32, VEC = PER_LANE / 8;   // uint4 = 8 entries
    static_assert(CAP % 256 == 0, "CAP");
    const int lane = threadIdx.x & 31;
    uint4* head4 = reinterpret_cast<uint4*>(head);
#pragma unroll
    for (int k = 0; k < VEC; ++k) head4[lane * VEC + k] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    unsigned int shift = 0u;   // SINGLE: head[] position of the first slot (its index in the output array, modulo 4)
    if constexpr (SINGLE) shift = (unsigned int)((slot_base - a.out_base) & 3ull);
    const unsigned int org = ws - shift;
    // run heads: element e (order r, lane, j) owns relative slots [n_prev - ws, n_e - ws)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        unsigned int prev = __shfl_up_sync(0xffffffffu, n[r][3], 1);
        unsigned int last_prev_round = __shfl_sync(0xffffffffu, n[(r + 3) & 3][3], 31);   // lane 31 of the previous round
        if (lane == 0) prev = (r == 0) ? ws : last_prev_round;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned int start = prev - org, end = n[r][j] - org;
            if constexpr (SINGLE) {
                if (end > start) head[start] = (unsigned short)(r * 128 + lane * 4 + j + 1);
            } else {
                if (end > start && start < chunk_lo + CAP && end > chunk_lo)
                    head[max(start, chunk_lo) - chunk_lo] = (unsigned short)(r * 128 + lane * 4 + j + 1);
            }
            prev = n[r][j];
        }
    }
    __syncwarp();
    // max-scan: lane owns slots [PER_LANE*lane, PER_LANE*lane + PER_LANE) of the chunk, two per 32-bit word; the scan runs on the
    // packed words (VIMNMX.U16x2): R[w] = running maxima over the even (low half) and over the odd (high half) slots up to word w
    constexpr int NW = PER_LANE / 2;
    unsigned int R[NW];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const uint4 h = head4[lane * VEC + k];
        R[4 * k + 0] = h.x; R[4 * k + 1] = h.y; R[4 * k + 2] = h.z; R[4 * k + 3] = h.w;
    }
#pragma unroll
    for (int w = 1; w < NW; ++w) R[w] = __vmaxu2(R[w], R[w - 1]);
    unsigned int incl = max(R[NW - 1] & 0xffffu, R[NW - 1] >> 16);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned int up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl = max(incl, up); }
    unsigned int pre = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) pre = 0;
    const unsigned int pre2 = pre | (pre << 16);
    // slot 2w: max(pre, evens up to w, odds up to w - 1); slot 2w + 1: max(pre, evens up to w, odds up to w)
#pragma unroll
    for (int w = NW - 1; w >= 0; --w)
        R[w] = __vimax3_u16x2(__byte_perm(R[w], 0u, 0x1010), __byte_perm(w ? R[w - 1] : 0u, R[w], 0x7632), pre2);
#pragma unroll
    for (int k = 0; k < VEC; ++k) head4[lane * VEC + k] = make_uint4(R[4 * k + 0], R[4 * k + 1], R[4 * k + 2], R[4 * k + 3]);
    __syncwarp();
    const unsigned int valid = min((unsigned int)CAP, total - chunk_lo);
    const unsigned long long slot0 = slot_base + chunk_lo;
    if constexpr (CLIP) {
        const unsigned long long rel0 = slot0 - a.out_base;   // (wraps for slots below the range: then rel0 + o >= n_out_local too, the totals being < 2^32)
#pragma unroll
        for (int k = 0; k < CAP / 32; ++k) {
            const unsigned int o = k * 32 + lane;
            const unsigned long long rel = rel0 + o;
            if (o < valid && rel < a.n_out_local) a.anc[rel] = src0 + (int32_t)head[o];
        }
        __syncwarp();
        return false;
    }
    if constexpr (SINGLE) {
        const unsigned int end_p = total + shift;
        int32_t* dst = a.anc + (slot0 - a.out_base) - shift;   // 16-byte aligned
        const uint2* head2 = reinterpret_cast<const uint2*>(head);
#pragma unroll
        for (int k = 0; k < CAP / 128; ++k) {
            const unsigned int p0 = 4u * (k * 32 + lane);
            if (p0 < end_p) {
                const uint2 h = head2[k * 32 + lane];
                const int4 v = make_int4(src0 + (int32_t)(h.x & 0xffffu), src0 + (int32_t)(h.x >> 16), src0 + (int32_t)(h.y & 0xffffu), src0 + (int32_t)(h.y >> 16));
                if (p0 >= shift && p0 + 4u <= end_p) *reinterpret_cast<int4*>(dst + p0) = v;
                else {   // the first and the last group of the range
                    if (p0 >= shift) dst[p0] = v.x;
                    if (p0 + 1u >= shift && p0 + 1u < end_p) dst[p0 + 1] = v.y;
                    if (p0 + 2u >= shift && p0 + 2u < end_p) dst[p0 + 2] = v.z;
                    if (p0 + 3u < end_p) dst[p0 + 3] = v.w;
                }
            }
        }
        __syncwarp();
        return false;
    }
    const bool local = slot0 >= a.out_base && slot0 + valid <= a.out_base + a.n_out_local;   // whole chunk lands in this shard (always, on one GPU)
    if (local) {
        int32_t* dst = a.anc + (slot0 - a.out_base);
#pragma unroll
        for (int k = 0; k < CAP / 32; ++k) {
            unsigned int o = k * 32 + lane;
            if (o < valid) dst[o] = src0 + (int32_t)head[o];
        }
    } else {
#pragma unroll
        for (int k = 0; k < CAP / 32; ++k) {   // slots of other shards: remote stores into the owner's array
            unsigned int o = k * 32 + lane;
            if (o < valid) {
                unsigned int slot = (unsigned int)(slot0 + o);
                unsigned int rk = peer_owner(a.peer, slot);
                a.peer.anc[rk][slot - rk * a.peer.n_loc] = src0 + (int32_t)head[o];
            }
        }
    }
    __syncwarp();
    return !local;   // stores went to another GPU's array
}

template <typename Real, bool STORED>
__global__ void __launch_bounds__(kScanThreads, 4) fixed_scan2_kernel(FixedArgs<Real> a, unsigned int num_tiles, OverflowEntry2* overflow) {
    __shared__ Scan2Shared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DeviceStats* st = a.stats;
    const unsigned int tile = blockIdx.x;
    pdl_wait();
    const long long epoch = a.epoch < 0 ? st->t : a.epoch;
    gate_weights(a.peer, st, epoch, a.dynamic, a.ess_threshold);
    if (a.dynamic && !st->do_resample) {   // ESS above the threshold: keep the population, weights keep accumulating
        if (tid == 0) a.desc[tile] = 0ull;
        if (tile == 0 && tid == 0) st->resampled_flag[epoch & 1] = 0;
        return;
    }
    const unsigned long long W = st->W;
    const float mx = fixed_max<Real>(a);
    if (W == 0ull) {   // degenerate: every weight is -inf (or NaN): identity ancestors, flagged
        for (size_t i = (size_t)tile * kScanTile + tid; i < min((size_t)(tile + 1) * kScanTile, a.n); i += kScanThreads) a.anc[i] = a.src_base + (int32_t)i;
        if (tile == 0 && tid == 0) { st->degenerate = 1; st->lse = -INFINITY; st->resampled = 1; st->resampled_flag[epoch & 1] = 1; }
        if (tid == 0) a.desc[tile] = 0ull;
        return;
    }
    const double inv_w = 1. / (double)W;
    if (tid == 0) {   // exact slot base of this tile from its exclusive prefix (reduce pass) and the shard's offset
        const unsigned long long U = __umul64hi(resample_rand_word(a.seed, a.rt, st), W);
        sh.base = tile_base_exact(st->c_offset + a.desc[tile], W, U, a.n_out, inv_w);
        a.desc[tile] = 0ull;   // ready for the next reduce pass
    }
    const size_t wt_base = (size_t)tile * kScanTile + (size_t)warp * kWarpTile;
    unsigned long long q[4][4], incl[4], own[4], tot[4];
    warp_tile_load_scan<Real, STORED>(a, wt_base, mx, q, incl, own, tot);
    if (lane == 0) sh.warp_tot[warp] = tot[0] + tot[1] + tot[2] + tot[3];
    __syncthreads();   // the only block barrier: warp totals and the tile base
    pdl_trigger();
    unsigned long long wp = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) if (w < warp) wp += sh.warp_tot[w];
    const TileBase base = sh.base;
    if (tile == 0 && tid == 0) {   // scalar bookkeeping of resample(): particle_filter.rs:104-105,114
        double lse = (double)mx + log((double)W) - (double)a.kbits * 0.6931471805599453;
        st->lse = lse;
        st->ess_stale = st->ess;
        if (a.accumulate_lml) st->lml_acc += lse - a.log_n_global;
        st->resampled = 1;
        st->resampled_flag[epoch & 1] = 1;
        st->n_resamples += 1;
        st->degenerate = 0;
    }
    unsigned int n[4][4];
    warp_tile_counts(wp, q, incl, own, tot, base, W, inv_w, a.n_out, n);
    const unsigned int ws = local_count(wp, base.rem, (double)base.rem, W, (double)a.n_out, a.n_out, inv_w);
    const unsigned int we = __shfl_sync(0xffffffffu, n[3][3], 31);
    const unsigned int total = we - ws;
    if (total == 0u) return;
    if (total > kWarpHeavyCap) {   // a few particles own a large share of the offspring
        if (lane == 0 && a.overflow_seen_host) *(volatile int*)a.overflow_seen_host = 1;
        if (a.overflow_follows) {   // queue the warp tile for the whole-grid pass
            if (lane == 0) {
                unsigned int slot = atomicAdd(&st->overflow_count, 1u);
                overflow[slot] = OverflowEntry2{base.rem, base.n_start, wp, tile, (unsigned int)warp, ws, total};
            }
            return;
        }
        // no overflow pass was launched for this step (none had been needed so far): this warp does it alone, and the
        // raised host word makes every later step launch the pass
    }
    const int32_t src0 = a.src_base + (int32_t)(tile * (unsigned int)kScanTile + warp * kWarpTile) - 1;
    for (unsigned int chunk_lo = 0; chunk_lo < total; chunk_lo += kWarpChunk)
        warp_expand_chunk<Real>(a, sh.head[warp], n, ws, total, chunk_lo, base.n_start + ws, src0);
}

// heavy warp tiles: every warp of the grid recomputes the tile's counts (512 particles) and expands its share of chunks
template <typename Real, bool STORED>
__global__ void __launch_bounds__(kScanThreads) fixed_overflow2_kernel(FixedArgs<Real> a, const OverflowEntry2* overflow) {
    __shared__ Scan2Shared sh;
    DeviceStats* st = a.stats;
    pdl_wait();
    pdl_trigger();
    if (a.dynamic && !st->do_resample) return;
    const unsigned int count = st->overflow_count;
    if (count == 0u && a.peer.world <= 1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (count != 0u) {
        const unsigned long long W = st->W;
        const float mx = fixed_max<Real>(a);
        const double inv_w = 1. / (double)W;
        const unsigned int gw = blockIdx.x * (kScanThreads / 32) + warp, nw = gridDim.x * (kScanThreads / 32);
        for (unsigned int k = 0; k < count; ++k) {
            const OverflowEntry2 e = overflow[k];
            const size_t wt_base = (size_t)e.tile * kScanTile + (size_t)e.warp * kWarpTile;
            unsigned long long q[4][4], incl[4], own[4], tot[4];
            warp_tile_load_scan<Real, STORED>(a, wt_base, mx, q, incl, own, tot);
            unsigned int n[4][4];
            warp_tile_counts(e.wp, q, incl, own, tot, TileBase{e.n_start, e.rem}, W, inv_w, a.n_out, n);
            const int32_t src0 = a.src_base + (int32_t)(e.tile * (unsigned int)kScanTile + e.warp * kWarpTile) - 1;
            for (unsigned long long chunk_lo = (unsigned long long)gw * kWarpChunk; chunk_lo < e.total; chunk_lo += (unsigned long long)nw * kWarpChunk)
                warp_expand_chunk<Real>(a, sh.head[warp], n, e.ws, e.total, (unsigned int)chunk_lo, e.n_start + e.ws, src0);
        }
    }
    if (a.peer.world > 1) {   // every ancestor this shard owes anybody is written -> tell every rank
        bool signal = false;
        if (count == 0u) signal = (blockIdx.x == 0 && threadIdx.x == 0);
        else {
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence_system();
                if (atomicAdd(&st->ticket, 1u) == gridDim.x - 1) { st->ticket = 0; signal = true; }
            }
        }
        if (signal) {
            st->trace[8] = global_ns();
            const long long epoch = a.epoch < 0 ? st->t : a.epoch;
            for (int h = 0; h < a.peer.world; ++h) {
                if (count == 0u) *(volatile long long*)&a.peer.mail[h]->flag_done[a.peer.rank] = epoch;
                else st_release_sys(&a.peer.mail[h]->flag_done[a.peer.rank], epoch);
            }
        }
    }
}

}  // namespace mpl
