// pf.cu -- ParticleSystem host driver + C ABI (reference modppl/src/inference/particle_filter.rs:8-121).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include "engine.h"
#include "cumsum_exact.cuh"

namespace mpl {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

int model_from_name(const std::string& name) {
    if (name == "lgssm4") return M_LGSSM4;
    if (name == "spiral") return M_SPIRAL;
    if (name == "sv") return M_SV;
    if (name == "hmm") return M_HMM;
    if (name == "line") return M_LINE;
    if (name == "hierarchical") return M_HIER;
    if (name == "pointed") return M_POINTED;
    return -1;
}

// ---- device functor construction from the parameter vector ---------------------------------------------
template <typename Real> Lgssm4<Real> make_lgssm4(const mpl_model& m) {
    const auto& p = m.params;
    Lgssm4<Real> f;
    f.q = (Real)(p.size() > 0 ? p[0] : 0.1);
    f.r = (Real)(p.size() > 1 ? p[1] : 0.5);
    f.x0 = (Real)(p.size() > 2 ? p[2] : 1.0);
    f.ln_r = (Real)std::log((double)f.r);
    f.inv_r = (Real)1 / f.r;
    f.lw_const = (Real)(1.8378770664093453 + 2. * std::log((double)f.r));
    return f;
}
template <typename Real> Spiral<Real> make_spiral(const mpl_model& m) {
    const auto& p = m.params;
    Spiral<Real> f;
    double dr = p.size() > 0 ? p[0] : 0.1, dm = p.size() > 1 ? p[1] : 0.4, ds = p.size() > 2 ? p[2] : 0.2, ov = p.size() > 3 ? p[3] : 0.001;
    f.dr_std = (Real)dr; f.dth_mean = (Real)dm; f.dth_std = (Real)ds;
    // nalgebra 2x2 closed forms (mvnormal.rs:17-18), hoisted
    double det = ov * ov - 0. * 0.;
    f.prec[0] = ov / det; f.prec[1] = -0. / det; f.prec[2] = -0. / det; f.prec[3] = ov / det;
    f.log_norm = 2. * std::log(2. * kPi) + std::log(det);
    return f;
}
template <typename Real> StochVol<Real> make_sv(const mpl_model& m) {
    const auto& p = m.params;
    StochVol<Real> f;
    f.mu = (Real)(p.size() > 0 ? p[0] : -1.024);
    f.phi = (Real)(p.size() > 1 ? p[1] : 0.9702);
    f.sig = (Real)(p.size() > 2 ? p[2] : 0.178);
    f.sd0 = f.sig / std::sqrt(1 - f.phi * f.phi);
    return f;
}
template <typename Real> Hmm<Real> make_hmm(const mpl_model& m) {
    const auto& p = m.params;
    Hmm<Real> f;
    std::memset(&f, 0, sizeof f);
    f.K = (int)p[0]; f.M = (int)p[1];
    for (int k = 0; k < f.K; ++k) f.prior[k] = p[2 + k];
    for (int i = 0; i < f.M * f.K; ++i) f.log_emis[i] = std::log(p[2 + f.K + i]);
    for (int i = 0; i < f.K * f.K; ++i) f.trans[i] = p[2 + f.K + f.M * f.K + i];
    return f;
}

// ---- launches with programmatic stream serialisation (the kernels call pdl_wait() first thing) -------------------
static bool g_use_pdl = getenv("MPL_NO_PDL") == nullptr;
static int g_inline_level1 = -1;   // nested scheme: -1 = by shard size, 0 = always a plan pass, 1 = never (mpl_test_set_inline_level1)
template <typename... KArgs, typename... Args>
static cudaError_t pdl_launch(void (*kernel)(KArgs...), unsigned int grid, unsigned int block, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- launch bookkeeping ----------------------------------------------------------------------------------
struct ScopedLaunch {
    mpl_ps* ps; const char* name; cudaEvent_t e0 = nullptr, e1 = nullptr;
    ScopedLaunch(mpl_ps* p, const char* n) : ps(p), name(n) {
        ps->launch_count++;
        if (ps->profile) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ps->stream); }
    }
    ~ScopedLaunch() {
        if (ps->profile) { cudaEventRecord(e1, ps->stream); ps->timers[name].pending.push_back({e0, e1}); }
    }
};

static int flush_timers(mpl_ps* ps) {
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    for (auto& kv : ps->timers) {
        for (auto& pr : kv.second.pending) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, pr.first, pr.second);
            kv.second.total_ms += ms; kv.second.launches++;
            cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
        }
        kv.second.pending.clear();
    }
    return MPL_OK;
}

static void refresh_chunk_records(mpl_ps* ps);
static NestedPrefixes make_prefixes(mpl_ps* ps);
static size_t elem_size(const mpl_ps* ps) { return ps->dtype == MPL_F64 ? 8 : 4; }

static int grid_for(size_t work_items, int per_block, int max_blocks) {
    size_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > (size_t)max_blocks) b = max_blocks;
    return (int)b;
}

// ---- extend dispatch ----------------------------------------------------------------------------------------
template <typename Real>
static ExtendArgs<Real> build_extend_args(mpl_ps* ps, int mode, const Obs& obs, bool from_dev_obs, bool dev_t) {
    ExtendArgs<Real> a;
    a.state_in = (const Real*)ps->state[ps->cur];
    a.state_out = (Real*)ps->state[mode == EXT_INIT || mode == EXT_ACCUM ? ps->cur : ps->cur ^ 1];
    a.lw_in = (const Real*)ps->lw;
    if (ps->world > 1 && ps->lw_alt) { std::swap(ps->lw, ps->lw_alt); ps->par ^= 1; refresh_chunk_records(ps); }   // sharded: the other buffer of the pair (PeerTable::lw)
    a.lw = (Real*)ps->lw;
    a.wait_done = ps->anc_pushed ? 1 : 0;
    a.anc = ps->anc;
    a.n = ps->n; a.ld = ps->ld;
    a.seed = ps->seed; a.gid_offset = ps->gid_offset;
    a.t = dev_t ? -1 : ps->t;
    a.obs = obs;
    a.obs_dev = from_dev_obs ? ps->obs_dev : nullptr;
    a.nobs = ps->model.obs_dim;
    a.stats = ps->stats;
    a.partials = ps->partials;
    a.peer = ps->peer;
    a.cur = ps->cur;
    a.rec = ChunkRecords{ps->rec_e, ps->rec_S, ps->rec_sq};
    a.kbits = fixed_kbits(ps->n_global);
    if (mode == EXT_DYNAMIC) a.state_out = (Real*)ps->state[ps->cur ^ 1];
    return a;
}

// which instantiation pf_extend_kernel<Model, Real, MODE, SHARDED, NESTED> a call maps to.  NESTED (fp32): the kernel's epilogue
// quantises each chunk (1: the integer weights replace the log-weights; 2: chunk records only, ESS-triggered loop).
static void extend_variant(const mpl_ps* ps, int mode, int nested, bool f32, bool& sharded, int& k_nested) {
    k_nested = 0;
    if (f32 && nested == 1 && (mode == EXT_INIT || mode == EXT_GATHER)) k_nested = 1;
    if (f32 && nested == 2 && (mode == EXT_INIT || mode == EXT_DYNAMIC)) k_nested = 2;
    sharded = ps->world > 1 && (mode == EXT_GATHER || mode == EXT_DYNAMIC);
}

template <class Model, typename Real, int NESTED>
static void launch_extend_kernel(mpl_ps* ps, const Model& model, const ExtendArgs<Real>& a, int mode, bool sharded) {
    const int grid = ps->grid_extend;
    switch (mode) {
        case EXT_INIT: pdl_launch(pf_extend_kernel<Model, Real, EXT_INIT, false, NESTED>, grid, kExtendThreads, ps->stream, a, model); break;
        case EXT_ACCUM: if constexpr (NESTED == 0) pdl_launch(pf_extend_kernel<Model, Real, EXT_ACCUM, false, 0>, grid, kExtendThreads, ps->stream, a, model); break;
        case EXT_GATHER:
            if constexpr (NESTED != 2) {
                if (sharded) pdl_launch(pf_extend_kernel<Model, Real, EXT_GATHER, true, NESTED>, grid, kExtendThreads, ps->stream, a, model);
                else pdl_launch(pf_extend_kernel<Model, Real, EXT_GATHER, false, NESTED>, grid, kExtendThreads, ps->stream, a, model);
            }
            break;
        default:
            if constexpr (NESTED != 1) {
                if (sharded) pdl_launch(pf_extend_kernel<Model, Real, EXT_DYNAMIC, true, NESTED>, grid, kExtendThreads, ps->stream, a, model);
                else pdl_launch(pf_extend_kernel<Model, Real, EXT_DYNAMIC, false, NESTED>, grid, kExtendThreads, ps->stream, a, model);
            }
            break;
    }
}

// Model == void*: a model compiled from a spec at run time (jit.cu) -- the same kernel template, instantiated by NVRTC
template <class Model, typename Real>
static int launch_extend_t(mpl_ps* ps, const Model& model, int mode, const Obs& obs, bool from_dev_obs, bool dev_t, int nested = 0) {
    ExtendArgs<Real> a = build_extend_args<Real>(ps, mode, obs, from_dev_obs, dev_t);
    bool sharded;
    int k_nested;
    extend_variant(ps, mode, nested, sizeof(Real) == 4, sharded, k_nested);
    if (mode == EXT_INIT) MPL_CUDA_OK(cudaMemsetAsync(ps->stats->max_bits, 0, sizeof(ps->stats->max_bits), ps->stream));
    {
        ScopedLaunch sl(ps, mode == EXT_INIT ? "init" : "extend");
        if constexpr (std::is_same<Model, const mpl_model*>::value) {
            int rc = jit_launch_extend(*model, ps->dtype, mode, sharded, k_nested, &a, (unsigned int)ps->grid_extend, kExtendThreads, ps->stream, g_use_pdl);
            if (rc) return rc;
        } else if constexpr (sizeof(Real) == 4) {
            if (k_nested == 1) launch_extend_kernel<Model, Real, 1>(ps, model, a, mode, sharded);
            else if (k_nested == 2) launch_extend_kernel<Model, Real, 2>(ps, model, a, mode, sharded);
            else launch_extend_kernel<Model, Real, 0>(ps, model, a, mode, sharded);
        } else launch_extend_kernel<Model, Real, 0>(ps, model, a, mode, sharded);
    }
    MPL_CUDA_OK(cudaGetLastError());
    if (mode == EXT_GATHER || mode == EXT_DYNAMIC) ps->cur ^= 1;
    ps->prequantised = k_nested;
    if (ps->hist_cap) {   // trajectory log: the state this step produced (kernel time index ps->t, not yet incremented by the caller)
        const size_t tt = (size_t)ps->t;
        if (tt < ps->hist_cap) {
            const size_t bytes = (size_t)ps->D * ps->ld * sizeof(Real);
            MPL_CUDA_OK(cudaMemcpyAsync((char*)ps->hist_state + tt * bytes, ps->state[ps->cur], bytes, cudaMemcpyDeviceToDevice, ps->stream));
            if (ps->hist_resampled.size() <= tt) ps->hist_resampled.resize(tt + 1, 0);
            ps->hist_resampled[tt] = 0;
        }
    }
    return MPL_OK;
}

static int launch_extend(mpl_ps* ps, int mode, const Obs& obs, bool from_dev_obs, bool dev_t, int nested = 0) {
    const mpl_model& m = ps->model;
    if (ps->dtype == MPL_F32) {
        switch (m.kind) {
            case M_LGSSM4: return launch_extend_t<Lgssm4<float>, float>(ps, make_lgssm4<float>(m), mode, obs, from_dev_obs, dev_t, nested);
            case M_SPIRAL: return launch_extend_t<Spiral<float>, float>(ps, make_spiral<float>(m), mode, obs, from_dev_obs, dev_t, nested);
            case M_SV: return launch_extend_t<StochVol<float>, float>(ps, make_sv<float>(m), mode, obs, from_dev_obs, dev_t, nested);
            case M_HMM: return launch_extend_t<Hmm<float>, float>(ps, make_hmm<float>(m), mode, obs, from_dev_obs, dev_t, nested);
            case M_JIT: return launch_extend_t<const mpl_model*, float>(ps, &ps->model, mode, obs, from_dev_obs, dev_t, nested);
        }
    } else {
        switch (m.kind) {
            case M_LGSSM4: return launch_extend_t<Lgssm4<double>, double>(ps, make_lgssm4<double>(m), mode, obs, from_dev_obs, dev_t);
            case M_SPIRAL: return launch_extend_t<Spiral<double>, double>(ps, make_spiral<double>(m), mode, obs, from_dev_obs, dev_t);
            case M_SV: return launch_extend_t<StochVol<double>, double>(ps, make_sv<double>(m), mode, obs, from_dev_obs, dev_t);
            case M_HMM: return launch_extend_t<Hmm<double>, double>(ps, make_hmm<double>(m), mode, obs, from_dev_obs, dev_t);
            case M_JIT: return launch_extend_t<const mpl_model*, double>(ps, &ps->model, mode, obs, from_dev_obs, dev_t);
        }
    }
    return fail(MPL_ERR_INVALID, "model is not an Unfold model");
}

static int ensure_stats(mpl_ps* ps) {
    if (ps->stats_valid) return MPL_OK;
    if (ps->world > 1) return fail(MPL_ERR_UNSUPPORTED, "sharded particle system: weight statistics come from the step itself (injected weights are single-GPU only)");
    {
        ScopedLaunch sl(ps, "weight_reduce");
        if (ps->dtype == MPL_F32) weight_reduce_kernel<float><<<ps->grid_reduce, 256, 0, ps->stream>>>((const float*)ps->lw, ps->n, ps->stats, ps->partials);
        else weight_reduce_kernel<double><<<ps->grid_reduce, 256, 0, ps->stream>>>((const double*)ps->lw, ps->n, ps->stats, ps->partials);
    }
    MPL_CUDA_OK(cudaGetLastError());
    ps->stats_valid = true;
    return MPL_OK;
}

static int fetch_stats(mpl_ps* ps) {
    MPL_CUDA_OK(cudaMemcpyAsync(ps->stats_host, ps->stats, sizeof(DeviceStats), cudaMemcpyDeviceToHost, ps->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    return MPL_OK;
}

// materialise a pending ancestor gather (only needed when the host looks at / overwrites state between resample and step)
int materialise(mpl_ps* ps) {
    if (!ps->pending_gather) return MPL_OK;
    if (ps->world > 1) return fail(MPL_ERR_UNSUPPORTED, "sharded particle system: read or write state after the next step (ancestors point into other shards)");
    const int grid = grid_for(ps->n, 256, kNumSMs * 8);
    {
        ScopedLaunch sl(ps, "gather");
        if (ps->dtype == MPL_F32) {
            gather_kernel<float><<<grid, 256, 0, ps->stream>>>((const float*)ps->state[ps->cur], (float*)ps->state[ps->cur ^ 1], ps->anc, ps->n, ps->ld, ps->D);
            fill_kernel<float><<<grid, 256, 0, ps->stream>>>((float*)ps->lw, ps->n, 0.f);
        } else {
            gather_kernel<double><<<grid, 256, 0, ps->stream>>>((const double*)ps->state[ps->cur], (double*)ps->state[ps->cur ^ 1], ps->anc, ps->n, ps->ld, ps->D);
            fill_kernel<double><<<grid, 256, 0, ps->stream>>>((double*)ps->lw, ps->n, 0.);
        }
    }
    ps->launch_count++;
    MPL_CUDA_OK(cudaGetLastError());
    // the gather is applied: the device-side "ancestors pending" flags must not make a later ESS-triggered extend
    // (EXT_DYNAMIC reads stats->resampled_flag[t & 1]) gather a second time through the stale ancestors
    MPL_CUDA_OK(cudaMemsetAsync((char*)ps->stats + offsetof(DeviceStats, resampled_flag), 0, sizeof(int) * 2, ps->stream));
    MPL_CUDA_OK(cudaMemsetAsync((char*)ps->stats + offsetof(DeviceStats, resampled), 0, sizeof(int), ps->stream));
    ps->cur ^= 1;
    ps->pending_gather = false;
    ps->stats_valid = false; ps->max_valid = false;
    return MPL_OK;
}

template <typename Real>
static FixedArgs<Real> fixed_args(mpl_ps* ps, bool dynamic, bool dev_t) {
    FixedArgs<Real> a;
    a.lw = (Real*)ps->lw;
    a.n = ps->n;
    a.kbits = fixed_kbits(ps->n_global);
    a.n_out = ps->n_global;
    a.c_offset = 0;
    a.out_base = ps->gid_offset;
    a.n_out_local = ps->n;
    a.log_n_global = std::log((double)ps->n_global);
    a.anc = ps->anc;
    a.src_base = (int32_t)ps->gid_offset;   // ancestors are global ids (== local indices on a single GPU)
    a.peer = ps->peer;
    a.epoch = dev_t ? -1 : ps->t;
    a.max_slot = (ps->world > 1 || ps->stats_valid) ? -1 : (int)((ps->t - 1) & 1);
    a.overflow_follows = (ps->world > 1 || ps->host_flags[0] != 0) ? 1 : 0;
    a.overflow_seen_host = ps->host_flags_dev;
    a.host_seq = 0;
    a.sq_partials = ps->sq_partials;
    a.ess_threshold = ps->ess_threshold_abs;
    a.desc = ps->desc;
    a.stats = ps->stats;
    a.partials = ps->ipartials;
    a.seed = ps->seed;
    a.rt = dev_t ? -1 : ps->t - 1;
    a.accumulate_lml = 1;
    a.dynamic = dynamic ? 1 : 0;
    a.inline_level1 = 0;
    return a;
}

template <typename Real>
static int resample_fixed_t(mpl_ps* ps, int scheme, bool dynamic, bool dev_t) {
    FixedArgs<Real> a = fixed_args<Real>(ps, dynamic, dev_t);
    const size_t num_tiles = (ps->n + kScanTile - 1) / kScanTile;
    {
        ScopedLaunch sl(ps, "fixed_reduce");
        if (dynamic) pdl_launch(fixed_reduce_kernel<Real, false>, (unsigned int)num_tiles, kScanThreads, ps->stream, a, (unsigned int)num_tiles);
        else pdl_launch(fixed_reduce_kernel<Real, true>, (unsigned int)num_tiles, kScanThreads, ps->stream, a, (unsigned int)num_tiles);
    }
    MPL_CUDA_OK(cudaGetLastError());
    if (scheme == MPL_RESAMPLE_SYSTEMATIC_FIXED) {
        {
            ScopedLaunch sl(ps, "fixed_scan");
            if (dynamic) pdl_launch(fixed_scan2_kernel<Real, false>, (unsigned int)num_tiles, kScanThreads, ps->stream, a, (unsigned int)num_tiles, (OverflowEntry2*)ps->overflow);
            else pdl_launch(fixed_scan2_kernel<Real, true>, (unsigned int)num_tiles, kScanThreads, ps->stream, a, (unsigned int)num_tiles, (OverflowEntry2*)ps->overflow);
        }
        MPL_CUDA_OK(cudaGetLastError());
        ps->anc_pushed = ps->world > 1;
        if (a.overflow_follows) {
            ScopedLaunch sl(ps, "fixed_overflow");
            if (dynamic) pdl_launch(fixed_overflow2_kernel<Real, false>, kNumSMs * 2, kScanThreads, ps->stream, a, (const OverflowEntry2*)ps->overflow);
            else pdl_launch(fixed_overflow2_kernel<Real, true>, kNumSMs * 2, kScanThreads, ps->stream, a, (const OverflowEntry2*)ps->overflow);
        }
        MPL_CUDA_OK(cudaGetLastError());
    } else {
        if (!ps->icum) MPL_CUDA_OK(cudaMalloc(&ps->icum, ps->ld * sizeof(unsigned long long)));
        {
            ScopedLaunch sl(ps, "fixed_cumsum");
            fixed_cumsum_kernel<Real><<<(unsigned int)num_tiles, kScanThreads, 0, ps->stream>>>(a, ps->icum);
        }
        MPL_CUDA_OK(cudaGetLastError());
        {
            ScopedLaunch sl(ps, "fixed_search");
            fixed_multinomial_search_kernel<<<grid_for(ps->n, 256, kNumSMs * 16), 256, 0, ps->stream>>>(ps->icum, ps->n, ps->n, ps->stats, ps->seed, ps->gid_offset,
                                                                                                      (uint32_t)(ps->t - 1), ps->anc);
        }
        MPL_CUDA_OK(cudaGetLastError());
    }
    return MPL_OK;
}


// the chunk records in use: parity ps->par of the pair (only sharded runs alternate)
static void refresh_chunk_records(mpl_ps* ps) {
    if (!ps->rec_e2) return;
    const size_t nch = ps->ld / kChunk;
    ps->rec_e = ps->rec_e2 + (size_t)ps->par * nch;
    ps->rec_S = ps->rec_S2 + (size_t)ps->par * nch;
}

int ensure_chunk_records(mpl_ps* ps) {
    if (ps->rec_e2) return MPL_OK;
    const size_t nch = ps->ld / kChunk;
    MPL_CUDA_OK(cudaMalloc(&ps->rec_e2, 2 * nch * sizeof(int)));
    MPL_CUDA_OK(cudaMalloc(&ps->rec_S2, 2 * nch * sizeof(unsigned int)));
    MPL_CUDA_OK(cudaMalloc(&ps->rec_sq, nch * sizeof(float)));
    refresh_chunk_records(ps);
    // section records and top-level results of EVERY shard: [E int | T u64 | sq f64 | pre u64 | M u64 | a u64 | n u64] x kMaxSections
    MPL_CUDA_OK(cudaMalloc(&ps->nest_sec, (size_t)kMaxSections * 7 * sizeof(unsigned long long)));
    MPL_CUDA_OK(cudaMemset(ps->nest_sec, 0, (size_t)kMaxSections * 7 * sizeof(unsigned long long)));
    const size_t nsec = (ps->ld + kSection - 1) / kSection;
    MPL_CUDA_OK(cudaMalloc(&ps->nest_tile_pre, 2 * nsec * kTilesPerSection * sizeof(unsigned long long)));   // (a pair, like the chunk records)
    // plan of the WHOLE population (a shard plans the sections whose slots it fills, wherever their particles live)
    const size_t nch_g = (ps->n_global + kChunk - 1) / kChunk, nt_g = (ps->n_global + kScanTile - 1) / kScanTile;
    MPL_CUDA_OK(cudaMalloc(&ps->nest_P, (nch_g + 2) * sizeof(unsigned int)));
    MPL_CUDA_OK(cudaMalloc(&ps->nest_F, (nt_g + 1) * sizeof(unsigned int)));
    if (ps->world <= 1) {   // one GPU: the "peer" tables point at this system's own arrays (kernels use one code path)
        for (int p = 0; p < 2; ++p) { ps->peer.lw[p][0] = ps->lw; ps->peer.rec_e[p][0] = ps->rec_e2; ps->peer.rec_S[p][0] = ps->rec_S2; ps->peer.tile_pre[p][0] = ps->nest_tile_pre; }
        ps->peer.n_loc = (unsigned int)ps->n;
    }
    return MPL_OK;
}

static NestedPrefixes make_prefixes(mpl_ps* ps) {
    unsigned long long* sec = ps->nest_sec;
    const size_t n_sec_global = (ps->n_global + kSection - 1) / kSection;
    const size_t n_tp = ((ps->ld + kSection - 1) / kSection) * kTilesPerSection;
    return NestedPrefixes{ps->nest_tile_pre + (size_t)ps->par * n_tp, (int*)sec, sec + kMaxSections, (double*)(sec + 2 * kMaxSections), sec + 3 * kMaxSections, sec + 4 * kMaxSections,
                          sec + 5 * kMaxSections, sec + 6 * kMaxSections, ps->nest_P, ps->nest_F,
                          (unsigned int)(ps->gid_offset / kSection), (unsigned int)((ps->n + kSection - 1) / kSection), (unsigned int)n_sec_global};
}

// phases: 1 = quantise (unless the extend did it) + chunk pass, 2 = expansion (+ the peers' "done" flag); 3 = both
// dynamic: ESS-triggered (decision on the device; the log-weights survive: quantisation keeps them, the expansion re-quantises)
template <typename Real>
static int resample_nested_t(mpl_ps* ps, int phases = 3, bool dynamic = false) {
    int rc = ensure_chunk_records(ps);
    if (rc) return rc;
    if (ps->world > 1 && ((ps->n % kSection) || (ps->gid_offset % kSection) || ps->n * (uint64_t)ps->world != ps->n_global))
        return fail(MPL_ERR_UNSUPPORTED, "nested scheme, sharded: equal shards of whole sections (multiples of 131072 particles)");
    const size_t n_sec_global = (ps->n_global + kSection - 1) / kSection;
    if (n_sec_global > (size_t)kMaxSections) return fail(MPL_ERR_UNSUPPORTED, "nested scheme: at most 2^28 particles");
    FixedArgs<Real> a = fixed_args<Real>(ps, dynamic, false);
    a.overflow_follows = ps->host_flags[0] != 0 ? 1 : 0;   // the heavy-tile pass runs only once a heavy warp tile has been seen
    // small populations on one GPU: a kernel boundary costs more than level 1 recomputed by every warp of the expansion -- no plan pass
    // (measured: on 8 GPUs the walk over whole edge sections costs more than the plan pass it saves -- 62.6 against 52 us per step;
    //  a single GPU with 2^21 particles gains 3 us per step)
    a.inline_level1 = g_inline_level1 >= 0 ? g_inline_level1 : (ps->world <= 1 && ps->n <= ((size_t)1 << 22) ? 1 : 0);   // (2^24: 51 us against 42 + 4.6 for the pass of its own)
    const bool post = (phases & 2) && !dynamic && !ps->in_device_loop;   // the call-per-step API polls the result in mapped host memory
    if (post) { ps->host_seq += 1; if (ps->host_seq == 0) ps->host_seq = 1; }
    a.host_seq = post ? ps->host_seq : 0u;
    if (ps->world <= 1) for (int p = 0; p < 2; ++p) ps->peer.lw[p][0] = ps->lw;   // (the log-weight array may have been reallocated: keep the self-table current)
    a.peer = ps->peer;
    ChunkRecords rec{ps->rec_e, ps->rec_S, ps->rec_sq};
    NestedPrefixes nb = make_prefixes(ps);
    const unsigned int num_tiles = (unsigned int)((ps->n + kScanTile - 1) / kScanTile);
    const unsigned int num_chunks = (unsigned int)((ps->n + kChunk - 1) / kChunk);
    const unsigned int n_chunks_global = (unsigned int)((ps->n_global + kChunk - 1) / kChunk), n_tiles_global = (unsigned int)((ps->n_global + kScanTile - 1) / kScanTile);
    if (dynamic && ps->prequantised == 1) return fail(MPL_ERR_INVALID, "ESS-triggered nested resampling after a step that dropped the log-weights");
    if ((phases & 1) && !ps->prequantised) {
        ScopedLaunch sl(ps, "nested_quantise");
        if (dynamic) pdl_launch(nested_quantise_kernel<Real, true>, num_tiles, kScanThreads, ps->stream, a, rec);
        else pdl_launch(nested_quantise_kernel<Real, false>, num_tiles, kScanThreads, ps->stream, a, rec);
    }
    {
        ScopedLaunch sl(ps, "nested_sections");
        if (phases == 3) pdl_launch(nested_sections_kernel<Real, 3>, nb.n_sec, kScanThreads, ps->stream, a, rec, nb, num_tiles, num_chunks);
        else if (phases == 1) pdl_launch(nested_sections_kernel<Real, 1>, nb.n_sec, kScanThreads, ps->stream, a, rec, nb, num_tiles, num_chunks);
        else pdl_launch(nested_sections_kernel<Real, 2>, 1, kScanThreads, ps->stream, a, rec, nb, num_tiles, num_chunks);
    }
    if ((phases & 2) && !a.inline_level1) {
        ScopedLaunch sl(ps, "nested_plan");
        if (ps->world > 1) pdl_launch(nested_plan_kernel<Real>, (unsigned int)n_sec_global, kScanThreads, ps->stream, a, nb, ps->par, n_chunks_global);
        else pdl_launch(nested_level1_kernel<Real>, (num_tiles + kScanThreads / 32 - 1) / (kScanThreads / 32), kScanThreads, ps->stream, a, nb, rec, num_tiles, n_chunks_global);
    }
    if (phases & 2) {
        ScopedLaunch sl(ps, "nested_expand");
        NestedHeavyEntry* hv = (NestedHeavyEntry*)ps->overflow;
        if (ps->world > 1) {   // the chunks that own this shard's slots: about its own tiles, one more at each edge, more when the weights are lopsided (grid-stride)
            const unsigned int grid = num_tiles + (a.inline_level1 ? 2u * kTilesPerSection : 2u);
            if (dynamic) pdl_launch(nested_expand_kernel<Real, true, true>, grid, kScanThreads, ps->stream, a, nb, rec, ps->par, num_tiles, n_tiles_global, n_chunks_global, hv);
            else pdl_launch(nested_expand_kernel<Real, false, true>, grid, kScanThreads, ps->stream, a, nb, rec, ps->par, num_tiles, n_tiles_global, n_chunks_global, hv);
        } else {
            if (dynamic) pdl_launch(nested_expand_kernel<Real, true, false>, num_tiles, kScanThreads, ps->stream, a, nb, rec, ps->par, num_tiles, n_tiles_global, n_chunks_global, hv);
            else pdl_launch(nested_expand_kernel<Real, false, false>, num_tiles, kScanThreads, ps->stream, a, nb, rec, ps->par, num_tiles, n_tiles_global, n_chunks_global, hv);
        }
    }
    if ((phases & 2) && a.overflow_follows) {
        ScopedLaunch sl(ps, "nested_heavy");
        const NestedHeavyEntry* hv = (const NestedHeavyEntry*)ps->overflow;
        if (ps->world > 1) {
            if (dynamic) pdl_launch(nested_heavy_kernel<Real, true, true>, kNumSMs * 2, kScanThreads, ps->stream, a, nb, rec, ps->par, n_chunks_global, hv);
            else pdl_launch(nested_heavy_kernel<Real, false, true>, kNumSMs * 2, kScanThreads, ps->stream, a, nb, rec, ps->par, n_chunks_global, hv);
        } else {
            if (dynamic) pdl_launch(nested_heavy_kernel<Real, true, false>, kNumSMs * 2, kScanThreads, ps->stream, a, nb, rec, ps->par, n_chunks_global, hv);
            else pdl_launch(nested_heavy_kernel<Real, false, false>, kNumSMs * 2, kScanThreads, ps->stream, a, nb, rec, ps->par, n_chunks_global, hv);
        }
    }
    MPL_CUDA_OK(cudaGetLastError());
    if (phases & 1) ps->prequantised = 0;
    if (phases & 2) { ps->host_lse_posted = post; ps->anc_pushed = false; }
    return MPL_OK;
}

static int resample_exact(mpl_ps* ps, int scheme) {
    if (!ps->probs) {
        MPL_CUDA_OK(cudaMalloc(&ps->probs, ps->ld * sizeof(double)));
        MPL_CUDA_OK(cudaMalloc(&ps->cums, ps->ld * sizeof(double)));
    }
    const int grid = grid_for(ps->n, 256, kNumSMs * 8);
    {
        ScopedLaunch sl(ps, "normalize");
        if (ps->dtype == MPL_F32) normalize_kernel<float><<<grid, 256, 0, ps->stream>>>((const float*)ps->lw, ps->n, ps->probs, ps->stats, std::log((double)ps->n_global), 1);
        else normalize_kernel<double><<<grid, 256, 0, ps->stream>>>((const double*)ps->lw, ps->n, ps->probs, ps->stats, std::log((double)ps->n_global), 1);
    }
    MPL_CUDA_OK(cudaGetLastError());
    int rc;
    {
        ScopedLaunch sl(ps, "cumsum_exact");
        rc = launch_cumsum_exact(ps, ps->probs, ps->n, ps->cums, ps->stream);
    }
    if (rc) return rc;
    {
        ScopedLaunch sl(ps, "search");
        search_kernel<int32_t><<<grid, 256, 0, ps->stream>>>(ps->cums, ps->n, ps->n, scheme == MPL_RESAMPLE_MULTINOMIAL ? 2 : 3, nullptr, ps->seed, ps->gid_offset,
                                                             (uint32_t)(ps->t - 1), ps->anc);
    }
    MPL_CUDA_OK(cudaGetLastError());
    return MPL_OK;
}

int launch_cumsum_exact(mpl_ps* ps, const double* probs, size_t n, double* out, cudaStream_t stream) {
    // categorical.rs:25-30: S_k = fl(S_{k-1} + p_k) in index order.  Short inputs: one thread adds in order.  Long inputs:
    // the parallel exact emulation of cumsum_exact.cuh (same bits).
    if (n < (size_t)4 * kCxTile) {
        if (ps) ps->launch_count++;
        cumsum_seq_kernel<<<1, 256, 0, stream>>>(probs, n, out);
        MPL_CUDA_OK(cudaGetLastError());
        return MPL_OK;
    }
    const unsigned int num_tiles = (unsigned int)((n + kCxTile - 1) / kCxTile);
    CxTile* tiles = nullptr;
    int* bad = nullptr;
    {   // the tile records come from the device's stream-ordered pool; by default the pool hands its memory back to the driver at
        // every synchronisation, and a caller that synchronises after each resample would pay a fresh cudaMalloc (~0.8 ms) per call
        static thread_local int pool_kept_for = -1;
        int dev = 0;
        MPL_CUDA_OK(cudaGetDevice(&dev));
        if (pool_kept_for != dev) {
            cudaMemPool_t pool;
            unsigned long long keep = ~0ull;
            MPL_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
            MPL_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            pool_kept_for = dev;
        }
    }
    MPL_CUDA_OK(cudaMallocAsync(&tiles, (size_t)num_tiles * sizeof(CxTile) + 64, stream));
    bad = reinterpret_cast<int*>(reinterpret_cast<char*>(tiles) + (size_t)num_tiles * sizeof(CxTile));
    MPL_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    cx_tilesum_kernel<<<num_tiles, kCxThreads, 0, stream>>>(probs, n, tiles, bad);
    cx_classify_kernel<<<1, 1024, 0, stream>>>(tiles, num_tiles, bad);
    cx_pairs_kernel<<<num_tiles, kCxThreads, 0, stream>>>(probs, n, tiles);
    cx_walk_kernel<<<1, kCxThreads, 0, stream>>>(probs, n, tiles, num_tiles, out, bad);
    cx_fill_kernel<<<num_tiles, kCxThreads, 0, stream>>>(probs, n, tiles, out);
    if (ps) ps->launch_count += 5;
    MPL_CUDA_OK(cudaGetLastError());
    MPL_CUDA_OK(cudaFreeAsync(tiles, stream));
    return MPL_OK;
}

static int do_resample(mpl_ps* ps, int scheme) {
    if (!ps->initialised) return fail(MPL_ERR_INVALID, "resample before init_step");
    if (ps->world > 1 && scheme != MPL_RESAMPLE_SYSTEMATIC_FIXED && scheme != MPL_RESAMPLE_SYSTEMATIC_NESTED) return fail(MPL_ERR_UNSUPPORTED, "sharded particle systems support the integer systematic schemes only");
    int rc = materialise(ps);   // resample twice in a row: apply the first one
    if (rc) return rc;
    const bool exact = scheme == MPL_RESAMPLE_MULTINOMIAL || scheme == MPL_RESAMPLE_SYSTEMATIC;
    if (exact || (!ps->max_valid && scheme != MPL_RESAMPLE_SYSTEMATIC_NESTED)) {   // the reference scheme needs log-sum-exp; the single-level integer schemes only the max (left by the extend); the nested one nothing
        rc = ensure_stats(ps);
        if (rc) return rc;
    }
    switch (scheme) {
        case MPL_RESAMPLE_MULTINOMIAL:
        case MPL_RESAMPLE_SYSTEMATIC: rc = resample_exact(ps, scheme); break;
        case MPL_RESAMPLE_SYSTEMATIC_FIXED:
        case MPL_RESAMPLE_MULTINOMIAL_FIXED:
            rc = ps->dtype == MPL_F32 ? resample_fixed_t<float>(ps, scheme, false, false) : resample_fixed_t<double>(ps, scheme, false, false);
            break;
        case MPL_RESAMPLE_SYSTEMATIC_NESTED:
            rc = ps->dtype == MPL_F32 ? resample_nested_t<float>(ps) : resample_nested_t<double>(ps);
            break;
        default: return fail(MPL_ERR_INVALID, "unknown resampling scheme");
    }
    if (rc) return rc;
    ps->pending_gather = true;
    ps->stats_valid = false; ps->max_valid = false;
    if (ps->hist_cap && ps->t >= 1 && (size_t)(ps->t - 1) < ps->hist_cap) {   // ancestors chosen after step t-1
        const size_t tt = (size_t)(ps->t - 1);
        MPL_CUDA_OK(cudaMemcpyAsync(ps->hist_anc + tt * ps->ld, ps->anc, ps->ld * sizeof(int32_t), cudaMemcpyDeviceToDevice, ps->stream));
        if (ps->hist_resampled.size() <= tt) ps->hist_resampled.resize(tt + 1, 0);
        ps->hist_resampled[tt] = 1;
    }
    return MPL_OK;
}

int ps_phase_extend(mpl_ps* ps, bool init, bool fuse_nested) {
    Obs dummy; std::memset(&dummy, 0, sizeof dummy);
    int rc;
    if (fuse_nested && (rc = ensure_chunk_records(ps))) return rc;
    if (init) {
        ps->t = 0; ps->pending_gather = false;
        rc = launch_extend(ps, EXT_INIT, dummy, true, false, fuse_nested);
        ps->t = 1; ps->initialised = true; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
    } else {
        rc = launch_extend(ps, ps->pending_gather ? EXT_GATHER : EXT_ACCUM, dummy, true, false, fuse_nested);
        ps->pending_gather = false; ps->t += 1; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
    }
    return rc;
}
int ps_phase_nested(mpl_ps* ps, int phase) {
    int rc = ps->dtype == MPL_F32 ? resample_nested_t<float>(ps, phase) : resample_nested_t<double>(ps, phase);
    if (rc == MPL_OK && phase == 2) { ps->pending_gather = true; ps->stats_valid = false; ps->max_valid = false; }
    return rc;
}
int ps_phase_reduce(mpl_ps* ps) {
    const size_t num_tiles = (ps->n + kScanTile - 1) / kScanTile;
    ScopedLaunch sl(ps, "fixed_reduce");
    if (ps->dtype == MPL_F32) { auto a = fixed_args<float>(ps, false, false); fixed_reduce_kernel<float, true><<<(unsigned int)num_tiles, kScanThreads, 0, ps->stream>>>(a, (unsigned int)num_tiles); }
    else { auto a = fixed_args<double>(ps, false, false); fixed_reduce_kernel<double, true><<<(unsigned int)num_tiles, kScanThreads, 0, ps->stream>>>(a, (unsigned int)num_tiles); }
    MPL_CUDA_OK(cudaGetLastError());
    return MPL_OK;
}
int ps_phase_scan(mpl_ps* ps) {
    const size_t num_tiles = (ps->n + kScanTile - 1) / kScanTile;
    if (ps->dtype == MPL_F32) {
        auto a = fixed_args<float>(ps, false, false);
        { ScopedLaunch sl(ps, "fixed_scan"); fixed_scan2_kernel<float, true><<<(unsigned int)num_tiles, kScanThreads, 0, ps->stream>>>(a, (unsigned int)num_tiles, (OverflowEntry2*)ps->overflow); }
        if (a.overflow_follows) { ScopedLaunch sl(ps, "fixed_overflow"); fixed_overflow2_kernel<float, true><<<kNumSMs * 2, kScanThreads, 0, ps->stream>>>(a, (const OverflowEntry2*)ps->overflow); }
    } else {
        auto a = fixed_args<double>(ps, false, false);
        { ScopedLaunch sl(ps, "fixed_scan"); fixed_scan2_kernel<double, true><<<(unsigned int)num_tiles, kScanThreads, 0, ps->stream>>>(a, (unsigned int)num_tiles, (OverflowEntry2*)ps->overflow); }
        if (a.overflow_follows) { ScopedLaunch sl(ps, "fixed_overflow"); fixed_overflow2_kernel<double, true><<<kNumSMs * 2, kScanThreads, 0, ps->stream>>>(a, (const OverflowEntry2*)ps->overflow); }
    }
    MPL_CUDA_OK(cudaGetLastError());
    ps->pending_gather = true; ps->stats_valid = false; ps->max_valid = false;
    ps->anc_pushed = ps->world > 1;
    return MPL_OK;
}

}  // namespace mpl

using namespace mpl;

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" const char* mpl_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* mpl_version(void) { return "modppl_b200 0.1 (sm_100a)"; }
extern "C" int mpl_device_count(int* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; return fail(MPL_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *count = c;
    return MPL_OK;
}

extern "C" mpl_model* mpl_model_create(const char* name, const double* params, size_t n_params) {
    if (!name) { fail(MPL_ERR_INVALID, "null model name"); return nullptr; }
    int kind = model_from_name(name);
    if (kind < 0) { fail(MPL_ERR_INVALID, std::string("unknown model: ") + name); return nullptr; }
    auto* m = new mpl_model;
    m->kind = kind; m->name = name;
    if (params && n_params) m->params.assign(params, params + n_params);
    m->num_latents = 0;
    switch (kind) {
        case M_LGSSM4: m->state_dim = 4; m->obs_dim = 2; break;
        case M_SPIRAL: m->state_dim = 2; m->obs_dim = 2; break;
        case M_SV: m->state_dim = 1; m->obs_dim = 1; break;
        case M_HMM: {
            m->state_dim = 1; m->obs_dim = 1;
            bool ok = n_params >= 2;
            int K = ok ? (int)params[0] : 0, M = ok ? (int)params[1] : 0;
            ok = ok && K >= 1 && K <= kHmmMaxK && M >= 1 && M <= kHmmMaxK && n_params == (size_t)(2 + K + M * K + K * K);
            if (!ok) { delete m; fail(MPL_ERR_INVALID, "hmm params: {K, M, prior[K], emission[M*K], transition[K*K]} with K, M <= 8"); return nullptr; }
            break;
        }
        case M_LINE: m->state_dim = 0; m->obs_dim = (int)n_params; m->num_latents = 2; break;
        case M_HIER: m->state_dim = 0; m->obs_dim = (int)n_params; m->num_latents = 4; break;
        case M_POINTED:
            m->state_dim = 0; m->obs_dim = 2; m->num_latents = 2;
            if (n_params != 8) { delete m; fail(MPL_ERR_INVALID, "pointed params: {xmin,xmax,ymin,ymax, cov[4]}"); return nullptr; }
            if (!(params[1] > params[0]) || !(params[3] > params[2])) { delete m; fail(MPL_ERR_INVALID, "pointed bounds: max must exceed min (types_2d.rs:24-25)"); return nullptr; }
            break;
    }
    return m;
}
extern "C" void mpl_model_destroy(mpl_model* m) { delete m; }
extern "C" int mpl_model_state_dim(const mpl_model* m) { return m ? m->state_dim : MPL_ERR_INVALID; }
extern "C" int mpl_model_obs_dim(const mpl_model* m) { return m ? m->obs_dim : MPL_ERR_INVALID; }
extern "C" int mpl_model_num_latents(const mpl_model* m) { return m ? m->num_latents : MPL_ERR_INVALID; }

extern "C" mpl_ps* mpl_particle_system_new(const mpl_model* model, uint64_t num_particles, const mpl_pf_config* cfg) {
    if (!model || model->state_dim <= 0) { fail(MPL_ERR_INVALID, "particle system needs an Unfold model"); return nullptr; }
    if (num_particles == 0 || num_particles > (1ull << 31)) { fail(MPL_ERR_INVALID, "num_particles must be in [1, 2^31]"); return nullptr; }
    mpl_pf_config c;
    if (cfg) c = *cfg; else { c.dtype = MPL_F64; c.device = -1; c.seed = 0; c.gid_offset = 0; c.n_global = 0; }
    if (c.dtype != MPL_F32 && c.dtype != MPL_F64) { fail(MPL_ERR_INVALID, "dtype"); return nullptr; }
    if (model->kind == M_SV && (c.gid_offset % 4)) { fail(MPL_ERR_INVALID, "stochastic-volatility model: shards start at multiples of 4 (4 particles share a Philox block)"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); fail(MPL_ERR_CUDA, "no CUDA device: modppl_b200 has no CPU fallback"); return nullptr; }
    if (c.device >= 0) { if (cudaSetDevice(c.device) != cudaSuccess) { fail(MPL_ERR_CUDA, "cudaSetDevice failed"); return nullptr; } }
    auto* ps = new mpl_ps();
    ps->model_ref = model; ps->model = *model;
    ps->dtype = c.dtype;
    cudaGetDevice(&ps->device);
    ps->n = num_particles;
    ps->ld = (num_particles + kScanTile - 1) / kScanTile * kScanTile;
    ps->seed = c.seed; ps->gid_offset = c.gid_offset; ps->n_global = c.n_global ? c.n_global : num_particles;
    ps->D = model->state_dim;
    ps->cur = 0; ps->t = 0; ps->initialised = false; ps->pending_gather = false; ps->stats_valid = false; ps->max_valid = false;
    ps->hist_state = nullptr; ps->hist_anc = nullptr; ps->hist_cap = 0;
    ps->rec_e = nullptr; ps->rec_S = nullptr; ps->rec_sq = nullptr; ps->prequantised = 0; ps->host_seq = 0; ps->host_lse_posted = false; ps->in_device_loop = false;
    ps->nest_tile_pre = nullptr; ps->nest_sec = nullptr; ps->nest_P = nullptr; ps->nest_F = nullptr;
    ps->rec_e2 = nullptr; ps->rec_S2 = nullptr; ps->lw_alt = nullptr; ps->par = 0; ps->anc_pushed = false;
    ps->sq_partials = nullptr; ps->host_flags = nullptr; ps->host_flags_dev = nullptr; ps->ess_threshold_abs = 0.; ps->dynamic_state_known = false; ps->hist_broken = false;
    ps->profile = false; ps->launch_count = 0; ps->rank = 0; ps->world = 1; ps->mailbox = nullptr; ps->peer_virtual = false;
    std::memset(&ps->peer, 0, sizeof ps->peer); ps->peer.world = 1; std::memset(ps->ipc_opened, 0, sizeof ps->ipc_opened);
    ps->barrier_seq = 0; ps->n_islands = 1; ps->island_rank = 0; std::memset(ps->island_state, 0, sizeof ps->island_state); std::memset(ps->island_opened, 0, sizeof ps->island_opened);
    ps->probs = nullptr; ps->cums = nullptr; ps->icum = nullptr; ps->obs_dev = nullptr; ps->obs_steps = 0; ps->staging = nullptr;
    const size_t es = elem_size(ps);
    const size_t num_tiles = ps->ld / kScanTile;
    ps->overflow_cap = (ps->n_global / kWarpHeavyCap + 8) * 2;   // entries (OverflowEntry2 is the larger record)
    const int V = ps->dtype == MPL_F32 ? 4 : 2;
    ps->grid_extend = grid_for(ps->n, kExtendThreads * V, kNumSMs * 8);
    ps->grid_reduce = grid_for(ps->n, 256 * 4, kNumSMs * 8);
    bool ok = cudaStreamCreateWithFlags(&ps->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->state[0], ps->D * ps->ld * es) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->state[1], ps->D * ps->ld * es) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->lw, ps->ld * es) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->anc, ps->ld * sizeof(int32_t)) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->desc, num_tiles * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemset(ps->desc, 0, num_tiles * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->overflow, ps->overflow_cap * sizeof(OverflowEntry2)) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->stats, sizeof(DeviceStats)) == cudaSuccess;
    ok = ok && cudaMallocHost(&ps->stats_host, sizeof(DeviceStats)) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->partials, kNumSMs * 8 * sizeof(Lse3<double>)) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->ipartials, kNumSMs * 8 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&ps->sq_partials, num_tiles * sizeof(double)) == cudaSuccess;
    ok = ok && cudaHostAlloc(&ps->host_flags, 64, cudaHostAllocMapped) == cudaSuccess;
    if (ok) { std::memset(ps->host_flags, 0, 64); ok = cudaHostGetDevicePointer(&ps->host_flags_dev, ps->host_flags, 0) == cudaSuccess; }
    if (ok) {
        // particle_filter.rs:44-57: zero log-weights, zero log-ML.  ESS of the all-zero stale buffer is 1/N (quirk Q1).
        DeviceStats init;
        std::memset(&init, 0, sizeof init);
        init.ess_stale = 1. / (double)ps->n_global;
        init.max = 0.; init.sumexp = (double)ps->n_global; init.sumexp2 = (double)ps->n_global; init.ess = (double)ps->n_global;
        ok = cudaMemcpyAsync(ps->stats, &init, sizeof init, cudaMemcpyHostToDevice, ps->stream) == cudaSuccess;
        ok = ok && cudaMemsetAsync(ps->state[0], 0, ps->D * ps->ld * es, ps->stream) == cudaSuccess;
        ok = ok && cudaMemsetAsync(ps->state[1], 0, ps->D * ps->ld * es, ps->stream) == cudaSuccess;
        ok = ok && cudaMemsetAsync(ps->lw, 0, ps->ld * es, ps->stream) == cudaSuccess;
        ok = ok && cudaMemsetAsync(ps->anc, 0, ps->ld * sizeof(int32_t), ps->stream) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(ps->stream) == cudaSuccess;
    }
    if (!ok) {
        fail(MPL_ERR_CUDA, std::string("particle system allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        mpl_ps_destroy(ps);
        return nullptr;
    }
    return ps;
}

extern "C" void mpl_ps_destroy(mpl_ps* ps) {
    if (!ps) return;
    if (ps->stream) cudaStreamSynchronize(ps->stream);
    for (auto& kv : ps->timers) for (auto& pr : kv.second.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    cudaFree(ps->state[0]); cudaFree(ps->state[1]); cudaFree(ps->lw); cudaFree(ps->anc); cudaFree(ps->desc); cudaFree(ps->overflow);
    cudaFree(ps->stats); cudaFreeHost(ps->stats_host); cudaFree(ps->partials); cudaFree(ps->ipartials);
    cudaFree(ps->sq_partials); if (ps->host_flags) cudaFreeHost(ps->host_flags);
    cudaFree(ps->hist_state); cudaFree(ps->hist_anc);
    cudaFree(ps->rec_e2); cudaFree(ps->rec_S2); cudaFree(ps->rec_sq); cudaFree(ps->nest_tile_pre); cudaFree(ps->nest_sec); cudaFree(ps->nest_P); cudaFree(ps->nest_F);
    cudaFree(ps->lw_alt);
    if (ps->world > 1 && !ps->peer_virtual) mpl_ps_peer_detach(ps);
    for (int k = 0; k < 2; ++k) for (int h = 0; h < kMaxPeers; ++h) if (ps->island_opened[k][h]) cudaIpcCloseMemHandle(ps->island_opened[k][h]);
    cudaFree(ps->mailbox);
    cudaFree(ps->probs); cudaFree(ps->cums); cudaFree(ps->icum); cudaFree(ps->obs_dev); cudaFree(ps->staging);
    if (ps->stream) cudaStreamDestroy(ps->stream);
    delete ps;
}

static int pack_obs(const mpl_ps* ps, const double* obs, size_t n_obs, Obs& o) {
    if (!obs || n_obs < (size_t)ps->model.obs_dim) return fail(MPL_ERR_INVALID, "observation vector too short for this model");
    for (int k = 0; k < 4; ++k) o.v[k] = (k < ps->model.obs_dim) ? obs[k] : 0.;
    if (ps->model.kind == M_HMM) {
        if (!(obs[0] >= 0. && obs[0] < ps->model.params[1]) || obs[0] != std::floor(obs[0])) return fail(MPL_ERR_INVALID, "hmm observation symbol out of range");
    }
    return MPL_OK;
}

extern "C" int mpl_ps_init_step(mpl_ps* ps, const double* obs, size_t n_obs) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    Obs o;
    int rc = pack_obs(ps, obs, n_obs, o);
    if (rc) return rc;
    if (ps->world > 1 && ps->initialised) return fail(MPL_ERR_UNSUPPORTED, "sharded particle system: init_step once per attach (mailbox epochs restart with t)");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    ps->t = 0;   // quirk Q4: init_step resets instead of pushing a second population
    ps->pending_gather = false;
    rc = launch_extend(ps, EXT_INIT, o, false, false);
    if (rc) return rc;
    ps->t = 1; ps->initialised = true; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
    return MPL_OK;
}

extern "C" int mpl_ps_step(mpl_ps* ps, const double* obs, size_t n_obs) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    if (!ps->initialised) return fail(MPL_ERR_INVALID, "step before init_step");
    Obs o;
    int rc = pack_obs(ps, obs, n_obs, o);
    if (rc) return rc;
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    rc = launch_extend(ps, ps->pending_gather ? EXT_GATHER : EXT_ACCUM, o, false, false);
    if (rc) return rc;
    ps->pending_gather = false;
    ps->t += 1; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
    return MPL_OK;
}

// step() immediately followed by resample(): the pair of calls of the reference's filtering loop (tests/smc.rs:78-81) as one,
// which lets the extend kernel hand the resampler its weights already quantised (nested scheme, fp32)
extern "C" int mpl_ps_step_resample(mpl_ps* ps, const double* obs, size_t n_obs, int scheme, double* log_total_weight) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    if (!ps->initialised) return fail(MPL_ERR_INVALID, "step before init_step");
    Obs o;
    int rc = pack_obs(ps, obs, n_obs, o);
    if (rc) return rc;
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    const bool fuse = scheme == MPL_RESAMPLE_SYSTEMATIC_NESTED && ps->dtype == MPL_F32 && ps->pending_gather &&
                      !(ps->world > 1 && (ps->n % kSection || ps->gid_offset % kSection));
    if (fuse && (rc = ensure_chunk_records(ps))) return rc;
    rc = launch_extend(ps, ps->pending_gather ? EXT_GATHER : EXT_ACCUM, o, false, false, fuse);
    if (rc) return rc;
    ps->pending_gather = false;
    ps->t += 1; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
    return mpl_ps_resample(ps, scheme, log_total_weight);
}

extern "C" int mpl_ps_effective_sample_size(mpl_ps* ps, int stale_like_reference, double* out) {
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc;
    if (!stale_like_reference) {
        if (ps->pending_gather) { *out = (double)ps->n_global; return MPL_OK; }   // all weights are zero
        if ((rc = ensure_stats(ps))) return rc;
    }
    if ((rc = fetch_stats(ps))) return rc;
    *out = stale_like_reference ? ps->stats_host->ess_stale : ps->stats_host->ess;
    return MPL_OK;
}

// The nested resampler posts its log total weight into mapped host memory as soon as the level-1 pass knows it (tagged
// words, nested.cuh: nested_post_to_host): poll that instead of draining the stream, so that the caller can queue the next
// step while the expansion kernel is still running.  Returns false if the words never showed up (then: synchronise).
static bool poll_host_lse(mpl_ps* ps, double* lse, int* degenerate, int* rc) {
    volatile unsigned long long* hm = reinterpret_cast<volatile unsigned long long*>(ps->host_flags) + 2;
    const unsigned long long seq = ps->host_seq;
    *rc = MPL_OK;
    for (unsigned long spins = 1;; ++spins) {
        const unsigned long long w0 = hm[0], w1 = hm[1], w2 = hm[2];
        if ((w0 >> 32) == seq && (w1 >> 32) == seq && (w2 >> 32) == seq) {
            const unsigned long long bits = (w0 & 0xffffffffull) | (w1 << 32);
            std::memcpy(lse, &bits, sizeof bits);
            *degenerate = (int)(w2 & 0xffffffffull);
            return true;
        }
        if ((spins & 0xfff) == 0) {   // every few microseconds: has the stream finished (or failed) without posting?
            const cudaError_t e = cudaStreamQuery(ps->stream);
            if (e == cudaSuccess) return false;
            if (e != cudaErrorNotReady) { *rc = fail(MPL_ERR_CUDA, std::string("resample: ") + cudaGetErrorString(e)); return false; }
        }
    }
}

extern "C" int mpl_ps_resample(mpl_ps* ps, int scheme, double* log_total_weight) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    ps->host_lse_posted = false;
    int rc = do_resample(ps, scheme);
    if (rc) return rc;
    if (log_total_weight) {
        double lse = 0.;
        int degenerate = 0;
        if (ps->host_lse_posted && !ps->profile && poll_host_lse(ps, &lse, &degenerate, &rc)) {
            *log_total_weight = lse;
            if (degenerate) return fail(MPL_ERR_DEGENERATE, "all particle weights are -inf");
            return MPL_OK;
        }
        if (rc) return rc;
        if ((rc = fetch_stats(ps))) return rc;
        *log_total_weight = ps->stats_host->lse;
        if (ps->stats_host->degenerate) return fail(MPL_ERR_DEGENERATE, "all particle weights are -inf");
    }
    return MPL_OK;
}

extern "C" int mpl_ps_log_marginal_likelihood_estimate(mpl_ps* ps, double* out) {
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc;
    if (!ps->pending_gather && (rc = ensure_stats(ps))) return rc;
    if ((rc = fetch_stats(ps))) return rc;
    const DeviceStats& s = *ps->stats_host;
    // particle_filter.rs:119-121 ; after a resample the weights are all zero: logsumexp(0..0) - ln N = 0
    *out = ps->pending_gather ? s.lml_acc : s.lml_acc + (s.max + std::log(s.sumexp)) - std::log((double)ps->n_global);
    return MPL_OK;
}

static int ensure_staging(mpl_ps* ps) {
    if (!ps->staging) MPL_CUDA_OK(cudaMalloc(&ps->staging, (size_t)ps->D * ps->ld * sizeof(double)));
    return MPL_OK;
}

extern "C" int mpl_ps_read(mpl_ps* ps, int what, void* host_dst, size_t bytes) {
    if (!ps || !host_dst) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc;
    const int grid = grid_for(ps->n, 256, kNumSMs * 8);
    if (what == MPL_READ_PARENTS) {
        if (bytes != ps->n * sizeof(int64_t)) return fail(MPL_ERR_INVALID, "parents buffer must be int64[N]");
        if ((rc = ensure_staging(ps))) return rc;
        i32_to_i64_kernel<<<grid, 256, 0, ps->stream>>>(ps->anc, (long long*)ps->staging, ps->n);
        MPL_CUDA_OK(cudaGetLastError());
        MPL_CUDA_OK(cudaMemcpyAsync(host_dst, ps->staging, bytes, cudaMemcpyDeviceToHost, ps->stream));
        MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
        return MPL_OK;
    }
    if ((rc = materialise(ps))) return rc;
    if ((rc = ensure_staging(ps))) return rc;
    if (what == MPL_READ_LOG_WEIGHTS) {
        if (bytes != ps->n * sizeof(double)) return fail(MPL_ERR_INVALID, "log-weight buffer must be double[N]");
        if (ps->dtype == MPL_F32) to_f64_kernel<float><<<grid, 256, 0, ps->stream>>>((const float*)ps->lw, ps->staging, ps->n);
        else to_f64_kernel<double><<<grid, 256, 0, ps->stream>>>((const double*)ps->lw, ps->staging, ps->n);
        MPL_CUDA_OK(cudaGetLastError());
        MPL_CUDA_OK(cudaMemcpyAsync(host_dst, ps->staging, bytes, cudaMemcpyDeviceToHost, ps->stream));
    } else if (what == MPL_READ_STATE) {
        if (bytes != (size_t)ps->D * ps->n * sizeof(double)) return fail(MPL_ERR_INVALID, "state buffer must be double[D*N]");
        if (ps->dtype == MPL_F32) state_to_f64_kernel<float><<<grid, 256, 0, ps->stream>>>((const float*)ps->state[ps->cur], ps->staging, ps->n, ps->ld, ps->D);
        else state_to_f64_kernel<double><<<grid, 256, 0, ps->stream>>>((const double*)ps->state[ps->cur], ps->staging, ps->n, ps->ld, ps->D);
        for (int d = 0; d < ps->D; ++d)
            MPL_CUDA_OK(cudaMemcpyAsync((double*)host_dst + (size_t)d * ps->n, ps->staging + (size_t)d * ps->ld, ps->n * sizeof(double), cudaMemcpyDeviceToHost, ps->stream));
        MPL_CUDA_OK(cudaGetLastError());
    } else return fail(MPL_ERR_INVALID, "unknown read selector");
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    return MPL_OK;
}

extern "C" int mpl_ps_write(mpl_ps* ps, int what, const void* host_src, size_t bytes) {
    if (!ps || !host_src) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc;
    if ((rc = materialise(ps))) return rc;
    if ((rc = ensure_staging(ps))) return rc;
    const int grid = grid_for(ps->n, 256, kNumSMs * 8);
    if (what == MPL_READ_LOG_WEIGHTS) {
        if (bytes != ps->n * sizeof(double)) return fail(MPL_ERR_INVALID, "log-weight buffer must be double[N]");
        MPL_CUDA_OK(cudaMemcpyAsync(ps->staging, host_src, bytes, cudaMemcpyHostToDevice, ps->stream));
        if (ps->dtype == MPL_F32) from_f64_kernel<float><<<grid, 256, 0, ps->stream>>>(ps->staging, (float*)ps->lw, ps->n);
        else from_f64_kernel<double><<<grid, 256, 0, ps->stream>>>(ps->staging, (double*)ps->lw, ps->n);
        ps->stats_valid = false; ps->max_valid = false;
    } else if (what == MPL_READ_STATE) {
        if (bytes != (size_t)ps->D * ps->n * sizeof(double)) return fail(MPL_ERR_INVALID, "state buffer must be double[D*N]");
        for (int d = 0; d < ps->D; ++d)
            MPL_CUDA_OK(cudaMemcpyAsync(ps->staging + (size_t)d * ps->ld, (const double*)host_src + (size_t)d * ps->n, ps->n * sizeof(double), cudaMemcpyHostToDevice, ps->stream));
        if (ps->dtype == MPL_F32) state_from_f64_kernel<float><<<grid, 256, 0, ps->stream>>>(ps->staging, (float*)ps->state[ps->cur], ps->n, ps->ld, ps->D);
        else state_from_f64_kernel<double><<<grid, 256, 0, ps->stream>>>(ps->staging, (double*)ps->state[ps->cur], ps->n, ps->ld, ps->D);
        if (!ps->initialised) { ps->initialised = true; if (ps->t == 0) ps->t = 1; }
    } else return fail(MPL_ERR_INVALID, "unknown write selector");
    MPL_CUDA_OK(cudaGetLastError());
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    return MPL_OK;
}

// ---- checkpoint / resume ----------------------------------------------------------------------------------
// A checkpoint is the filter's whole resumable state in its native precision: header, live state (D x ld), log-weights (ld).
// A pending resample is applied first (the gather it stands for gives the same particles as the fused gather of the next
// step would), so no ancestors need to be kept.  Single GPU; the trajectory log is not part of it.
namespace {
struct CkptHeader {
    uint64_t magic, n, ld, n_global, seed, gid_offset;
    int32_t D, dtype, model_kind, pad;
    int64_t t;
    double lml_acc, ess, ess_stale, lse;
    uint64_t n_resamples;
    uint64_t param_hash;   // FNV-1a over the model's parameter vector: a checkpoint continues the SAME filter only
};
constexpr uint64_t kCkptMagic = 0x334b434c504d6f6dull;   // "moMPLCK3" (3: particle-major state)
uint64_t model_param_hash(const mpl_model& m) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (double v : m.params) {
        unsigned char b[8];
        std::memcpy(b, &v, 8);
        for (int i = 0; i < 8; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    }
    return h;
}
size_t ckpt_bytes(const mpl_ps* ps) { return sizeof(CkptHeader) + ((size_t)ps->D + 1) * ps->ld * (ps->dtype == MPL_F32 ? 4 : 8); }
}  // namespace

extern "C" int mpl_ps_checkpoint_size(mpl_ps* ps, uint64_t* bytes) {
    if (!ps || !bytes) return fail(MPL_ERR_INVALID, "null argument");
    *bytes = ckpt_bytes(ps);
    return MPL_OK;
}

extern "C" int mpl_ps_checkpoint(mpl_ps* ps, void* dst, uint64_t bytes) {
    if (!ps || !dst) return fail(MPL_ERR_INVALID, "null argument");
    if (!ps->initialised) return fail(MPL_ERR_INVALID, "checkpoint before init_step");
    if (ps->world > 1) return fail(MPL_ERR_UNSUPPORTED, "checkpoint: single GPU only");
    if (bytes != ckpt_bytes(ps)) return fail(MPL_ERR_INVALID, "checkpoint buffer size (see mpl_ps_checkpoint_size)");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc;
    if ((rc = materialise(ps))) return rc;
    if ((rc = fetch_stats(ps))) return rc;
    CkptHeader h;
    std::memset(&h, 0, sizeof h);
    h.magic = kCkptMagic; h.n = ps->n; h.ld = ps->ld; h.n_global = ps->n_global; h.seed = ps->seed; h.gid_offset = ps->gid_offset;
    h.D = ps->D; h.dtype = ps->dtype; h.model_kind = (int32_t)ps->model.kind; h.t = ps->t;
    h.lml_acc = ps->stats_host->lml_acc; h.ess = ps->stats_host->ess; h.ess_stale = ps->stats_host->ess_stale; h.lse = ps->stats_host->lse;
    h.n_resamples = ps->stats_host->n_resamples;
    h.param_hash = model_param_hash(ps->model);
    std::memcpy(dst, &h, sizeof h);
    const size_t es = elem_size(ps);
    char* out = (char*)dst + sizeof h;
    MPL_CUDA_OK(cudaMemcpyAsync(out, ps->state[ps->cur], (size_t)ps->D * ps->ld * es, cudaMemcpyDeviceToHost, ps->stream));
    MPL_CUDA_OK(cudaMemcpyAsync(out + (size_t)ps->D * ps->ld * es, ps->lw, ps->ld * es, cudaMemcpyDeviceToHost, ps->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    return MPL_OK;
}

extern "C" int mpl_ps_restore(mpl_ps* ps, const void* src, uint64_t bytes) {
    if (!ps || !src) return fail(MPL_ERR_INVALID, "null argument");
    if (ps->world > 1) return fail(MPL_ERR_UNSUPPORTED, "restore: single GPU only");
    if (bytes != ckpt_bytes(ps)) return fail(MPL_ERR_INVALID, "checkpoint does not fit this particle system (size)");
    CkptHeader h;
    std::memcpy(&h, src, sizeof h);
    if (h.magic != kCkptMagic || h.n != ps->n || h.ld != ps->ld || h.n_global != ps->n_global || h.D != ps->D || h.dtype != ps->dtype ||
        h.model_kind != (int32_t)ps->model.kind || h.gid_offset != ps->gid_offset)
        return fail(MPL_ERR_INVALID, "checkpoint does not fit this particle system (model, particle count, precision or shard differ)");
    if (h.param_hash != model_param_hash(ps->model)) return fail(MPL_ERR_INVALID, "checkpoint was taken with other model parameters: the continuation would be a different filter");
    if (h.t <= 0) return fail(MPL_ERR_INVALID, "checkpoint header: time index must be positive");
    if (h.seed != ps->seed) return fail(MPL_ERR_INVALID, "checkpoint was taken with another seed: the continuation would not reproduce the original run");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc;
    if ((rc = fetch_stats(ps))) return rc;   // (also drains the stream)
    const size_t es = elem_size(ps);
    const char* in = (const char*)src + sizeof h;
    MPL_CUDA_OK(cudaMemcpyAsync(ps->state[ps->cur], in, (size_t)ps->D * ps->ld * es, cudaMemcpyHostToDevice, ps->stream));
    MPL_CUDA_OK(cudaMemcpyAsync(ps->lw, in + (size_t)ps->D * ps->ld * es, ps->ld * es, cudaMemcpyHostToDevice, ps->stream));
    DeviceStats& st = *ps->stats_host;
    st.lml_acc = h.lml_acc; st.ess = h.ess; st.ess_stale = h.ess_stale; st.lse = h.lse; st.n_resamples = h.n_resamples; st.t = h.t;
    st.resampled = 0; st.resampled_flag[0] = st.resampled_flag[1] = 0; st.degenerate = 0; st.do_resample = 0;
    st.max_bits[0] = st.max_bits[1] = 0ull; st.blocks_done = 0; st.ticket = 0; st.overflow_count = 0;
    MPL_CUDA_OK(cudaMemcpyAsync(ps->stats, ps->stats_host, sizeof(DeviceStats), cudaMemcpyHostToDevice, ps->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    ps->t = h.t; ps->initialised = true; ps->pending_gather = false; ps->stats_valid = false; ps->max_valid = false;
    ps->prequantised = 0; ps->dynamic_state_known = true;   // (an ESS-triggered run may continue from here: nothing is pending)
    ps->hist_broken = ps->hist_cap != 0;                    // the trajectory log does not describe the restored population
    return MPL_OK;
}

extern "C" int mpl_ps_history_enable(mpl_ps* ps, uint64_t max_steps) {
    if (!ps || max_steps == 0) return fail(MPL_ERR_INVALID, "bad argument");
    if (ps->world > 1) return fail(MPL_ERR_UNSUPPORTED, "trajectory log: single GPU only");
    if (ps->initialised) return fail(MPL_ERR_INVALID, "enable the trajectory log before init_step");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    cudaFree(ps->hist_state); cudaFree(ps->hist_anc);
    ps->hist_state = nullptr; ps->hist_anc = nullptr; ps->hist_cap = 0;
    const size_t es = elem_size(ps);
    cudaError_t e = cudaMalloc(&ps->hist_state, (size_t)max_steps * ps->D * ps->ld * es);
    if (e == cudaSuccess) e = cudaMalloc(&ps->hist_anc, (size_t)max_steps * ps->ld * sizeof(int32_t));
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(ps->hist_state); ps->hist_state = nullptr; return fail(MPL_ERR_CUDA, "trajectory log does not fit in device memory (it is max_steps x (D+1) x N)"); }
    ps->hist_cap = (size_t)max_steps;
    ps->hist_resampled.clear();
    return MPL_OK;
}

extern "C" int mpl_ps_trajectories(mpl_ps* ps, const int64_t* ids, uint64_t n_ids, double* out, size_t bytes, uint64_t* n_steps) {
    // out[k][t][d]: the state at step t of the lineage of particle ids[k] (`traces[ids[k]].retv`, dynunfold.rs:91)
    if (!ps || !ids || !out) return fail(MPL_ERR_INVALID, "null argument");
    if (!ps->hist_cap) return fail(MPL_ERR_INVALID, "trajectory log not enabled");
    if (ps->hist_broken) return fail(MPL_ERR_UNSUPPORTED, "trajectory log: not after an ESS-triggered device loop or a restore (the log does not cover those steps)");
    const size_t T = (size_t)ps->t;
    if (T == 0 || T > ps->hist_cap) return fail(MPL_ERR_INVALID, "no logged steps, or more steps than the log holds");
    if (bytes != n_ids * T * ps->D * sizeof(double)) return fail(MPL_ERR_INVALID, "trajectory buffer must be double[n_ids * T * D]");
    for (uint64_t k = 0; k < n_ids; ++k) if (ids[k] < 0 || (uint64_t)ids[k] >= ps->n) return fail(MPL_ERR_INVALID, "particle id out of range");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    long long* dids = nullptr; int* dres = nullptr; double* dout = nullptr;
    std::vector<int> res(ps->hist_resampled);
    res.resize(T, 0);
    MPL_CUDA_OK(cudaMalloc(&dids, n_ids * 8));
    MPL_CUDA_OK(cudaMalloc(&dres, T * sizeof(int)));
    MPL_CUDA_OK(cudaMalloc(&dout, bytes));
    MPL_CUDA_OK(cudaMemcpyAsync(dids, ids, n_ids * 8, cudaMemcpyHostToDevice, ps->stream));
    MPL_CUDA_OK(cudaMemcpyAsync(dres, res.data(), T * sizeof(int), cudaMemcpyHostToDevice, ps->stream));
    // ids name the particles as they are NOW: after a resample that followed the last step they are post-resample particles,
    // whether or not the gather has been applied yet (the next extend resets the entry)
    const int after = res[T - 1] ? 1 : 0;
    const int grid = grid_for(n_ids, 128, kNumSMs * 8);
    if (ps->dtype == MPL_F32) backtrace_kernel<float><<<grid, 128, 0, ps->stream>>>((const float*)ps->hist_state, ps->hist_anc, dres, ps->ld, ps->D, (int)T, after, dids, n_ids, dout);
    else backtrace_kernel<double><<<grid, 128, 0, ps->stream>>>((const double*)ps->hist_state, ps->hist_anc, dres, ps->ld, ps->D, (int)T, after, dids, n_ids, dout);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ps->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ps->stream);
    cudaFree(dids); cudaFree(dres); cudaFree(dout);
    if (e != cudaSuccess) return fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    if (n_steps) *n_steps = T;
    return MPL_OK;
}

extern "C" int mpl_ps_num_particles(const mpl_ps* ps, uint64_t* out) {
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    *out = ps->n;
    return MPL_OK;
}

extern "C" int mpl_ps_sync(mpl_ps* ps) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    return MPL_OK;
}

extern "C" int mpl_ps_upload_observations(mpl_ps* ps, const double* obs, size_t n_steps, size_t n_obs) {
    if (!ps || !obs) return fail(MPL_ERR_INVALID, "null argument");
    if (n_obs != (size_t)ps->model.obs_dim) return fail(MPL_ERR_INVALID, "n_obs must equal the model's observation dimension");
    if (ps->model.kind == M_HMM) {   // the kernel indexes its emission table with the symbol (models.cuh): same check as pack_obs
        const double M = ps->model.params[1];
        for (size_t i = 0; i < n_steps * n_obs; ++i)
            if (!(obs[i] >= 0. && obs[i] < M) || obs[i] != std::floor(obs[i])) return fail(MPL_ERR_INVALID, "hmm observation symbol out of range (step " + std::to_string(i) + ")");
    }
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    cudaFree(ps->obs_dev); ps->obs_dev = nullptr;
    MPL_CUDA_OK(cudaMalloc(&ps->obs_dev, n_steps * n_obs * sizeof(double)));
    MPL_CUDA_OK(cudaMemcpy(ps->obs_dev, obs, n_steps * n_obs * sizeof(double), cudaMemcpyHostToDevice));
    ps->obs_steps = n_steps;
    return MPL_OK;
}

extern "C" int mpl_ps_run(mpl_ps* ps, size_t first_step, size_t n_steps, int scheme, double ess_threshold, float* elapsed_ms) {
    // n_steps x (step(obs[t]); resample) entirely on the device: observation t is read from HBM by the extend kernel.
    // first_step == 0 starts with init_step(obs[0]) followed by a resample, like tests/smc.rs:63-70.
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    if (!ps->obs_dev) return fail(MPL_ERR_INVALID, "mpl_ps_upload_observations first");
    if (first_step + n_steps > ps->obs_steps) return fail(MPL_ERR_INVALID, "run exceeds the uploaded observations");
    if (first_step > 0 && (long long)first_step != ps->t) return fail(MPL_ERR_INVALID, "first_step must equal the filter's current time index");
    if (first_step == 0 && ps->world > 1 && ps->initialised)
        return fail(MPL_ERR_UNSUPPORTED, "sharded particle system: one run from step 0 per attach (mailbox words are validated by step number and would look current)");
    const bool dynamic = ess_threshold > 0.;
    int rc0 = MPL_OK;
    if (dynamic && scheme != MPL_RESAMPLE_SYSTEMATIC_FIXED && scheme != MPL_RESAMPLE_SYSTEMATIC_NESTED)
        return fail(MPL_ERR_UNSUPPORTED, "ESS-triggered device loop: MPL_RESAMPLE_SYSTEMATIC_FIXED or MPL_RESAMPLE_SYSTEMATIC_NESTED");
    const bool dyn_nested = dynamic && scheme == MPL_RESAMPLE_SYSTEMATIC_NESTED;
    const int dyn_records = dyn_nested && ps->dtype == MPL_F32 ? 2 : 0;   // fp32: the extend leaves the chunk records (log-weights kept)
    if (dyn_nested && (rc0 = ensure_chunk_records(ps))) return rc0;
    if (dynamic && first_step > 0 && !ps->dynamic_state_known) return fail(MPL_ERR_INVALID, "ESS-triggered run must start at step 0 or continue a previous ESS-triggered run");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    ps->ess_threshold_abs = dynamic ? ess_threshold * (double)ps->n_global : 0.;
    const bool fuse_nested = !dynamic && scheme == MPL_RESAMPLE_SYSTEMATIC_NESTED && ps->dtype == MPL_F32 && !(ps->world > 1 && (ps->n % kChunk));
    if (fuse_nested && (rc0 = ensure_chunk_records(ps))) return rc0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (elapsed_ms) { MPL_CUDA_OK(cudaEventCreate(&e0)); MPL_CUDA_OK(cudaEventCreate(&e1)); MPL_CUDA_OK(cudaEventRecord(e0, ps->stream)); }
    Obs dummy; std::memset(&dummy, 0, sizeof dummy);
    int rc = MPL_OK;
    struct LoopFlag { mpl_ps* p; explicit LoopFlag(mpl_ps* q) : p(q) { p->in_device_loop = true; } ~LoopFlag() { p->in_device_loop = false; } } loop_flag(ps);
    for (size_t k = 0; k < n_steps && rc == MPL_OK; ++k) {
        size_t tt = first_step + k;
        if (tt == 0) {
            ps->t = 0; ps->pending_gather = false;
            rc = launch_extend(ps, EXT_INIT, dummy, true, false, fuse_nested ? 1 : dyn_records);
            ps->t = 1; ps->initialised = true; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
        } else if (dynamic) {
            // whether the previous step resampled is only known on the device: the extend reads stats->resampled_flag[t & 1]
            rc = launch_extend(ps, EXT_DYNAMIC, dummy, true, false, dyn_records);
            ps->pending_gather = false; ps->t += 1; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
        } else {
            rc = launch_extend(ps, ps->pending_gather ? EXT_GATHER : EXT_ACCUM, dummy, true, false, fuse_nested);
            ps->pending_gather = false; ps->t += 1; ps->stats_valid = false; ps->max_valid = !ps->prequantised;
        }
        if (rc != MPL_OK) break;
        if (dyn_nested) {
            rc = ps->dtype == MPL_F32 ? resample_nested_t<float>(ps, 3, true) : resample_nested_t<double>(ps, 3, true);
        } else if (dynamic) {
            rc = ps->dtype == MPL_F32 ? resample_fixed_t<float>(ps, scheme, true, false) : resample_fixed_t<double>(ps, scheme, true, false);
        } else rc = do_resample(ps, scheme);
    }
    if (elapsed_ms) {
        cudaEventRecord(e1, ps->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) cudaEventElapsedTime(elapsed_ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (e != cudaSuccess) return fail(MPL_ERR_CUDA, std::string("run: ") + cudaGetErrorString(e));
    }
    if (rc == MPL_OK && dynamic) {   // learn from the device whether the last step resampled
        if ((rc = fetch_stats(ps))) return rc;
        ps->pending_gather = ps->stats_host->resampled_flag[ps->t & 1] != 0;
        ps->stats_valid = false; ps->max_valid = !ps->pending_gather && !dyn_nested;
        ps->dynamic_state_known = true;
        ps->hist_broken = ps->hist_cap != 0;   // which steps resampled is only known on the device
    }
    return rc;
}

extern "C" int mpl_test_set_inline_level1(int mode) {
    if (mode < -1 || mode > 1) return fail(MPL_ERR_INVALID, "mode: -1 (by shard size), 0, 1");
    g_inline_level1 = mode;
    return MPL_OK;
}

extern "C" int mpl_ps_num_resamples(mpl_ps* ps, uint64_t* out) {
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    int rc = fetch_stats(ps);
    if (rc) return rc;
    *out = ps->stats_host->n_resamples;
    return MPL_OK;
}

extern "C" int mpl_ps_profile_enable(mpl_ps* ps, int on) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    int rc = flush_timers(ps);
    ps->profile = on != 0;
    if (on) ps->timers.clear();
    return rc;
}
extern "C" int mpl_ps_profile_get(mpl_ps* ps, const char* kernel, double* total_ms, uint64_t* launches) {
    if (!ps || !kernel) return fail(MPL_ERR_INVALID, "null argument");
    int rc = flush_timers(ps);
    if (rc) return rc;
    auto it = ps->timers.find(kernel);
    if (total_ms) *total_ms = it == ps->timers.end() ? 0. : it->second.total_ms;
    if (launches) *launches = it == ps->timers.end() ? 0 : it->second.launches;
    return MPL_OK;
}
extern "C" int mpl_ps_launch_count(mpl_ps* ps, uint64_t* out) {
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    *out = ps->launch_count;
    return MPL_OK;
}

// ---- parity hooks -----------------------------------------------------------------------------------------------
static int require_device() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(MPL_ERR_CUDA, "no CUDA device: modppl_b200 has no CPU fallback"); }
    return MPL_OK;
}

extern "C" int mpl_cumsum_sequential(const double* probs, uint64_t n, double* out) {
    if (!probs || !out || n == 0) return fail(MPL_ERR_INVALID, "bad argument");
    int rc = require_device();
    if (rc) return rc;
    double *dp = nullptr, *ds = nullptr;
    MPL_CUDA_OK(cudaMalloc(&dp, n * 8));
    MPL_CUDA_OK(cudaMalloc(&ds, n * 8));
    MPL_CUDA_OK(cudaMemcpy(dp, probs, n * 8, cudaMemcpyHostToDevice));
    rc = launch_cumsum_exact(nullptr, dp, n, ds, 0);
    if (rc == MPL_OK) { cudaError_t e = cudaMemcpy(out, ds, n * 8, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = fail(MPL_ERR_CUDA, cudaGetErrorString(e)); }
    cudaFree(dp); cudaFree(ds);
    return rc;
}

extern "C" int mpl_resample_indices(const double* probs, const double* uniforms, uint64_t n, uint64_t n_draws, int scheme, int64_t* parents) {
    if (!probs || !uniforms || !parents || n == 0) return fail(MPL_ERR_INVALID, "bad argument");
    if (scheme != MPL_RESAMPLE_MULTINOMIAL && scheme != MPL_RESAMPLE_SYSTEMATIC) return fail(MPL_ERR_INVALID, "scheme must be MULTINOMIAL or SYSTEMATIC");
    int rc = require_device();
    if (rc) return rc;
    if (n_draws == 0) return MPL_OK;
    double *dp = nullptr, *ds = nullptr, *du = nullptr; long long* dpar = nullptr;
    size_t nu = scheme == MPL_RESAMPLE_MULTINOMIAL ? n_draws : 1;
    MPL_CUDA_OK(cudaMalloc(&dp, n * 8));
    MPL_CUDA_OK(cudaMalloc(&ds, n * 8));
    MPL_CUDA_OK(cudaMalloc(&du, nu * 8));
    MPL_CUDA_OK(cudaMalloc(&dpar, n_draws * 8));
    MPL_CUDA_OK(cudaMemcpy(dp, probs, n * 8, cudaMemcpyHostToDevice));
    MPL_CUDA_OK(cudaMemcpy(du, uniforms, nu * 8, cudaMemcpyHostToDevice));
    rc = launch_cumsum_exact(nullptr, dp, n, ds, 0);
    if (rc == MPL_OK) {
        search_kernel<long long><<<grid_for(n_draws, 256, kNumSMs * 8), 256>>>(ds, n, n_draws, scheme == MPL_RESAMPLE_MULTINOMIAL ? 0 : 1, du, 0, 0, 0, dpar);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(parents, dpar, n_draws * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(dp); cudaFree(ds); cudaFree(du); cudaFree(dpar);
    return rc;
}

extern "C" int mpl_logsumexp_stats(const void* lw, uint64_t n, int dtype, double* lse, double* ess, double* mx) {
    if (!lw || n == 0) return fail(MPL_ERR_INVALID, "bad argument");
    int rc = require_device();
    if (rc) return rc;
    size_t es = dtype == MPL_F64 ? 8 : 4;
    void* d = nullptr; DeviceStats* st = nullptr; Lse3<double>* part = nullptr;
    int grid = grid_for(n, 256 * 4, kNumSMs * 8);
    MPL_CUDA_OK(cudaMalloc(&d, n * es));
    MPL_CUDA_OK(cudaMalloc(&st, sizeof(DeviceStats)));
    MPL_CUDA_OK(cudaMalloc(&part, grid * sizeof(Lse3<double>)));
    MPL_CUDA_OK(cudaMemset(st, 0, sizeof(DeviceStats)));
    MPL_CUDA_OK(cudaMemcpy(d, lw, n * es, cudaMemcpyHostToDevice));
    if (dtype == MPL_F64) weight_reduce_kernel<double><<<grid, 256>>>((const double*)d, n, st, part);
    else weight_reduce_kernel<float><<<grid, 256>>>((const float*)d, n, st, part);
    DeviceStats h;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(&h, st, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d); cudaFree(st); cudaFree(part);
    if (e != cudaSuccess) return fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    if (lse) *lse = (h.max == -INFINITY) ? -INFINITY : h.max + std::log(h.sumexp);   // lib.rs:36-37
    if (ess) *ess = h.ess;
    if (mx) *mx = h.max;
    return MPL_OK;
}

extern "C" int mpl_fixed_resample(const float* lw, uint64_t n, int scheme, uint64_t rand_word_or_seed, uint32_t t, int32_t* anc, double* lse, uint64_t* total_weight) {
    // integer-weight resampling of injected f32 log-weights through a scratch particle system (D = 1)
    if (!lw || !anc || n == 0) return fail(MPL_ERR_INVALID, "bad argument");
    if (scheme != MPL_RESAMPLE_SYSTEMATIC_FIXED && scheme != MPL_RESAMPLE_MULTINOMIAL_FIXED && scheme != MPL_RESAMPLE_SYSTEMATIC_NESTED) return fail(MPL_ERR_INVALID, "scheme must be an integer-weight scheme");
    int rc = require_device();
    if (rc) return rc;
    double svp[3] = {0., 0.5, 1.};
    mpl_model* m = mpl_model_create("sv", svp, 3);
    mpl_pf_config c; c.dtype = MPL_F32; c.device = -1; c.seed = rand_word_or_seed; c.gid_offset = 0; c.n_global = 0;
    mpl_ps* ps = mpl_particle_system_new(m, n, &c);
    if (!ps) { mpl_model_destroy(m); return MPL_ERR_CUDA; }
    ps->initialised = true; ps->t = (long long)t + 1;
    cudaError_t e = cudaMemcpyAsync(ps->lw, lw, n * 4, cudaMemcpyHostToDevice, ps->stream);
    if (e == cudaSuccess) {
        ps->stats_valid = false;
        rc = do_resample(ps, scheme);
        if (rc == MPL_OK) rc = fetch_stats(ps);
        if (rc == MPL_OK) {
            e = cudaMemcpy(anc, ps->anc, n * 4, cudaMemcpyDeviceToHost);
            if (lse) *lse = ps->stats_host->lse;
            if (total_weight) *total_weight = ps->stats_host->W;
        }
    }
    if (e != cudaSuccess) rc = fail(MPL_ERR_CUDA, cudaGetErrorString(e));
    mpl_ps_destroy(ps); mpl_model_destroy(m);
    return rc;
}
