// multi_gpu.cu -- sharded particle systems: one process per GPU, peers reached over NVLink through CUDA-IPC mapped
// pointers.  No NCCL call sits on the data path: the per-step exchange is
//   (1) each shard's (max, sum exp, sum exp^2)  -> remote stores into every rank's mailbox from the extend epilogue,
//   (2) each shard's integer weight total       -> remote stores from the reduce pass' last block,
//   (3) ancestors that land in another shard    -> remote stores from the scan's expansion,
//   (4) parents that live in another shard      -> remote loads in the next extend's fused gather,
// with step-numbered flags that the consuming kernels spin on locally.  The caller only has to all-gather one
// MPL_PEER_BLOB_BYTES blob per rank once (bench.py does it with torch.distributed).
#include <cstddef>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "engine.h"

using namespace mpl;

namespace {
constexpr int kPeerHandles = 9;
struct PeerBlob {   // must fit MPL_PEER_BLOB_BYTES
    cudaIpcMemHandle_t h[kPeerHandles];   // state0, state1, anc, mailbox, lw pair, chunk-record pairs (e, S), tile-prefix pair: 9 x 64 bytes
};
void* const* peer_buffers(mpl_ps* ps, void* (&out)[kPeerHandles]) {
    out[0] = ps->state[0]; out[1] = ps->state[1]; out[2] = ps->anc; out[3] = ps->mailbox;
    out[4] = ps->par == 0 ? ps->lw : ps->lw_alt; out[5] = ps->par == 0 ? ps->lw_alt : ps->lw; out[6] = ps->rec_e2; out[7] = ps->rec_S2; out[8] = ps->nest_tile_pre;
    return out;
}
// fills rank h's column of the peer table from its 8 buffers (own pointers or IPC-mapped ones)
void set_peer_column(mpl_ps* ps, PeerTable& t, int h, void* const* b) {
    const size_t nch = ps->ld / kChunk;
    t.state[0][h] = b[0]; t.state[1][h] = b[1]; t.anc[h] = (int32_t*)b[2]; t.mail[h] = (Mailbox*)b[3];
    t.lw[0][h] = b[4]; t.lw[1][h] = b[5];
    const size_t n_tp = ((ps->ld + kSection - 1) / kSection) * kTilesPerSection;
    for (int p = 0; p < 2; ++p) {
        t.rec_e[p][h] = (const int*)b[6] + (size_t)p * nch; t.rec_S[p][h] = (const unsigned int*)b[7] + (size_t)p * nch;
        t.tile_pre[p][h] = (const unsigned long long*)b[8] + (size_t)p * n_tp;
    }
}
// the second log-weight buffer and the chunk records exist before anything is exported (sharded runs alternate them)
int ensure_pairs(mpl_ps* ps) {
    int rc = ensure_chunk_records(ps);
    if (rc) return rc;
    if (!ps->lw_alt) {
        const size_t bytes = ps->ld * (ps->dtype == MPL_F64 ? 8 : 4);
        MPL_CUDA_OK(cudaMalloc(&ps->lw_alt, bytes));
        MPL_CUDA_OK(cudaMemset(ps->lw_alt, 0, bytes));
    }
    return MPL_OK;
}
static_assert(sizeof(PeerBlob) <= MPL_PEER_BLOB_BYTES, "peer blob too large");

int ensure_mailbox(mpl_ps* ps) {
    if (ps->mailbox) return MPL_OK;
    MPL_CUDA_OK(cudaMalloc(&ps->mailbox, sizeof(Mailbox)));
    MPL_CUDA_OK(cudaMemset(ps->mailbox, 0, sizeof(Mailbox)));
    return MPL_OK;
}

int check_shard_shape(const mpl_ps* ps, int rank, int world) {
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(MPL_ERR_INVALID, "peer attach: world must be 1..8 and 0 <= rank < world");
    if (ps->n_global != ps->n * (uint64_t)world || ps->gid_offset != ps->n * (uint64_t)rank)
        return fail(MPL_ERR_INVALID, "peer attach: shards must be equal contiguous blocks (n_global == n * world, gid_offset == n * rank)");
    if (ps->n_global > (1ull << 31)) return fail(MPL_ERR_INVALID, "peer attach: at most 2^31 particles in total");
    return MPL_OK;
}

void fill_table_header(mpl_ps* ps, int rank, int world) {
    ps->rank = rank; ps->world = world;
    ps->peer.world = world; ps->peer.rank = rank;
    ps->peer.n_loc = (unsigned int)ps->n;
    int sh = -1;
    for (int b = 0; b < 32; ++b) if ((1ull << b) == ps->n) sh = b;
    ps->peer.shift = sh;
}
}  // namespace

extern "C" int mpl_ps_peer_export(mpl_ps* ps, void* blob) {
    if (!ps || !blob) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc = ensure_mailbox(ps);
    if (rc) return rc;
    // an export precedes the all-gather of the blobs, hence every peer's first store: a (re-)attach starts from a clean
    // mailbox (its words are validated by step number only)
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    MPL_CUDA_OK(cudaMemset(ps->mailbox, 0, sizeof(Mailbox)));
    if ((rc = ensure_pairs(ps))) return rc;
    PeerBlob b;
    std::memset(&b, 0, sizeof b);
    void* bufs[kPeerHandles];
    peer_buffers(ps, bufs);
    for (int k = 0; k < kPeerHandles; ++k) MPL_CUDA_OK(cudaIpcGetMemHandle(&b.h[k], bufs[k]));
    std::memset(blob, 0, MPL_PEER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    return MPL_OK;
}

extern "C" int mpl_ps_peer_attach(mpl_ps* ps, int rank, int world, const void* blobs) {
    if (!ps || !blobs) return fail(MPL_ERR_INVALID, "null argument");
    if (ps->world > 1) return fail(MPL_ERR_INVALID, "peer attach: already attached");
    int rc = check_shard_shape(ps, rank, world);
    if (rc) return rc;
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    if ((rc = ensure_mailbox(ps))) return rc;
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    // open every handle first (recording each as it is opened); only a complete table is committed
    PeerTable table = ps->peer;
    auto rollback = [&](cudaError_t e) {
        for (int k = 0; k < kPeerHandles; ++k)
            for (int h = 0; h < kMaxPeers; ++h)
                if (ps->ipc_opened[k][h]) { cudaIpcCloseMemHandle(ps->ipc_opened[k][h]); ps->ipc_opened[k][h] = nullptr; }
        cudaGetLastError();
        return fail(MPL_ERR_CUDA, std::string("peer attach: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    };
    if ((rc = ensure_pairs(ps))) return rc;
    for (int h = 0; h < world; ++h) {
        void* bufs[kPeerHandles];
        if (h == rank) { set_peer_column(ps, table, h, peer_buffers(ps, bufs)); continue; }
        PeerBlob b;
        std::memcpy(&b, (const char*)blobs + (size_t)h * MPL_PEER_BLOB_BYTES, sizeof b);
        for (int k = 0; k < kPeerHandles; ++k) {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, b.h[k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return rollback(e);
            ps->ipc_opened[k][h] = p;
            bufs[k] = p;
        }
        set_peer_column(ps, table, h, bufs);
    }
    ps->peer = table;
    fill_table_header(ps, rank, world);
    ps->peer_virtual = false;
    return MPL_OK;
}

extern "C" int mpl_ps_peer_detach(mpl_ps* ps) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    if (ps->world <= 1) return MPL_OK;
    cudaStreamSynchronize(ps->stream);
    if (!ps->peer_virtual)
        for (int k = 0; k < kPeerHandles; ++k)
            for (int h = 0; h < kMaxPeers; ++h)
                if (ps->ipc_opened[k][h]) { cudaIpcCloseMemHandle(ps->ipc_opened[k][h]); ps->ipc_opened[k][h] = nullptr; }
    std::memset(&ps->peer, 0, sizeof ps->peer);
    ps->peer.world = 1; ps->world = 1; ps->rank = 0; ps->anc_pushed = false;
    void* bufs[kPeerHandles];
    PeerTable t = ps->peer;
    set_peer_column(ps, t, 0, peer_buffers(ps, bufs));   // back to one GPU: the self-table of ensure_chunk_records
    for (int p = 0; p < 2; ++p) { t.lw[p][0] = ps->lw; t.rec_e[p][0] = ps->rec_e; t.rec_S[p][0] = ps->rec_S; t.tile_pre[p][0] = ps->nest_tile_pre + (size_t)ps->par * (((ps->ld + kSection - 1) / kSection) * kTilesPerSection); }
    t.n_loc = (unsigned int)ps->n;
    ps->peer = t;
    return MPL_OK;
}

// ---- islands (the local-resample variant of SURVEY 8e): every GPU runs its own filter on N / G particles and resamples locally;
// nothing is exchanged per step.  Now and then the host compares the islands' weights (their log-ML increments) and, when the
// island-level ESS has dropped, resamples whole islands: an island that is selected twice is copied over NVLink into the place of
// one that was not selected.  These calls are that copy; the orchestration is host code (modppl_b200/distributed.py).
namespace {
struct IslandBlob { cudaIpcMemHandle_t state[2]; };
int island_install(mpl_ps* dst, const void* src_state) {   // src_state: the source island's live state, D x ld elements, same shape
    int rc = MPL_OK;
    const size_t bytes = (size_t)dst->D * dst->ld * (dst->dtype == MPL_F64 ? 8 : 4);
    MPL_CUDA_OK(cudaMemcpyAsync(dst->state[dst->cur ^ 1], src_state, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    MPL_CUDA_OK(cudaMemsetAsync(dst->lw, 0, dst->ld * (dst->dtype == MPL_F64 ? 8 : 4), dst->stream));
    MPL_CUDA_OK(cudaMemsetAsync((char*)dst->stats + offsetof(DeviceStats, resampled_flag), 0, sizeof(int) * 2, dst->stream));
    MPL_CUDA_OK(cudaMemsetAsync((char*)dst->stats + offsetof(DeviceStats, resampled), 0, sizeof(int), dst->stream));
    MPL_CUDA_OK(cudaStreamSynchronize(dst->stream));
    dst->cur ^= 1;
    dst->pending_gather = false; dst->stats_valid = false; dst->max_valid = false; dst->prequantised = 0;
    return rc;
}
}  // namespace

extern "C" int mpl_ps_island_export(mpl_ps* ps, void* blob) {
    if (!ps || !blob) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    IslandBlob b;
    std::memset(&b, 0, sizeof b);
    MPL_CUDA_OK(cudaIpcGetMemHandle(&b.state[0], ps->state[0]));
    MPL_CUDA_OK(cudaIpcGetMemHandle(&b.state[1], ps->state[1]));
    std::memset(blob, 0, MPL_PEER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    return MPL_OK;
}

extern "C" int mpl_ps_island_attach(mpl_ps* ps, int rank, int n_islands, const void* blobs) {
    if (!ps || !blobs) return fail(MPL_ERR_INVALID, "null argument");
    if (n_islands < 1 || n_islands > kMaxPeers || rank < 0 || rank >= n_islands) return fail(MPL_ERR_INVALID, "island attach: 1..8 islands, 0 <= rank < islands");
    if (ps->world > 1) return fail(MPL_ERR_INVALID, "island attach: this particle system is a shard of a global one");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    for (int h = 0; h < n_islands; ++h) {
        if (h == rank) { ps->island_state[0][h] = ps->state[0]; ps->island_state[1][h] = ps->state[1]; continue; }
        IslandBlob b;
        std::memcpy(&b, (const char*)blobs + (size_t)h * MPL_PEER_BLOB_BYTES, sizeof b);
        for (int k = 0; k < 2; ++k) {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, b.state[k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); return fail(MPL_ERR_CUDA, std::string("island attach: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
            ps->island_state[k][h] = p;
            ps->island_opened[k][h] = p;
        }
    }
    ps->n_islands = n_islands; ps->island_rank = rank;
    return MPL_OK;
}

extern "C" int mpl_ps_live_buffer(mpl_ps* ps, int* out) {   // which of the two state buffers is live, with any pending resample applied
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    int rc = materialise(ps);   // a pending resample is applied (one gather on the device)
    if (rc) return rc;
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    *out = ps->cur;
    return MPL_OK;
}

extern "C" int mpl_ps_island_copy_from(mpl_ps* ps, int src_island, int src_live_buffer) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    if (src_island < 0 || src_island >= ps->n_islands || (src_live_buffer != 0 && src_live_buffer != 1)) return fail(MPL_ERR_INVALID, "island copy: bad source");
    if (src_island == ps->island_rank) return MPL_OK;
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    return island_install(ps, ps->island_state[src_live_buffer][src_island]);
}

extern "C" int mpl_ps_copy_state(mpl_ps* dst, mpl_ps* src) {   // islands living in one process (tests): same device, same shape
    if (!dst || !src) return fail(MPL_ERR_INVALID, "null handle");
    if (dst->n != src->n || dst->D != src->D || dst->dtype != src->dtype || dst->ld != src->ld) return fail(MPL_ERR_INVALID, "copy_state: the two particle systems differ in shape");
    int live = 0, rc = mpl_ps_live_buffer(src, &live);
    if (rc) return rc;
    return island_install(dst, src->state[live]);
}

extern "C" int mpl_ps_trace(mpl_ps* ps, long long* out16) {
    // %globaltimer stamps (ns) of the last step: [0] extend gate entered, [1] passed, [2] extend's last block done,
    // [3] reduce gate entered, [4] passed, [5] reduce's last block done, [6] scan gate entered, [7] passed, [8] done flag sent
    if (!ps || !out16) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    DeviceStats h;
    MPL_CUDA_OK(cudaMemcpy(&h, ps->stats, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; ++i) out16[i] = h.trace[i];
    return MPL_OK;
}

// device-side rendezvous of all ranks' streams (no host involved): every rank raises its own barrier word and polls the peers'.
// Queued in front of a timed run it makes all GPUs start the run together -- the host-side barrier before it releases the ranks
// hundreds of microseconds apart, which a loop of 50 us steps that meets at a gate every step would otherwise count as step time.
static __global__ void peer_barrier_kernel(PeerTable p, unsigned long long seq) {
    if (threadIdx.x == 0) {
        *(volatile unsigned long long*)&p.mail[p.rank]->barrier_seq = seq;
        SpinGuard g(p);
        for (int h = 0; h < p.world; ++h)
            while (*(volatile unsigned long long*)&p.mail[h]->barrier_seq < seq) if (g.give_up()) return;
    }
}
extern "C" int mpl_ps_peer_barrier(mpl_ps* ps) {
    if (!ps) return fail(MPL_ERR_INVALID, "null handle");
    if (ps->world <= 1 || ps->peer_virtual) return MPL_OK;
    MPL_CUDA_OK(cudaSetDevice(ps->device));
    ps->barrier_seq += 1;
    peer_barrier_kernel<<<1, 32, 0, ps->stream>>>(ps->peer, ps->barrier_seq);
    MPL_CUDA_OK(cudaGetLastError());
    return MPL_OK;
}

extern "C" int mpl_ps_nvlink_bytes(mpl_ps* ps, uint64_t* out) {
    // payload bytes this GPU has requested from its peers' memory so far: parent states gathered across a shard edge, integer
    // weights and chunk records of chunks that own some of this GPU's slots, the peers' section records
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    DeviceStats h;
    MPL_CUDA_OK(cudaMemcpy(&h, ps->stats, sizeof h, cudaMemcpyDeviceToHost));
    unsigned long long polled = 0;
    if (ps->mailbox) MPL_CUDA_OK(cudaMemcpy(&polled, (const char*)ps->mailbox + offsetof(Mailbox, nvlink_polled), sizeof polled, cudaMemcpyDeviceToHost));
    *out = h.nvlink_bytes + polled;
    return MPL_OK;
}

extern "C" int mpl_ps_peer_error(mpl_ps* ps, int* out) {
    if (!ps || !out) return fail(MPL_ERR_INVALID, "null argument");
    *out = 0;
    if (!ps->mailbox) return MPL_OK;
    MPL_CUDA_OK(cudaStreamSynchronize(ps->stream));
    MPL_CUDA_OK(cudaMemcpy(out, (const char*)ps->mailbox + offsetof(Mailbox, error), sizeof(int), cudaMemcpyDeviceToHost));
    return MPL_OK;
}

// ---- shards emulated on ONE GPU (test hook) ---------------------------------------------------------------------------
// `world` particle systems in this process share the device; their peer tables point straight at each other's
// buffers and every phase of a step is issued for all shards before the next phase starts, so no kernel ever has to
// wait for one that has not been launched (B200_PROFILING.md: never co-schedule mutually waiting kernels on one GPU).
// The kernels executed are exactly the multi-GPU ones (remote loads/stores become local ones).
extern "C" int mpl_test_virtual_shards(const mpl_model* model, uint64_t n_global, int world, int dtype, uint64_t seed, const double* obs, size_t n_steps,
                                       size_t n_obs, double* state_out, double* lw_out, double* lml_out, double* loop_ms) {
    return mpl_test_virtual_shards_scheme(model, n_global, world, dtype, seed, MPL_RESAMPLE_SYSTEMATIC_FIXED, obs, n_steps, n_obs, state_out, lw_out, lml_out, loop_ms);
}

extern "C" int mpl_test_virtual_shards_scheme(const mpl_model* model, uint64_t n_global, int world, int dtype, uint64_t seed, int scheme, const double* obs,
                                              size_t n_steps, size_t n_obs, double* state_out, double* lw_out, double* lml_out, double* loop_ms) {
    if (scheme != MPL_RESAMPLE_SYSTEMATIC_FIXED && scheme != MPL_RESAMPLE_SYSTEMATIC_NESTED) return fail(MPL_ERR_INVALID, "scheme: single-level or nested systematic");
    const bool nested = scheme == MPL_RESAMPLE_SYSTEMATIC_NESTED;
    const bool fuse = nested && dtype == MPL_F32;   // the extend kernel quantises in its epilogue, like mpl_ps_run
    // runs init, resample, then (n_steps - 1) x (step, resample) except that the last step is NOT followed by a
    // resample; returns the final state [D * n_global], log-weights [n_global] and the log-ML estimate.
    if (!model || !obs || n_steps < 1 || world < 1 || world > kMaxPeers || n_global % world) return fail(MPL_ERR_INVALID, "bad argument");
    std::vector<mpl_ps*> sh(world, nullptr);
    const uint64_t n_loc = n_global / world;
    int rc = MPL_OK;
    for (int g = 0; g < world && rc == MPL_OK; ++g) {
        mpl_pf_config c; c.dtype = dtype; c.device = -1; c.seed = seed; c.gid_offset = n_loc * g; c.n_global = n_global;
        sh[g] = mpl_particle_system_new(model, n_loc, &c);
        if (!sh[g]) rc = MPL_ERR_CUDA;
        else rc = mpl_ps_upload_observations(sh[g], obs, n_steps, n_obs);
        if (rc == MPL_OK) rc = ensure_mailbox(sh[g]);
        if (rc == MPL_OK && world > 1) rc = ensure_pairs(sh[g]);
    }
    if (rc == MPL_OK && world > 1) {
        for (int g = 0; g < world; ++g) {
            fill_table_header(sh[g], g, world);
            sh[g]->peer_virtual = true;
            for (int h = 0; h < world; ++h) {
                void* bufs[kPeerHandles];
                set_peer_column(sh[g], sh[g]->peer, h, peer_buffers(sh[h], bufs));
            }
        }
    }
    auto wall0 = std::chrono::steady_clock::now();
    for (size_t t = 0; t < n_steps && rc == MPL_OK; ++t) {
        const bool last = t + 1 == n_steps;   // (the last step keeps its log-weights: no fused quantisation)
        for (int g = 0; g < world && rc == MPL_OK; ++g) { rc = ps_phase_extend(sh[g], t == 0, fuse && !last); if (rc == MPL_OK) rc = mpl_ps_sync(sh[g]); }
        if (last) break;
        for (int g = 0; g < world && rc == MPL_OK; ++g) { rc = nested ? ps_phase_nested(sh[g], 1) : ps_phase_reduce(sh[g]); if (rc == MPL_OK) rc = mpl_ps_sync(sh[g]); }
        for (int g = 0; g < world && rc == MPL_OK; ++g) { rc = nested ? ps_phase_nested(sh[g], 2) : ps_phase_scan(sh[g]); if (rc == MPL_OK) rc = mpl_ps_sync(sh[g]); }
    }
    if (loop_ms) *loop_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    const int D = model->state_dim;
    std::vector<double> tmp;
    for (int g = 0; g < world && rc == MPL_OK; ++g) {
        int err = 0;
        mpl_ps_peer_error(sh[g], &err);
        if (err) { rc = fail(MPL_ERR_CUDA, "peer wait timed out"); break; }
        tmp.resize((size_t)D * n_loc);
        rc = mpl_ps_read(sh[g], MPL_READ_STATE, tmp.data(), tmp.size() * 8);
        if (rc) break;
        for (int d = 0; d < D; ++d) std::memcpy(state_out + (size_t)d * n_global + n_loc * g, tmp.data() + (size_t)d * n_loc, n_loc * 8);
        rc = mpl_ps_read(sh[g], MPL_READ_LOG_WEIGHTS, lw_out + n_loc * g, n_loc * 8);
    }
    if (rc == MPL_OK && lml_out) {
        // particle_filter.rs:119-121 over all shards: lml_acc + logsumexp(all log-weights) - ln N
        DeviceStats h;
        cudaMemcpy(&h, sh[0]->stats, sizeof h, cudaMemcpyDeviceToHost);
        double mx = -INFINITY;
        for (uint64_t i = 0; i < n_global; ++i) mx = std::fmax(mx, lw_out[i]);
        double s = 0.;
        for (uint64_t i = 0; i < n_global; ++i) s += std::exp(lw_out[i] - mx);
        *lml_out = h.lml_acc + mx + std::log(s) - std::log((double)n_global);
    }
    for (int g = 0; g < world; ++g) if (sh[g]) { sh[g]->world = 1; mpl_ps_destroy(sh[g]); }
    return rc;
}
