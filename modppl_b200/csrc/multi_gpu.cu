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
struct PeerBlob {   // must fit MPL_PEER_BLOB_BYTES
    cudaIpcMemHandle_t state0, state1, anc, mail;   // 4 x 64 bytes
};
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
    PeerBlob b;
    std::memset(&b, 0, sizeof b);
    MPL_CUDA_OK(cudaIpcGetMemHandle(&b.state0, ps->state[0]));
    MPL_CUDA_OK(cudaIpcGetMemHandle(&b.state1, ps->state[1]));
    MPL_CUDA_OK(cudaIpcGetMemHandle(&b.anc, ps->anc));
    MPL_CUDA_OK(cudaIpcGetMemHandle(&b.mail, ps->mailbox));
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
        for (int k = 0; k < 4; ++k)
            for (int h = 0; h < kMaxPeers; ++h)
                if (ps->ipc_opened[k][h]) { cudaIpcCloseMemHandle(ps->ipc_opened[k][h]); ps->ipc_opened[k][h] = nullptr; }
        cudaGetLastError();
        return fail(MPL_ERR_CUDA, std::string("peer attach: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    };
    for (int h = 0; h < world; ++h) {
        if (h == rank) {
            table.state[0][h] = ps->state[0]; table.state[1][h] = ps->state[1];
            table.anc[h] = ps->anc; table.mail[h] = ps->mailbox;
            continue;
        }
        PeerBlob b;
        std::memcpy(&b, (const char*)blobs + (size_t)h * MPL_PEER_BLOB_BYTES, sizeof b);
        const cudaIpcMemHandle_t* hs[4] = {&b.state0, &b.state1, &b.anc, &b.mail};
        for (int k = 0; k < 4; ++k) {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, *hs[k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return rollback(e);
            ps->ipc_opened[k][h] = p;
        }
        table.state[0][h] = ps->ipc_opened[0][h]; table.state[1][h] = ps->ipc_opened[1][h];
        table.anc[h] = (int32_t*)ps->ipc_opened[2][h]; table.mail[h] = (Mailbox*)ps->ipc_opened[3][h];
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
        for (int k = 0; k < 4; ++k)
            for (int h = 0; h < kMaxPeers; ++h)
                if (ps->ipc_opened[k][h]) { cudaIpcCloseMemHandle(ps->ipc_opened[k][h]); ps->ipc_opened[k][h] = nullptr; }
    std::memset(&ps->peer, 0, sizeof ps->peer);
    ps->peer.world = 1; ps->world = 1; ps->rank = 0;
    return MPL_OK;
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
    }
    if (rc == MPL_OK && world > 1) {
        for (int g = 0; g < world; ++g) {
            fill_table_header(sh[g], g, world);
            sh[g]->peer_virtual = true;
            for (int h = 0; h < world; ++h) {
                sh[g]->peer.state[0][h] = sh[h]->state[0]; sh[g]->peer.state[1][h] = sh[h]->state[1];
                sh[g]->peer.anc[h] = sh[h]->anc; sh[g]->peer.mail[h] = sh[h]->mailbox;
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
