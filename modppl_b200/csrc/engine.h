// engine.h -- host-side objects behind the opaque handles of include/modppl_b200.h
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../include/modppl_b200.h"
#include "pf_kernels.cuh"

namespace mpl { struct JitProgram; }

struct mpl_model {
    int kind;
    std::string name;
    std::vector<double> params;
    int state_dim, obs_dim, num_latents;
    std::shared_ptr<mpl::JitProgram> jit;   // M_JIT: the compiled spec (shared by the copies particle systems keep)
};

namespace mpl {

struct KernelTimer {
    double total_ms = 0.;
    uint64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

int model_from_name(const std::string& name);

}  // namespace mpl

struct mpl_ps {
    const mpl_model* model_ref;
    mpl_model model;   // copy
    int dtype, device;
    size_t n, ld;
    uint64_t seed, gid_offset, n_global;
    int D;
    cudaStream_t stream;
    void* state[2];   // D x ld Real
    int cur;          // index of the live state buffer
    void* lw;         // ld Real: the log-weights in use
    void* lw_alt;     // sharded runs only: the other buffer of the pair (every extend writes the one the previous step did not, see PeerTable::lw)
    int par;          // which of the pair `lw` is (index into PeerTable::lw / rec_e / rec_S; 0 on one GPU)
    bool anc_pushed;  // sharded: the last resample's ancestors were pushed by the peers (single-level scheme) -> the next extend waits for their flags
    int32_t* anc;     // ld
    double* probs;    // ld (exact schemes), lazily allocated
    double* cums;     // ld
    unsigned long long* icum;   // ld (integer multinomial)
    unsigned long long* desc;   // scan tile descriptors
    void* overflow;   // OverflowEntry2 records (scan2.cuh)
    size_t overflow_cap;
    mpl::DeviceStats* stats;    // device
    mpl::DeviceStats* stats_host;   // pinned
    mpl::Lse3<double>* partials;
    unsigned long long* ipartials;
    double* obs_dev;
    size_t obs_steps;
    double* staging;   // ld*max(D,1) doubles for read/write conversions
    int grid_extend, grid_reduce;
    long long t;       // next kernel time index (host mirror)
    bool initialised, pending_gather;
    bool stats_valid;   // stats->{max,sumexp,sumexp2,ess} describe the current log-weights (weight_reduce ran)
    bool max_valid;     // stats->max_bits[(t-1)&1] holds their exact max (left by the last extend)
    double* sq_partials;   // per-tile sums of squared weights (ESS in the integer resampler)
    int* host_flags;       // pinned + mapped: [0] = a heavy tile was seen (launch the overflow pass from now on)
    int* host_flags_dev;
    unsigned int host_seq;  // tag of the last nested resample whose log total weight is posted to host_flags (bytes 16..39)
    bool in_device_loop;    // inside mpl_ps_run: results are not posted to the host per step
    bool host_lse_posted;   // the last resample posts its result there (nested scheme, not ESS-triggered)
    // trajectory reconstruction (reference keeps traces[i].retv as a Vec<State>, dynunfold.rs:91-92): optional log of the
    // per-step states and ancestors, back-traced on demand
    int* rec_e; unsigned int* rec_S; float* rec_sq;   // chunk records of the nested scheme in use (ld / 128 entries), lazily allocated ...
    int* rec_e2; unsigned int* rec_S2;               // ... as parity `par` of these pairs
    unsigned long long* nest_tile_pre; unsigned long long* nest_sec;   // nested scheme: tile prefixes inside a section; section records + top-level prefixes
    unsigned int* nest_P; unsigned int* nest_F;      // nested scheme: the plan pass' output (first slot of every GLOBAL chunk, first chunk of every output tile)
    int prequantised;            // the last extend's fused epilogue left 1: integer weights + chunk records, 2: chunk records only (log-weights kept)
    void* hist_state;            // [hist_cap][D][ld] Real
    int32_t* hist_anc;           // [hist_cap][ld]
    size_t hist_cap;
    std::vector<int> hist_resampled;   // [t]: a resample followed step t
    double ess_threshold_abs;   // ESS-triggered device loop: threshold in particles
    bool dynamic_state_known;
    bool hist_broken;            // the trajectory log no longer describes the population (ESS-triggered device loop ran, or a restore)
    bool profile;
    std::map<std::string, mpl::KernelTimer> timers;
    uint64_t launch_count;
    // multi-GPU (multi_gpu.cu): peer table handed to every kernel; world == 1 on a single GPU
    int rank, world;
    mpl::PeerTable peer;
    mpl::Mailbox* mailbox;            // device memory of this rank, written by every rank
    void* ipc_opened[9][mpl::kMaxPeers];   // pointers obtained from cudaIpcOpenMemHandle (to close on detach)
    bool peer_virtual;                // peers live in this process (single-GPU emulation used by the tests)
    unsigned long long barrier_seq;   // mpl_ps_peer_barrier calls so far (every rank makes the same sequence of calls)
    // islands (local resampling; multi_gpu.cu): the other islands' state buffers, for the occasional island-level resampling
    int n_islands, island_rank;
    const void* island_state[2][mpl::kMaxPeers];
    void* island_opened[2][mpl::kMaxPeers];
};

namespace mpl {
// step phases, callable separately so that shards emulated on one GPU can be advanced phase by phase
int ps_phase_extend(mpl_ps* ps, bool init, bool fuse_nested = false);
int ps_phase_nested(mpl_ps* ps, int phase);
int ps_phase_reduce(mpl_ps* ps);
int ps_phase_scan(mpl_ps* ps);
int ensure_chunk_records(mpl_ps* ps);
int materialise(mpl_ps* ps);   // apply a pending ancestor gather
// jit.cu: launches pf_extend_kernel<JitModel<Real>, Real, mode, sharded, nested> of a model compiled from a spec
int jit_launch_extend(const mpl_model& m, int dtype, int mode, bool sharded, int nested, const void* extend_args, unsigned int grid, unsigned int block, cudaStream_t stream, bool pdl);
// categorical.rs:25-30: the sequential f64 running sum of `probs`, bit for bit (parallel emulation for long inputs); ps may be null
int launch_cumsum_exact(mpl_ps* ps, const double* probs, size_t n, double* out, cudaStream_t stream);
}
