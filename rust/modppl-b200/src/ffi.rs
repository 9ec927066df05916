//! `extern "C"` declarations of include/modppl_b200.h, one for one (tests/test_abi.py checks every name and argument
//! count against the header and the built library).  Status codes: 0 ok, < 0 error; `mpl_last_error()` has the message.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_longlong, c_void};

pub const MPL_OK: c_int = 0;
pub const MPL_ERR_INVALID: c_int = -1;
pub const MPL_ERR_CUDA: c_int = -2;
pub const MPL_ERR_DEGENERATE: c_int = -3;
pub const MPL_ERR_UNSUPPORTED: c_int = -4;
pub const MPL_F32: c_int = 0;
pub const MPL_F64: c_int = 1;
pub const MPL_RESAMPLE_MULTINOMIAL: c_int = 0; // the reference's: particle_filter.rs:37-41 -> categorical.rs:22-32
pub const MPL_RESAMPLE_SYSTEMATIC: c_int = 1;
pub const MPL_RESAMPLE_SYSTEMATIC_FIXED: c_int = 2;
pub const MPL_RESAMPLE_MULTINOMIAL_FIXED: c_int = 3;
pub const MPL_RESAMPLE_SYSTEMATIC_NESTED: c_int = 4; // the throughput scheme (fp32)
pub const MPL_READ_STATE: c_int = 0;
pub const MPL_READ_LOG_WEIGHTS: c_int = 1;
pub const MPL_READ_PARENTS: c_int = 2;
pub const MPL_MOVE_MH: i32 = 0;
pub const MPL_MOVE_REGEN: i32 = 1;
pub const MPL_PEER_BLOB_BYTES: usize = 1024;

#[repr(C)]
pub struct mpl_model { _opaque: [u8; 0] }
#[repr(C)]
pub struct mpl_ps { _opaque: [u8; 0] }
#[repr(C)]
pub struct mpl_chains { _opaque: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct mpl_pf_config {
    pub dtype: c_int,
    pub device: c_int,
    pub seed: u64,
    pub gid_offset: u64,
    pub n_global: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct mpl_move {
    pub kind: i32,
    pub proposal: i32,
    pub arg: f64,
    pub mask: u32,
    pub repeat: u32,
}

extern "C" {
    pub fn mpl_last_error() -> *const c_char;
    pub fn mpl_version() -> *const c_char;
    pub fn mpl_device_count(count: *mut c_int) -> c_int;

    // models: the registry of device functors, and specs compiled at run time (stands in for dyngen!)
    pub fn mpl_model_create(name: *const c_char, params: *const c_double, n_params: usize) -> *mut mpl_model;
    pub fn mpl_model_compile(spec_json: *const c_char) -> *mut mpl_model;
    pub fn mpl_model_jit_compile(m: *mut mpl_model, dtype: c_int, log: *mut c_char, log_bytes: usize) -> c_int;
    pub fn mpl_model_jit_source(m: *const mpl_model, dtype: c_int) -> *const c_char;
    pub fn mpl_model_destroy(m: *mut mpl_model);
    pub fn mpl_model_state_dim(m: *const mpl_model) -> c_int;
    pub fn mpl_model_obs_dim(m: *const mpl_model) -> c_int;

    // ParticleSystem (modppl/src/inference/particle_filter.rs)
    pub fn mpl_particle_system_new(m: *const mpl_model, num_particles: u64, cfg: *const mpl_pf_config) -> *mut mpl_ps; // ::new :44-57
    pub fn mpl_ps_destroy(ps: *mut mpl_ps);
    pub fn mpl_ps_init_step(ps: *mut mpl_ps, obs: *const c_double, n_obs: usize) -> c_int; // :60-70
    pub fn mpl_ps_step(ps: *mut mpl_ps, obs: *const c_double, n_obs: usize) -> c_int; // :73-95
    pub fn mpl_ps_effective_sample_size(ps: *mut mpl_ps, stale_like_reference: c_int, out: *mut c_double) -> c_int; // :98-100
    pub fn mpl_ps_resample(ps: *mut mpl_ps, scheme: c_int, log_total_weight: *mut c_double) -> c_int; // :103-116
    pub fn mpl_ps_step_resample(ps: *mut mpl_ps, obs: *const c_double, n_obs: usize, scheme: c_int, log_total_weight: *mut c_double) -> c_int;
    pub fn mpl_ps_log_marginal_likelihood_estimate(ps: *mut mpl_ps, out: *mut c_double) -> c_int; // :119-121
    pub fn mpl_ps_checkpoint_size(ps: *mut mpl_ps, bytes: *mut u64) -> c_int;
    pub fn mpl_ps_checkpoint(ps: *mut mpl_ps, dst: *mut c_void, bytes: u64) -> c_int;
    pub fn mpl_ps_restore(ps: *mut mpl_ps, src: *const c_void, bytes: u64) -> c_int;
    pub fn mpl_ps_read(ps: *mut mpl_ps, what: c_int, host_dst: *mut c_void, bytes: usize) -> c_int; // `pub traces` :13
    pub fn mpl_ps_write(ps: *mut mpl_ps, what: c_int, host_src: *const c_void, bytes: usize) -> c_int;
    pub fn mpl_ps_num_particles(ps: *const mpl_ps, out: *mut u64) -> c_int;
    pub fn mpl_ps_history_enable(ps: *mut mpl_ps, max_steps: u64) -> c_int;
    pub fn mpl_ps_trajectories(ps: *mut mpl_ps, ids: *const i64, n_ids: u64, out: *mut c_double, bytes: usize, n_steps: *mut u64) -> c_int;
    pub fn mpl_ps_sync(ps: *mut mpl_ps) -> c_int;
    pub fn mpl_ps_upload_observations(ps: *mut mpl_ps, obs: *const c_double, n_steps: usize, n_obs: usize) -> c_int;
    pub fn mpl_ps_run(ps: *mut mpl_ps, first_step: usize, n_steps: usize, scheme: c_int, ess_threshold: c_double, elapsed_ms: *mut c_float) -> c_int;
    pub fn mpl_ps_num_resamples(ps: *mut mpl_ps, out: *mut u64) -> c_int;
    pub fn mpl_ps_profile_enable(ps: *mut mpl_ps, on: c_int) -> c_int;
    pub fn mpl_ps_profile_get(ps: *mut mpl_ps, kernel: *const c_char, total_ms: *mut c_double, launches: *mut u64) -> c_int;
    pub fn mpl_ps_launch_count(ps: *mut mpl_ps, out: *mut u64) -> c_int;

    // importance sampling (modppl/src/inference/importance.rs)
    pub fn mpl_importance_sampling(m: *const mpl_model, obs: *const c_double, n_obs: usize, num_samples: u32, seed: u64, batch: u64,
                                   latents: *mut c_double, log_norm_weights: *mut c_double, lml: *mut c_double) -> c_int; // :12-28
    pub fn mpl_importance_resampling(m: *const mpl_model, obs: *const c_double, n_obs: usize, num_samples: u32, num_ret_samples: u32, seed: u64,
                                     batch: u64, latents: *mut c_double, resampled_indices: *mut i64, lml: *mut c_double) -> c_int; // :37-51
    pub fn mpl_model_num_latents(m: *const mpl_model) -> c_int;

    // Metropolis-Hastings over many chains (modppl/src/inference/mh.rs)
    pub fn mpl_model_num_proposals(m: *const mpl_model) -> c_int;
    pub fn mpl_model_proposal_name(m: *const mpl_model, index: c_int) -> *const c_char;
    pub fn mpl_model_proposal_index(m: *const mpl_model, name: *const c_char) -> c_int;
    pub fn mpl_chains_new(m: *const mpl_model, obs: *const c_double, n_obs: usize, n_chains: u64, seed: u64, chain_offset: u64, device: c_int) -> *mut mpl_chains;
    pub fn mpl_chains_destroy(c: *mut mpl_chains);
    pub fn mpl_mh(c: *mut mpl_chains, proposal: *const c_char, proposal_arg: c_double, n_steps: u32, n_accepted: *mut u64) -> c_int; // :9-50
    pub fn mpl_regen_mh(c: *mut mpl_chains, mask_bits: u32, n_steps: u32, n_accepted: *mut u64) -> c_int; // :54-76
    pub fn mpl_mh_schedule(c: *mut mpl_chains, moves: *const mpl_move, n_moves: u32, n_sweeps: u32, n_accepted: *mut u64, elapsed_ms: *mut c_float) -> c_int;
    pub fn mpl_chains_num_slots(c: *const mpl_chains) -> c_int;
    pub fn mpl_chains_read(c: *mut mpl_chains, host_dst: *mut c_double, bytes: usize) -> c_int;
    pub fn mpl_chains_write(c: *mut mpl_chains, host_src: *const c_double, bytes: usize) -> c_int;

    // parity hooks (injected inputs, no RNG)
    pub fn mpl_resample_indices(probs: *const c_double, uniforms: *const c_double, n: u64, n_draws: u64, scheme: c_int, parents: *mut i64) -> c_int;
    pub fn mpl_cumsum_sequential(probs: *const c_double, n: u64, out: *mut c_double) -> c_int;
    pub fn mpl_logsumexp_stats(lw: *const c_void, n: u64, dtype: c_int, lse: *mut c_double, ess: *mut c_double, max: *mut c_double) -> c_int;
    pub fn mpl_logpdf(dist: *const c_char, x: *const c_double, params: *const c_double, n_params: usize, out: *mut c_double) -> c_int;

    // several GPUs, one process each: peers' memory through CUDA IPC handles the caller exchanges (any transport)
    pub fn mpl_ps_peer_export(ps: *mut mpl_ps, blob: *mut c_void) -> c_int;
    pub fn mpl_ps_peer_attach(ps: *mut mpl_ps, rank: c_int, world: c_int, blobs: *const c_void) -> c_int;
    pub fn mpl_ps_peer_detach(ps: *mut mpl_ps) -> c_int;
    pub fn mpl_ps_peer_barrier(ps: *mut mpl_ps) -> c_int;
    pub fn mpl_ps_peer_error(ps: *mut mpl_ps, out: *mut c_int) -> c_int;
    pub fn mpl_ps_nvlink_bytes(ps: *mut mpl_ps, out: *mut u64) -> c_int;
    pub fn mpl_ps_trace(ps: *mut mpl_ps, out16: *mut c_longlong) -> c_int;
    pub fn mpl_ps_island_export(ps: *mut mpl_ps, blob: *mut c_void) -> c_int;
    pub fn mpl_ps_island_attach(ps: *mut mpl_ps, rank: c_int, n_islands: c_int, blobs: *const c_void) -> c_int;
    pub fn mpl_ps_live_buffer(ps: *mut mpl_ps, out: *mut c_int) -> c_int;
    pub fn mpl_ps_island_copy_from(ps: *mut mpl_ps, src_island: c_int, src_live_buffer: c_int) -> c_int;
    pub fn mpl_ps_copy_state(dst: *mut mpl_ps, src: *mut mpl_ps) -> c_int;
}
