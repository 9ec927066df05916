//! modppl's inference entry points on the B200 engine (`libmodppl_b200.so`, C ABI in `include/modppl_b200.h`).
//!
//! The reference reaches models only through `trait GenFn` (modppl/src/gfi.rs:49-92) and executes them one particle at a
//! time on the host.  Here a model is a *device functor* -- one of the engine's registry or one compiled from a spec at run
//! time -- and every entry point below keeps the reference's name, argument order and return tuple, with the heap
//! `Trace`s replaced by flat `f64` arrays:
//!
//! | reference (modppl/src/inference/)                    | here                                                      |
//! |-------------------------------------------------------|-----------------------------------------------------------|
//! | `ParticleSystem::new(model, n, rng)`  particle_filter.rs:44  | `ParticleSystem::new(&model, n, seed)`               |
//! | `init_step(args, constraints)`  :60                  | `init_step(&obs)`                                         |
//! | `step(self, constraints) -> Self`  :73               | `step(self, &obs) -> Self`                                |
//! | `effective_sample_size()`  :98                       | same (the stale value the reference returns, quirk Q1)    |
//! | `resample() -> f64`  :103                            | same (`resample_with` picks another scheme)               |
//! | `log_marginal_likelihood_estimate()`  :119           | same                                                      |
//! | `importance_sampling(model, args, constraints, n)`  importance.rs:12 | `importance_sampling(&model, &obs, n, seed)` |
//! | `importance_resampling(.., n, n_ret)`  :37           | `importance_resampling(&model, &obs, n, n_ret, seed)`     |
//! | `mh(model, trace, proposal, proposal_args)`  mh.rs:9 | `Chains::mh("proposal name", arg, steps)` over all chains |
//! | `regen_mh(model, trace, mask)`  :54                  | `Chains::regen_mh(mask_bits, steps)`                      |
//!
//! Errors: the reference panics (gfi.rs:72, dyngenfn.rs:526-529); so do these wrappers, with the library's message.
//! Threading: one handle, one thread at a time (the reference's `ThreadRng` pins a `ParticleSystem` to its thread too).
//!
//! NOTE: the image this repository is built in has no Rust toolchain, so this crate has not been compiled there; the
//! declarations in `ffi.rs` are checked name by name and argument by argument against the header and the built library
//! (tests/test_abi.py), and the same calls are exercised through Python/ctypes by the GPU test-suite.
pub mod ffi;

use std::ffi::{CStr, CString};
use std::os::raw::{c_int, c_void};

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::mpl_last_error()) }.to_string_lossy().into_owned()
}

fn check(rc: c_int) {
    if rc != ffi::MPL_OK {
        panic!("modppl_b200: {}", last_error());
    }
}

/// Resampling schemes (`MPL_RESAMPLE_*`).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Scheme {
    /// the reference's own: normalised f64 weights, sequential f64 running sum, `parents[i] = min{k : S_k >= u_i}`
    Multinomial = 0,
    Systematic = 1,
    SystematicFixed = 2,
    MultinomialFixed = 3,
    /// nested systematic resampling on integer weights: the throughput scheme (fp32 particle systems)
    SystematicNested = 4,
}

/// Arithmetic type of a particle system.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum DType {
    F32 = 0,
    F64 = 1,
}

/// A model the engine has a device functor for.
pub struct DeviceModel {
    h: *mut ffi::mpl_model,
    state_dim: usize,
    obs_dim: usize,
}

impl DeviceModel {
    /// One of the registry: "spiral", "lgssm4", "sv", "hmm" (Unfold-style), "line", "hierarchical", "pointed" (static).
    pub fn builtin(name: &str, params: &[f64]) -> Self {
        let c = CString::new(name).expect("model name");
        let h = unsafe { ffi::mpl_model_create(c.as_ptr(), params.as_ptr(), params.len()) };
        assert!(!h.is_null(), "modppl_b200: {}", last_error());
        Self::wrap(h)
    }

    /// `spiral_model` of the reference's tests (tests/dyngenfns/unfold.rs:14-33).
    pub fn spiral(dr_std: f64, dtheta_mean: f64, dtheta_std: f64, obs_var: f64) -> Self {
        Self::builtin("spiral", &[dr_std, dtheta_mean, dtheta_std, obs_var])
    }

    /// A model written in the spec language of `include/modppl_b200.h` (the restricted form of a `dyngen!` Unfold
    /// kernel): compiled with NVRTC against the engine's own kernels, no rebuild of the library.
    pub fn from_spec(json: &str) -> Self {
        let c = CString::new(json).expect("spec");
        let h = unsafe { ffi::mpl_model_compile(c.as_ptr()) };
        assert!(!h.is_null(), "modppl_b200: {}", last_error());
        Self::wrap(h)
    }

    fn wrap(h: *mut ffi::mpl_model) -> Self {
        let state_dim = unsafe { ffi::mpl_model_state_dim(h) }.max(0) as usize;
        let obs_dim = unsafe { ffi::mpl_model_obs_dim(h) }.max(0) as usize;
        DeviceModel { h, state_dim, obs_dim }
    }

    pub fn state_dim(&self) -> usize {
        self.state_dim
    }
    pub fn obs_dim(&self) -> usize {
        self.obs_dim
    }
    /// number of latent slots of a static model (importance sampling / MH)
    pub fn num_latents(&self) -> usize {
        unsafe { ffi::mpl_model_num_latents(self.h) }.max(0) as usize
    }
    /// the proposals a static model registers, by the names of the reference's fixtures
    pub fn proposals(&self) -> Vec<String> {
        let n = unsafe { ffi::mpl_model_num_proposals(self.h) }.max(0);
        (0..n)
            .map(|i| unsafe { CStr::from_ptr(ffi::mpl_model_proposal_name(self.h, i)) }.to_string_lossy().into_owned())
            .collect()
    }
}

impl Drop for DeviceModel {
    fn drop(&mut self) {
        unsafe { ffi::mpl_model_destroy(self.h) }
    }
}

/// Drop-in for `modppl::ParticleSystem` (particle_filter.rs:8-24) when the model is a `DeviceModel`.
pub struct ParticleSystem {
    h: *mut ffi::mpl_ps,
    num_particles: usize,
    state_dim: usize,
}

impl ParticleSystem {
    /// `ParticleSystem::new(model, num_particles, rng)` (particle_filter.rs:44-57); `seed` stands in for the `ThreadRng`.
    /// fp64 like the reference; `with_dtype` gives the fp32 throughput path.
    pub fn new(model: &DeviceModel, num_particles: usize, seed: u64) -> Self {
        Self::with_dtype(model, num_particles, seed, DType::F64)
    }

    pub fn with_dtype(model: &DeviceModel, num_particles: usize, seed: u64, dtype: DType) -> Self {
        let cfg = ffi::mpl_pf_config { dtype: dtype as c_int, device: -1, seed, gid_offset: 0, n_global: 0 };
        let h = unsafe { ffi::mpl_particle_system_new(model.h, num_particles as u64, &cfg) };
        assert!(!h.is_null(), "modppl_b200: {}", last_error());
        ParticleSystem { h, num_particles, state_dim: model.state_dim }
    }

    /// `init_step(args, constraints)` (:60-70): N x `generate`; the constraints are the first observation.
    pub fn init_step(&mut self, constraints: &[f64]) {
        check(unsafe { ffi::mpl_ps_init_step(self.h, constraints.as_ptr(), constraints.len()) })
    }

    /// `step(self, constraints) -> Self` (:73-95): N x `update(.., Extend, ..)`; consumes and returns the system like the reference.
    pub fn step(self, constraints: &[f64]) -> Self {
        check(unsafe { ffi::mpl_ps_step(self.h, constraints.as_ptr(), constraints.len()) });
        self
    }

    /// `effective_sample_size()` (:98-100): as of the last `normalize_weights` (the reference's stale value).
    pub fn effective_sample_size(&self) -> f64 {
        let mut v = 0.0;
        check(unsafe { ffi::mpl_ps_effective_sample_size(self.h, 1, &mut v) });
        v
    }

    /// the ESS of the current weights
    pub fn effective_sample_size_fresh(&self) -> f64 {
        let mut v = 0.0;
        check(unsafe { ffi::mpl_ps_effective_sample_size(self.h, 0, &mut v) });
        v
    }

    /// `resample() -> f64` (:103-116) with the reference's multinomial scheme; returns the log total weight.
    pub fn resample(&mut self) -> f64 {
        self.resample_with(Scheme::Multinomial)
    }

    pub fn resample_with(&mut self, scheme: Scheme) -> f64 {
        let mut v = 0.0;
        check(unsafe { ffi::mpl_ps_resample(self.h, scheme as c_int, &mut v) });
        v
    }

    /// The loop body of tests/smc.rs:78-81 (`filter = filter.step(..); filter.resample();`) as one call.
    pub fn step_resample(self, constraints: &[f64], scheme: Scheme) -> (Self, f64) {
        let mut v = 0.0;
        check(unsafe { ffi::mpl_ps_step_resample(self.h, constraints.as_ptr(), constraints.len(), scheme as c_int, &mut v) });
        (self, v)
    }

    /// `log_marginal_likelihood_estimate()` (:119-121)
    pub fn log_marginal_likelihood_estimate(&self) -> f64 {
        let mut v = 0.0;
        check(unsafe { ffi::mpl_ps_log_marginal_likelihood_estimate(self.h, &mut v) });
        v
    }

    /// `traces[i].retv.last()` of every particle: `[state_dim][num_particles]`
    pub fn states(&self) -> Vec<f64> {
        let mut v = vec![0.0f64; self.state_dim * self.num_particles];
        check(unsafe { ffi::mpl_ps_read(self.h, ffi::MPL_READ_STATE, v.as_mut_ptr() as *mut c_void, v.len() * 8) });
        v
    }

    /// `log_weights` (:15)
    pub fn log_weights(&self) -> Vec<f64> {
        let mut v = vec![0.0f64; self.num_particles];
        check(unsafe { ffi::mpl_ps_read(self.h, ffi::MPL_READ_LOG_WEIGHTS, v.as_mut_ptr() as *mut c_void, v.len() * 8) });
        v
    }

    /// `parents` (:20)
    pub fn parents(&self) -> Vec<i64> {
        let mut v = vec![0i64; self.num_particles];
        check(unsafe { ffi::mpl_ps_read(self.h, ffi::MPL_READ_PARENTS, v.as_mut_ptr() as *mut c_void, v.len() * 8) });
        v
    }

    /// A whole filtering run on the device: observations `[n_steps][obs_dim]` resident in HBM, `init_step` + resample, then
    /// `step` + resample for every further row, no host round trip in between.  `ess_threshold > 0`: resample only when
    /// the ESS falls below it (decided on the GPU).  Returns the elapsed milliseconds (CUDA events).
    pub fn run(&mut self, observations: &[f64], obs_dim: usize, scheme: Scheme, ess_threshold: f64) -> f32 {
        assert!(obs_dim > 0 && observations.len() % obs_dim == 0);
        let n_steps = observations.len() / obs_dim;
        check(unsafe { ffi::mpl_ps_upload_observations(self.h, observations.as_ptr(), n_steps, obs_dim) });
        let mut ms = 0.0f32;
        check(unsafe { ffi::mpl_ps_run(self.h, 0, n_steps, scheme as c_int, ess_threshold, &mut ms) });
        ms
    }

    pub fn num_particles(&self) -> usize {
        self.num_particles
    }
}

impl Drop for ParticleSystem {
    fn drop(&mut self) {
        unsafe { ffi::mpl_ps_destroy(self.h) }
    }
}

/// `importance_sampling(model, args, constraints, num_samples)` (importance.rs:12-28): returns the latents of every
/// proposal (`[num_latents][num_samples]`, in place of `Vec<Trace>`), the log normalised weights and the log-ML estimate.
pub fn importance_sampling(model: &DeviceModel, constraints: &[f64], num_samples: u32, seed: u64) -> (Vec<f64>, Vec<f64>, f64) {
    let mut latents = vec![0.0f64; model.num_latents() * num_samples as usize];
    let mut lnw = vec![0.0f64; num_samples as usize];
    let mut lml = 0.0;
    check(unsafe {
        ffi::mpl_importance_sampling(model.h, constraints.as_ptr(), constraints.len(), num_samples, seed, 0, latents.as_mut_ptr(), lnw.as_mut_ptr(), &mut lml)
    });
    (latents, lnw, lml)
}

/// `importance_resampling(model, args, constraints, num_samples, num_ret_samples)` (importance.rs:37-51): the latents, the
/// resampled indices (`categorical` draws from the normalised weights, bit-exact running sum) and the log-ML estimate.
pub fn importance_resampling(model: &DeviceModel, constraints: &[f64], num_samples: u32, num_ret_samples: u32, seed: u64) -> (Vec<f64>, Vec<usize>, f64) {
    let mut latents = vec![0.0f64; model.num_latents() * num_samples as usize];
    let mut idx = vec![0i64; num_ret_samples as usize];
    let mut lml = 0.0;
    check(unsafe {
        ffi::mpl_importance_resampling(model.h, constraints.as_ptr(), constraints.len(), num_samples, num_ret_samples, seed, 0, latents.as_mut_ptr(), idx.as_mut_ptr(), &mut lml)
    });
    (latents, idx.into_iter().map(|i| i as usize).collect(), lml)
}

/// Many independent MCMC chains over a static model: `trace = model.generate(args, constraints).0` per chain.
pub struct Chains {
    h: *mut ffi::mpl_chains,
    n_chains: usize,
}

/// One entry of a sweep (`mpl_move`): a proposal-based move or a regeneration, repeated `repeat` times in a row.
pub enum Move<'a> {
    Mh { proposal: &'a str, arg: f64, repeat: u32 },
    Regen { mask_bits: u32, repeat: u32 },
}

impl Chains {
    pub fn new(model: &DeviceModel, constraints: &[f64], n_chains: usize, seed: u64) -> Self {
        let h = unsafe { ffi::mpl_chains_new(model.h, constraints.as_ptr(), constraints.len(), n_chains as u64, seed, 0, -1) };
        assert!(!h.is_null(), "modppl_b200: {}", last_error());
        Chains { h, n_chains }
    }

    /// `mh(model, trace, proposal, proposal_args)` (mh.rs:9-50) `n_steps` times on every chain; the proposal is one of the
    /// model's registered device functors, chosen by the name of the reference's fixture.  Returns the accepted moves.
    pub fn mh(&mut self, proposal: &str, proposal_arg: f64, n_steps: u32) -> u64 {
        let c = CString::new(proposal).expect("proposal name");
        let mut acc = 0u64;
        check(unsafe { ffi::mpl_mh(self.h, c.as_ptr(), proposal_arg, n_steps, &mut acc) });
        acc
    }

    /// `regen_mh(model, trace, mask)` (mh.rs:54-76); `mask_bits` over the model's latent slots, 0 = everything.
    pub fn regen_mh(&mut self, mask_bits: u32, n_steps: u32) -> u64 {
        let mut acc = 0u64;
        check(unsafe { ffi::mpl_regen_mh(self.h, mask_bits, n_steps, &mut acc) });
        acc
    }

    /// The body of an MCMC loop (e.g. tests/mh.rs:93-106) handed over as a whole: run `n_sweeps` times per chain in one launch.
    pub fn sweeps(&mut self, model: &DeviceModel, moves: &[Move], n_sweeps: u32) -> u64 {
        let raw: Vec<ffi::mpl_move> = moves
            .iter()
            .map(|m| match m {
                Move::Mh { proposal, arg, repeat } => {
                    let c = CString::new(*proposal).expect("proposal name");
                    let index = unsafe { ffi::mpl_model_proposal_index(model.h, c.as_ptr()) };
                    assert!(index >= 0, "modppl_b200: {}", last_error());
                    ffi::mpl_move { kind: ffi::MPL_MOVE_MH, proposal: index, arg: *arg, mask: 0, repeat: *repeat }
                }
                Move::Regen { mask_bits, repeat } => ffi::mpl_move { kind: ffi::MPL_MOVE_REGEN, proposal: 0, arg: 0.0, mask: *mask_bits, repeat: *repeat },
            })
            .collect();
        let mut acc = 0u64;
        let mut ms = 0.0f32;
        check(unsafe { ffi::mpl_mh_schedule(self.h, raw.as_ptr(), raw.len() as u32, n_sweeps, &mut acc, &mut ms) });
        acc
    }

    /// the chains' latent slots: `[slots][n_chains]`
    pub fn read(&mut self) -> Vec<f64> {
        let slots = unsafe { ffi::mpl_chains_num_slots(self.h) }.max(0) as usize;
        let mut v = vec![0.0f64; slots * self.n_chains];
        check(unsafe { ffi::mpl_chains_read(self.h, v.as_mut_ptr(), v.len() * 8) });
        v
    }
}

impl Drop for Chains {
    fn drop(&mut self) {
        unsafe { ffi::mpl_chains_destroy(self.h) }
    }
}
