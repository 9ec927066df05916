//! The reference's SMC test (modppl/tests/smc.rs:63-85) on the engine: spiral model, N particles, resample every step.
//!   cargo run --release --example smc_spiral
use modppl_b200::{DeviceModel, ParticleSystem};

fn main() {
    // tests/dyngenfns/unfold.rs:14-33: dr ~ N(0, .1), dtheta ~ N(.4, .2), obs ~ mvnormal(x, .001 I)
    let model = DeviceModel::spiral(0.1, 0.4, 0.2, 0.001);
    // a noiseless spiral as the data
    let (mut r, mut theta) = (1.0f64, 0.0f64);
    let mut observations: Vec<[f64; 2]> = Vec::new();
    for _ in 0..100 {
        observations.push([r * theta.cos(), r * theta.sin()]);
        r += 0.01;
        theta += 0.4;
    }
    let mut filter = ParticleSystem::new(&model, 1000, 1);
    filter.init_step(&observations[0]);
    filter.resample();
    for y in &observations[1..] {
        filter = filter.step(y);
        let _ess = filter.effective_sample_size();
        filter.resample();
    }
    println!("log marginal likelihood estimate: {}", filter.log_marginal_likelihood_estimate());
}
