//! Raw bindings to `include/sde_b200.h` (libsde_b200.so): every entry point, same names, same order.
//! The library talks to the GPU through the CUDA driver API and NVRTC (both dlopen'ed), so linking needs no CUDA SDK.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const SDE_OK: c_int = 0;
pub const SDE_ERR_VALUE: c_int = 1; // the pyo3 layer's ValueError (src/py_binding.rs:20-32)
pub const SDE_ERR_RUNTIME: c_int = 2; // its RuntimeError (src/py_binding.rs:47-53)

#[repr(C)]
pub struct sde_universe {
    _p: [u8; 0],
}
#[repr(C)]
pub struct sde_plan {
    _p: [u8; 0],
}
#[repr(C)]
pub struct sde_result {
    _p: [u8; 0],
}
#[repr(C)]
pub struct sde_device_plans {
    _p: [u8; 0],
}

pub const SDE_OUT_PATHS: i32 = 0;
pub const SDE_OUT_TERMINAL: i32 = 1;
pub const SDE_OUT_MOMENTS: i32 = 2;
pub const SDE_LAYOUT_NTP: i32 = 0;
pub const SDE_LAYOUT_TPN: i32 = 1;
pub const SDE_SCRAMBLE_CP_SHIFT_PER_PATH: i32 = 0;
pub const SDE_SCRAMBLE_XOR: i32 = 1;
pub const SDE_SCRAMBLE_NONE: i32 = 2;
pub const SDE_ICDF_REFERENCE: i32 = 0;
pub const SDE_ICDF_FAST: i32 = 1;
pub const SDE_ICDF_SINGLE: i32 = 2;
pub const SDE_ARITH_STRICT: i32 = 0;
pub const SDE_ARITH_FAST: i32 = 1;
pub const SDE_DTYPE_F64: i32 = 0;
pub const SDE_DTYPE_F32: i32 = 1;
pub const SDE_GEN_CHACHA8: i32 = 0;
pub const SDE_GEN_PHILOX: i32 = 1;
pub const SDE_RK_REFERENCE: i32 = 0;
pub const SDE_RK_TEXTBOOK: i32 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct sde_options {
    pub struct_size: u32,
    pub device: i32,
    pub seed: u64,
    pub scenario_offset: u64,
    pub output: i32,
    pub layout: i32,
    pub scramble: i32,
    pub icdf: i32,
    pub arith: i32,
    pub rk_variant: i32,
    pub stream: *mut c_void,
    pub inject: *const c_double,
    pub tile_steps: i32,
    pub block_threads: i32,
    pub min_blocks: i32,
    pub ntp_direct: i32,
    pub dtype: i32,
    pub wide_mma: i32,
    pub generator: i32,
}

extern "C" {
    // ---- model
    pub fn sde_universe_parse(equations: *const *const c_char, n_equations: usize, times: *const c_double, n_times: usize, out: *mut *mut sde_universe) -> c_int;
    pub fn sde_universe_free(u: *mut sde_universe);
    pub fn sde_universe_num_processes(u: *const sde_universe) -> usize;
    pub fn sde_universe_num_factors(u: *const sde_universe) -> usize;
    pub fn sde_universe_num_times(u: *const sde_universe) -> usize;
    pub fn sde_universe_process_name(u: *const sde_universe, i: usize) -> *const c_char;
    pub fn sde_universe_process_is_levy(u: *const sde_universe, i: usize) -> c_int;
    pub fn sde_universe_process_num_terms(u: *const sde_universe, i: usize) -> usize;
    pub fn sde_universe_factor_name(u: *const sde_universe, k: usize) -> *const c_char;
    // ---- options
    pub fn sde_options_default(o: *mut sde_options);
    // ---- plans
    pub fn sde_plan_create(u: *const sde_universe, scheme: *const c_char, rng_method: *const c_char, opt: *const sde_options, out: *mut *mut sde_plan) -> c_int;
    pub fn sde_plan_free(p: *mut sde_plan);
    pub fn sde_plan_source(p: *const sde_plan) -> *const c_char;
    pub fn sde_plan_is_prelowered(p: *const sde_plan) -> c_int;
    pub fn sde_plan_output_elems(p: *const sde_plan, n_scenarios: u64) -> usize;
    pub fn sde_lower_only(u: *const sde_universe, scheme: *const c_char, rng_method: *const c_char, opt: *const sde_options, compile: c_int, source_out: *mut *mut c_char, cubin_bytes: *mut usize) -> c_int;
    pub fn sde_free_string(s: *mut c_char);
    pub fn sde_plan_run_device(p: *mut sde_plan, init_names: *const *const c_char, init_vals: *const c_double, n_init: usize, n_scenarios: u64, seed: u64, scenario_offset: u64, d_out: *mut c_double, stream: *mut c_void, n_launches: *mut c_int) -> c_int;
    pub fn sde_plan_run_host(p: *mut sde_plan, init_names: *const *const c_char, init_vals: *const c_double, n_init: usize, n_scenarios: u64, seed: u64, scenario_offset: u64, h_out: *mut c_double, n_launches: *mut c_int) -> c_int;
    // ---- one-shot simulate and its result
    pub fn sde_simulate(u: *const sde_universe, init_names: *const *const c_char, init_vals: *const c_double, n_init: usize, n_scenarios: u64, scheme: *const c_char, rng_method: *const c_char, opt: *const sde_options, out: *mut *mut sde_result) -> c_int;
    pub fn sde_result_free(r: *mut sde_result);
    pub fn sde_result_shape(r: *const sde_result, n_scenarios: *mut u64, n_times: *mut usize, n_processes: *mut usize);
    pub fn sde_result_num_elems(r: *const sde_result) -> usize;
    pub fn sde_result_values_device(r: *const sde_result) -> *const c_double;
    pub fn sde_result_values_host(r: *const sde_result, dst: *mut c_double, n_elems: usize) -> c_int;
    pub fn sde_result_kernel_ms(r: *const sde_result) -> c_double;
    pub fn sde_result_process_name(r: *const sde_result, i: usize) -> *const c_char;
    pub fn sde_result_times(r: *const sde_result, dst: *mut c_double, n_times: usize) -> c_int;
    pub fn sde_result_scenario_offset(r: *const sde_result) -> u64;
    pub fn sde_result_device(r: *const sde_result) -> c_int;
    pub fn sde_result_output(r: *const sde_result) -> c_int;
    pub fn sde_result_moments(r: *const sde_result, dst: *mut c_double) -> c_int;
    // ---- several GPUs, one host thread
    pub fn sde_shard_range(n_scenarios: u64, part: usize, n_parts: usize, lo: *mut u64, hi: *mut u64);
    pub fn sde_device_plans_create(u: *const sde_universe, scheme: *const c_char, rng_method: *const c_char, opt: *const sde_options, devices: *const i32, n_devices: usize, out: *mut *mut sde_device_plans) -> c_int;
    pub fn sde_device_plans_free(ps: *mut sde_device_plans);
    pub fn sde_device_plans_count(ps: *const sde_device_plans) -> usize;
    pub fn sde_device_plans_device(ps: *const sde_device_plans, i: usize) -> c_int;
    pub fn sde_device_plans_collective(ps: *const sde_device_plans) -> c_int;
    pub fn sde_plan_run_devices(ps: *mut sde_device_plans, init_names: *const *const c_char, init_vals: *const c_double, n_init: usize, n_scenarios: u64, seed: u64, scenario_offset: u64, d_out: *const *mut c_double, n_launches: *mut c_int, collective_ms: *mut c_double) -> c_int;
    pub fn sde_simulate_devices(u: *const sde_universe, init_names: *const *const c_char, init_vals: *const c_double, n_init: usize, n_scenarios: u64, scheme: *const c_char, rng_method: *const c_char, opt: *const sde_options, devices: *const i32, n_devices: usize, out: *mut *mut sde_result) -> c_int;
    // ---- building blocks
    pub fn sde_sobol_points(device: c_int, dims: u32, first: u64, count: u64, h_out: *mut u64) -> c_int;
    pub fn sde_sobol_cp_shift_uniforms(device: c_int, dims: u32, seed: u64, first_scenario: u64, count: u64, h_out: *mut c_double) -> c_int;
    pub fn sde_joe_kuo_params(dims: u32, poly: *mut u32, minit: *mut u32) -> c_int;
    pub fn sde_chacha8_u64(device: c_int, seed: u64, n: usize, h_out: *mut u64) -> c_int;
    pub fn sde_icdf_normal(device: c_int, mode: c_int, h_p: *const c_double, n: usize, h_out: *mut c_double) -> c_int;
    pub fn sde_icdf_poisson(device: c_int, h_u: *const c_double, h_lambda: *const c_double, n: usize, h_out: *mut c_double) -> c_int;
    pub fn sde_moments_merge(shards: *const c_double, n_shards: usize, n_processes: usize, out: *mut c_double) -> c_int;
    pub fn sde_moments_merge_device(device: c_int, d_shards: *const c_double, n_shards: usize, n_processes: usize, d_out: *mut c_double, stream: *mut c_void) -> c_int;
    pub fn sde_measure_peaks(device: c_int, fill_gbs: *mut c_double, dfma_tflops: *mut c_double, ffma_tflops: *mut c_double) -> c_int;
    pub fn sde_last_error() -> *const c_char;
    pub fn sde_version() -> *const c_char;
    pub fn sde_cuda_available() -> c_int;
    pub fn sde_device_count() -> c_int;
}
