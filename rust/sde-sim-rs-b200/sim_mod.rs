//! Replacement for `src/sim/mod.rs` of sde-sim-rs 0.5.1: the same public `simulate`, its body one FFI call into
//! libsde_b200.so (include/sde_b200.h) instead of the rayon loop over scenarios (reference src/sim/mod.rs:41-88).
//! Source only (no Rust toolchain in this repository's build image); see README.md in this directory for how it is
//! applied.  `euler` / `runge_kutta` stay in the crate unchanged (the golden-vector harness calls them).
pub mod euler;
pub mod runge_kutta;

use crate::proc::ProcessUniverse;
use ordered_float::OrderedFloat;
use polars::prelude::*;
use rand::Rng;
use sde_b200_sys as ffi;
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int};

fn check(rc: c_int) -> PolarsResult<()> {
    if rc == ffi::SDE_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::sde_last_error()) }.to_string_lossy().into_owned();
    Err(PolarsError::ComputeError(msg.into())) // py_binding.rs:47 maps this to RuntimeError
}

/// Same signature and result as the reference's `simulate` (src/sim/mod.rs:20-27): a LazyFrame with the columns
/// `scenario:i32, time:f64, process_name:str, value:f64` in (scenario, time, process) order (src/filtration.rs:108-113).
/// Every visible GPU takes a contiguous shard of the scenarios (sde_simulate_devices).
pub fn simulate(
    process_universe: &ProcessUniverse,
    timesteps: Vec<OrderedFloat<f64>>,
    initial_values: HashMap<String, f64>,
    num_scenarios: u64,
    scheme: &str,
    rng_method: &str,
) -> PolarsResult<LazyFrame> {
    let times: Vec<f64> = timesteps.iter().map(|t| t.0).collect();
    // the device model is lowered from exactly the strings parse_equations accepted (one new field, see the patch)
    let eqs: Vec<CString> = process_universe.equation_strings.iter().map(|s| CString::new(s.as_str()).unwrap()).collect();
    let eq_ptrs: Vec<*const c_char> = eqs.iter().map(|s| s.as_ptr()).collect();
    let mut u = std::ptr::null_mut();
    check(unsafe { ffi::sde_universe_parse(eq_ptrs.as_ptr(), eq_ptrs.len(), times.as_ptr(), times.len(), &mut u) })?;

    let (names, vals): (Vec<CString>, Vec<f64>) = initial_values.iter().map(|(k, v)| (CString::new(k.as_str()).unwrap(), *v)).unzip();
    let name_ptrs: Vec<*const c_char> = names.iter().map(|s| s.as_ptr()).collect();
    let mut opt: ffi::sde_options = unsafe { std::mem::zeroed() };
    unsafe { ffi::sde_options_default(&mut opt) };
    opt.seed = rand::rng().random(); // the reference's per-call entropy (src/sim/mod.rs:28-29)
    let (sch, rng) = (CString::new(scheme).unwrap(), CString::new(rng_method).unwrap());

    let n_dev = unsafe { ffi::sde_device_count() }.max(1) as usize;
    let mut shards = vec![std::ptr::null_mut(); n_dev];
    let rc = unsafe {
        ffi::sde_simulate_devices(u, name_ptrs.as_ptr(), vals.as_ptr(), vals.len(), num_scenarios, sch.as_ptr(), rng.as_ptr(), &opt,
                                  std::ptr::null(), n_dev, shards.as_mut_ptr())
    };
    if rc != ffi::SDE_OK {
        unsafe { ffi::sde_universe_free(u) };
        return check(rc).map(|_| unreachable!());
    }
    let (t, p) = (times.len(), process_universe.processes.len());
    let mut value = vec![0f64; num_scenarios as usize * t * p];
    let mut at = 0usize;
    for r in shards.iter() {
        let ne = unsafe { ffi::sde_result_num_elems(*r) };
        if ne > 0 {
            check(unsafe { ffi::sde_result_values_host(*r, value[at..].as_mut_ptr(), ne) })?;
        }
        at += ne;
        unsafe { ffi::sde_result_free(*r) };
    }
    unsafe { ffi::sde_universe_free(u) };

    // rows are already in (scenario, time, process) order — the order concat() produced (src/sim/mod.rs:88-91)
    let n = num_scenarios as usize;
    let scenario: Vec<i32> = (0..n).flat_map(|s| std::iter::repeat_n(s as i32, t * p)).collect(); // `s_idx as i32`, mod.rs:47
    let time: Vec<f64> = (0..n).flat_map(|_| times.iter().flat_map(|x| std::iter::repeat_n(*x, p))).collect();
    let name: Vec<&str> = (0..n * t).flat_map(|_| process_universe.processes.iter().map(|q| q.name())).collect();
    Ok(df!["scenario" => scenario, "time" => time, "process_name" => name, "value" => value]?.lazy())
}
