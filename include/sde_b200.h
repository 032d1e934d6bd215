/* sde_b200.h — C-ABI of the B200-native SDE path-simulation engine.
 *
 * Drop-in boundary for the hot path of Aschii85/sde-sim-rs (crate v0.5.1).  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree).
 * Plain pointers and sizes only; no C++/torch types.  The library (libsde_b200.so) talks to
 * the GPU through the CUDA *driver* API (dlopen'ed libcuda.so.1 + libnvrtc.so.12), so it
 * loads on a machine without a GPU; every compute entry point then fails with
 * SDE_ERR_RUNTIME — there is no CPU fallback.
 *
 * Error model (replaces Rust Result / panics, src/py_binding.rs:20-53): every function
 * returns 0 on success, SDE_ERR_VALUE for bad arguments / unparsable equations (the
 * pyo3 layer's ValueError) or SDE_ERR_RUNTIME for simulation / CUDA failures (its
 * RuntimeError).  sde_last_error() returns the thread-local message.  Nothing unwinds
 * across this boundary.
 */
#ifndef SDE_B200_H
#define SDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDE_OK 0
#define SDE_ERR_VALUE 1
#define SDE_ERR_RUNTIME 2

/* ---- model -------------------------------------------------------------------------- */

/* Opaque, immutable, shareable across threads: replaces proc::ProcessUniverse
 * (src/proc/mod.rs:61-90). */
typedef struct sde_universe sde_universe;

/* Replaces proc::util::parse_equations(&[String], Vec<OrderedFloat<f64>>)
 * (src/proc/util.rs:52-66).  Same grammar, same registry order (src/proc/util.rs:68-166).
 * Deviations (documented in DESIGN.md): `times` must be finite and strictly increasing;
 * an undefined variable is reported here instead of panicking at the first eval
 * (src/sim/euler.rs:22). */
int sde_universe_parse(const char* const* equations, size_t n_equations,
                       const double* times, size_t n_times, sde_universe** out);
void sde_universe_free(sde_universe* u);

size_t sde_universe_num_processes(const sde_universe* u);           /* processes.len()            */
size_t sde_universe_num_factors(const sde_universe* u);             /* stochastic_registry.len()  */
size_t sde_universe_num_times(const sde_universe* u);
const char* sde_universe_process_name(const sde_universe* u, size_t i);   /* Process::name(), mod.rs:53-58 */
int sde_universe_process_is_levy(const sde_universe* u, size_t i);
size_t sde_universe_process_num_terms(const sde_universe* u, size_t i);
const char* sde_universe_factor_name(const sde_universe* u, size_t k);

/* ---- options -------------------------------------------------------------------------- */

enum sde_output {
    SDE_OUT_PATHS = 0,     /* every row of every ScenarioFiltration (src/filtration.rs:16,112)     */
    SDE_OUT_TERMINAL = 1,  /* last row only: [N][P]                                                */
    SDE_OUT_MOMENTS = 2    /* per process (count, mean, M2) of the last row: [P][3]                */
};
enum sde_layout {
    SDE_LAYOUT_NTP = 0,    /* [scenario][time][process] — the reference's row order (filtration.rs:87-113, sim/mod.rs:88-91) */
    SDE_LAYOUT_TPN = 1     /* [time][process][scenario] — transposed, naturally coalesced          */
};
enum sde_scramble {
    SDE_SCRAMBLE_CP_SHIFT_PER_PATH = 0, /* reference behaviour: u = fract(x + ChaCha8(seed+s)) (src/rng/sobol.rs:45-47,73-76) */
    SDE_SCRAMBLE_XOR = 1,               /* one 32-bit digital-shift mask per dimension per run on the 32-bit Sobol integers
                                           (what README.md:13 describes): u = ((x ^ mask) + 1/2) * 2^-32                     */
    SDE_SCRAMBLE_NONE = 2
};
enum sde_generator {
    SDE_GEN_CHACHA8 = 0,    /* the reference's stream: ChaCha8Rng::seed_from_u64(seed + s) f64 draws (src/rng/pseudo.rs:14-31), bit-exact */
    SDE_GEN_PHILOX = 1      /* Philox4x32-10, counter = (scenario, draw block), 32-bit uniforms (w + 1/2) 2^-32: not in the reference —
                               a cheaper counter-based stream for the pseudo-random MC path, which has to agree with the reference
                               statistically only; rng_method != "sobol" only */
};
enum sde_icdf {
    SDE_ICDF_REFERENCE = 0, /* A&S 26.2.23 exactly as src/proc/increment.rs:161-179, IEEE log/sqrt/div, no contraction */
    SDE_ICDF_FAST = 1,      /* same formula; table-driven log + one-step cubic sqrt/div in f64; |dz| <= 5e-13 vs REFERENCE */
    SDE_ICDF_SINGLE = 2     /* same formula evaluated in FP32 (MUFU log2/rsqrt/rcp); |dz| <= 4e-6 vs REFERENCE — a precision
                               tier of its own (paths stay f64), never selected implicitly                            */
};
enum sde_arith {
    SDE_ARITH_STRICT = 0,   /* separate mul/add roundings in the reference's evaluation order       */
    SDE_ARITH_FAST = 1      /* FMA contraction allowed (differs by <= ~1 ulp per step)              */
};
enum sde_dtype {
    SDE_DTYPE_F64 = 0,      /* the reference's precision (default)                                 */
    SDE_DTYPE_F32 = 1       /* f32 variant: stated tolerance vs the f64 oracle 2e-4 relative on the tested models (tests/test_gpu_f32.py) */
};
enum sde_rk_variant {
    SDE_RK_REFERENCE = 0,   /* bug-compatible stale-cache semantics (src/sim/runge_kutta.rs + src/func.rs:37-39) */
    SDE_RK_TEXTBOOK = 1     /* k1 evaluated at the settled state                                    */
};

typedef struct sde_options {
    uint32_t struct_size;     /* = sizeof(sde_options); allows extension                            */
    int32_t device;           /* CUDA ordinal                                                       */
    uint64_t seed;            /* replaces rand::rng().random() (src/sim/mod.rs:28-29)               */
    uint64_t scenario_offset; /* global index of local scenario 0 (multi-GPU shards, disjoint Sobol index ranges / RNG keys) */
    int32_t output;           /* enum sde_output                                                    */
    int32_t layout;           /* enum sde_layout                                                    */
    int32_t scramble;         /* enum sde_scramble                                                  */
    int32_t icdf;             /* enum sde_icdf                                                      */
    int32_t arith;            /* enum sde_arith                                                     */
    int32_t rk_variant;       /* enum sde_rk_variant                                                */
    void* stream;             /* sde_simulate only: CUstream to launch on (NULL = a library-owned stream); the call waits for
                               * the result either way.  Plans take their stream per run (sde_plan_run_device)        */
    const double* inject;     /* DEVICE pointer [N][S][K+1] or NULL.  Test hook for the "identical normal draws" parity check:
                                 entry k<K is the normal z (Wiener factor) or uniform u (Poisson factor) of factor k, entry K is u[t][0] (RK's sk). */
    int32_t tile_steps;       /* 0 = auto; time-tile length override (tuning)                       */
    int32_t block_threads;    /* 0 = auto                                                           */
    int32_t min_blocks;       /* 0 = auto; CTAs per SM promised to the compiler (tuning)            */
    int32_t ntp_direct;       /* SDE_LAYOUT_NTP paths: 0 = auto; 1 = shared-memory transpose; 2 = direct sector stores from the
                               * time-tiled kernel; 3 = persistent-warp kernel with resident tables (Sobol xor / none only);
                               * 4 = per-lane bulk copies shared -> global (f64, even number of processes, 16-byte aligned buffer);
                               * 5 = per-warp 2-D tensor-map stores (f64, 2 or 4 processes, K <= 2, 16-byte aligned buffer)        */
    int32_t dtype;            /* enum sde_dtype: element type of the state, the model arithmetic and the stored rows.
                               * SDE_DTYPE_F32 needs arith = SDE_ARITH_FAST; paths / terminal buffers are then float,
                               * moments stay [P][3] f64 (accumulated in f64 from the f32 terminal values)                        */
    int32_t wide_mma;         /* wide linear models (Cholesky-loaded baskets; SDE_SCRAMBLE_XOR, SDE_ARITH_FAST, f64; [N][T][P] paths,
                               * terminal values or moments):
                               * 0 = auto (FP64 tensor-core kernel sde_sim_wide.cuh when the model qualifies); 1 = off (time-tiled
                               * kernel); 2 = required (plan creation fails when the model does not qualify)                      */
    int32_t generator;        /* enum sde_generator: the pseudo-random stream of rng_method != "sobol" (default: the reference's)  */
} sde_options;

void sde_options_default(sde_options* o);

/* ---- simulation ----------------------------------------------------------------------- */

/* A plan = one model lowered to device code for one (scheme, rng_method, options) choice,
 * compiled for sm_100a, with its direction-number tables resident on the device.
 * Replaces the per-call setup of sim::simulate (src/sim/mod.rs:28-39).
 * Threading: a sde_universe is immutable and may be shared freely (the reference shares &ProcessUniverse across rayon
 * threads, src/sim/mod.rs:21,45); a sde_plan owns scratch buffers (moment partials, chunk buffers, digital-shift masks), so
 * its runs must be serialised — one stream at a time per plan; create one plan per thread / stream for concurrent work.
 * sde_simulate builds its own plan per call and is re-entrant. */
typedef struct sde_plan sde_plan;

int sde_plan_create(const sde_universe* u, const char* scheme, const char* rng_method,
                    const sde_options* opt, sde_plan** out);
void sde_plan_free(sde_plan* p);
/* Generated CUDA source of the plan (for inspection / offline nvcc + cuobjdump). */
const char* sde_plan_source(const sde_plan* p);
/* 0/1: was the cubin found in the ahead-of-time cache (SDE_B200_CACHE directory, filled by `build()` for the BASELINE
 * configs and by earlier runs) instead of being NVRTC-compiled at plan creation? */
int sde_plan_is_prelowered(const sde_plan* p);
/* Number of elements (f64, or f32 for SDE_DTYPE_F32 paths / terminal values) a run over N scenarios writes. */
size_t sde_plan_output_elems(const sde_plan* p, uint64_t n_scenarios);

/* Lowering without a device (works on a CPU-only machine; used by the build check and tests):
 * returns the generated CUDA translation unit in *source_out (free with sde_free_string) and,
 * when `compile` is non-zero, NVRTC-compiles it for sm_100a and reports the cubin size. */
int sde_lower_only(const sde_universe* u, const char* scheme, const char* rng_method,
                   const sde_options* opt, int compile, char** source_out, size_t* cubin_bytes);
void sde_free_string(char* s);

/* Run N scenarios; result written to DEVICE memory `d_out` (caller-owned, e.g. a torch
 * tensor).  Asynchronous on `stream` (a CUstream; NULL = the legacy default stream, so the launch is ordered with the
 * caller's default-stream work and nothing is synchronised).  seed / scenario_offset may differ per run.
 * For SDE_OUT_MOMENTS d_out receives [P][3] (count, mean, M2).  kernel launches are
 * counted in *n_launches when non-NULL. */
int sde_plan_run_device(sde_plan* p, const char* const* init_names, const double* init_vals,
                        size_t n_init, uint64_t n_scenarios, uint64_t seed,
                        uint64_t scenario_offset, double* d_out, void* stream, int* n_launches);

/* Same, HOST buffers: device memory is managed by the library, the result is copied to
 * `h_out` (chunked, copy overlapped with compute).  Synchronous. */
int sde_plan_run_host(sde_plan* p, const char* const* init_names, const double* init_vals,
                      size_t n_init, uint64_t n_scenarios, uint64_t seed,
                      uint64_t scenario_offset, double* h_out, int* n_launches);

/* Opaque result of the one-shot call below. */
typedef struct sde_result sde_result;

/* Replaces sim::simulate(&ProcessUniverse, Vec<OrderedFloat<f64>>, HashMap<String,f64>,
 * u64, &str, &str) -> PolarsResult<LazyFrame> (src/sim/mod.rs:20-27).  Unknown scheme ->
 * SDE_ERR_VALUE (reference: unimplemented!() panic, src/sim/mod.rs:82); any rng_method
 * other than "sobol" -> pseudo (src/sim/mod.rs:65).  The result stays on the device until
 * asked for. */
int sde_simulate(const sde_universe* u, const char* const* init_names, const double* init_vals,
                 size_t n_init, uint64_t n_scenarios, const char* scheme, const char* rng_method,
                 const sde_options* opt, sde_result** out);
void sde_result_free(sde_result* r);
void sde_result_shape(const sde_result* r, uint64_t* n_scenarios, size_t* n_times, size_t* n_processes);
size_t sde_result_num_elems(const sde_result* r);
const double* sde_result_values_device(const sde_result* r);
int sde_result_values_host(const sde_result* r, double* dst, size_t n_elems);   /* the `value` column, filtration.rs:112 */
double sde_result_kernel_ms(const sde_result* r);
/* What the reference's frame carries besides `value` (src/filtration.rs:108-113): the process names in equation order
 * (`process_name`, src/proc/mod.rs:75-76), the time grid (`time`) and the global index of the first scenario (`scenario`). */
const char* sde_result_process_name(const sde_result* r, size_t i);
int sde_result_times(const sde_result* r, double* dst, size_t n_times);
uint64_t sde_result_scenario_offset(const sde_result* r);
int sde_result_device(const sde_result* r);
int sde_result_output(const sde_result* r);                  /* enum sde_output */
/* SDE_OUT_MOMENTS results: the [P][3] (count, mean, M2) triples, host copy. */
int sde_result_moments(const sde_result* r, double* dst);

/* ---- several GPUs of one box, one host thread -------------------------------------------
 * Replaces the rayon `into_par_iter` over scenarios inside sim::simulate (src/sim/mod.rs:41-43,88): device i of G
 * simulates the scenarios sde_shard_range(N, i, G) — disjoint Sobol index ranges / ChaCha keys, so the union of the
 * shards is bit-identical to a one-device run — and every launch is issued before any device is synchronised.
 * SDE_OUT_MOMENTS: every device reduces its shard; the 3 P doubles per device are all-gathered over NCCL
 * (ncclAllGather inside one group on the compute streams; libnccl.so.2 is dlopen'ed, peer-to-peer copies stand in when
 * it is absent) and Chan-merged by a device kernel right behind, so every device holds the merged [P][3].  No host hop. */
typedef struct sde_device_plans sde_device_plans;
void sde_shard_range(uint64_t n_scenarios, size_t part, size_t n_parts, uint64_t* lo, uint64_t* hi);
/* One plan per device of `devices` (NULL: every visible device) + the communicator the moment merge needs. */
int sde_device_plans_create(const sde_universe* u, const char* scheme, const char* rng_method, const sde_options* opt,
                            const int32_t* devices, size_t n_devices, sde_device_plans** out);
void sde_device_plans_free(sde_device_plans* ps);
size_t sde_device_plans_count(const sde_device_plans* ps);
int sde_device_plans_device(const sde_device_plans* ps, size_t i);
int sde_device_plans_collective(const sde_device_plans* ps);  /* 0 none (one device / no moments), 1 NCCL, 2 peer-to-peer copies */
/* d_out[i]: DEVICE memory on device i receiving shard i ([n_i][T][P] paths / [n_i][P] terminal values) or, for
 * SDE_OUT_MOMENTS, the merged [P][3] (NULL entries are skipped for moments).  Synchronous.  collective_ms (optional):
 * device time from the first gather call to the end of the merge kernel on device 0. */
int sde_plan_run_devices(sde_device_plans* ps, const char* const* init_names, const double* init_vals, size_t n_init,
                         uint64_t n_scenarios, uint64_t seed, uint64_t scenario_offset, double* const* d_out,
                         int* n_launches, double* collective_ms);
/* sim::simulate over several devices in one call: out[i] receives shard i (sde_result_scenario_offset / _shape tell
 * which scenarios; for SDE_OUT_MOMENTS every out[i] holds the merged triples).  `out` has room for n_devices results
 * (the number of visible devices when `devices` is NULL). */
int sde_simulate_devices(const sde_universe* u, const char* const* init_names, const double* init_vals, size_t n_init,
                         uint64_t n_scenarios, const char* scheme, const char* rng_method, const sde_options* opt,
                         const int32_t* devices, size_t n_devices, sde_result** out);

/* ---- building blocks exposed for parity tests and measurement ------------------------- */

/* Integer Sobol points n = first .. first+count-1, all `dims` dimensions, as u64
 * (device kernel; host output [count][dims]).  Replaces sobol::Sobol::<f64>::new(dims,
 * JoeKuoD6::extended()) as used at src/rng/sobol.rs:15-25. */
int sde_sobol_points(int device, uint32_t dims, uint64_t first, uint64_t count, uint64_t* h_out);
/* The uniforms the reference's Sobol mode feeds a path (device kernel; host output [count][dims]):
 * u[i][d] = fract(raw_d(point first_scenario + i + 5) + shift_d), shift = f64 draws of ChaCha8Rng::seed_from_u64(seed +
 * scenario) — SobolRng::new + RandomShiftScrambler (src/rng/sobol.rs:35-53,62-79), scenario seeding src/sim/mod.rs:56. */
int sde_sobol_cp_shift_uniforms(int device, uint32_t dims, uint64_t seed, uint64_t first_scenario, uint64_t count, double* h_out);
/* Joe–Kuo parameters shipped with the library (for checking against an independent copy). */
int sde_joe_kuo_params(uint32_t dims, uint32_t* poly /*[dims]*/, uint32_t* minit /*[dims][18]*/);
/* u64 / f64 stream of ChaCha8Rng::seed_from_u64(seed) (src/rng/pseudo.rs:18,25), generated on the device. */
int sde_chacha8_u64(int device, uint64_t seed, size_t n, uint64_t* h_out);
/* fast_inverse_normal_cdf (src/proc/increment.rs:161-179) on the device; mode = enum sde_icdf. */
int sde_icdf_normal(int device, int mode, const double* h_p, size_t n, double* h_out);
/* fast_inverse_poisson_cdf (src/proc/increment.rs:182-200) on the device. */
int sde_icdf_poisson(int device, const double* h_u, const double* h_lambda, size_t n, double* h_out);
/* Merge per-shard (count, mean, M2) triples [n_shards][P][3] -> [P][3] (host, Chan et al.). */
int sde_moments_merge(const double* shards, size_t n_shards, size_t n_processes, double* out);
/* The same merge on the device (shards and result in DEVICE memory), asynchronous on `stream` (CUstream, NULL = legacy
 * default stream): what follows the all-gather of the triples when one process drives one GPU (torch.distributed / MPI). */
int sde_moments_merge_device(int device, const double* d_shards, size_t n_shards, size_t n_processes, double* d_out, void* stream);
/* Device microbenchmarks recorded beside MEASURED_PEAKS.json: pure-write GB/s, DFMA / FFMA TFLOP/s. */
int sde_measure_peaks(int device, double* fill_gbs, double* dfma_tflops, double* ffma_tflops);

const char* sde_last_error(void);
const char* sde_version(void);
/* 1 when libcuda + a device are usable from this process. */
int sde_cuda_available(void);
/* Number of visible CUDA devices (0 without a driver / GPU). */
int sde_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SDE_B200_H */
