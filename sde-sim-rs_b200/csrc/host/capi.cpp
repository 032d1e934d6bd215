// capi.cpp — the extern "C" boundary declared in include/sde_b200.h.
// Exceptions never cross it: every entry point maps them to SDE_ERR_VALUE / SDE_ERR_RUNTIME
// plus a thread-local message (the pyo3 layer's ValueError / RuntimeError, src/py_binding.rs:20-53).
// the library is built with -fvisibility=hidden; only the C-ABI is exported
#pragma GCC visibility push(default)
#include "../../../include/sde_b200.h"
#pragma GCC visibility pop

#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>

#include "engine.h"

using namespace sde;

struct sde_universe { Universe u; };
struct sde_plan { std::unique_ptr<Plan> plan; const double* inject = nullptr; };
struct sde_result {
    std::shared_ptr<Plan> plan;
    DeviceBuffer values;
    uint64_t n = 0;
    uint64_t scenario_offset = 0;
    size_t elems = 0;
    double kernel_ms = 0.0;
};
struct sde_device_plans { std::unique_ptr<DevicePlans> ps; };

namespace {
thread_local std::string g_error;

template <class F>
int guarded(F&& f) {
    try {
        f();
        return SDE_OK;
    } catch (const ExprError& e) { g_error = e.msg; return SDE_ERR_VALUE; }
    catch (const CudaError& e) { g_error = e.msg; return SDE_ERR_RUNTIME; }
    catch (const std::bad_alloc&) { g_error = "out of host memory"; return SDE_ERR_RUNTIME; }
    catch (const std::exception& e) { g_error = e.what(); return SDE_ERR_RUNTIME; }
    catch (...) { g_error = "unknown error"; return SDE_ERR_RUNTIME; }
}

PlanOptions plan_options(const Universe& u, const char* scheme, const char* rng_method, const sde_options* o_in) {
    sde_options o;
    sde_options_default(&o);
    if (o_in) std::memcpy(&o, o_in, std::min<size_t>(sizeof o, o_in->struct_size ? o_in->struct_size : sizeof o));
    if (!scheme || !rng_method) throw ExprError{"scheme / rng_method must not be NULL"};
    PlanOptions po;
    po.device = o.device;
    std::string sch(scheme), rng(rng_method);
    if (sch == "euler") po.lower.scheme = SCHEME_EULER;
    else if (sch == "runge-kutta") po.lower.scheme = SCHEME_RK;
    else throw ExprError{"unknown scheme '" + sch + "' (expected \"euler\" or \"runge-kutta\"; the reference panics with unimplemented!(), src/sim/mod.rs:82)"};
    if (o.inject) po.lower.rng = RNG_INJECT;
    else if (rng == "sobol") {
        switch (o.scramble) {
            case SDE_SCRAMBLE_CP_SHIFT_PER_PATH: po.lower.rng = RNG_SOBOL_CP; break;
            case SDE_SCRAMBLE_XOR: po.lower.rng = RNG_SOBOL_XOR; break;
            case SDE_SCRAMBLE_NONE: po.lower.rng = RNG_SOBOL_RAW; break;
            default: throw ExprError{"unknown scramble mode"};
        }
    } else if (o.generator == SDE_GEN_PHILOX) po.lower.rng = RNG_PHILOX;
    else if (o.generator == SDE_GEN_CHACHA8) po.lower.rng = RNG_PSEUDO;   // anything else -> pseudo (src/sim/mod.rs:65)
    else throw ExprError{"unknown generator"};
    switch (o.output) {
        case SDE_OUT_PATHS: po.lower.out = o.layout == SDE_LAYOUT_TPN ? OUT_PATHS_TPN : OUT_PATHS_NTP; break;
        case SDE_OUT_TERMINAL: po.lower.out = OUT_TERMINAL; break;
        case SDE_OUT_MOMENTS: po.lower.out = OUT_MOMENTS; break;
        default: throw ExprError{"unknown output mode"};
    }
    if (o.layout != SDE_LAYOUT_NTP && o.layout != SDE_LAYOUT_TPN) throw ExprError{"unknown layout"};
    if (o.icdf != SDE_ICDF_REFERENCE && o.icdf != SDE_ICDF_FAST && o.icdf != SDE_ICDF_SINGLE) throw ExprError{"unknown icdf mode"};
    po.lower.icdf = o.icdf;
    po.lower.strict = o.arith != SDE_ARITH_FAST;
    po.lower.rk_textbook = o.rk_variant == SDE_RK_TEXTBOOK;
    if (o.dtype != SDE_DTYPE_F64 && o.dtype != SDE_DTYPE_F32) throw ExprError{"unknown dtype"};
    po.lower.f32 = o.dtype == SDE_DTYPE_F32;
    po.lower.block = o.block_threads;
    po.lower.tile_steps = o.tile_steps;
    po.lower.min_blocks = o.min_blocks;
    po.lower.direct = o.ntp_direct == 1 ? 0 : (o.ntp_direct == 2 ? 1 : (o.ntp_direct == 3 ? 2 : (o.ntp_direct == 4 ? 3 : (o.ntp_direct == 5 ? 4 : -1))));
    if (o.wide_mma < 0 || o.wide_mma > 2) throw ExprError{"unknown wide_mma mode"};
    po.lower.wide_mma = o.wide_mma == 1 ? 0 : (o.wide_mma == 2 ? 1 : -1);
    (void)u;
    return po;
}

std::vector<std::pair<std::string, double>> init_pairs(const char* const* names, const double* vals, size_t n) {
    std::vector<std::pair<std::string, double>> v;
    if (n && (!names || !vals)) throw ExprError{"initial values: NULL array"};
    for (size_t i = 0; i < n; ++i) v.emplace_back(names[i] ? names[i] : "", vals[i]);
    return v;
}
}  // namespace

extern "C" {

const char* sde_last_error(void) { return g_error.c_str(); }
const char* sde_version(void) { return "sde_b200 0.1.0 (sm_100a; reference sde-sim-rs 0.5.1)"; }
int sde_device_count(void) { try { std::string why; return driver_available(&why) ? device_count() : 0; } catch (...) { return 0; } }
int sde_cuda_available(void) { std::string why; bool ok = driver_available(&why); if (!ok) g_error = why; return ok ? 1 : 0; }

void sde_options_default(sde_options* o) {
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->struct_size = (uint32_t)sizeof *o;
    o->output = SDE_OUT_PATHS;
    o->layout = SDE_LAYOUT_NTP;
    o->scramble = SDE_SCRAMBLE_CP_SHIFT_PER_PATH;
    o->icdf = SDE_ICDF_REFERENCE;
    o->arith = SDE_ARITH_STRICT;
    o->rk_variant = SDE_RK_REFERENCE;
    o->dtype = SDE_DTYPE_F64;
}

int sde_universe_parse(const char* const* equations, size_t n_equations, const double* times, size_t n_times, sde_universe** out) {
    return guarded([&] {
        if (!out) throw ExprError{"out must not be NULL"};
        *out = nullptr;
        if ((n_equations && !equations) || (n_times && !times)) throw ExprError{"NULL input array"};
        std::vector<std::string> eqs;
        for (size_t i = 0; i < n_equations; ++i) eqs.emplace_back(equations[i] ? equations[i] : "");
        std::vector<double> tt(times, times + n_times);
        auto h = std::make_unique<sde_universe>();
        h->u = parse_equations(eqs, tt);
        *out = h.release();
    });
}
void sde_universe_free(sde_universe* u) { delete u; }
size_t sde_universe_num_processes(const sde_universe* u) { return u ? u->u.processes.size() : 0; }
size_t sde_universe_num_factors(const sde_universe* u) { return u ? u->u.factor_names.size() : 0; }
size_t sde_universe_num_times(const sde_universe* u) { return u ? u->u.times.size() : 0; }
const char* sde_universe_process_name(const sde_universe* u, size_t i) { return (u && i < u->u.processes.size()) ? u->u.processes[i].name.c_str() : nullptr; }
int sde_universe_process_is_levy(const sde_universe* u, size_t i) { return (u && i < u->u.processes.size() && u->u.processes[i].levy) ? 1 : 0; }
size_t sde_universe_process_num_terms(const sde_universe* u, size_t i) { return (u && i < u->u.processes.size()) ? u->u.processes[i].terms.size() : 0; }
const char* sde_universe_factor_name(const sde_universe* u, size_t k) { return (u && k < u->u.factor_names.size()) ? u->u.factor_names[k].c_str() : nullptr; }

int sde_plan_create(const sde_universe* u, const char* scheme, const char* rng_method, const sde_options* opt, sde_plan** out) {
    return guarded([&] {
        if (!u || !out) throw ExprError{"NULL argument"};
        *out = nullptr;
        PlanOptions po = plan_options(u->u, scheme, rng_method, opt);
        auto h = std::make_unique<sde_plan>();
        h->plan = std::make_unique<Plan>(u->u, po);
        h->inject = opt ? opt->inject : nullptr;
        *out = h.release();
    });
}
void sde_plan_free(sde_plan* p) { delete p; }

int sde_lower_only(const sde_universe* u, const char* scheme, const char* rng_method, const sde_options* opt, int compile,
                   char** source_out, size_t* cubin_bytes) {
    return guarded([&] {
        if (!u) throw ExprError{"NULL argument"};
        if (source_out) *source_out = nullptr;
        PlanOptions po = plan_options(u->u, scheme, rng_method, opt);
        Lowered low = lower_model(u->u, po.lower);
        if (compile) {
            std::string log;
            std::vector<char> cubin = nvrtc_compile(low.source, "sde_plan.cu", &log);
            if (cubin_bytes) *cubin_bytes = cubin.size();
        }
        if (source_out) {
            char* s = static_cast<char*>(std::malloc(low.source.size() + 1));
            if (!s) throw std::bad_alloc();
            std::memcpy(s, low.source.c_str(), low.source.size() + 1);
            *source_out = s;
        }
    });
}
void sde_free_string(char* s) { std::free(s); }
const char* sde_plan_source(const sde_plan* p) { return p ? p->plan->lowered().source.c_str() : nullptr; }
int sde_plan_is_prelowered(const sde_plan* p) { return (p && p->plan->prelowered()) ? 1 : 0; }
size_t sde_plan_output_elems(const sde_plan* p, uint64_t n) { return p ? p->plan->output_elems(n) : 0; }

int sde_plan_run_device(sde_plan* p, const char* const* init_names, const double* init_vals, size_t n_init, uint64_t n_scenarios,
                        uint64_t seed, uint64_t scenario_offset, double* d_out, void* stream, int* n_launches) {
    return guarded([&] {
        if (!p || !d_out) throw ExprError{"NULL argument"};
        if (n_scenarios == 0) throw ExprError{"scenarios must be a positive integer"};
        p->plan->run_device(init_pairs(init_names, init_vals, n_init), n_scenarios, seed, scenario_offset, d_out, p->inject, (CUstream)stream, n_launches);
    });
}

int sde_plan_run_host(sde_plan* p, const char* const* init_names, const double* init_vals, size_t n_init, uint64_t n_scenarios,
                      uint64_t seed, uint64_t scenario_offset, double* h_out, int* n_launches) {
    return guarded([&] {
        if (!p || !h_out) throw ExprError{"NULL argument"};
        if (n_scenarios == 0) throw ExprError{"scenarios must be a positive integer"};
        p->plan->run_host(init_pairs(init_names, init_vals, n_init), n_scenarios, seed, scenario_offset, h_out, n_launches);
    });
}

int sde_simulate(const sde_universe* u, const char* const* init_names, const double* init_vals, size_t n_init, uint64_t n_scenarios,
                 const char* scheme, const char* rng_method, const sde_options* opt, sde_result** out) {
    return guarded([&] {
        if (!u || !out) throw ExprError{"NULL argument"};
        *out = nullptr;
        if (n_scenarios == 0) throw ExprError{"scenarios must be a positive integer"};   // py_binding.rs:20-24
        PlanOptions po = plan_options(u->u, scheme, rng_method, opt);
        sde_options o;
        sde_options_default(&o);
        if (opt) std::memcpy(&o, opt, std::min<size_t>(sizeof o, opt->struct_size ? opt->struct_size : sizeof o));
        auto r = std::make_unique<sde_result>();
        r->plan = std::make_shared<Plan>(u->u, po);
        r->n = n_scenarios;
        r->scenario_offset = o.scenario_offset;
        r->elems = r->plan->output_elems(n_scenarios);
        use_device(po.device);
        r->values.alloc(r->plan->output_bytes(n_scenarios));
        r->plan->run_timed(init_pairs(init_names, init_vals, n_init), n_scenarios, o.seed, o.scenario_offset,
                           r->values.as<double>(), o.inject, (CUstream)o.stream, nullptr);
        r->kernel_ms = r->plan->last_kernel_ms();
        *out = r.release();
    });
}
void sde_result_free(sde_result* r) {
    if (!r) return;
    try { use_device(r->plan->device()); } catch (...) {}
    delete r;
}
void sde_result_shape(const sde_result* r, uint64_t* n, size_t* T, size_t* P) {
    if (!r) return;
    if (n) *n = r->n;
    if (T) *T = r->plan->universe().times.size();
    if (P) *P = r->plan->universe().processes.size();
}
size_t sde_result_num_elems(const sde_result* r) { return r ? r->elems : 0; }
const double* sde_result_values_device(const sde_result* r) { return r ? r->values.as<double>() : nullptr; }
int sde_result_values_host(const sde_result* r, double* dst, size_t n_elems) {
    return guarded([&] {
        if (!r || !dst) throw ExprError{"NULL argument"};
        if (n_elems < r->elems) throw ExprError{"destination too small"};
        use_device(r->plan->device());
        cu_check(driver().cuMemcpyDtoH(dst, r->values.ptr(), r->plan->output_bytes(r->n)), "cuMemcpyDtoH");
    });
}
double sde_result_kernel_ms(const sde_result* r) { return r ? r->kernel_ms : 0.0; }
int sde_result_device(const sde_result* r) { return r ? r->plan->device() : -1; }
uint64_t sde_result_scenario_offset(const sde_result* r) { return r ? r->scenario_offset : 0; }
int sde_result_output(const sde_result* r) {
    if (!r) return -1;
    const int out = r->plan->options().lower.out;
    return out == OUT_TERMINAL ? SDE_OUT_TERMINAL : (out == OUT_MOMENTS ? SDE_OUT_MOMENTS : SDE_OUT_PATHS);
}
const char* sde_result_process_name(const sde_result* r, size_t i) {
    return (r && i < r->plan->universe().processes.size()) ? r->plan->universe().processes[i].name.c_str() : nullptr;
}
int sde_result_times(const sde_result* r, double* dst, size_t n) {
    return guarded([&] {
        if (!r || !dst) throw ExprError{"NULL argument"};
        const std::vector<double>& t = r->plan->universe().times;
        if (n < t.size()) throw ExprError{"destination too small"};
        std::memcpy(dst, t.data(), t.size() * sizeof(double));
    });
}
int sde_result_moments(const sde_result* r, double* dst) {
    return guarded([&] {
        if (!r || !dst) throw ExprError{"NULL argument"};
        if (r->plan->options().lower.out != OUT_MOMENTS) throw ExprError{"sde_result_moments needs output = SDE_OUT_MOMENTS"};
        use_device(r->plan->device());
        cu_check(driver().cuMemcpyDtoH(dst, r->values.ptr(), r->plan->universe().processes.size() * 3 * sizeof(double)), "cuMemcpyDtoH");
    });
}

// ---- several GPUs, one host thread
void sde_shard_range(uint64_t n_scenarios, size_t part, size_t n_parts, uint64_t* lo, uint64_t* hi) {
    if (n_parts == 0 || part >= n_parts) { if (lo) *lo = 0; if (hi) *hi = 0; return; }
    shard_range(n_scenarios, part, n_parts, lo, hi);
}
int sde_device_plans_create(const sde_universe* u, const char* scheme, const char* rng_method, const sde_options* opt,
                            const int32_t* devices, size_t n_devices, sde_device_plans** out) {
    return guarded([&] {
        if (!u || !out) throw ExprError{"NULL argument"};
        *out = nullptr;
        if (opt && opt->inject) throw ExprError{"injected draws are a single-device test hook"};
        PlanOptions po = plan_options(u->u, scheme, rng_method, opt);
        std::vector<int> devs;
        if (devices) devs.assign(devices, devices + n_devices);
        else for (int i = 0; i < device_count(); ++i) devs.push_back(i);          // NULL: every visible device
        auto h = std::make_unique<sde_device_plans>();
        h->ps = std::make_unique<DevicePlans>(u->u, po, devs);
        *out = h.release();
    });
}
void sde_device_plans_free(sde_device_plans* ps) { delete ps; }
size_t sde_device_plans_count(const sde_device_plans* ps) { return ps ? ps->ps->size() : 0; }
int sde_device_plans_device(const sde_device_plans* ps, size_t i) { return (ps && i < ps->ps->size()) ? ps->ps->device(i) : -1; }
int sde_device_plans_collective(const sde_device_plans* ps) { return ps ? ps->ps->collective() : 0; }
int sde_plan_run_devices(sde_device_plans* ps, const char* const* init_names, const double* init_vals, size_t n_init,
                         uint64_t n_scenarios, uint64_t seed, uint64_t scenario_offset, double* const* d_out, int* n_launches,
                         double* collective_ms) {
    return guarded([&] {
        if (!ps || !d_out) throw ExprError{"NULL argument"};
        if (n_scenarios == 0) throw ExprError{"scenarios must be a positive integer"};
        ps->ps->run(init_pairs(init_names, init_vals, n_init), n_scenarios, seed, scenario_offset, d_out, n_launches, collective_ms);
    });
}
int sde_simulate_devices(const sde_universe* u, const char* const* init_names, const double* init_vals, size_t n_init,
                         uint64_t n_scenarios, const char* scheme, const char* rng_method, const sde_options* opt,
                         const int32_t* devices, size_t n_devices, sde_result** out) {
    return guarded([&] {
        if (!u || !out) throw ExprError{"NULL argument"};
        if (n_scenarios == 0) throw ExprError{"scenarios must be a positive integer"};   // py_binding.rs:20-24
        if (opt && opt->inject) throw ExprError{"injected draws are a single-device test hook"};
        PlanOptions po = plan_options(u->u, scheme, rng_method, opt);
        sde_options o;
        sde_options_default(&o);
        if (opt) std::memcpy(&o, opt, std::min<size_t>(sizeof o, opt->struct_size ? opt->struct_size : sizeof o));
        std::vector<int> devs;
        if (devices) devs.assign(devices, devices + n_devices);
        else for (int i = 0; i < device_count(); ++i) devs.push_back(i);
        for (size_t i = 0; i < devs.size(); ++i) out[i] = nullptr;
        DevicePlans ps(u->u, po, devs);
        const bool moments = po.lower.out == OUT_MOMENTS;
        std::vector<std::unique_ptr<sde_result>> rs;
        std::vector<double*> d_out;
        for (size_t i = 0; i < ps.size(); ++i) {
            uint64_t lo, hi;
            shard_range(n_scenarios, i, ps.size(), &lo, &hi);
            auto r = std::make_unique<sde_result>();
            r->plan = ps.plan(i);
            r->n = moments ? n_scenarios : hi - lo;          // merged moments describe the whole run
            r->scenario_offset = o.scenario_offset + (moments ? 0 : lo);
            r->elems = moments ? r->plan->output_elems(1) : (hi > lo ? r->plan->output_elems(hi - lo) : 0);
            use_device(ps.device(i));
            if (r->elems) r->values.alloc(moments ? r->elems * 8 : r->plan->output_bytes(hi - lo));
            d_out.push_back(r->values.as<double>());
            rs.push_back(std::move(r));
        }
        ps.run(init_pairs(init_names, init_vals, n_init), n_scenarios, o.seed, o.scenario_offset, d_out.data(), nullptr, nullptr);
        for (size_t i = 0; i < rs.size(); ++i) out[i] = rs[i].release();
    });
}
int sde_moments_merge_device(int device, const double* d_shards, size_t n_shards, size_t n_processes, double* d_out, void* stream) {
    return guarded([&] {
        if (!d_shards || !d_out) throw ExprError{"NULL argument"};
        use_device(device);
        moments_merge_device(device, d_shards, n_shards, n_processes, d_out, (CUstream)stream);
    });
}

int sde_sobol_points(int device, uint32_t dims, uint64_t first, uint64_t count, uint64_t* h_out) {
    return guarded([&] { util_sobol_points(device, dims, first, count, h_out); });
}
int sde_sobol_cp_shift_uniforms(int device, uint32_t dims, uint64_t seed, uint64_t first_scenario, uint64_t count, double* h_out) {
    return guarded([&] { util_sobol_cp_uniforms(device, dims, seed, first_scenario, count, h_out); });
}
int sde_joe_kuo_params(uint32_t dims, uint32_t* poly, uint32_t* minit) {
    return guarded([&] { joe_kuo_params(dims, poly, minit); });
}
int sde_chacha8_u64(int device, uint64_t seed, size_t n, uint64_t* h_out) {
    return guarded([&] { util_chacha8_u64(device, seed, n, h_out); });
}
int sde_icdf_normal(int device, int mode, const double* h_p, size_t n, double* h_out) {
    return guarded([&] { util_icdf_normal(device, mode, h_p, n, h_out); });
}
int sde_icdf_poisson(int device, const double* h_u, const double* h_lambda, size_t n, double* h_out) {
    return guarded([&] { util_icdf_poisson(device, h_u, h_lambda, n, h_out); });
}
int sde_moments_merge(const double* shards, size_t n_shards, size_t n_processes, double* out) {
    return guarded([&] { moments_merge(shards, n_shards, n_processes, out); });
}
int sde_measure_peaks(int device, double* fill_gbs, double* dfma_tflops, double* ffma_tflops) {
    return guarded([&] { util_measure_peaks(device, fill_gbs, dfma_tflops, ffma_tflops); });
}

}  // extern "C"
