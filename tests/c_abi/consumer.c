/* consumer.c — a plain C caller of libsde_b200.so, the way the reference's Rust crate would call it through a -sys
 * binding: parse_equations -> simulate -> read the `value` column.  Built and run by
 * tests/test_c_abi.py with `gcc -std=c99` (so the header is checked as C, not C++).
 *
 *   consumer <n_scenarios> <n_steps> <seed> <rng_method> <scheme>
 * prints one line per check; on a machine without a GPU the simulate call must fail with SDE_ERR_RUNTIME
 * (no CPU fallback) and the program says so and exits 3.                                                     */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sde_b200.h"

int main(int argc, char** argv) {
    const unsigned long long n = argc > 1 ? strtoull(argv[1], NULL, 10) : 1000ull;
    const size_t steps = argc > 2 ? (size_t)strtoull(argv[2], NULL, 10) : 16;
    const unsigned long long seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 42ull;
    const char* rng = argc > 4 ? argv[4] : "pseudo";
    const char* scheme = argc > 5 ? argv[5] : "euler";

    /* the equations of examples/example_gbm.py */
    const char* eqs[1] = {"dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"};
    double* times = (double*)malloc((steps + 1) * sizeof(double));
    for (size_t k = 0; k <= steps; ++k) times[k] = (double)k / 252.0;

    printf("version: %s\n", sde_version());

    /* parse errors are value errors (py_binding.rs:30-32) */
    const char* bad[1] = {"dX1 = ( 0.05 * X1 * dt"};
    sde_universe* u = NULL;
    int rc = sde_universe_parse(bad, 1, times, steps + 1, &u);
    printf("bad equation: rc=%d\n", rc);
    if (rc != SDE_ERR_VALUE || u != NULL) return 1;

    rc = sde_universe_parse(eqs, 1, times, steps + 1, &u);
    if (rc != SDE_OK) { printf("parse failed: %s\n", sde_last_error()); return 1; }
    printf("universe: P=%zu K=%zu T=%zu name=%s factor=%s\n", sde_universe_num_processes(u), sde_universe_num_factors(u),
           sde_universe_num_times(u), sde_universe_process_name(u, 0), sde_universe_factor_name(u, 0));

    sde_options opt;
    sde_options_default(&opt);
    if (opt.struct_size != sizeof(sde_options)) { printf("struct size mismatch\n"); return 1; }
    opt.seed = seed;
    if (strcmp(rng, "sobol") == 0) opt.scramble = SDE_SCRAMBLE_XOR;

    const char* names[1] = {"X1"};
    const double vals[1] = {1.0};
    sde_result* res = NULL;
    rc = sde_simulate(u, names, vals, 1, n, scheme, rng, &opt, &res);
    if (rc == SDE_ERR_RUNTIME && !sde_cuda_available()) {
        printf("no GPU: simulate rc=%d (%s) — no CPU fallback\n", rc, sde_last_error());
        sde_universe_free(u);
        free(times);
        return 3;
    }
    if (rc != SDE_OK) { printf("simulate failed rc=%d: %s\n", rc, sde_last_error()); return 1; }

    uint64_t rn = 0; size_t rt = 0, rp = 0;
    sde_result_shape(res, &rn, &rt, &rp);
    const size_t ne = sde_result_num_elems(res);
    printf("result: N=%llu T=%zu P=%zu elems=%zu kernel_ms=%.4f\n", (unsigned long long)rn, rt, rp, ne, sde_result_kernel_ms(res));
    if (rn != n || rt != steps + 1 || rp != 1 || ne != (size_t)n * (steps + 1)) return 1;
    double* v = (double*)malloc(ne * sizeof(double));
    rc = sde_result_values_host(res, v, ne);
    if (rc != SDE_OK) { printf("copy failed: %s\n", sde_last_error()); return 1; }
    double sum = 0.0, sum_t = 0.0;
    int ok = 1;
    for (size_t s = 0; s < (size_t)n; ++s) {
        if (v[s * rt] != 1.0) ok = 0;                         /* t0 row = initial value (filtration.rs:42-50) */
        for (size_t t = 0; t < rt; ++t) { if (!(v[s * rt + t] > 0.0) || !isfinite(v[s * rt + t])) ok = 0; sum += v[s * rt + t]; }
        sum_t += v[s * rt + rt - 1];
    }
    printf("checks: %s\n", ok ? "ok" : "FAILED");
    printf("sum=%.17g\nterminal_mean=%.17g\nfirst_path_terminal=%.17g\n", sum, sum_t / (double)n, v[rt - 1]);
    /* what else the reference's frame carries (filtration.rs:108-113): names, times, first scenario */
    {
        double* tt = (double*)malloc(rt * sizeof(double));
        if (sde_result_times(res, tt, rt) != SDE_OK || tt[0] != times[0] || tt[rt - 1] != times[rt - 1]) ok = 0;
        if (strcmp(sde_result_process_name(res, 0), "X1") != 0 || sde_result_scenario_offset(res) != 0 || sde_result_output(res) != SDE_OUT_PATHS) ok = 0;
        free(tt);
    }

    /* ---- the same run sharded over every GPU of the box in ONE call (rayon's par_iter over scenarios, sim/mod.rs:41-43):
     *      the union of the shards must be bit-identical to the single-device result; then the moments, merged on the devices */
    {
        const int n_dev = sde_device_count();
        sde_result** shards = (sde_result**)calloc((size_t)n_dev, sizeof(sde_result*));
        rc = sde_simulate_devices(u, names, vals, 1, n, scheme, rng, &opt, NULL, (size_t)n_dev, shards);
        if (rc != SDE_OK) { printf("simulate_devices failed rc=%d: %s\n", rc, sde_last_error()); return 1; }
        size_t at = 0;
        int same = 1;
        for (int i = 0; i < n_dev; ++i) {
            uint64_t lo = 0, hi = 0, sn = 0;
            sde_shard_range(n, (size_t)i, (size_t)n_dev, &lo, &hi);
            sde_result_shape(shards[i], &sn, NULL, NULL);
            if (sn != hi - lo || sde_result_scenario_offset(shards[i]) != lo || sde_result_device(shards[i]) != i) same = 0;
            const size_t se = sde_result_num_elems(shards[i]);
            if (se) {
                double* sv = (double*)malloc(se * sizeof(double));
                if (sde_result_values_host(shards[i], sv, se) != SDE_OK || memcmp(sv, v + at, se * sizeof(double)) != 0) same = 0;
                free(sv);
            }
            at += se;
            sde_result_free(shards[i]);
        }
        if (at != ne) same = 0;
        printf("devices: %d shard union %s\n", n_dev, same ? "bit-identical" : "DIFFERS");
        if (!same) ok = 0;

        sde_options mo = opt;
        mo.output = SDE_OUT_MOMENTS;
        rc = sde_simulate_devices(u, names, vals, 1, n, scheme, rng, &mo, NULL, (size_t)n_dev, shards);
        if (rc != SDE_OK) { printf("simulate_devices(moments) failed rc=%d: %s\n", rc, sde_last_error()); return 1; }
        double mom[3] = {0, 0, 0}, mean = sum_t / (double)n, m2 = 0.0;
        for (size_t s = 0; s < (size_t)n; ++s) { const double d = v[s * rt + rt - 1] - mean; m2 += d * d; }
        int mom_ok = 1;
        for (int i = 0; i < n_dev; ++i) {                      /* every device holds the merged triple */
            if (sde_result_moments(shards[i], mom) != SDE_OK) mom_ok = 0;
            if (mom[0] != (double)n || fabs(mom[1] / mean - 1.0) > 1e-13 || fabs(mom[2] / m2 - 1.0) > 1e-10) mom_ok = 0;
            sde_result_free(shards[i]);
        }
        printf("devices: merged moments %s (count=%.0f mean=%.17g)\n", mom_ok ? "ok" : "FAILED", mom[0], mom[1]);
        if (!mom_ok) ok = 0;
        free(shards);
    }
    printf("all checks: %s\n", ok ? "ok" : "FAILED");
    free(v);
    sde_result_free(res);
    sde_universe_free(u);
    free(times);
    return ok ? 0 : 1;
}
