// sde_util_kernels.cu — model-independent kernels, compiled ahead of time for sm_100a
// (nvcc -cubin) and embedded in libsde_b200.so.  They expose each building block of the
// fused kernel on its own so that every one has its own parity test, and hold the
// microbenchmarks whose results sit beside MEASURED_PEAKS.json.
#include "sde_device_rng.cuh"
#include "sde_device_icdf.cuh"

// K1: integer Sobol points out[i][d] (u64, low 32 bits zero) for n = first + i.
// Same decomposition as the fused kernel: CTA-uniform part in shared memory, lane part from
// the x_d(lane) table.  Replaces sobol::Sobol::next (src/rng/sobol.rs:23-25).
extern "C" __global__ void __launch_bounds__(256) sde_k_sobol_points(const sde_u32* __restrict__ V, const sde_u32* __restrict__ lane_tab,
                                                                      sde_u32 dims, sde_u64 n_base, sde_u64 first, sde_u64 count,
                                                                      sde_u64* __restrict__ out) {
    extern __shared__ sde_u32 s_bw[];                    // [dims_tile][8]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const sde_u64 n_cta = n_base + (sde_u64)blockIdx.x * 256;
    const sde_u64 n = n_cta + tid;
    const bool valid = n >= first && n - first < count;
    const sde_u32 DT = 256;                              // dimensions per tile
    for (sde_u32 d0 = 0; d0 < dims; d0 += DT) {
        const sde_u32 nd = min(DT, dims - d0);
        __syncthreads();
        for (sde_u32 e = tid; e < nd * 8; e += 256) {
            const sde_u32 dl = e >> 3, w = e & 7;
            s_bw[e] = sde_sobol_point32(V + (size_t)(d0 + dl) * 32, (sde_u32)n_cta + 32u * w);
        }
        __syncthreads();
        if (valid) {
            for (sde_u32 dl = 0; dl < nd; ++dl) {
                const sde_u32 x = s_bw[dl * 8 + warp] ^ __ldg(lane_tab + (size_t)(d0 + dl) * 32 + lane);
                out[(n - first) * dims + d0 + dl] = ((sde_u64)x) << 32;
            }
        }
    }
}

// K1s: the uniforms of the reference's Sobol mode, u[i][d] = fract(x_d(n) + shift_d) for scenario first_scenario + i
// (n = scenario + 5): raw point rendered as x / 2^32 (exact: the points only touch the top 32 bits), shift_d = f64 draw #d of
// ChaCha8Rng::seed_from_u64(seed + scenario) (src/rng/sobol.rs:45-47,67-76; src/sim/mod.rs:56) — the same expression the
// fused kernel evaluates for SDE_RNG == 1, on its own for the bit-exact test.
extern "C" __global__ void __launch_bounds__(128) sde_k_sobol_cp_uniforms(const sde_u32* __restrict__ V, sde_u32 dims, sde_u64 seed,
                                                                          sde_u64 first_scenario, sde_u64 count, double* __restrict__ out) {
    const sde_u64 i = (sde_u64)blockIdx.x * 128 + threadIdx.x;
    if (i >= count) return;
    const sde_u64 s = first_scenario + i;
    SdeChaCha8Stream cha;
    cha.init(seed + s);
    for (sde_u32 d0 = 0; d0 < dims; d0 += 8) {
        cha.refill();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const sde_u32 d = d0 + q;
            if (d < dims) {
                const sde_u32 x = sde_sobol_point32(V + (size_t)d * 32, (sde_u32)(s + 5ull));
                const double v = (double)x * 2.3283064365386963e-10 + (double)(long long)cha.bits53(q) * 1.1102230246251565e-16;
                out[i * dims + d] = (v >= 1.0) ? v - 1.0 : v;
            }
        }
    }
}

// K2: first n u64 outputs of ChaCha8Rng::seed_from_u64(seed); thread i produces block i.
extern "C" __global__ void sde_k_chacha8_u64(sde_u64 seed, sde_u64 n, sde_u64* __restrict__ out) {
    const sde_u64 b = (sde_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 8 >= n) return;
    sde_u32 key[8], buf[16];
    sde_seed_from_u64(seed, key);
    sde_chacha_block<8>(key, b, buf);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (b * 8 + i < n) out[b * 8 + i] = ((sde_u64)buf[2 * i + 1] << 32) | buf[2 * i];
}

// K3: inverse normal CDF.  mode 0 reference evaluation, 1 fast (f64 entry), 2 fast through the 32-bit integer
// front end of the digital-shift path with the arithmetic exponent term (sde_icdf_normal_fast_k32s): p[i] must be
// (k + 1/2) 2^-32 for a 32-bit integer k.  mode 3 single (FP32 evaluation, f64 entry), 4 single through the 32-bit front end.
extern "C" __global__ void __launch_bounds__(256) sde_k_icdf_normal(const double* __restrict__ p, sde_u64 n, int mode, double* __restrict__ out) {
    __shared__ double4 s_raw[SDE_ICDF_TABLE_DOUBLES / 4];
    double* s_table = reinterpret_cast<double*>(s_raw);
    sde_icdf_table_load(s_table, threadIdx.x, 256, mode == 2 ? SDE_ICDF_Y_OFFSET_K32 : 0.0);
    __syncthreads();
    const sde_u64 i = (sde_u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (mode == 3) { out[i] = sde_icdf_normal_single(p[i]); return; }
    if (mode == 4) { out[i] = sde_icdf_normal_single_k32((sde_u32)(unsigned long long)(p[i] * 4294967296.0)); return; }
    if (mode == 2) {
        const sde_u32 k = (sde_u32)(unsigned long long)(p[i] * 4294967296.0);     // exact: p 2^32 = k + 1/2, truncated
        out[i] = sde_icdf_normal_fast_k32s(k, (sde_u32)__cvta_generic_to_shared(s_table + 2 * (threadIdx.x & (SDE_ICDF_TABLE_REPL - 1))));
    } else {
        out[i] = mode == 1 ? sde_icdf_normal_fast(p[i], s_table, threadIdx.x & 31) : sde_icdf_normal_reference(p[i]);
    }
}

// K3w: the 32-bit front end with the 1024-entry log table of the persistent kernel (128 KB of dynamic shared memory).
// f32seed = 1: the variant with FP32-unit seeds and quadratic steps (sde_icdf_as_tail_f32seed), through the sign-folded entry.
extern "C" __global__ void __launch_bounds__(256) sde_k_icdf_normal_wide(const double* __restrict__ p, sde_u64 n, double* __restrict__ out, int f32seed) {
    extern __shared__ double4 s_wide_raw[];
    double* s_table = reinterpret_cast<double*>(s_wide_raw);
    sde_icdf_wide_table_build(s_table, threadIdx.x, 256, SDE_ICDF_Y_OFFSET_K32, f32seed ? 16.0 : 1.0);
    __syncthreads();
    const sde_u32 tab_lane = (sde_u32)__cvta_generic_to_shared(s_table + 2 * (threadIdx.x & (SDE_ICDF_TABLE_REPL - 1)));
    for (sde_u64 i = (sde_u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (sde_u64)gridDim.x * 256) {
        const sde_u32 k = (sde_u32)(unsigned long long)(p[i] * 4294967296.0);
        const sde_u32 y = k ^ ((sde_u32)((int)k >> 31) & 0x7fffffffu);          // sign-folded word (sde_sim_resident.cuh)
        sde_u32 j;
        asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(j) : "r"(y));
        out[i] = f32seed ? sde_icdf_fast_j32w_t<1>(j, ~y, tab_lane) : sde_icdf_fast_j32w_t<0>(j, ~y, tab_lane);
    }
}

extern "C" __global__ void sde_k_icdf_poisson(const double* __restrict__ u, const double* __restrict__ lambda, sde_u64 n, double* __restrict__ out) {
    const sde_u64 i = (sde_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sde_icdf_poisson(u[i], lambda[i]);
}

// K8 (merge stage): per-shard (count, mean, M2) triples [n_shards][P][3] -> [P][3], Chan et al., shards in order with
// separately rounded operations — the same arithmetic, bit for bit, as the host's sde_moments_merge.  Runs on the compute
// stream right behind the all-gather of the triples (one thread per process; 3 P doubles per shard).
extern "C" __global__ void __launch_bounds__(128) sde_k_moments_merge(const double* __restrict__ shards, sde_u64 n_shards, int P,
                                                                       double* __restrict__ out) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        double n = 0.0, mean = 0.0, m2 = 0.0;
        for (sde_u64 s = 0; s < n_shards; ++s) {
            const double* b = shards + (s * (sde_u64)P + p) * 3;
            const double nb = b[0];
            if (nb == 0.0) continue;
            const double nn = __dadd_rn(n, nb), dlt = __dsub_rn(b[1], mean), f = __ddiv_rn(nb, nn);
            m2 = __dadd_rn(__dadd_rn(m2, b[2]), __dmul_rn(__dmul_rn(__dmul_rn(dlt, dlt), n), f));
            mean = __dadd_rn(mean, __dmul_rn(dlt, f));
            n = nn;
        }
        out[p * 3] = n; out[p * 3 + 1] = mean; out[p * 3 + 2] = m2;
    }
}

// ---- microbenchmarks ------------------------------------------------------------------
// pure-write bandwidth: 16-byte stores, grid-stride
extern "C" __global__ void __launch_bounds__(256) sde_k_fill(double2* __restrict__ dst, sde_u64 n_vec, double v) {
    const sde_u64 stride = (sde_u64)gridDim.x * 256;
    for (sde_u64 i = (sde_u64)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += stride) dst[i] = make_double2(v, v);
}
// FP64 FMA issue peak: 8 independent chains per thread
extern "C" __global__ void __launch_bounds__(256) sde_k_dfma(double* __restrict__ out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * 256 + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
extern "C" __global__ void __launch_bounds__(256) sde_k_ffma(float* __restrict__ out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[(size_t)blockIdx.x * 256 + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
