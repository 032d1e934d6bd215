// Micro-benchmark: how good are the MUFU seeds the fast inverse normal polishes?  Decides between the cubic and a
// quadratic Newton step for sqrt / reciprocal in sde_device_icdf.cuh (sde_icdf_as_tail).  B200, sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_seed mufu_seed.cu && ./mufu_seed
// For w2 = -2 ln w over the range the digital-shift uniforms reach ([1.386, 45.75]) and D(t) in [2.9, 20]:
//   rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64 (MUFU.RSQ64H / RCP64H: only the high word of the operand is read),
//   the same with the operand's low word cleared (the truncation alone), and the FP32 units on the operand truncated
//   to 24 bits (bit tricks instead of conversions), each against the correctly rounded f64 result.
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__device__ double rsq64h(double a) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); return r; }
__device__ double rcp64h(double a) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); return r; }
__device__ float rsq32(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ float rcp32(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

// f64 -> f32 by truncation with integer instructions (hi << 3 | lo >> 29, exponent re-biased), and back
__device__ float trunc_to_f32(double a) {
    const unsigned hi = (unsigned)__double2hiint(a), lo = (unsigned)__double2loint(a);
    return __uint_as_float(__funnelshift_l(lo, hi, 3) - 0xC0000000u);      // (e - 896) << 23: e in [897, 1150]
}
__device__ double f32_bits_to_f64(float f) {
    const unsigned b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}

struct Stat { double max_abs, min_signed, max_signed; };

__global__ void k_err(int mode, double lo, double hi, unsigned long long n, Stat* out) {
    double mx = 0, mn = 0, ma = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        // geometric sweep with a golden-ratio dither: every binade of [lo, hi] densely, all low-word patterns
        const double u = (double)i / (double)n;
        double a = lo * exp(u * log(hi / lo));
        a = __longlong_as_double(__double_as_longlong(a) ^ (long long)((i * 0x9E3779B97F4A7C15ull) >> 34));
        double seed, exact;
        switch (mode) {
            case 0: seed = rsq64h(a); exact = 1.0 / sqrt(a); break;
            case 1: seed = rsq64h(__hiloint2double(__double2hiint(a), 0)); exact = 1.0 / sqrt(__hiloint2double(__double2hiint(a), 0)); break;
            case 2: seed = f32_bits_to_f64(rsq32(trunc_to_f32(a))); exact = 1.0 / sqrt(a); break;
            case 3: seed = rcp64h(a); exact = 1.0 / a; break;
            case 4: seed = rcp64h(__hiloint2double(__double2hiint(a), 0)); exact = 1.0 / __hiloint2double(__double2hiint(a), 0); break;
            default: seed = f32_bits_to_f64(rcp32(trunc_to_f32(a))); exact = 1.0 / a; break;
        }
        // only the high word of the 64H seeds is defined: compare the high-word value
        if (mode == 0 || mode == 1 || mode == 3 || mode == 4) seed = __hiloint2double(__double2hiint(seed), 0);
        const double e = (seed - exact) / exact;
        mx = fmax(mx, e); mn = fmin(mn, e); ma = fmax(ma, fabs(e));
    }
    for (int o = 16; o; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // (non-atomic max over warps through a per-warp slot)
        Stat* s = out + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
        s->max_abs = ma; s->min_signed = mn; s->max_signed = mx;
    }
}

int main() {
    const int grid = 148 * 4, block = 256, slots = grid * block / 32;
    Stat* st;
    cudaMallocManaged(&st, slots * sizeof(Stat));
    const char* names[] = {"rsqrt.approx.f64 (MUFU.RSQ64H), full operand", "rsqrt.approx.f64, operand low word = 0", "MUFU.RSQ f32 on the operand truncated to 24 bits",
                           "rcp.approx.f64 (MUFU.RCP64H), full operand", "rcp.approx.f64, operand low word = 0", "MUFU.RCP f32 on the operand truncated to 24 bits"};
    for (int mode = 0; mode < 6; ++mode) {
        const double lo = mode < 3 ? 1.386 : 2.9, hi = mode < 3 ? 45.75 : 20.0;
        k_err<<<grid, block>>>(mode, lo, hi, 1ull << 32, st);
        cudaDeviceSynchronize();
        double ma = 0, mn = 0, mx = 0;
        for (int i = 0; i < slots; ++i) { ma = fmax(ma, st[i].max_abs); mn = fmin(mn, st[i].min_signed); mx = fmax(mx, st[i].max_signed); }
        printf("%-55s rel err in [%+.3e, %+.3e]   max |e| = 2^%.2f\n", names[mode], mn, mx, log2(ma));
    }
    return 0;
}
