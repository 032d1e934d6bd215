// sde_sim_common.cuh — pieces shared by the two variants of the fused path-simulation kernel
// (sde_sim_kernel.cuh: time-tiled CTAs; sde_sim_resident.cuh: persistent warps with resident tables):
// the kernel parameter block, the (count, mean, M2) moment algebra and the second reduction stage.
#pragma once
#include "sde_device_rng.cuh"
#include "sde_device_icdf.cuh"

struct SdeParams {
    sde_u64 n_paths;       // local scenario count N
    sde_u64 scen_offset;   // global index of local scenario 0
    sde_u64 n_base;        // point index of thread 0 of CTA 0 (multiple of SDE_BLOCK)
    sde_u64 seed;
    int n_steps;           // S = T - 1
    int reserved;
    const double* times;      // [T]
    const double* dts;        // [S]   times[t+1] - times[t]            (increment.rs:38-41)
    const double* sqrt_dts;   // [S]   sqrt(dts[t])                     (increment.rs:75-79)
    const double* x0;         // [P]   row 0                            (filtration.rs:42-50)
    const sde_u32* sobol_nib;   // [S*K][8][16]  XOR of direction numbers selected by nibble i of gray(n)
                                //               (persistent kernel: transposed, [8][16][S*K rounded up to 32])
    const sde_u32* sobol_lane;  // [S*K][32]     x_d(lane)
    const sde_u32* xor_masks;   // [S*K]         32-bit digital-shift masks
    const double* inject;     // [N][S][K+1]
    double* out;
    double* partials;         // moments: [grid][P][3]; persistent kernel: scratch row [T*P + 4] for the pad lanes' stores
};

struct SdeMoments { double n, mean, m2; };

// Chan et al. pairwise merge of (count, mean, M2).
__device__ __forceinline__ SdeMoments sde_mom_merge(const SdeMoments a, const SdeMoments b) {
    SdeMoments r;
    r.n = a.n + b.n;
    if (r.n == 0.0) { r.mean = 0.0; r.m2 = 0.0; return r; }
    const double d = b.mean - a.mean;
    const double f = b.n / r.n;
    r.mean = a.mean + d * f;
    r.m2 = a.m2 + b.m2 + d * d * a.n * f;
    return r;
}

// Second stage of the moment reduction: one CTA folds [n_partials][P][3] -> [P][3], fixed order.
extern "C" __global__ void __launch_bounds__(256) sde_moments_finalize(const double* __restrict__ partials, sde_u64 n_partials,
                                                                        double* __restrict__ out) {
    __shared__ double s[256 * 3];
    const int tid = threadIdx.x;
    for (int p = 0; p < SDE_P; ++p) {
        SdeMoments acc; acc.n = 0.0; acc.mean = 0.0; acc.m2 = 0.0;
        for (sde_u64 i = tid; i < n_partials; i += 256) {
            const double* src = partials + (i * SDE_P + p) * 3;
            SdeMoments o; o.n = src[0]; o.mean = src[1]; o.m2 = src[2];
            acc = sde_mom_merge(acc, o);
        }
        __syncthreads();
        s[tid * 3] = acc.n; s[tid * 3 + 1] = acc.mean; s[tid * 3 + 2] = acc.m2;
        __syncthreads();
        for (int stride = 128; stride > 0; stride >>= 1) {
            if (tid < stride) {
                SdeMoments a, b;
                a.n = s[tid * 3]; a.mean = s[tid * 3 + 1]; a.m2 = s[tid * 3 + 2];
                b.n = s[(tid + stride) * 3]; b.mean = s[(tid + stride) * 3 + 1]; b.m2 = s[(tid + stride) * 3 + 2];
                a = sde_mom_merge(a, b);
                s[tid * 3] = a.n; s[tid * 3 + 1] = a.mean; s[tid * 3 + 2] = a.m2;
            }
            __syncthreads();
        }
        if (tid == 0) { out[p * 3] = s[0]; out[p * 3 + 1] = s[1]; out[p * 3 + 2] = s[2]; }
    }
}
