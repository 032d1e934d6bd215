// sde_sim_wide.cuh — tensor-core variant of the fused path-simulation kernel for wide linear models (sm_100a).
//
// Same contract as sde_sim_kernel.cuh (one launch = the parallel region of sim::simulate, src/sim/mod.rs:41-88),
// selected by the lowering when the model is a Cholesky-loaded basket — every process Levy with coefficients
// a_j X_i on dt / dW_k only (src/sim/euler.rs:15-28 applied to `( a * S_i ) * dW_k` terms) — driven by scrambled
// Sobol points: BASELINE config C4 (64 assets, 64 factors, terminal moments), and the same models with terminal values
// or full paths in reference order.  One Euler step of such a model is
//     X_i <- X_i * (1 + a_i dt + sqrt(dt) * sum_k M[i][k] z_k),
// i.e. per step a [paths x K] * [K x P] matrix product in f64: GEMM-shaped work, so it runs on the FP64 tensor
// path (DMMA, mma.sync.m8n8k4.f64) instead of P*K scalar FMAs per path with one constant-memory operand each.
//
// Mapping.  A warp owns 8*SDE_WMT paths (SDE_WMT row tiles of 8) for the whole time grid; with r = lane >> 2 and
// c = lane & 3 the m8n8k4 fragments are
//     A (draws)      z[path 8m + r][factor 4kk + c]                       one value per lane per (m, kk)
//     B (loadings)   M[process 8j + r][factor 4kk + c]                    shared memory, fragment order, CTA prologue
//     C (sums)       acc[path 8m + r][process 8j + 2c + {0,1}]            two values per lane per (m, j)
// so a lane generates exactly the draws its A fragments need (no shuffles): the Sobol integer of (path, factor) by
// index, x_d(n0 + 8m + r) = x_d(n0) ^ x_d(8m + r) (GF(2) linearity, sde_device_rng.cuh), the digital shift and the
// inverse normal of sde_device_icdf.cuh.  The state X lives in C-fragment layout in registers.  Triangular loadings
// (M[i][k] = 0 for k beyond the process block's last factor) skip their k-steps at compile time.
// Warps are autonomous (no block barrier after the CTA prologue): each folds the warp part x_d(n0) of the step's
// dimensions itself from the dimension-fastest nibble table (coalesced, L2/L1 resident) and reads the lane part
// from a global table laid out in fragment order (one 128-byte line per (kk, m)).
//
// Macros expected from the generated prelude (besides SDE_P, SDE_K, SDE_RNG, SDE_OUT, SDE_ICDF, SDE_BLOCK):
//   SDE_S                 steps (compile time: the plan owns the time grid)
//   SDE_WNB, SDE_WNKK     process tiles of 8, factor steps of 4 (P and K rounded up)
//   SDE_WMT               row tiles per warp (1 or 2)
//   sde_wm[SDE_WNB * SDE_WNKK * 32]   loadings in B-fragment order;  sde_wa[8 * SDE_WNB]  drift coefficients a_i
//   sde_wkk_end(j)        constexpr: factor steps process tile j needs
#pragma once
#include "sde_sim_common.cuh"

#if SDE_RNG != 2 || !(SDE_OUT == 0 || SDE_OUT == 2 || SDE_OUT == 3)
#error "sde_sim_wide.cuh: Sobol with XOR digital shift; full paths [N][T][P], terminal values or moments"
#endif
#ifndef SDE_ICDF_WIDE
#define SDE_ICDF_WIDE 0
#endif
#define SDE_NW (SDE_BLOCK / 32)
#define SDE_WPATHS (8 * SDE_WMT)
#define SDE_WKP (4 * SDE_WNKK)
#define SDE_WPP (8 * SDE_WNB)
#define SDE_SK (SDE_S * SDE_K)
#define SDE_NIB_LD ((SDE_SK + 31) & ~31)      /* leading dimension of the transposed nibble table */
// shared-memory carve-up (bytes); mirrored by the host in lower.cpp
#define SDE_SMEM_ICDF_BYTES ((SDE_ICDF == 1) ? ((SDE_ICDF_WIDE ? SDE_ICDF_WIDE_DOUBLES : SDE_ICDF_TABLE_DOUBLES) * 8) : 0)
#define SDE_SMEM_WM_BYTES (SDE_WNB * SDE_WNKK * 32 * 8)
#define SDE_SMEM_WA_BYTES (SDE_WPP * 8)
#define SDE_SMEM_BW_BYTES (SDE_NW * SDE_WKP * 4)
#define SDE_SMEM_MOM_BYTES ((SDE_OUT == 3) ? (SDE_NW * SDE_WPP * 3 * 8) : 0)
#define SDE_SMEM_BYTES (SDE_SMEM_ICDF_BYTES + SDE_SMEM_WM_BYTES + SDE_SMEM_WA_BYTES + SDE_SMEM_MOM_BYTES + SDE_SMEM_BW_BYTES)

__device__ __forceinline__ void sde_dmma884(double (&c)[2], const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 1) sde_sim_kernel(const SdeParams prm) {
    extern __shared__ double4 sde_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(sde_smem_raw);
    double* s_icdf = reinterpret_cast<double*>(smem);
    double* s_wm = reinterpret_cast<double*>(smem + SDE_SMEM_ICDF_BYTES);
    double* s_wa = s_wm + SDE_WNB * SDE_WNKK * 32;
    double* s_mom = s_wa + SDE_WPP;
    sde_u32* s_bw = reinterpret_cast<sde_u32*>(smem + SDE_SMEM_BYTES - SDE_SMEM_BW_BYTES);
    (void)s_icdf; (void)s_mom;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;                  // fragment row / column group of this lane
    constexpr int S = SDE_S;

    // ---- CTA prologue: inverse-normal table, loadings in fragment order, drift coefficients, zeroed moment slots
#if SDE_ICDF == 1
#if SDE_ICDF_WIDE
    sde_icdf_wide_table_build(s_icdf, tid, SDE_BLOCK, SDE_ICDF_Y_OFFSET_K32);
#else
    sde_icdf_table_load(s_icdf, tid, SDE_BLOCK, SDE_ICDF_Y_OFFSET_K32);
#endif
#endif
    for (int e = tid; e < SDE_WNB * SDE_WNKK * 32; e += SDE_BLOCK) s_wm[e] = sde_wm[e];
    for (int e = tid; e < SDE_WPP; e += SDE_BLOCK) s_wa[e] = sde_wa[e];
#if SDE_OUT == 3
    for (int e = tid; e < SDE_NW * SDE_WPP * 3; e += SDE_BLOCK) s_mom[e] = 0.0;
#endif
    for (int e = tid; e < SDE_NW * SDE_WKP; e += SDE_BLOCK) s_bw[e] = 0u;     // pad factors keep a (harmless) zero
    __syncthreads();

    sde_u32* const my_bw = s_bw + warp * SDE_WKP;
#if SDE_OUT == 3
    double* const my_mom = s_mom + warp * SDE_WPP * 3;
#endif
#if SDE_ICDF == 1
    const sde_u32 tab_lane = (sde_u32)__cvta_generic_to_shared(s_icdf + 2 * (lane & (SDE_ICDF_TABLE_REPL - 1)));
#endif

    const sde_u64 first_n = prm.scen_offset + 5ull;           // Sobol::new(..).skip(5)  (sobol.rs:17)
    const sde_u64 n_base = first_n & ~(sde_u64)(SDE_WPATHS - 1);
    const sde_u64 n_items = (first_n + prm.n_paths - n_base + (SDE_WPATHS - 1)) / SDE_WPATHS;
    const sde_u64 item_stride = (sde_u64)gridDim.x * SDE_NW;

#pragma unroll 1
    for (sde_u64 item = (sde_u64)blockIdx.x * SDE_NW + warp; item < n_items; item += item_stride) {
        const sde_u64 n0 = n_base + item * SDE_WPATHS;        // multiple of 8 SDE_WMT: x_d(n0 + j) = x_d(n0) ^ x_d(j)
        bool valid[SDE_WMT];
        long long s_local[SDE_WMT];
#pragma unroll
        for (int m = 0; m < SDE_WMT; ++m) {
            const sde_u64 n = n0 + (sde_u64)(8 * m + fr);
            valid[m] = (n >= first_n) && (n - first_n < prm.n_paths);
            s_local[m] = (long long)(n - first_n);
        }
        // nibble-table rows of gray(n0): x_d(n0) = XOR over the 8 nibbles of a 16-entry table (independent loads)
        sde_u32 off[8];
        {
            const sde_u32 g = (sde_u32)n0 ^ ((sde_u32)n0 >> 1);
#pragma unroll
            for (int q = 0; q < 8; ++q) off[q] = (sde_u32)(q * 16 + ((g >> (4 * q)) & 15u)) * SDE_NIB_LD + lane;
        }
        // ScenarioFiltration::new — row 0 from initial_values (filtration.rs:42-50), in C-fragment layout
        double X[SDE_WMT][SDE_WNB][2];
#pragma unroll
        for (int j = 0; j < SDE_WNB; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int p = 8 * j + 2 * fc + e;
                const double v = p < SDE_P ? __ldg(prm.x0 + p) : 0.0;
#pragma unroll
                for (int m = 0; m < SDE_WMT; ++m) X[m][j][e] = v;
            }

#if SDE_OUT == 0
        // full paths in reference order [N][T][P] (filtration.rs:87-113): the 4 lanes of a fragment row write 8 consecutive
        // processes of one path's row, 64 contiguous bytes = two whole sectors per (m, j)
        double* row_ptr[SDE_WMT];
#pragma unroll
        for (int m = 0; m < SDE_WMT; ++m) {
            row_ptr[m] = prm.out + (size_t)(valid[m] ? s_local[m] : 0) * (S + 1) * SDE_P + 2 * fc;
            if (valid[m]) {
#pragma unroll
                for (int j = 0; j < SDE_WNB; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int p = 8 * j + 2 * fc + e;
                        if (p < SDE_P) row_ptr[m][8 * j + e] = __ldg(prm.x0 + p);
                    }
            }
        }
#endif
#pragma unroll 1
        for (int t = 0; t < S; ++t) {
            // ---- warp part of this step's dimensions, digital shift folded in: my_bw[k] = x_d(n0) ^ mask_d, d = t K + k
            __syncwarp();
#pragma unroll
            for (int k0 = 0; k0 < SDE_K; k0 += 32) {
                if (k0 + lane < SDE_K) {
                    const sde_u32 d = (sde_u32)(t * SDE_K + k0);
                    sde_u32 v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = __ldg(prm.sobol_nib + off[q] + d);
                    const sde_u32 mask = __ldg(prm.xor_masks + d + lane);
                    my_bw[k0 + lane] = (((v[0] ^ v[1]) ^ (v[2] ^ v[3])) ^ ((v[4] ^ v[5]) ^ (v[6] ^ v[7]))) ^ mask;
                }
            }
            __syncwarp();
            const double dt = __ldg(prm.dts + t), sq = __ldg(prm.sqrt_dts + t);

            // ---- draws in A-fragment layout: z[m][kk] = N^-1(u) of (path 8m + r, factor 4kk + c)
            double z[SDE_WMT][SDE_WNKK];
            {
                const sde_u32* lt = prm.sobol_lane + (size_t)t * (SDE_WNKK * SDE_WMT * 32) + lane;
#pragma unroll
                for (int kk = 0; kk < SDE_WNKK; ++kk) {
                    const sde_u32 bw = my_bw[4 * kk + fc];
#pragma unroll
                    for (int m = 0; m < SDE_WMT; ++m) {
                        const sde_u32 x = bw ^ __ldg(lt + (kk * SDE_WMT + m) * 32);   // u = (x + 1/2) 2^-32
#if SDE_ICDF == 1 && SDE_ICDF_WIDE
                        z[m][kk] = sde_icdf_normal_fast_k32w(x, tab_lane);
#elif SDE_ICDF == 1
                        z[m][kk] = sde_icdf_normal_fast_k32s(x, tab_lane);
#elif SDE_ICDF == 2
                        z[m][kk] = (double)sde_icdf_normal_single_k32(x);
#else
                        z[m][kk] = sde_icdf_normal_reference(fma((double)x, 2.3283064365386963e-10, 1.1641532182693481e-10));
#endif
                    }
                }
            }

            // ---- correlation product on the FP64 tensor path and the state update, one process tile at a time
#pragma unroll
            for (int j = 0; j < SDE_WNB; ++j) {
                double acc[SDE_WMT][2];
#pragma unroll
                for (int m = 0; m < SDE_WMT; ++m) acc[m][0] = acc[m][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < SDE_WNKK; ++kk) {
                    if (kk < sde_wkk_end(j)) {
                        const double b = s_wm[(j * SDE_WNKK + kk) * 32 + lane];
#pragma unroll
                        for (int m = 0; m < SDE_WMT; ++m) sde_dmma884(acc[m], z[m][kk], b);
                    }
                }
                const double2 a = *reinterpret_cast<const double2*>(s_wa + 8 * j + 2 * fc);
                const double g0 = fma(a.x, dt, 1.0), g1 = fma(a.y, dt, 1.0);
#pragma unroll
                for (int m = 0; m < SDE_WMT; ++m) {
                    X[m][j][0] *= fma(acc[m][0], sq, g0);
                    X[m][j][1] *= fma(acc[m][1], sq, g1);
#if SDE_OUT == 0
                    if (valid[m]) {
                        double* dst = row_ptr[m] + (size_t)(t + 1) * SDE_P + 8 * j;
                        if ((SDE_P & 1) == 0 && 8 * j + 8 <= SDE_P) {
                            *reinterpret_cast<double2*>(dst) = make_double2(X[m][j][0], X[m][j][1]);   // P even: 16-byte aligned
                        } else {
                            if (8 * j + 2 * fc < SDE_P) dst[0] = X[m][j][0];
                            if (8 * j + 2 * fc + 1 < SDE_P) dst[1] = X[m][j][1];
                        }
                    }
#endif
                }
            }
        }

#if SDE_OUT == 0
        // rows already stored step by step
#elif SDE_OUT == 2
#pragma unroll
        for (int m = 0; m < SDE_WMT; ++m)
            if (valid[m]) {
                double* dst = prm.out + (size_t)s_local[m] * SDE_P + 2 * fc;
#pragma unroll
                for (int j = 0; j < SDE_WNB; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (8 * j + 2 * fc + e < SDE_P) dst[8 * j + e] = X[m][j][e];
            }
#else
        // (count, mean, M2) of the item's paths per process: row tiles in the lane, rows by shuffle, then into the warp's
        // running moments; fixed order => deterministic
#pragma unroll
        for (int j = 0; j < SDE_WNB; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                SdeMoments mm;
                mm.n = valid[0] ? 1.0 : 0.0; mm.mean = valid[0] ? X[0][j][e] : 0.0; mm.m2 = 0.0;
#pragma unroll
                for (int m = 1; m < SDE_WMT; ++m) {
                    SdeMoments o;
                    o.n = valid[m] ? 1.0 : 0.0; o.mean = valid[m] ? X[m][j][e] : 0.0; o.m2 = 0.0;
                    mm = sde_mom_merge(mm, o);
                }
#pragma unroll
                for (int sh = 4; sh < 32; sh <<= 1) {
                    SdeMoments o;
                    o.n = __shfl_xor_sync(0xffffffffu, mm.n, sh);
                    o.mean = __shfl_xor_sync(0xffffffffu, mm.mean, sh);
                    o.m2 = __shfl_xor_sync(0xffffffffu, mm.m2, sh);
                    mm = ((lane & sh) == 0) ? sde_mom_merge(mm, o) : sde_mom_merge(o, mm);
                }
                const int p = 8 * j + 2 * fc + e;
                if (fr == 0 && p < SDE_P) {
                    SdeMoments run;
                    run.n = my_mom[p * 3]; run.mean = my_mom[p * 3 + 1]; run.m2 = my_mom[p * 3 + 2];
                    run = sde_mom_merge(run, mm);
                    my_mom[p * 3] = run.n; my_mom[p * 3 + 1] = run.mean; my_mom[p * 3 + 2] = run.m2;
                }
            }
#endif
    }

#if SDE_OUT == 3
    // one partial per warp: [grid * SDE_NW][P][3], folded by sde_moments_finalize in a fixed order
    __syncwarp();
    {
        double* dst = prm.partials + (size_t)(blockIdx.x * SDE_NW + warp) * SDE_P * 3;
        for (int e = lane; e < SDE_P * 3; e += 32) dst[e] = my_mom[e];
    }
#endif
}
