// sde_sim_resident.cuh — persistent-warp variant of the fused path-simulation kernel (sm_100a).
//
// Same contract as sde_sim_kernel.cuh (one launch = the parallel region of sim::simulate,
// src/sim/mod.rs:41-88; rows in the reference order of src/filtration.rs:87-113), selected by the
// lowering for Sobol-driven full-path output [N][T][P] when everything a path reads fits in shared
// memory for the WHOLE time grid:
//     lane table   x_d(4 lane) ^ mask_d        S K x 32 u32   (digital shift folded in once per CTA)
//     step records {t, t+dt, dt, sqrt dt, hoisted model constants}   S x (4 + NSLOT) f64
//     inverse-normal tables (sde_device_icdf.cuh)
//     per warp: x_d(n0) of the warp's current 32 paths             S K u32
// One CTA per SM stays resident and its warps walk the work items on their own:
//   item i = 32 paths  n = n0 + 4 lane,  n0 = n_base + 128 (i >> 2) + (i & 3)   (point index; scenario = n - 5,
//   src/rng/sobol.rs:17), i.e. the lanes of a warp own paths 4 apart, so all 32 rows share one phase modulo a
//   32-byte sector and every 4-step group leaves as one aligned 256-bit store per lane (st.global.v4.f64),
//   exactly like the direct path of the tiled kernel.
// There is no block barrier after the CTA prologue: a warp folds the Sobol part of its next item itself
// (8 nibble-table loads per dimension, L1/L2 resident), so warps drift apart and the FP64, integer and
// load/store phases of different warps overlap instead of marching in step; no tile prologue/epilogue code,
// no staged prefetch registers.
//
// Work split: a CTA only takes items of ONE class c = blockIdx.x & 3 (item = 4 m + c), so every warp of the CTA has
// the same step shift gamma, and the Sobol words a 4-step group reads are stored so that they are ONE aligned 128-bit
// shared-memory load per table:
//     lane table   [quad][lane][4]   word of dimension d at quad (d + off) >> 2, component (d + off) & 3
//     warp part    [d + off]         (broadcast load)
// with off = (-gamma K) mod 4 chosen per CTA so that the dimensions of a group start on a quad boundary.  The grid is
// a multiple of 4 CTAs (lower.cpp / engine.cpp).
// SDE_RES_FOLD (digital shift + fast inverse normal, all factors Wiener): both tables hold the words in sign-folded form
//     y = x ^ ((x >>s 31) & 0x7fffffff)        bit 31 = (p >= 1/2), bits 30..0 = the bits of min(p, 1 - p)
// which is GF(2)-linear in x like the Sobol map itself, so the XOR of the two table words IS the folded integer of the
// path's uniform: the inverse normal starts from it with no sign-mask / conditional-complement instructions.
//
// Macros expected from the generated prelude (as for sde_sim_kernel.cuh) plus
//   SDE_S                          number of steps S = T - 1 (compile time: the plan owns the time grid)
//   SDE_ALL_WIENER                 1 when every stochastic factor is a Wiener increment
// Requires SDE_RNG in {2, 3} (Sobol with XOR digital shift / raw), SDE_OUT == 0, SDE_UNR == 4.
#pragma once
#include "sde_sim_common.cuh"

#if !(SDE_RNG == 2 || SDE_RNG == 3) || SDE_OUT != 0 || SDE_UNR != 4
#error "sde_sim_resident.cuh: Sobol (xor / raw) full-path NTP output with 4-step groups only"
#endif
#ifndef SDE_KK
#define SDE_KK (SDE_K > 0 ? SDE_K : 1)
#endif
#ifndef SDE_NSLOT
#define SDE_NSLOT 0
#endif
#ifndef SDE_ST256
#define SDE_ST256 1
#endif

#ifndef SDE_RES_GRP
#define SDE_RES_GRP 4                          /* steps per unrolled group = one 32-byte sector per lane and process */
#endif
#if SDE_RES_GRP != 4
#error "sde_sim_resident.cuh: the row heads / tails assume groups of 4 steps"
#endif
#ifndef SDE_RES_PIPE
#define SDE_RES_PIPE 1                         /* draws of group g+1 overlap the state updates of group g */
#endif
#ifndef SDE_FULL_SECTORS
#ifdef SDE_DEBUG_NOSCALAR
#define SDE_FULL_SECTORS 0
#else
#define SDE_FULL_SECTORS (SDE_P == 1 && SDE_S >= 8 && SDE_ST256)   /* row heads / tails leave as whole sectors */
#endif
#endif
#ifndef SDE_ST_HINT
#define SDE_ST_HINT ""                         /* cache operator of the group stores (tuning: ".cs", ".wt", ".cg") */
#endif
#ifndef SDE_ALL_WIENER
#define SDE_ALL_WIENER 0
#endif
#ifndef SDE_RES_FOLD
#define SDE_RES_FOLD (SDE_RNG == 2 && SDE_ICDF == 1 && !SDE_NEEDS_U0 && SDE_ALL_WIENER)
#endif
#define SDE_NW (SDE_BLOCK / 32)
#define SDE_SK (SDE_S * SDE_K)
#define SDE_STEP_LD (4 + SDE_NSLOT)
#define SDE_NQ ((SDE_SK + 3 + 3) / 4)          /* quads of 4 dimensions, with room for the per-CTA offset 0..3 */
#define SDE_BW_LD (SDE_NQ * 4)
#define SDE_QPG ((SDE_RES_GRP * SDE_K) / 4)    /* quads per step group */
#define SDE_NIB_LD ((SDE_SK + 31) & ~31)      /* leading dimension of the transposed nibble table */
// shared-memory carve-up (bytes); mirrored by the host in lower.cpp
#ifndef SDE_ICDF_WIDE
#define SDE_ICDF_WIDE 0                         /* 1: 1024-entry log table (128 KB), set by the lowering when it fits */
#endif
#define SDE_SMEM_ICDF_BYTES ((SDE_ICDF == 1) ? ((SDE_ICDF_WIDE ? SDE_ICDF_WIDE_DOUBLES : SDE_ICDF_TABLE_DOUBLES) * 8) : 0)
// SDE_RES_LANE_GLOBAL: the lane table stays in global memory (time grids whose table does not fit in shared memory: C3 has
// 2000 dimensions = 256 KB).  The host prepares it per seed in the layout the CTA prologue would build — digital shift
// folded in, sign-folded where the plan folds, [off][quad][lane][4] for every quad offset 0..3 — and prm.sobol_lane points at
// it; a group's words are then one 128-bit read-only global load per quad (L1 / L2 resident: every warp of the GPU reads the
// same 256 KB) issued a whole group ahead by the software pipeline, instead of a shared-memory load.
#ifndef SDE_RES_LANE_GLOBAL
#define SDE_RES_LANE_GLOBAL 0
#endif
#define SDE_SMEM_LANE_BYTES (SDE_RES_LANE_GLOBAL ? 0 : (SDE_NQ * 32 * 16))
#define SDE_SMEM_STEP_BYTES (SDE_S * SDE_STEP_LD * 8)
#define SDE_SMEM_BW_BYTES (SDE_NW * SDE_BW_LD * 4)
#define SDE_SMEM_BYTES (SDE_SMEM_ICDF_BYTES + SDE_SMEM_STEP_BYTES + SDE_SMEM_LANE_BYTES + SDE_SMEM_BW_BYTES)

// sign-folded form of a 32-bit word (see SDE_RES_FOLD above); the identity when the plan does not fold
__device__ __forceinline__ sde_u32 sde_res_fold(sde_u32 x) {
#if SDE_RES_FOLD
    return x ^ ((sde_u32)((int)x >> 31) & 0x7fffffffu);
#else
    return x;
#endif
}

// The item loop is instantiated once per step shift gamma = 0..3 (a CTA only takes items of one class, so the shift is
// CTA-uniform): the <= 3 head steps, the <= 3 tail steps and the <= 2 recomputed steps of the next row's head then have
// compile-time trip counts, unroll, and their uniform -> normal chains overlap like those of a group instead of running
// one latency-exposed step at a time.  SDE_RES_GAMMA_SPECIALISE = 0 keeps one copy with a run-time shift.
#ifndef SDE_RES_GAMMA_SPECIALISE
#define SDE_RES_GAMMA_SPECIALISE 1
#endif
template <int V> struct sde_int_tag { static constexpr int value = V; };

extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_MIN_BLOCKS) sde_sim_kernel(const SdeParams prm) {
    extern __shared__ double4 sde_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(sde_smem_raw);
    double* s_icdf = reinterpret_cast<double*>(smem);
    double* s_step = reinterpret_cast<double*>(smem + SDE_SMEM_ICDF_BYTES);
    sde_u32* s_lane = reinterpret_cast<sde_u32*>(smem + SDE_SMEM_ICDF_BYTES + SDE_SMEM_STEP_BYTES);
    sde_u32* s_bw = s_lane + SDE_SMEM_LANE_BYTES / 4;
    (void)s_icdf;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int S = SDE_S, T = SDE_S + 1;

    const sde_u64 first_n = prm.scen_offset + 5ull;           // Sobol::new(..).skip(5)  (sobol.rs:17)
    const sde_u64 n_base = first_n & ~127ull;
    const sde_u64 n_blocks = (first_n + prm.n_paths - n_base + 127ull) >> 7;   // blocks of 128 paths = items per class
    // this CTA's item class and the step shift all of its rows share: row s starts at element s T P, and the group that
    // starts at step gamma writes elements from (gamma + 1) P on — P (s T + gamma + 1) = 0 (mod 4) puts that on a
    // 32-byte boundary (the output base is 32-byte aligned); s = n - first_n = cls - first_n (mod 4) for every lane
    const int cls = (int)(blockIdx.x & 3u);
    const int gamma_rt = (4 - (int)((((long long)(n_base + (sde_u64)cls) - (long long)first_n) * T + 1) & 3)) & 3;
    const int off = (4 - ((gamma_rt * SDE_K) & 3)) & 3;       // (gamma K + off) = 0 (mod 4): groups start on a quad

    // ---- CTA prologue: the tables every path of every item reads
#if SDE_ICDF == 1
#if SDE_ICDF_WIDE
    sde_icdf_wide_table_build(s_icdf, tid, SDE_BLOCK, SDE_ICDF_Y_OFFSET_K32);
#else
    sde_icdf_table_load(s_icdf, tid, SDE_BLOCK, SDE_RNG == 2 ? SDE_ICDF_Y_OFFSET_K32 : 0.0);
#endif
#endif
#if !SDE_RES_LANE_GLOBAL
    for (int e = tid; e < SDE_SK * 32; e += SDE_BLOCK) {
        const int d = e >> 5, l = e & 31;
        sde_u32 v = __ldg(prm.sobol_lane + e);
#if SDE_RNG == 2
        v ^= __ldg(prm.xor_masks + d);                        // u = ((x ^ mask) + 1/2) 2^-32: one mask per dimension
#endif
        const int dd = d + off;
        s_lane[(((dd >> 2) << 5) + l) * 4 + (dd & 3)] = sde_res_fold(v);
    }
#endif
    for (int e = tid; e < S; e += SDE_BLOCK) {
        const double t_cur = __ldg(prm.times + e), t_next = __ldg(prm.times + e + 1);
        const double dt = __ldg(prm.dts + e), sq = __ldg(prm.sqrt_dts + e);
        double* rec = s_step + e * SDE_STEP_LD;
        rec[0] = t_cur; rec[1] = t_next; rec[2] = dt; rec[3] = sq;
#if SDE_NSLOT > 0
        double slots[SDE_NSLOT];
        sde_model_step_consts(t_cur, t_next, dt, sq, slots);
#pragma unroll
        for (int q = 0; q < SDE_NSLOT; ++q) rec[4 + q] = slots[q];
#endif
    }
    __syncthreads();

    sde_u32* const my_bw = s_bw + warp * SDE_BW_LD;
    // shared-window byte addresses of the words of "step 0" in this lane's column of the lane table / in this warp's part
    // (explicit ld.shared.v4.u32 in draw_group): the quad of step tg is 128 K tg / 4 K tg bytes further on
#if SDE_RES_LANE_GLOBAL
    // this class's copy of the prepared table (quad offset `off` already applied by the host), this lane's column
    const sde_u32* const g_col = prm.sobol_lane + (size_t)off * (SDE_NQ * 128) + lane * 4;
    const sde_u32* const g_lane = g_col + off * 32;           // words of "step 0": a group's first quad is 128 K tg bytes further on
#else
    const sde_u32 lane_t0 = (sde_u32)__cvta_generic_to_shared(s_lane + lane * 4) + (sde_u32)off * 128u;
#endif
    const sde_u32 bw_t0 = (sde_u32)__cvta_generic_to_shared(my_bw) + (sde_u32)off * 4u;
#if SDE_ICDF == 1
    const sde_u32 tab_lane = (sde_u32)__cvta_generic_to_shared(s_icdf + 2 * (lane & (SDE_ICDF_TABLE_REPL - 1)));
#endif
    // word of dimension d: lane part / warp part (scalar accesses: row heads and tails)
#if SDE_RES_LANE_GLOBAL
    auto lane_word = [&](const int d) __attribute__((always_inline)) { const int dd = d + off; return __ldg(g_col + ((dd >> 2) << 7) + (dd & 3)); };
#else
    auto lane_word = [&](const int d) __attribute__((always_inline)) { const int dd = d + off; return s_lane[(((dd >> 2) << 5) + lane) * 4 + (dd & 3)]; };
#endif
    auto bw_word = [&](const int d) __attribute__((always_inline)) { return my_bw[d + off]; };

    double x0[SDE_P];
#pragma unroll
    for (int p = 0; p < SDE_P; ++p) x0[p] = __ldg(prm.x0 + p);
    const double t_first = __ldg(prm.times);
    const sde_u64 m_stride = (sde_u64)(gridDim.x >> 2) * SDE_NW;

    auto run_items = [&](auto gamma_tag) __attribute__((always_inline)) {
#if SDE_RES_GAMMA_SPECIALISE
    constexpr int gamma = decltype(gamma_tag)::value;         // this CTA's step shift, compile time in this instance
#else
    const int gamma = gamma_rt;
    (void)gamma_tag;
#endif
#pragma unroll 1
    for (sde_u64 m = (sde_u64)(blockIdx.x >> 2) * SDE_NW + warp; m < n_blocks; m += m_stride) {
        // item = 32 paths n = n0 + 4 lane of block m
        const sde_u64 n0 = n_base + (m << 7) + (sde_u64)cls;
        const sde_u64 n = n0 + (sde_u64)(4 * lane);
        const bool valid = (n >= first_n) && (n - first_n < prm.n_paths);
        if (!__any_sync(0xffffffffu, valid)) continue;
        const long long s_local = (long long)(n - first_n);   // "negative" for the <= 5 leading pad lanes

        // ---- Sobol part of this item, x_d(n0): XOR over the nibbles of gray(n0) of 16-entry tables (independent loads)
        __syncwarp();                                         // the previous item's reads of my_bw are complete
#ifndef SDE_DEBUG_NOCOMPUTE
        {
            // prm.sobol_nib arrives transposed for this kernel, [8][16][SDE_NIB_LD] (dimension fastest): the 32 lanes of
            // a load read 32 consecutive dimensions of one (nibble position, nibble value) row — one 128-byte line
            const sde_u32 g = (sde_u32)n0 ^ ((sde_u32)n0 >> 1);
            sde_u32 noff[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) noff[q] = (sde_u32)(q * 16 + ((g >> (4 * q)) & 15u)) * SDE_NIB_LD + lane;
            // all loads of up to 8 chunks (256 dimensions) in flight at once: one L2 round trip per item instead of one
            // per pair of chunks (folding ahead inside the step loop was measured slower: it costs the loop 30 registers)
#pragma unroll 8
            for (int d = 0; d < SDE_NIB_LD; d += 32) {
                sde_u32 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = __ldg(prm.sobol_nib + noff[q] + d);
                if (d + lane < SDE_SK) my_bw[d + lane + off] = sde_res_fold(((v[0] ^ v[1]) ^ (v[2] ^ v[3])) ^ ((v[4] ^ v[5]) ^ (v[6] ^ v[7])));
            }
        }
#endif
        __syncwarp();

        // ---- ScenarioFiltration::new — row 0 from initial_values, cache loaded from row 0 (filtration.rs:42-51)
        double row[SDE_P], cache[SDE_P];
        double ct = t_first;
#pragma unroll
        for (int p = 0; p < SDE_P; ++p) { row[p] = x0[p]; cache[p] = x0[p]; }
        // this path's rows [T][P].  Pad lanes (they only exist in the first and last block of a launch) are pointed at a
        // scratch row with the same phase modulo a sector, so the group stores of the step loop need no predicate
        double* const my_row = valid ? prm.out + (size_t)s_local * T * SDE_P
                                     : prm.partials + (size_t)(((long long)s_local * T * SDE_P) & 3);

        // draw_word: the (digitally shifted, possibly sign-folded) 32-bit word of factor k -> normal / Poisson draw
        auto draw_word = [&](const sde_u32 w, const int k, double& z, sde_u0_t& u0) __attribute__((always_inline)) {
#if SDE_RES_FOLD
            // w is the folded integer of p = (x + 1/2) 2^-32: every factor is Wiener, nothing else reads the uniform
#if SDE_ICDF_WIDE
            z = sde_icdf_normal_fast_y32w(w, tab_lane);
#else
            z = sde_icdf_normal_fast_y32s(w, tab_lane);
#endif
            (void)k; (void)u0;
#elif SDE_RNG == 2
            const sde_u32 x = w;
#if SDE_U0_BITS
            if (k == 0 && SDE_NEEDS_U0) u0 = x;               // the step reads u0 > 1/2 = the top bit of x (sde_u0_t)
#else
            if (k == 0 && SDE_NEEDS_U0) u0 = fma((double)x, 2.3283064365386963e-10, 1.1641532182693481e-10);
#endif
            // digital shift: u = (x + 1/2) * 2^-32 in (0, 1)
            if (sde_factor_is_wiener(k)) {
#if SDE_ICDF == 1 && SDE_ICDF_WIDE
                z = sde_icdf_normal_fast_k32w(x, tab_lane);
#elif SDE_ICDF == 1
                z = sde_icdf_normal_fast_k32s(x, tab_lane);
#elif SDE_ICDF == 2
                z = sde_icdf_normal_single_k32(x);
#else
                z = sde_icdf_normal_reference(fma((double)x, 2.3283064365386963e-10, 1.1641532182693481e-10));
#endif
            } else {
                z = fma((double)x, 2.3283064365386963e-10, 1.1641532182693481e-10);
            }
#else
            {   // raw points: u = x / 2^32 (can be 0: the reference's ln(0) path gives NaN)
                const double u = (double)w * 2.3283064365386963e-10;
                if (k == 0) u0 = u;
                if (sde_factor_is_wiener(k)) {
#if SDE_ICDF == 1
                    z = sde_icdf_normal_fast(u, s_icdf, lane);
#elif SDE_ICDF == 2
                    z = sde_icdf_normal_single(u);
#else
                    z = sde_icdf_normal_reference(u);
#endif
                } else {
                    z = u;
                }
            }
#endif
        };
        // draw_x: draws of step t on its own (row heads / tails).  `flip[k]` is XORed into the word of factor k: zero for
        // the lane's own path, (the folded form of) V_d[ctz(n + 1)] for the path that follows it
        // (x_d(n+1) = x_d(n) ^ V_d[ctz(n+1)], and folding is linear)
        auto draw_x = [&](const int t, const sde_u32 (&flip)[SDE_KK], double (&zu)[SDE_KK], sde_u0_t& u0) __attribute__((always_inline)) {
            u0 = (sde_u0_t)0;
            zu[0] = 0.0;
#ifdef SDE_DEBUG_NOCOMPUTE
            zu[0] = (double)t * 1e-4;                         // profiling aid: stores only (no Sobol reads, no inverse normal)
            return;
#endif
#pragma unroll
            for (int k = 0; k < SDE_K; ++k) {
                const int d = t * SDE_K + k;
                draw_word(bw_word(d) ^ lane_word(d) ^ flip[k], k, zu[k], u0);
            }
        };
        auto draw = [&](const int t, double (&zu)[SDE_KK], sde_u0_t& u0) __attribute__((always_inline)) {
            sde_u32 none[SDE_KK];
#pragma unroll
            for (int k = 0; k < SDE_KK; ++k) none[k] = 0u;
            draw_x(t, none, zu, u0);
        };
        // draw_group: the draws of the SDE_RES_GRP steps from tg on (tg = gamma mod 4): their words sit in SDE_QPG quads of
        // either table — one 128-bit load per quad (lane table: conflict free; warp part: broadcast).  Quad q of the lane
        // table is 512 bytes, of the warp part 16 bytes, and q = (tg K + off) / 4: both addresses are one multiply-add on
        // the step counter (written as opaque PTX so that they are not turned into extra loop-carried pointers)
        auto draw_group = [&](const int tg, double (&zu)[SDE_RES_GRP][SDE_KK], sde_u0_t (&u0)[SDE_RES_GRP]) __attribute__((always_inline)) {
#ifdef SDE_DEBUG_NOCOMPUTE
#pragma unroll
            for (int j = 0; j < SDE_RES_GRP; ++j) { u0[j] = (sde_u0_t)0; zu[j][0] = (double)(tg + j) * 1e-4; }
            return;
#endif
            sde_u32 ba;
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ba) : "r"(tg), "n"(4 * SDE_K), "r"(bw_t0));
            sde_u32 lw[4 * SDE_QPG], bw[4 * SDE_QPG];
#if SDE_RES_LANE_GLOBAL
            const sde_u32* gl;                                // the group's first quad in this lane's column (bytes: 128 K tg further on)
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(gl) : "r"(tg), "n"(128 * SDE_K), "l"(g_lane));
#else
            sde_u32 la;
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(la) : "r"(tg), "n"(128 * SDE_K), "r"(lane_t0));
#endif
#pragma unroll
            for (int i = 0; i < SDE_QPG; ++i) {
#if SDE_RES_LANE_GLOBAL
                asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(lw[4 * i]), "=r"(lw[4 * i + 1]), "=r"(lw[4 * i + 2]), "=r"(lw[4 * i + 3]) : "l"(gl + i * 128) : "memory");
#else
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(lw[4 * i]), "=r"(lw[4 * i + 1]), "=r"(lw[4 * i + 2]), "=r"(lw[4 * i + 3]) : "r"(la + i * 512u) : "memory");
#endif
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bw[4 * i]), "=r"(bw[4 * i + 1]), "=r"(bw[4 * i + 2]), "=r"(bw[4 * i + 3]) : "r"(ba + i * 16u) : "memory");
            }
#pragma unroll
            for (int j = 0; j < SDE_RES_GRP; ++j) {
                u0[j] = (sde_u0_t)0;
                zu[j][0] = 0.0;
#pragma unroll
                for (int k = 0; k < SDE_K; ++k) draw_word(lw[j * SDE_K + k] ^ bw[j * SDE_K + k], k, zu[j][k], u0[j]);
            }
        };
        // one step on its own (the <= 3 steps before the first and after the last aligned group)
        auto single = [&](const int t) __attribute__((always_inline)) {
            double zu[SDE_KK];
            sde_u0_t u0;
            draw(t, zu, u0);
            sde_model_step(row, cache, ct, zu, u0, s_step + t * SDE_STEP_LD);
#ifndef SDE_DEBUG_NOSCALAR
            if (valid) {
#pragma unroll
                for (int p = 0; p < SDE_P; ++p) my_row[(size_t)(t + 1) * SDE_P + p] = row[p];
            }
#endif
        };

        (void)single;
        int t = 0;
        const int g_eff = gamma < S ? gamma : S;
#if SDE_FULL_SECTORS
        // Whole sectors only.  A row is T doubles and T is not a multiple of 4, so the 32-byte sector at a row boundary
        // holds the last r elements of row s and the first 4 - r of row s + 1.  Written as 8-byte pieces by two lanes
        // it costs a quarter of the kernel's store throughput (partial-sector requests; measured 460 vs 616 G
        // path-steps/s for the store stream alone).  Instead the lane of row s writes that sector once, complete: the
        // head of row s + 1 is x0 and at most 2 steps of path s + 1, which it recomputes from its own Sobol integers
        // (x_d(n + 1) = x_d(n) ^ V_d[ctz(n + 1)]).  A row whose head fills a sector (gamma = 3) writes it itself; only
        // the first row's head and the last row's tail of a launch go out as scalars.
        double hv[4];
        hv[0] = row[0];
#if SDE_RES_GAMMA_SPECIALISE
        {
            // the g_eff head steps: all draws first (independent chains), then the serial state updates
            double hz[3][SDE_KK];
            sde_u0_t hu[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j < g_eff) draw(j, hz[j], hu[j]);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (j < g_eff) {
                    sde_model_step(row, cache, ct, hz[j], hu[j], s_step + j * SDE_STEP_LD);
                    hv[j + 1] = row[0];
                }
            t = g_eff;
        }
#else
#pragma unroll 1
        for (; t < g_eff; ++t) {
            double zu[SDE_KK];
            sde_u0_t u0;
            draw(t, zu, u0);
            sde_model_step(row, cache, ct, zu, u0, s_step + t * SDE_STEP_LD);
            if (t == 0) hv[1] = row[0]; else if (t == 1) hv[2] = row[0]; else hv[3] = row[0];
        }
#endif
        if (gamma == 3 && g_eff == 3) {
            const int lv = valid ? 1 : 0;
            asm volatile("{ .reg .pred p; setp.ne.s32 p, %5, 0; @p st.global.v4.f64 [%0], {%1, %2, %3, %4}; }"
                         ::"l"(my_row), "d"(hv[0]), "d"(hv[1]), "d"(hv[2]), "d"(hv[3]), "r"(lv) : "memory");
        } else if (valid && s_local == 0) {
            for (int j = 0; j <= g_eff; ++j) my_row[j] = j == 0 ? hv[0] : (j == 1 ? hv[1] : (j == 2 ? hv[2] : hv[3]));
        }
#else
#ifndef SDE_DEBUG_NOSCALAR
        if (valid) {
#pragma unroll
            for (int p = 0; p < SDE_P; ++p) my_row[p] = row[p];
        }
#endif
#pragma unroll 1
        for (; t < g_eff; ++t) single(t);
#endif
        const int n_groups = (S - g_eff) / SDE_RES_GRP;
        const int t_last = g_eff + n_groups * SDE_RES_GRP;    // end of the grouped steps
        double* const row1 = my_row + SDE_P;                  // element (tg + 1) P of the row: 32-byte aligned for tg = gamma (mod 4)
#ifdef SDE_DEBUG_NOSTORE
        const int live = (valid && prm.reserved == 12345) ? 1 : 0;   // profiling aid: group stores predicated off
#else
        const int live = valid ? 1 : 0;
        (void)live;
#endif
        // advance_group: the sequential state updates of one group (rows t+1 .. t+GRP collected in output order) and
        // their full-sector stores (pad lanes write the scratch row)
        auto advance_group = [&](const int tg, const double (&zu)[SDE_RES_GRP][SDE_KK], const sde_u0_t (&u0)[SDE_RES_GRP]) __attribute__((always_inline)) {
            double vals[SDE_RES_GRP * SDE_P];
#pragma unroll
            for (int j = 0; j < SDE_RES_GRP; ++j) {
                sde_model_step(row, cache, ct, zu[j], u0[j], s_step + (tg + j) * SDE_STEP_LD);
#pragma unroll
                for (int p = 0; p < SDE_P; ++p) vals[j * SDE_P + p] = row[p];
            }
            double* dst;                                      // row1 + tg P (one wide multiply-add on the step counter)
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(dst) : "r"(tg), "n"(8 * SDE_P), "l"(row1));
#pragma unroll
            for (int q = 0; q < SDE_P * (SDE_RES_GRP / 4); ++q) {
#if SDE_ST256 && !defined(SDE_DEBUG_NOSTORE)
                asm volatile("st.global" SDE_ST_HINT ".v4.f64 [%0], {%1, %2, %3, %4};"
                             ::"l"(dst + 4 * q), "d"(vals[4 * q]), "d"(vals[4 * q + 1]), "d"(vals[4 * q + 2]), "d"(vals[4 * q + 3]) : "memory");
#elif SDE_ST256
                asm volatile("{ .reg .pred p; setp.ne.s32 p, %5, 0; @p st.global.v4.f64 [%0], {%1, %2, %3, %4}; }"
                             ::"l"(dst + 4 * q), "d"(vals[4 * q]), "d"(vals[4 * q + 1]), "d"(vals[4 * q + 2]), "d"(vals[4 * q + 3]), "r"(live) : "memory");
#else
                asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p st.global.v2.f64 [%0], {%1, %2}; }"
                             ::"l"(dst + 4 * q), "d"(vals[4 * q]), "d"(vals[4 * q + 1]), "r"(live) : "memory");
                asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p st.global.v2.f64 [%0], {%1, %2}; }"
                             ::"l"(dst + 4 * q + 2), "d"(vals[4 * q + 2]), "d"(vals[4 * q + 3]), "r"(live) : "memory");
#endif
            }
        };
#if SDE_RES_PIPE
        // software pipeline: the state-independent uniform -> normal chains of group g+1 are issued together with the
        // (serial) state updates of group g, so the dependent multiply chain and the stores hide behind them
        // (two register sets, ping-pong: no copies on the loop back edge)
        if (n_groups > 0) {
            double za[SDE_RES_GRP][SDE_KK], zb[SDE_RES_GRP][SDE_KK];
            sde_u0_t ua[SDE_RES_GRP], ub[SDE_RES_GRP];
            draw_group(t, za, ua);
#pragma unroll 1
            for (; t + 3 * SDE_RES_GRP <= t_last; t += 2 * SDE_RES_GRP) {   // the step counter is the only loop-carried integer
                draw_group(t + SDE_RES_GRP, zb, ub);
                advance_group(t, za, ua);
                draw_group(t + 2 * SDE_RES_GRP, za, ua);
                advance_group(t + SDE_RES_GRP, zb, ub);
            }
            if (t + 2 * SDE_RES_GRP <= t_last) {
                draw_group(t + SDE_RES_GRP, zb, ub);
                advance_group(t, za, ua);
                advance_group(t + SDE_RES_GRP, zb, ub);
                t += 2 * SDE_RES_GRP;
            } else {
                advance_group(t, za, ua);
                t += SDE_RES_GRP;
            }
        }
#else
#pragma unroll 1
        for (int gi = 0; gi < n_groups; ++gi, t += SDE_RES_GRP) {
            // phase 1: the state-independent uniform -> normal chains of the group (independent instruction streams)
            double zu[SDE_RES_GRP][SDE_KK];
            sde_u0_t u0[SDE_RES_GRP];
            draw_group(t, zu, u0);
            advance_group(t, zu, u0);
        }
#endif
#if SDE_FULL_SECTORS
        {
#if SDE_RES_GAMMA_SPECIALISE
            constexpr int g_c = gamma < S ? gamma : S;
            constexpr int t_tail = g_c + ((S - g_c) / SDE_RES_GRP) * SDE_RES_GRP;
            constexpr int r = S - t_tail;                     // 0..3 tail elements; they start on a sector boundary
            double tv[3];
            tv[0] = tv[1] = tv[2] = 0.0;
            {
                double tz[3][SDE_KK];
                sde_u0_t tu[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) if (j < r) draw(t_tail + j, tz[j], tu[j]);
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < r) {
                        sde_model_step(row, cache, ct, tz[j], tu[j], s_step + (t_tail + j) * SDE_STEP_LD);
                        tv[j] = row[0];
                    }
            }
            t = S;
#else
            const int r = S - t;                              // 0..3 tail elements; they start on a sector boundary
            const int t_tail = t;
            double tv[3];
            tv[0] = tv[1] = tv[2] = 0.0;
#pragma unroll 1
            for (; t < S; ++t) {
                double zu[SDE_KK];
                sde_u0_t u0;
                draw(t, zu, u0);
                sde_model_step(row, cache, ct, zu, u0, s_step + t * SDE_STEP_LD);
                if (t == t_tail) tv[0] = row[0]; else if (t == t_tail + 1) tv[1] = row[0]; else tv[2] = row[0];
            }
#endif
            if (r > 0) {
                const bool has_next = valid && (sde_u64)(s_local + 1) < prm.n_paths;
                if (__any_sync(0xffffffffu, has_next)) {
                    // head of the next row: x0 and its first 3 - r steps (path n + 1)
                    const sde_u32 b = (sde_u32)__ffsll((long long)(n + 1ull)) - 1u;         // ctz(n + 1)
                    double nh[3];
                    nh[0] = x0[0]; nh[1] = nh[2] = 0.0;
                    double row2[SDE_P], cache2[SDE_P];
                    double ct2 = t_first;
                    row2[0] = x0[0]; cache2[0] = x0[0];
#if SDE_RES_GAMMA_SPECIALISE
                    {
                        double nz[2][SDE_KK];
                        sde_u0_t nu[2];
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            if (j < 3 - r) {
                                sde_u32 flip[SDE_KK];
#pragma unroll
                                for (int k = 0; k < SDE_K; ++k)   // V_d[b] = nibble-table entry of the single-bit nibble value
                                    flip[k] = sde_res_fold(__ldg(prm.sobol_nib + (size_t)((b >> 2) * 16u + (1u << (b & 3u))) * SDE_NIB_LD + (j * SDE_K + k)));
                                draw_x(j, flip, nz[j], nu[j]);
                            }
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            if (j < 3 - r) {
                                sde_model_step(row2, cache2, ct2, nz[j], nu[j], s_step + j * SDE_STEP_LD);
                                nh[j + 1] = row2[0];
                            }
                    }
#else
#pragma unroll 1
                    for (int j = 0; j < 3 - r; ++j) {
                        sde_u32 flip[SDE_KK];
#pragma unroll
                        for (int k = 0; k < SDE_K; ++k)       // V_d[b] = nibble-table entry of the single-bit nibble value
                            flip[k] = sde_res_fold(__ldg(prm.sobol_nib + (size_t)((b >> 2) * 16u + (1u << (b & 3u))) * SDE_NIB_LD + (j * SDE_K + k)));
                        double zu[SDE_KK];
                        sde_u0_t u0;
                        draw_x(j, flip, zu, u0);
                        sde_model_step(row2, cache2, ct2, zu, u0, s_step + j * SDE_STEP_LD);
                        if (j == 0) nh[1] = row2[0]; else nh[2] = row2[0];
                    }
#endif
                    // sector = [tail (r), next head (4 - r)]
                    const double o1 = r >= 2 ? tv[1] : nh[0];
                    const double o2 = r == 3 ? tv[2] : (r == 2 ? nh[0] : nh[1]);
                    const double o3 = r == 3 ? nh[0] : (r == 2 ? nh[1] : nh[2]);
                    const int lv = has_next ? 1 : 0;
                    asm volatile("{ .reg .pred p; setp.ne.s32 p, %5, 0; @p st.global.v4.f64 [%0], {%1, %2, %3, %4}; }"
                                 ::"l"(my_row + (t_tail + 1)), "d"(tv[0]), "d"(o1), "d"(o2), "d"(o3), "r"(lv) : "memory");
                }
                if (valid && !has_next) {                     // last row of the launch: nothing follows it in this buffer
                    for (int j = 0; j < r; ++j) my_row[t_tail + 1 + j] = j == 0 ? tv[0] : (j == 1 ? tv[1] : tv[2]);
                }
            }
        }
#else
#pragma unroll 1
        for (; t < S; ++t) single(t);
#endif
    }
    };
#if SDE_RES_GAMMA_SPECIALISE
    switch (gamma_rt) {
        case 0: run_items(sde_int_tag<0>{}); break;
        case 1: run_items(sde_int_tag<1>{}); break;
        case 2: run_items(sde_int_tag<2>{}); break;
        default: run_items(sde_int_tag<3>{}); break;
    }
#else
    run_items(sde_int_tag<0>{});
#endif
}
