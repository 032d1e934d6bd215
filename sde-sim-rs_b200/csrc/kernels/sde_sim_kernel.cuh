// sde_sim_kernel.cuh — the fused path-simulation kernel (sm_100a).
//
// One launch replaces the whole of sim::simulate's parallel region (src/sim/mod.rs:41-88):
// per-scenario RNG construction (src/rng/pseudo.rs:14-20, src/rng/sobol.rs:35-53), the time
// loop (src/sim/mod.rs:68-84) calling euler_iteration / runge_kutta_iteration, the
// Wiener/Poisson incrementors (src/proc/increment.rs:89-148) and the dense Filtration rows
// (src/filtration.rs:55-64,112).  One thread owns one path; its state row and the
// expression cache live in registers; the model's drift/diffusion expressions arrive as
// generated device code (`sde_model_step`, emitted by csrc/host/lower.cpp from the parsed
// equations) in the translation unit that includes this header.
//
// Structure of a CTA (SDE_BLOCK paths = SDE_NW warps):
//   time is cut into tiles of SDE_TT steps.  Everything a tile's steps read — {t, t+dt, dt, sqrt(dt)} and
//   model constants hoisted out of the path loop, the Sobol CTA/warp part x_d(n_cta + 32w) (with the
//   digital-shift mask folded in) and the lane part x_d(lane) of the tile's dimensions — is staged in
//   shared memory, double buffered: the loads for tile k+1 are issued before the last step group of tile k
//   and land in registers while it computes; one __syncthreads per tile.
//   Inside a tile, groups of SDE_UNR steps run in two phases: first the state-independent uniform ->
//   normal chains of the whole group (independent instruction streams for the scheduler), then the
//   sequential state updates.
//   Full paths, reference row order [N][T][P], leave in one of two ways:
//     SDE_DIRECT = 1  (Sobol / injected draws, groups of 4 steps): the lanes of a warp own paths that are 4 apart
//       (path = cta + 128 (w>>2) + 4 lane + (w&3)), so all 32 rows start at the same offset modulo a 32-byte
//       sector.  Each warp shifts its step groups by gamma in 0..3 steps so that a group's 4P values start
//       on a sector boundary, and every lane writes them straight from registers as P aligned 256-bit stores
//       (st.global.v4.f64): full sectors only, no shared-memory transpose, no flush.  The staged tables cover
//       SDE_TT + 3 steps per tile to absorb the shift.
//     SDE_DIRECT = 0  (ChaCha-driven modes, wide models): rows are transposed through a per-warp
//       shared-memory tile and each path's [t0+1, t0+TT] x P segment leaves as contiguous 8-byte-coalesced stores.
//     SDE_TMA = 1     (f64, P even: every row segment starts and ends on a 16-byte boundary): a lane stages its path's
//       [t0+1, t0+TT] x P segment in shared memory with 128-bit stores (conflict free: leading dimension = 2 mod 4
//       doubles) and hands it to the copy engine once per tile — cp.async.bulk shared -> global, one instruction per lane
//       and tile.  The store bytes never pass the load/store data pipe as scattered sectors: a 256-bit sector store of 32
//       rows costs ~26 data-pipe wavefronts (measured, profiles/r2_ncu_full_c3_summary.txt), the same 1 KB staged with
//       128-bit shared stores 8, and the pipe is shared with the inverse normal's table reads.  Opt-in (ntp_direct = 4):
//       the bulk copy takes uniform-register addresses, so the 32 lanes' copies serialise (~10 instructions each) and the
//       net effect on C3 is -6 %; a 2-D tensor-map store per warp would avoid that (not built).
//
// Macros expected from the generated prelude:
//   SDE_P, SDE_K, SDE_KK           processes / stochastic factors / max(K, 1)
//   SDE_RNG                        0 pseudo(ChaCha8)  1 sobol+per-path CP shift (reference)
//                                  2 sobol+XOR digital shift  3 sobol raw  4 injected draws
//                                  5 Philox4x32-10 (generator = "philox": counter-based MC tier, not in the reference)
//   SDE_OUT                        0 paths [N][T][P]  1 paths [T][P][N]  2 terminal [N][P]  3 moments
//   SDE_ICDF                       0 reference  1 fast  2 single (FP32 evaluation)
//   SDE_NEEDS_U0                   1 when the scheme consumes u[t][0] directly (Runge–Kutta sk)
//   SDE_BLOCK, SDE_MIN_BLOCKS      launch bounds
//   SDE_TT, SDE_UNR, SDE_CH        time tile, steps per unrolled group, ChaCha chunk (8 / gcd(8, K))
//   SDE_NSLOT                      per-step model constants hoisted into the tile prologue
//   SDE_DIRECT                     1: register -> HBM sector stores for SDE_OUT == 0 (needs SDE_UNR == 4, no ChaCha)
//   sde_factor_is_wiener(k)        constexpr predicate
//   sde_model_step_consts(t_cur, t_next, dt, sqrt_dt, slots)     fills SDE_NSLOT doubles
//   sde_model_step(row, cache, ct, zu, u0, ss)                   ss = {t_cur, t_next, dt, sqrt_dt, slots...}
#pragma once
#include "sde_sim_common.cuh"

#ifndef SDE_W64_RARE_SHIFT
#define SDE_W64_RARE_SHIFT 0   /* test hook: > 0 sends draws with min(p, 1-p) < 2^(shift-32) through the rare path as well */
#endif
#define SDE_NW (SDE_BLOCK / 32)
#define SDE_USES_CHACHA (SDE_RNG == 0 || SDE_RNG == 1)
#define SDE_USES_SOBOL (SDE_RNG == 1 || SDE_RNG == 2 || SDE_RNG == 3)
#define SDE_USES_PHILOX (SDE_RNG == 5)
#ifndef SDE_KK
#define SDE_KK (SDE_K > 0 ? SDE_K : 1)
#endif
#ifndef SDE_NSLOT
#define SDE_NSLOT 0
#endif
#ifndef SDE_UNR
#define SDE_UNR SDE_CH
#endif
#ifndef SDE_DIRECT
#define SDE_DIRECT 0
#endif
#ifndef SDE_TMA
#define SDE_TMA 0
#endif
// SDE_TMA == 2: one 2-D tensor-map store per warp and box of SDE_TBOX_STEPS steps (P = 2 or 4: a box row — one path's
// SDE_TBOX_STEPS x P values — is 128 bytes).  The 32 lanes of a warp own 32 consecutive paths; each writes its row of the box
// into shared memory with 128-bit stores in the 128-byte swizzle pattern of the tensor map (16-byte chunk c of row r at
// chunk c ^ (r & 7): conflict free), lane 0 issues cp.async.bulk.tensor.2d shared -> global.  The copy engine writes every
// row as one 128-byte piece and clips rows / columns outside [N] x [T P] (pad lanes, ragged ends), and none of the store
// bytes crosses the load/store data pipe as scattered sectors.  Two boxes per warp, ping-pong.
#define SDE_TBOX_STEPS (16 / SDE_P)
struct alignas(64) SdeTensorMap { unsigned long long v[16]; };
#ifndef SDE_ST256
#define SDE_ST256 1   /* 256-bit st.global.v4.f64 (PTX ISA 8.8 / CUDA 12.9 ptxas) */
#endif
// steps staged per tile buffer: the direct path lets a warp run up to 3 steps past the tile boundary
#define SDE_TS (SDE_TT + (SDE_DIRECT ? 3 : 0))
// leading dimension of a warp's staging row: odd => conflict-free column writes; bulk-copy mode: = 2 (mod 4) doubles =>
// 16-byte aligned rows whose 128-bit writes are conflict free (8 lanes x 16 B per wavefront land in 8 distinct bank groups)
#if SDE_TMA
#define SDE_TILE_LD ((((SDE_TT * SDE_P) + 3) & ~3) + 2)
#else
#define SDE_TILE_LD ((SDE_TT * SDE_P) | 1)
#endif
#define SDE_STEP_LD (4 + SDE_NSLOT)

// shared-memory carve-up (bytes); mirrored by the host in lower.cpp
//   icdf tables | output staging tile | 2 x { step records | Sobol CTA/warp part | Sobol lane part } | moment scratch
#if SDE_TMA == 2
#define SDE_SMEM_ICDF_BYTES (((SDE_ICDF == 1 && SDE_RNG != 4) ? ((SDE_ICDF_TABLE_DOUBLES * 8 + 1023) & ~1023) : 0))   /* boxes start 1 KB aligned */
#define SDE_SMEM_TILE_BYTES (SDE_NW * 2 * 4096)
#else
#define SDE_SMEM_ICDF_BYTES ((SDE_ICDF == 1 && SDE_RNG != 4) ? (SDE_ICDF_TABLE_DOUBLES * 8) : 0)
#define SDE_SMEM_TILE_BYTES ((SDE_OUT == 0 && !SDE_DIRECT) ? (((SDE_NW * 32 * SDE_TILE_LD * (int)sizeof(sde_real) + 7) & ~7)) : 0)
#endif
#define SDE_SMEM_STEP_BYTES (SDE_TS * SDE_STEP_LD * 8)
#define SDE_SMEM_BW_BYTES (SDE_USES_SOBOL ? (SDE_TS * SDE_KK * SDE_NW * 4) : 0)
#define SDE_SMEM_LANE_BYTES (SDE_USES_SOBOL ? (SDE_TS * SDE_KK * 32 * 4) : 0)
#define SDE_SMEM_STAGE_BYTES (SDE_SMEM_STEP_BYTES + SDE_SMEM_BW_BYTES + SDE_SMEM_LANE_BYTES)
#define SDE_SMEM_MOM_BYTES ((SDE_OUT == 3) ? (SDE_NW * 3 * 8) : 0)
// SDE_NSTAGE stage buffers.  2: tile k+1 is committed at the end of tile k, one __syncthreads per tile.  4: tile k+2 is
// committed at the end of tile k and announced on an mbarrier that a warp only waits for at the start of tile k+2 — a whole
// tile of slack, so warps whose store phases drift no longer wait for each other every tile (the barrier stall was 7 % of
// C3's issue slots with the row stores on, 2 % without; profiles/r2_ncu_full_c3_summary.txt).  Buffer (k+2) % 4 was last
// read in tile k-2, which every warp has left: a warp reaches the end of tile k only after the phase of tile k completed,
// i.e. after every warp arrived at the end of tile k-2.  Two mbarriers (even / odd tiles) keep a waiting warp at most one
// phase behind its barrier.
#ifndef SDE_NSTAGE
#define SDE_NSTAGE 2
#endif
#define SDE_SMEM_MBAR_BYTES (SDE_NSTAGE == 4 ? 16 : 0)
#define SDE_SMEM_BYTES (SDE_SMEM_ICDF_BYTES + SDE_SMEM_TILE_BYTES + SDE_NSTAGE * SDE_SMEM_STAGE_BYTES + SDE_SMEM_MOM_BYTES + SDE_SMEM_MBAR_BYTES)

// prefetch register counts (compile-time): entries of each staged table owned by one thread
#define SDE_PF_BW ((SDE_TS * SDE_KK * SDE_NW + SDE_BLOCK - 1) / SDE_BLOCK)
#define SDE_PF_LANE ((SDE_TS * SDE_KK * 32 + SDE_BLOCK - 1) / SDE_BLOCK)
#define SDE_PF_STEP ((SDE_TS + SDE_BLOCK - 1) / SDE_BLOCK)

__device__ __forceinline__ sde_real sde_uniform_to_draw(double u, bool wiener, const double* s_icdf, int lane) {
    if (!wiener) return u;                               // Poisson factors consume the uniform itself
#if SDE_ICDF == 1
    return sde_icdf_normal_fast(u, s_icdf, lane);
#elif SDE_ICDF == 2
    return sde_icdf_normal_single(u);
#else
    return sde_icdf_normal_reference(u);
#endif
}

// 32-bit-word uniforms (Sobol digital shift / raw-free modes, Philox) with the fast inverse normal: the shared-window entry
// points of sde_device_icdf.cuh (one LDS.128 per draw, exponent term by conversion, FP32-unit seeds).  For the Sobol
// digital shift the staged CTA/warp and lane words are SIGN-FOLDED, y = x ^ ((x >>s 31) & 0x7fffffff): GF(2)-linear like
// the Sobol map itself, so their XOR is the folded integer of the path's uniform (bit 31: p >= 1/2, bits 30..0: those of
// min(p, 1-p)) and the draw starts from it without sign-mask / conditional-complement instructions.
#define SDE_TILED_W32 ((SDE_RNG == 2 || SDE_RNG == 5) && SDE_ICDF == 1)
#define SDE_TILED_FOLD (SDE_RNG == 2 && SDE_ICDF == 1)
__device__ __forceinline__ sde_u32 sde_tile_fold(sde_u32 x) {
#if SDE_TILED_FOLD
    return x ^ ((sde_u32)((int)x >> 31) & 0x7fffffffu);
#else
    return x;
#endif
}

// Registers that carry the next tile's staged data from the moment its loads are issued to the moment
// they are written to shared memory (static indexing only).
struct SdeTilePrefetch {
    double st[SDE_PF_STEP][4];
#if SDE_USES_SOBOL
    sde_u32 bw[SDE_PF_BW][9];   // the 8 nibble-table words (+ digital-shift mask) of each entry; XORed at commit time so
                                // that nothing waits on these loads while the last step group of the tile runs
    sde_u32 ln[SDE_PF_LANE];
#endif
};

#if SDE_TMA == 2
extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_MIN_BLOCKS) sde_sim_kernel(const SdeParams prm, const __grid_constant__ SdeTensorMap tmap) {
    extern __shared__ __align__(1024) double4 sde_smem_raw[];
#else
extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_MIN_BLOCKS) sde_sim_kernel(const SdeParams prm) {
    extern __shared__ double4 sde_smem_raw[];
#endif
    unsigned char* smem = reinterpret_cast<unsigned char*>(sde_smem_raw);
    double* s_icdf = reinterpret_cast<double*>(smem);
    sde_real* s_tile = reinterpret_cast<sde_real*>(smem + SDE_SMEM_ICDF_BYTES);
    unsigned char* s_stage = smem + SDE_SMEM_ICDF_BYTES + SDE_SMEM_TILE_BYTES;     // two buffers of SDE_SMEM_STAGE_BYTES
    double* s_mom = reinterpret_cast<double*>(smem + SDE_SMEM_BYTES - SDE_SMEM_MBAR_BYTES - SDE_SMEM_MOM_BYTES);
#if SDE_NSTAGE == 4
    const unsigned s_mbar = (unsigned)__cvta_generic_to_shared(smem + SDE_SMEM_BYTES - SDE_SMEM_MBAR_BYTES);   // two 8-byte mbarriers
#endif
    (void)s_icdf; (void)s_tile; (void)s_mom;
    sde_real* const out_r = reinterpret_cast<sde_real*>(prm.out);   // rows are sde_real (f32 plans: float)
    (void)out_r;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int S = prm.n_steps;
    const int T = S + 1;

    // thread -> point index n -> scenario (n - 5): the reference skips 5 points (sobol.rs:17)
    const sde_u64 n_cta = prm.n_base + (sde_u64)blockIdx.x * SDE_BLOCK;
#if SDE_DIRECT
    // lanes own paths 4 apart so that every row of a warp has the same phase modulo a 32-byte sector
    const sde_u64 n = n_cta + (sde_u64)(128 * (warp >> 2) + 4 * lane + (warp & 3));
#else
    const sde_u64 n = n_cta + (sde_u64)tid;
#endif
    const sde_u64 first_n = prm.scen_offset + 5ull;
    const bool valid = (n >= first_n) && (n - first_n < prm.n_paths);
    const long long s_local = (long long)(n - first_n);           // may be "negative" for the <=5 leading pad threads
    const sde_u64 s_global = n - 5ull;
    (void)s_global;

#if SDE_ICDF == 1 && SDE_RNG != 4
    sde_icdf_table_load(s_icdf, tid, SDE_BLOCK, SDE_TILED_W32 ? SDE_ICDF_Y_OFFSET_K32 : 0.0);
#endif
#if SDE_TILED_W32
    const sde_u32 tab_lane = (sde_u32)__cvta_generic_to_shared(s_icdf + 2 * (lane & (SDE_ICDF_TABLE_REPL - 1)));
#endif

    // ---- staging of one tile's read-only data (see header).  issue(): global loads into registers;
    //      commit(): registers -> shared memory buffer `buf`.
    auto issue = [&](const int t0, SdeTilePrefetch& pf) __attribute__((always_inline)) {
        const int nt = min(SDE_TS, S - t0);
#pragma unroll
        for (int i = 0; i < SDE_PF_STEP; ++i) {
            const int e = tid + i * SDE_BLOCK;
            if (e < nt) {
                pf.st[i][0] = __ldg(prm.times + t0 + e);
                pf.st[i][1] = __ldg(prm.times + t0 + e + 1);
                pf.st[i][2] = __ldg(prm.dts + t0 + e);
                pf.st[i][3] = __ldg(prm.sqrt_dts + t0 + e);
            }
        }
#if SDE_USES_SOBOL
        const int nd = nt * SDE_K;
        const size_t d0 = (size_t)t0 * SDE_K;
#pragma unroll
        for (int i = 0; i < SDE_PF_BW; ++i) {
            const int e = tid + i * SDE_BLOCK;
#pragma unroll
            for (int q = 0; q < 9; ++q) pf.bw[i][q] = 0u;
            if (e < nd * SDE_NW) {
                // x_d(n_cta + warp part) = XOR over the nibbles of gray(n) of a 16-entry table: independent loads
                const int dl = e / SDE_NW, w = e - dl * SDE_NW;
#if SDE_DIRECT
                const sde_u32 nw = (sde_u32)n_cta + 128u * (sde_u32)(w >> 2) + (sde_u32)(w & 3);
#else
                const sde_u32 nw = (sde_u32)n_cta + 32u * (sde_u32)w;
#endif
                const sde_u32 g = nw ^ (nw >> 1);
                const sde_u32* tab = prm.sobol_nib + (d0 + dl) * 128;
#pragma unroll
                for (int q = (SDE_DIRECT ? 0 : 1); q < 8; ++q) pf.bw[i][q] = __ldg(tab + q * 16 + ((g >> (4 * q)) & 15u));
#if SDE_RNG == 2
                pf.bw[i][8] = __ldg(prm.xor_masks + d0 + dl);      // the digital shift of this dimension, folded in
#endif
            }
        }
#pragma unroll
        for (int i = 0; i < SDE_PF_LANE; ++i) {
            const int e = tid + i * SDE_BLOCK;
            pf.ln[i] = (e < nd * 32) ? __ldg(prm.sobol_lane + d0 * 32 + e) : 0u;
        }
#endif
    };
    auto commit = [&](const int t0, const SdeTilePrefetch& pf, const int buf) __attribute__((always_inline)) {
        const int nt = min(SDE_TS, S - t0);
        double* st = reinterpret_cast<double*>(s_stage + buf * SDE_SMEM_STAGE_BYTES);
#pragma unroll
        for (int i = 0; i < SDE_PF_STEP; ++i) {
            const int e = tid + i * SDE_BLOCK;
            if (e < nt) {
                double* rec = st + e * SDE_STEP_LD;
                rec[0] = pf.st[i][0]; rec[1] = pf.st[i][1]; rec[2] = pf.st[i][2]; rec[3] = pf.st[i][3];
#if SDE_NSLOT > 0
                double slots[SDE_NSLOT];
                sde_model_step_consts(pf.st[i][0], pf.st[i][1], pf.st[i][2], pf.st[i][3], slots);
#pragma unroll
                for (int q = 0; q < SDE_NSLOT; ++q) rec[4 + q] = slots[q];
#endif
            }
        }
#if SDE_USES_SOBOL
        sde_u32* bw = reinterpret_cast<sde_u32*>(s_stage + buf * SDE_SMEM_STAGE_BYTES + SDE_SMEM_STEP_BYTES);
        sde_u32* ln = bw + SDE_SMEM_BW_BYTES / 4;
#pragma unroll
        for (int i = 0; i < SDE_PF_BW; ++i) {
            const int e = tid + i * SDE_BLOCK;
            if (e < SDE_TS * SDE_KK * SDE_NW)
                bw[e] = sde_tile_fold(((pf.bw[i][0] ^ pf.bw[i][1]) ^ (pf.bw[i][2] ^ pf.bw[i][3])) ^ ((pf.bw[i][4] ^ pf.bw[i][5]) ^ (pf.bw[i][6] ^ pf.bw[i][7])) ^ pf.bw[i][8]);
        }
#pragma unroll
        for (int i = 0; i < SDE_PF_LANE; ++i) {
            const int e = tid + i * SDE_BLOCK;
            if (e < SDE_TS * SDE_KK * 32) ln[e] = sde_tile_fold(pf.ln[i]);
        }
#endif
    };

    // ScenarioFiltration::new — row 0 from initial_values, cache loaded from row 0 (filtration.rs:42-51)
    sde_real row[SDE_P], cache[SDE_P];
    double ct = __ldg(prm.times);
#pragma unroll
    for (int p = 0; p < SDE_P; ++p) { row[p] = (sde_real)__ldg(prm.x0 + p); cache[p] = row[p]; }

#if SDE_USES_CHACHA
    SdeChaCha8Stream cha;
    cha.init(prm.seed + s_global);                                 // sim/mod.rs:56,65 (wrapping add)
#endif
#if SDE_USES_PHILOX
    SdePhiloxStream phx;
    phx.init(prm.seed, s_global);
#endif
    sde_u32 rare_min = 0xffffffffu;                                // see draw_fixup
    (void)rare_min;

#if SDE_OUT == 0
    if (valid) {
#pragma unroll
        for (int p = 0; p < SDE_P; ++p) out_r[(size_t)s_local * T * SDE_P + p] = row[p];
    }
#if SDE_DIRECT
    sde_real* const my_row = out_r + (size_t)(valid ? s_local : 0) * T * SDE_P;      // this path's row [T][P]
    // step shift of this warp: the group that starts at step gamma writes elements from (gamma+1) P on, and
    // P (s T + gamma + 1) = 0 (mod 4) puts that on a 32-byte boundary (the output base is 32-byte aligned)
    const int gamma = __shfl_sync(0xffffffffu, (int)((4 - (int)(((long long)s_local * T + 1) & 3)) & 3), 0);
#else
    sde_real* my_tile = s_tile + (size_t)(warp * 32 + lane) * SDE_TILE_LD;
    const unsigned valid_mask = __ballot_sync(0xffffffffu, valid);
    const long long s_warp0 = s_local - lane;
    (void)valid_mask; (void)s_warp0;
#if SDE_TMA == 2
    // this warp's two boxes; box_row = this lane's 128-byte row with the lane's swizzle term folded in: chunk c of the row
    // lives at box_row ^ (c << 4) (+ 4096 for the second box)
    const unsigned box_warp = (unsigned)__cvta_generic_to_shared(s_tile) + (unsigned)warp * 8192u;
    const unsigned box_row = box_warp + (unsigned)lane * 128u + ((unsigned)(lane & 7) << 4);
    int box_open = -1;                                        // box (index = step / SDE_TBOX_STEPS) that holds unsent rows
    // box_send: all rows of box b are in shared memory -> one tensor store by lane 0 (columns (b STEPS + 1) P .., rows s_warp0 ..)
    auto box_send = [&](const int b) __attribute__((always_inline)) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&tmap), "r"((b * SDE_TBOX_STEPS + 1) * SDE_P), "r"((int)s_warp0), "r"(box_warp + (unsigned)(b & 1) * 4096u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    };
    // box_rows: the SDE_P values of `ns` consecutive steps starting at step t go into the lane's row of box t / STEPS
    auto box_put = [&](const int t, const int ns, const sde_real* vals) __attribute__((always_inline)) {
        if (s_warp0 < 0) {
            // the copy engine faults on a box that starts at a negative row (measured): the one warp of a launch that holds
            // the <= 5 leading pad lanes (Sobol::skip(5), sobol.rs:17) writes its rows itself
            if (valid) {
                sde_real* dst = out_r + ((size_t)s_local * T + (size_t)(t + 1)) * SDE_P;
#pragma unroll
                for (int q = 0; q < SDE_UNR * SDE_P; ++q) if (q < ns * SDE_P) dst[q] = vals[q];
            }
            return;
        }
        const int b = t / SDE_TBOX_STEPS, so = t - b * SDE_TBOX_STEPS;
        if (b != box_open) {
            // a new box: the store that last read this buffer (box b - 2) must be done with it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            box_open = b;
        }
        const unsigned base = (box_row + (unsigned)(b & 1) * 4096u) ^ ((unsigned)(so * SDE_P / 2) << 4);
#pragma unroll
        for (int q = 0; q < SDE_UNR * SDE_P / 2; ++q)
            if (q < ns * SDE_P / 2)
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base ^ ((unsigned)q << 4)), "d"(vals[2 * q]), "d"(vals[2 * q + 1]) : "memory");
        if (so + ns == SDE_TBOX_STEPS || t + ns == S) { box_send(b); box_open = -1; }
    };
#endif
#endif
#elif SDE_OUT == 1
    if (valid) {
#pragma unroll
        for (int p = 0; p < SDE_P; ++p) out_r[(size_t)p * prm.n_paths + s_local] = row[p];
    }
#endif

    {   // tile 0 (and tile 1 with four stage buffers) is staged synchronously
        SdeTilePrefetch pf;
        issue(0, pf);
        commit(0, pf, 0);
#if SDE_NSTAGE == 4
        if (SDE_TT < S) { issue(SDE_TT, pf); commit(SDE_TT, pf, 1); }
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_mbar), "r"(SDE_BLOCK) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_mbar + 8u), "r"(SDE_BLOCK) : "memory");
        }
#endif
    }
    __syncthreads();

#if SDE_DIRECT
    // next of the <= 3 steps that follow the warp's last full group; they run in whichever tile stages them (-1: not yet known)
    int tail_t = -1;
#endif
    int buf = 0;
#if SDE_NSTAGE == 4
    int tile = 0;
    for (int t0 = 0; t0 < S; t0 += SDE_TT, buf = (buf + 1) & 3, ++tile) {
        if (tile >= 2) {
            // tile `tile` was committed by every thread at the end of tile - 2: wait for that phase (parity of its count on
            // the even / odd barrier)
            const unsigned bar = s_mbar + 8u * (unsigned)(tile & 1), parity = (unsigned)(((tile - 2) >> 1) & 1);
            unsigned done;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            } while (!done);
        }
#else
    for (int t0 = 0; t0 < S; t0 += SDE_TT, buf ^= 1) {
#endif
        const int t_end = min(t0 + SDE_TT, S);
        const double* s_step = reinterpret_cast<const double*>(s_stage + buf * SDE_SMEM_STAGE_BYTES);
        const sde_u32* s_bw = reinterpret_cast<const sde_u32*>(s_stage + buf * SDE_SMEM_STAGE_BYTES + SDE_SMEM_STEP_BYTES);
        const sde_u32* s_lane = s_bw + SDE_SMEM_BW_BYTES / 4;
        (void)s_bw; (void)s_lane;

        //   draw(t, j, ..)    uniforms -> normal / Poisson draws of step t (j = position in the group; it must be a
        //                     compile-time constant after unrolling so the ChaCha buffer is indexed statically)
        //   advance(t, ..)    one step of the scheme + staging of the new row
        auto draw = [&](const int t, const int j, sde_real (&zu)[SDE_KK], sde_u0_t& u0) __attribute__((always_inline)) {
            const int tl = t - t0;
            u0 = (sde_u0_t)0;
            zu[0] = 0.0;
            (void)tl; (void)j;
#if SDE_RNG == 4
            {
                const double* src = prm.inject + ((size_t)(valid ? s_local : 0) * S + t) * (SDE_K + 1);
#pragma unroll
                for (int k = 0; k < SDE_K; ++k) zu[k] = __ldg(src + k);
                u0 = __ldg(src + SDE_K);
            }
#else
            // wide models: a partially unrolled loop keeps 4 inverse-normal chains in flight and lets the draws live in a
            // (L1-resident) local array instead of 2 K registers; ChaCha modes index their block buffer statically
#if SDE_K >= 16 && !SDE_USES_CHACHA && !SDE_USES_PHILOX
#pragma unroll 4
#else
#pragma unroll
#endif
            for (int k = 0; k < SDE_K; ++k) {
#if SDE_USES_CHACHA
                const int slot = (j * SDE_K + k) & 7;      // compile-time after unrolling
                if (slot == 0) cha.refill();
                const sde_u64 jc = cha.bits53(slot);       // rand: f64 = (next_u64 >> 11) * 2^-53
#endif
#if SDE_USES_SOBOL
                const int dl = tl * SDE_K + k;
                const sde_u32 x = s_bw[dl * SDE_NW + warp] ^ s_lane[dl * 32 + lane];    // (digitally shifted) Sobol integer
#endif
#if SDE_USES_PHILOX
                const int pslot = (j * SDE_K + k) & 3;     // compile-time after unrolling: groups start on a block boundary
                if (pslot == 0) phx.refill();
                const sde_u32 x = phx.buf[pslot];          // u = (x + 1/2) 2^-32
#endif
#if SDE_RNG == 0
                // uniforms of the form j * 2^-53 stay in the integer pipe until the draw is needed
                if (k == 0 && SDE_NEEDS_U0) u0 = (double)(long long)jc * 1.1102230246251565e-16;
                if (sde_factor_is_wiener(k)) {
#if SDE_ICDF == 1
                    {
                        sde_u32 xh;
                        zu[k] = sde_icdf_normal_fast_w64(cha.buf[2 * slot], cha.buf[2 * slot + 1], s_icdf, lane, &xh);
                        rare_min = min(rare_min, xh >> SDE_W64_RARE_SHIFT);   // 0 marks a draw with min(p, 1-p) < 2^-32: redone in draw_fixup
                    }
#elif SDE_ICDF == 2
                    zu[k] = sde_icdf_normal_single((double)(long long)jc * 1.1102230246251565e-16);
#else
                    zu[k] = sde_icdf_normal_reference((double)(long long)jc * 1.1102230246251565e-16);
#endif
                } else {
                    zu[k] = (double)(long long)jc * 1.1102230246251565e-16;
                }
#elif SDE_RNG == 2 || SDE_RNG == 5
                // digital shift: u = ((x ^ mask) + 1/2) * 2^-32, one 32-bit mask per dimension (already folded into x);
                // Philox words take the same 32-bit form.  Under SDE_TILED_FOLD `x` is the sign-folded word (see above).
#if SDE_U0_BITS
                if (k == 0 && SDE_NEEDS_U0) u0 = x;           // the step reads u0 > 1/2 = the top bit of x, folded or not (sde_u0_t)
#else
                if (k == 0 && SDE_NEEDS_U0) u0 = fma((double)sde_tile_fold(x), 2.3283064365386963e-10, 1.1641532182693481e-10);
#endif
                if (sde_factor_is_wiener(k)) {
#if SDE_TILED_FOLD
                    zu[k] = sde_icdf_normal_fast_y32s(x, tab_lane);
#elif SDE_TILED_W32
                    zu[k] = sde_icdf_normal_fast_k32s(x, tab_lane);
#elif SDE_ICDF == 2
                    zu[k] = sde_icdf_normal_single_k32(x);
#else
                    zu[k] = sde_icdf_normal_reference(fma((double)x, 2.3283064365386963e-10, 1.1641532182693481e-10));
#endif
                } else {
                    zu[k] = fma((double)sde_tile_fold(x), 2.3283064365386963e-10, 1.1641532182693481e-10);   // (folding is an involution)
                }
#else
                double u;
#if SDE_RNG == 1
                {   // RandomShiftScrambler::scramble: (raw + shift).fract()   (sobol.rs:73-76)
                    const double v = (double)x * 2.3283064365386963e-10 + (double)(long long)jc * 1.1102230246251565e-16;
                    u = (v >= 1.0) ? v - 1.0 : v;
                }
#else
                u = (double)x * 2.3283064365386963e-10;                                      // x / 2^32
#endif
                if (k == 0) u0 = u;
                zu[k] = sde_uniform_to_draw(u, sde_factor_is_wiener(k), s_icdf, lane);
#endif
            }
#endif
        };
#if SDE_RNG == 0 && SDE_ICDF == 1
        // Rare path of the 32-bit-word inverse normal (sde_icdf_normal_fast_w64): when some draw of the last `ns` steps had
        // min(p, 1-p) < 2^-32 (once per 2^32 draws) its block is regenerated (ChaCha is counter based) and the draw redone through
        // the general 53-bit entry.  One cold branch per step group keeps the draw code itself straight-line.
        auto draw_fixup = [&](const int tc, const int ns, sde_real (*zu)[SDE_KK]) __attribute__((always_inline)) {
            if (__builtin_expect(rare_min == 0u, 0)) {
                for (int j = 0; j < ns; ++j)
                    for (int k = 0; k < SDE_K; ++k) {
                        if (!sde_factor_is_wiener(k)) continue;
                        const sde_u64 i = (sde_u64)(tc + j) * SDE_K + k;                 // draw index in the path's stream
                        sde_u32 blk[16];
                        sde_chacha_block<8>(cha.key, i >> 3, blk);
                        const int s2 = 2 * (int)(i & 7);
                        sde_u32 lo = 0, hi = 0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) if (s2 == 2 * q) { lo = blk[2 * q]; hi = blk[2 * q + 1]; }
                        if ((sde_icdf_w64_folded_high(lo, hi) >> SDE_W64_RARE_SHIFT) == 0u)
                            zu[j][k] = sde_icdf_normal_fast_j53((((sde_u64)hi << 32) | (sde_u64)lo) >> 11, s_icdf, lane);
                    }
            }
            rare_min = 0xffffffffu;
        };
#else
        auto draw_fixup = [&](const int, const int, sde_real (*)[SDE_KK]) __attribute__((always_inline)) {};
#endif
        auto advance = [&](const int t, const sde_real (&zu)[SDE_KK], const sde_u0_t u0) __attribute__((always_inline)) {
            const int tl = t - t0;
            sde_model_step(row, cache, ct, zu, u0, s_step + tl * SDE_STEP_LD);
#if SDE_OUT == 0 && !SDE_DIRECT && !SDE_TMA
#pragma unroll
            for (int p = 0; p < SDE_P; ++p) my_tile[tl * SDE_P + p] = row[p];
#elif SDE_OUT == 1
            if (valid) {
#pragma unroll
                for (int p = 0; p < SDE_P; ++p)
                    out_r[((size_t)(t + 1) * SDE_P + p) * prm.n_paths + s_local] = row[p];
            }
#endif
        };
#if SDE_DIRECT
        sde_real* group_dst = my_row;                           // running store pointer of the step groups
#endif
        auto group = [&](const int tc) __attribute__((always_inline)) {
            sde_real zu[SDE_UNR][SDE_KK];
            sde_u0_t u0[SDE_UNR];
#pragma unroll
            for (int j = 0; j < SDE_UNR; ++j) draw(tc + j, j, zu[j], u0[j]);
            draw_fixup(tc, SDE_UNR, zu);
#if SDE_DIRECT
            sde_real vals[SDE_UNR * SDE_P];                     // rows tc+1 .. tc+4, in output order
#pragma unroll
            for (int j = 0; j < SDE_UNR; ++j) {
                advance(tc + j, zu[j], u0[j]);
#pragma unroll
                for (int p = 0; p < SDE_P; ++p) vals[j * SDE_P + p] = row[p];
            }
            {
                // predicated (not branched) stores: dead lanes only exist in the first and last CTA
                sde_real* dst = group_dst;                      // = my_row + (tc + 1) P: 32-byte aligned by the choice of gamma
                group_dst += 4 * SDE_P;
#ifdef SDE_DEBUG_NOSTORE
                const int live = (int)gridDim.z - 1;            // profiling aid: the stores stay in the code, predicated off at run time
#else
                const int live = valid ? 1 : 0;
#endif
#pragma unroll
                for (int q = 0; q < SDE_P; ++q) {
#if SDE_ST256
                    asm volatile("{ .reg .pred p; setp.ne.s32 p, %5, 0; @p st.global.v4.f64 [%0], {%1, %2, %3, %4}; }"
                                 ::"l"(dst + 4 * q), "d"(vals[4 * q]), "d"(vals[4 * q + 1]), "d"(vals[4 * q + 2]), "d"(vals[4 * q + 3]), "r"(live) : "memory");
#else
                    asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p st.global.v2.f64 [%0], {%1, %2}; }"
                                 ::"l"(dst + 4 * q), "d"(vals[4 * q]), "d"(vals[4 * q + 1]), "r"(live) : "memory");
                    asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p st.global.v2.f64 [%0], {%1, %2}; }"
                                 ::"l"(dst + 4 * q + 2), "d"(vals[4 * q + 2]), "d"(vals[4 * q + 3]), "r"(live) : "memory");
#endif
                }
            }
#elif SDE_OUT == 0 && SDE_TMA
            sde_real vals[SDE_UNR * SDE_P];                     // rows tc+1 .. tc+UNR, in output order
#pragma unroll
            for (int j = 0; j < SDE_UNR; ++j) {
                advance(tc + j, zu[j], u0[j]);
#pragma unroll
                for (int p = 0; p < SDE_P; ++p) vals[j * SDE_P + p] = row[p];
            }
#if SDE_TMA == 2
            box_put(tc, SDE_UNR, vals);
#else
            // the copy engine may still be reading the previous tile's segment from this row: its read is waited for here,
            // a whole group's draws and updates after it was issued (no-op from the second group of a tile on)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            {
                double2* dst = reinterpret_cast<double2*>(my_tile + (tc - t0) * SDE_P);
#pragma unroll
                for (int q = 0; q < SDE_UNR * SDE_P / 2; ++q) dst[q] = make_double2(vals[2 * q], vals[2 * q + 1]);
            }
#endif
#else
#pragma unroll
            for (int j = 0; j < SDE_UNR; ++j) advance(tc + j, zu[j], u0[j]);
#endif
        };
        // one step on its own (ragged ends); direct mode stores its row element-wise
        auto single = [&](const int t, const int j) __attribute__((always_inline)) {
            sde_real zu[1][SDE_KK];
            sde_u0_t u0;
            draw(t, j, zu[0], u0);
            draw_fixup(t, 1, zu);
            advance(t, zu[0], u0);
#if SDE_DIRECT
            if (valid) {
#pragma unroll
                for (int p = 0; p < SDE_P; ++p) my_row[(size_t)(t + 1) * SDE_P + p] = row[p];
            }
#elif SDE_OUT == 0 && SDE_TMA == 2
            {
                sde_real one[SDE_UNR * SDE_P];
#pragma unroll
                for (int p = 0; p < SDE_UNR * SDE_P; ++p) one[p] = p < SDE_P ? row[p] : (sde_real)0;
                box_put(t, 1, one);
            }
#elif SDE_OUT == 0 && SDE_TMA
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
            for (int p = 0; p < SDE_P; ++p) my_tile[(t - t0) * SDE_P + p] = row[p];
#endif
        };

#if SDE_NSTAGE == 4
        const bool more = t0 + 2 * SDE_TT < S;                // a tile after the next follows: prefetch it behind the last group
        const int t0_next = t0 + 2 * SDE_TT;
#else
        const bool more = t0 + SDE_TT < S;                    // another tile follows: prefetch it behind the last group
        const int t0_next = t0 + SDE_TT;
#endif
        SdeTilePrefetch pf;
#if SDE_DIRECT
        {
            // this warp's steps in tile k: [t0 + gamma, t0 + TT + gamma), clipped to the last full group; the first
            // gamma steps of the path and the <= 3 steps after the last full group run on their own
            const int g_eff = min(gamma, S);
            const int s_full = g_eff + ((S - g_eff) & ~3);
            if (t0 == 0) {
#pragma unroll
                for (int j = 0; j < 3; ++j) if (j < g_eff) single(j, j);
            }
            int tc = t0 + g_eff;
            group_dst = my_row + (size_t)(tc + 1) * SDE_P;
            const int t_hi = min(t0 + SDE_TT + g_eff, s_full);
            const int n_groups = t_hi > tc ? (t_hi - tc) >> 2 : 0;
#pragma unroll 1
            for (int gi = 0; gi + 1 < n_groups; ++gi, tc += 4) group(tc);
            if (more) issue(t0_next, pf);
            if (n_groups > 0) { group(tc); tc += 4; }
            // The <= 3 steps after the last full group run as soon as the groups are done, in the tile whose staged tables
            // [t0, t0 + SDE_TS) hold them.  (The last group can end up to 3 steps before a tile boundary; left to the last
            // tile, a tail that starts before that tile's t0 would index its staged tables with negative offsets.)
            if (t0 + SDE_TT + g_eff >= s_full) {
                if (tail_t < 0) tail_t = s_full;              // first such tile: t0 + g_eff < s_full held in the tile before
                const int staged_end = min(t0 + SDE_TS, S);
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (tail_t < staged_end) { single(tail_t, j); ++tail_t; }
            }
        }
#else
        {
            // full groups run unguarded (any tile length); at most SDE_UNR - 1 trailing steps take the guarded path
            const int n_groups = (t_end - t0) / SDE_UNR;
            int tc = t0;
#if SDE_UNR == 1 && SDE_P * SDE_KK >= 256
            // wide models (a step is thousands of instructions): ONE copy of the step in the instruction stream — a
            // peeled last group doubles the loop's footprint past the instruction cache (64-asset basket: 2 x 105 KB,
            // `no_instruction` was the second largest stall)
#pragma unroll 1
            for (int gi = 0; gi < n_groups; ++gi, tc += SDE_UNR) {
                if (more && gi + 1 == n_groups) issue(t0_next, pf);
                group(tc);
            }
            if (more && n_groups == 0) issue(t0_next, pf);
#else
#pragma unroll 1
            for (int gi = 0; gi + 1 < n_groups; ++gi, tc += SDE_UNR) group(tc);
            if (more) issue(t0_next, pf);
            if (n_groups > 0) { group(tc); tc += SDE_UNR; }
#pragma unroll
            for (int j = 0; j < SDE_UNR - 1; ++j)
                if (tc + j < t_end) single(tc + j, j);
#endif
        }
#endif
#if SDE_NSTAGE == 4
        if (more) {
            commit(t0_next, pf, (buf + 2) & 3);               // last read in tile k-2, which every warp has left (see SDE_NSTAGE)
            asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(s_mbar + 8u * (unsigned)(tile & 1)) : "memory");
        }
#else
        if (more) commit(t0_next, pf, buf ^ 1);               // the other buffer was last read in tile k-1 (barrier below)
#endif

#if SDE_OUT == 0 && SDE_TMA == 2
        // (boxes leave as they fill: nothing to do at the tile boundary)
#elif SDE_OUT == 0 && SDE_TMA
        // hand this lane's [t0+1, t_end] x P segment (contiguous in HBM, 16-byte aligned: P is even) to the copy engine
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the lane's own generic-proxy writes -> async proxy
        if (valid) {
            const sde_real* src_g = out_r + ((size_t)s_local * T + (t0 + 1)) * SDE_P;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(src_g), "r"((unsigned)__cvta_generic_to_shared(my_tile)), "r"((unsigned)((t_end - t0) * SDE_P * 8)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#elif SDE_OUT == 0 && !SDE_DIRECT
        // transpose through shared memory: each path's [t0+1, t_end] x P segment is contiguous in HBM
        __syncwarp();
        {
            const sde_real* wt = s_tile + (size_t)warp * 32 * SDE_TILE_LD;
            constexpr int NC = SDE_TT * SDE_P;                 // columns of a full tile
            const size_t row_stride = (size_t)T * SDE_P;
            constexpr bool kRegularNC = (NC <= 32 && 32 % NC == 0) || (NC % 32 == 0);
            if (kRegularNC && t_end - t0 == SDE_TT && valid_mask == 0xffffffffu) {
                // full tile, all 32 paths live
                sde_real* dst0 = out_r + ((size_t)s_warp0 * T + (t0 + 1)) * SDE_P;
                if constexpr (NC <= 32 && 32 % NC == 0) {
                    // a warp store covers 32/NC rows; running pointers, one 64-bit add per store
                    constexpr int RPI = 32 / NC;
                    const int r0 = lane / NC, i0 = lane % NC;
                    sde_real* p = dst0 + (size_t)r0 * row_stride + i0;
                    const sde_real* q = wt + r0 * SDE_TILE_LD + i0;
#pragma unroll
                    for (int it = 0; it < 32 / RPI; ++it) {
                        *p = q[it * RPI * SDE_TILE_LD];
                        p += RPI * row_stride;
                    }
                } else {
                    sde_real* p = dst0 + lane;
#pragma unroll 4
                    for (int r = 0; r < 32; ++r) {
#pragma unroll
                        for (int c = 0; c < NC / 32; ++c) p[c * 32] = wt[r * SDE_TILE_LD + c * 32 + lane];
                        p += row_stride;
                    }
                }
            } else {
                // partial tile and/or dead lanes: one row at a time, running pointers
                const int ncols = (t_end - t0) * SDE_P;
                sde_real* p = out_r + ((size_t)s_warp0 * T + (t0 + 1)) * SDE_P + lane;
                const sde_real* q = wt + lane;
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    if ((valid_mask >> r) & 1u) {
                        for (int i = lane; i < ncols; i += 32) p[i - lane] = q[i - lane];
                    }
                    p += row_stride;
                    q += SDE_TILE_LD;
                }
            }
        }
        __syncwarp();
#endif
#if SDE_NSTAGE != 4 && !defined(SDE_DEBUG_NOBARRIER)           // (NOBARRIER: profiling aid, results are garbage)
        __syncthreads();                                      // next tile's staged data visible; this tile's buffer free
#endif
    }
#if SDE_OUT == 0 && SDE_TMA
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // the last segment has left shared memory and reached global memory
#endif

#if SDE_OUT == 2
    if (valid) {
#pragma unroll
        for (int p = 0; p < SDE_P; ++p) out_r[(size_t)s_local * SDE_P + p] = row[p];
    }
#elif SDE_OUT == 3
    // warp-shuffle + block reduction of (count, mean, M2) per process; fixed order => deterministic
    for (int p = 0; p < SDE_P; ++p) {
        SdeMoments m;
        m.n = valid ? 1.0 : 0.0; m.mean = valid ? (double)row[p] : 0.0; m.m2 = 0.0;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            SdeMoments o;
            o.n = __shfl_xor_sync(0xffffffffu, m.n, off);
            o.mean = __shfl_xor_sync(0xffffffffu, m.mean, off);
            o.m2 = __shfl_xor_sync(0xffffffffu, m.m2, off);
            // merge in a lane-independent order so every lane holds the same bits
            m = ((lane & off) == 0) ? sde_mom_merge(m, o) : sde_mom_merge(o, m);
        }
        __syncthreads();
        if (lane == 0) { s_mom[warp * 3 + 0] = m.n; s_mom[warp * 3 + 1] = m.mean; s_mom[warp * 3 + 2] = m.m2; }
        __syncthreads();
        if (tid == 0) {
            SdeMoments acc; acc.n = s_mom[0]; acc.mean = s_mom[1]; acc.m2 = s_mom[2];
            for (int w = 1; w < SDE_NW; ++w) {
                SdeMoments o; o.n = s_mom[w * 3]; o.mean = s_mom[w * 3 + 1]; o.m2 = s_mom[w * 3 + 2];
                acc = sde_mom_merge(acc, o);
            }
            double* dst = prm.partials + ((size_t)blockIdx.x * SDE_P + p) * 3;
            dst[0] = acc.n; dst[1] = acc.mean; dst[2] = acc.m2;
        }
    }
#endif
}
