// lower.h — lowering of a parsed model to the CUDA translation unit of one plan.
//
// Replaces the interpretive inner loops of euler_iteration (src/sim/euler.rs:5-37) and
// runge_kutta_iteration (src/sim/runge_kutta.rs:5-107): the per-step control flow, the
// term order and the reference's time-keyed cache rule (src/func.rs:37-39) are resolved
// here, at lowering time, into straight-line device code over registers.
#pragma once
#include <string>

#include "universe.h"

namespace sde {

enum RngMode { RNG_PSEUDO = 0, RNG_SOBOL_CP = 1, RNG_SOBOL_XOR = 2, RNG_SOBOL_RAW = 3, RNG_INJECT = 4, RNG_PHILOX = 5 };
enum OutMode { OUT_PATHS_NTP = 0, OUT_PATHS_TPN = 1, OUT_TERMINAL = 2, OUT_MOMENTS = 3 };
enum SchemeId { SCHEME_EULER = 0, SCHEME_RK = 1 };

struct LowerOptions {
    int scheme = SCHEME_EULER;
    int rng = RNG_PSEUDO;
    int out = OUT_PATHS_NTP;
    int icdf = 0;            // 0 reference, 1 fast, 2 single (FP32 evaluation)
    bool strict = true;      // no FMA contraction in model arithmetic
    bool rk_textbook = false;
    bool u0_bits = false;    // set by lower_model: the RK probe uniform reaches the step as its 32-bit word (sde_u0_t)
    bool f32 = false;        // dtype f32: state, model arithmetic and stored rows in single precision (needs !strict)
    int block = 0;           // 0 = auto
    int tile_steps = 0;      // 0 = auto
    int min_blocks = 0;      // 0 = auto: CTAs per SM promised to the compiler (__launch_bounds__)
    int direct = -1;         // NTP paths: -1 = auto, 0 = shared-memory transpose, 1 = direct sector stores (tiled kernel),
                             // 2 = persistent-warp kernel with resident tables (sde_sim_resident.cuh),
                             // 3 = per-lane bulk copies shared -> global (cp.async.bulk; tiled kernel, f64, P even),
                             // 4 = per-warp 2-D tensor-map stores (cp.async.bulk.tensor; tiled kernel, f64, P = 2 or 4)
    int wide_mma = -1;       // wide linear models, terminal / moments: -1 = auto, 0 = time-tiled kernel, 1 = require the
                             // FP64 tensor-core kernel (sde_sim_wide.cuh)
};

struct Lowered {
    std::string source;      // complete translation unit (includes sde_sim_kernel.cuh)
    int block = 256;
    int min_blocks = 1;
    int tt = 32;
    int ch = 1;
    int unr = 1;             // steps unrolled per loop trip (multiple of ch)
    size_t smem_bytes = 0;   // dynamic shared memory of sde_sim_kernel
    bool direct = false;     // NTP full paths leave as 256-bit sector stores from registers (lane stride 4 mapping)
    int nstage = 2;          // time-tiled kernel: stage buffers (2: __syncthreads per tile; 4: mbarrier hand-over, SDE_NSTAGE)
    bool tma2 = false;       // ... as one 2-D tensor-map store per warp and box (SDE_TMA == 2): the launch takes a CUtensorMap
    bool tma = false;        // NTP full paths leave as per-lane bulk copies from a shared-memory staging row (16-byte aligned buffer)
    bool icdf_wide = false;  // persistent kernel: 1024-entry inverse-normal log table (128 KB of shared memory)
    bool lane_global = false; // persistent kernel: lane table prepared by the host in global memory (SDE_RES_LANE_GLOBAL)
    bool res_fold = false;   // persistent kernel: the tables hold sign-folded words (mirrors SDE_RES_FOLD)
    bool resident = false;   // persistent-warp kernel (sde_sim_resident.cuh): grid = SMs x min_blocks, whole time grid in shared memory
    bool wide = false;       // tensor-core kernel for wide linear models (sde_sim_wide.cuh): persistent warps, 8 wide_mt paths per warp
    int wide_mt = 0;         // row tiles (of 8 paths) per warp
    int wide_nb = 0, wide_nkk = 0;   // process tiles of 8 / factor steps of 4
    bool enter_eq = false;   // steady-state: cache.time == times[t] on entry to a step (stale-cache case)
};

Lowered lower_model(const Universe& u, const LowerOptions& opt);

}  // namespace sde
