// sde_device_icdf.cuh — inverse CDFs on the device (f64).
//
// Replaces fast_inverse_normal_cdf / fast_inverse_poisson_cdf (src/proc/increment.rs:160-200).
// The normal map is Abramowitz–Stegun 26.2.23 with the reference's constants; two
// evaluations of the SAME formula are provided:
//   REFERENCE  IEEE log / sqrt / divide and separately rounded mul/add in the reference's
//              operation order.  Differs from the CPU oracle only by CUDA's log (<= 1 ulp)
//              vs glibc's: |dz| <= 4 ulp(z) away from p = 0.5, <= 1e-15 absolute near it.
//   SINGLE     FP32 evaluation (see below): a separate precision tier, |dz| <= 4e-6.
//   FAST       table-driven log (128 x {1/c, -2 ln c}, degree-4 log1p), rsqrt.approx.f64 +
//              one cubic step, rcp.approx.f64 + one cubic step, FMA Horner: 21 FP64-pipe
//              instructions instead of ~60.  Stated tolerance: |dz| <= 5e-13 absolute vs
//              REFERENCE over p in [2^-53, 1 - 2^-53] (measured: tests/test_gpu_blocks.py).
#pragma once
#include "sde_icdf_tables.cuh"

#ifndef SDE_TYPES_DEFINED
#define SDE_TYPES_DEFINED
typedef unsigned int sde_u32;
typedef unsigned long long sde_u64;
#endif

#ifndef SDE_ICDF_F32SEED
#define SDE_ICDF_F32SEED 1                      /* FAST map: 1 = FP32-unit seeds + quadratic steps (sde_icdf_as_tail_f32seed), 0 = MUFU.*64H seeds + cubic steps */
#endif
#ifndef SDE_ICDF_HORNER
#define SDE_ICDF_HORNER 1                       /* FP32-seed variant: N(t), D(t) by Horner in t (0: the even/odd split in w = t^2) */
#endif

#define SDE_AS_C0 2.515517
#define SDE_AS_C1 0.802853
#define SDE_AS_C2 0.010328
#define SDE_AS_D1 1.432788
#define SDE_AS_D2 0.189269
#define SDE_AS_D3 0.001308

// increment.rs:161-179 verbatim in evaluation order; __d*_rn blocks FMA contraction.
__device__ __forceinline__ double sde_icdf_normal_reference(double p) {
    const bool lower = p < 0.5;
    const double w = lower ? p : __dsub_rn(1.0, p);
    const double t = sqrt(__dmul_rn(-2.0, log(w)));
    const double num = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(SDE_AS_C2, t), SDE_AS_C1), t), SDE_AS_C0);
    const double den = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(SDE_AS_D3, t), SDE_AS_D2), t), SDE_AS_D1), t), 1.0);
    const double x = __dsub_rn(t, __ddiv_rn(num, den));
    return lower ? -x : x;
}

__device__ __forceinline__ double sde_rsqrt_approx(double a) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}
__device__ __forceinline__ double sde_rcp_approx(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

// Shared-memory tables of the FAST path:
//   [0, 128*2*REPL)  log table {1/c, -2 ln c}; REPL = 8 replicates every entry once per 16-byte bank
//                    group so the 8 lanes of a quarter-warp never collide (LDS.128 = 4 clk/warp);
//   then 64 doubles  eln2[h] = (h - 53) * (-2 ln 2): the exponent term for w = 1.m * 2^(h-53)
//                    (avoids an int->f64 conversion per draw; lookups are bank-conflict free).
#ifndef SDE_ICDF_TABLE_REPL
#define SDE_ICDF_TABLE_REPL 8
#endif
#define SDE_ICDF_LOG_DOUBLES (128 * 2 * SDE_ICDF_TABLE_REPL)
#define SDE_ICDF_TABLE_DOUBLES (SDE_ICDF_LOG_DOUBLES + 64)

// y_offset is added to every -2 ln c entry: SDE_ICDF_Y_OFFSET_K32 for kernels that form the exponent term
// arithmetically (sde_icdf_normal_fast_k32s), 0 for the ones that read it from the eln2 table.
#ifndef SDE_ICDF_EXP_MAGIC
#define SDE_ICDF_EXP_MAGIC 0
#endif
#if SDE_ICDF_EXP_MAGIC
#define SDE_ICDF_Y_OFFSET_K32 134.47055302862938   /* 194 ln 2 = 66 ln 2 (w = 1.m 2^(pos-33)) + 128 ln 2 (D = 32 + pos/2) */
#else
#define SDE_ICDF_Y_OFFSET_K32 45.74771391695639    /* 66 ln 2: w = 1.m 2^(pos-33) */
#endif
// With SDE_ICDF_F32SEED the map works on 16 x (-2 ln w) (sde_icdf_as_tail_f32seed): both tables are stored times 16 (exact).
__device__ __forceinline__ void sde_icdf_table_load(double* s_table, int tid, int nthreads, double y_offset = 0.0) {
    const double sc = SDE_ICDF_F32SEED ? 16.0 : 1.0;
    for (int i = tid; i < 128 * SDE_ICDF_TABLE_REPL; i += nthreads) {
        const int idx = i / SDE_ICDF_TABLE_REPL;
        s_table[2 * i] = sde_icdf_log_table[idx][0];
        s_table[2 * i + 1] = (sde_icdf_log_table[idx][1] + y_offset) * sc;
    }
    for (int h = tid; h < 64; h += nthreads) s_table[SDE_ICDF_LOG_DOUBLES + h] = (double)(h - 53) * -1.3862943611198906 * sc;
}

// Constants of the FAST path live in constant memory so that FP64 instructions read them as
// c[bank][offset] operands instead of spending issue slots on 64-bit immediate moves.
__constant__ double sde_kc[24] = {
    0x1.0000999a03338p-1,   // 0  a3  } degree-3 minimax polynomial for -2 log1p(r) / r on |r| <= 2^-8,
    -0x1.55560888fbbc1p-1,  // 1  a2  } a1 = 1, a0 = -2 exact  (max |err| 4.9e-14 absolute in -2 ln w)
    SDE_AS_C2, SDE_AS_C1, SDE_AS_C0,       // 2..4
    SDE_AS_D3, SDE_AS_D2, SDE_AS_D1,       // 5..7
    0.375,                                 // 8   3/8 of the cubic square-root step
    -2.772588722239781,                    // 9   -4 ln 2: exponent term, D = 32 + pos/2 variant
    -1.3862943611198906,                   // 10  -2 ln 2: exponent term, converted-integer variant
    -0.6666666666666666,     // 12..: the FP32-seed variant works on 16 x (-2 ln w) (see sde_icdf_as_tail_f32seed): the same coefficients times powers of two
    -0.25 + 9.5e-15,                       // 12  -1/4 (+ half of the largest second-order term left out by the quadratic square-root step)
    -1.3862943611198906 * 16.0,            // 13  16 x (-2 ln 2)
    -0.6666666666666666 * 16.0,            // 14  16 x (-2/3)
    SDE_AS_C0 * 16.0, SDE_AS_C1 * 16.0,    // 15, 16
    SDE_AS_D1 * 16.0,                      // 17
    SDE_AS_C2 * 16.0, SDE_AS_D3 * 16.0, SDE_AS_D2 * 16.0,    // 18, 19, 20
    0x1.0000999a03338p-1 * 16.0, -0x1.55560888fbbc1p-1 * 16.0};   // 21, 22: a3, a2 times 16
// SDE_KC(i): coefficient i of the FAST path.
//   SDE_KC_MODE 0  the __constant__ array above (ptxas hoists the loads into vector registers)
//   SDE_KC_MODE 1  literals (same)
//   SDE_KC_MODE 2  uniform-datapath values, sde_uc() in sde_expr_helpers.cuh: uniform-register DFMA operands (default
//                  where that header is included, i.e. in the fused kernels)
#ifndef SDE_KC_MODE
#ifdef SDE_KC_LITERAL
#define SDE_KC_MODE 1
#elif defined(SDE_UC)
#define SDE_KC_MODE (SDE_UC ? 2 : 0)
#else
#define SDE_KC_MODE 0
#endif
#endif
#if SDE_KC_MODE
#define SDE_KCL_0 (0x1.0000999a03338p-1)
#define SDE_KCL_1 (-0x1.55560888fbbc1p-1)
#define SDE_KCL_2 (SDE_AS_C2)
#define SDE_KCL_3 (SDE_AS_C1)
#define SDE_KCL_4 (SDE_AS_C0)
#define SDE_KCL_5 (SDE_AS_D3)
#define SDE_KCL_6 (SDE_AS_D2)
#define SDE_KCL_7 (SDE_AS_D1)
#define SDE_KCL_8 (0.375)
#define SDE_KCL_9 (-2.772588722239781)
#define SDE_KCL_10 (-1.3862943611198906)
#define SDE_KCL_11 (-0.6666666666666666)
#define SDE_KCL_12 (-0.25 + 9.5e-15)
#define SDE_KCL_13 (-1.3862943611198906 * 16.0)
#define SDE_KCL_14 (-0.6666666666666666 * 16.0)
#define SDE_KCL_15 (SDE_AS_C0 * 16.0)
#define SDE_KCL_16 (SDE_AS_C1 * 16.0)
#define SDE_KCL_17 (SDE_AS_D1 * 16.0)
#define SDE_KCL_18 (SDE_AS_C2 * 16.0)
#define SDE_KCL_19 (SDE_AS_D3 * 16.0)
#define SDE_KCL_20 (SDE_AS_D2 * 16.0)
#define SDE_KCL_21 (0x1.0000999a03338p-1 * 16.0)
#define SDE_KCL_22 (-0x1.55560888fbbc1p-1 * 16.0)
#define SDE_KC(i) SDE_KCL_##i
#if SDE_KC_MODE == 2
// SDE_KU(i): a coefficient that is the ONLY non-register operand of its instruction (a DFMA takes one uniform-register,
// constant or immediate operand): those go through the uniform datapath; coefficients that share their instruction
// with an immediate or with another coefficient stay SDE_KC (a vector register)
#define SDE_KU(i) sde_uc(SDE_KCL_##i)
#else
#define SDE_KU(i) SDE_KCL_##i
#endif
#else
#define SDE_KC(i) sde_kc[i]
#define SDE_KU(i) sde_kc[i]
#endif

// Core: w = 1.mb * 2^e in (0, 0.5], given as mantissa bits (52 bits in hi:lo, leading one removed) and the
// byte offset `eoff` of the exponent term in the eln2 table.  Returns A&S x(w) (caller applies the sign).
// 20 FP64-pipe instructions + 2 MUFU:
//   -2 ln w   table {1/c, -2 ln c} on the top 7 mantissa bits, r = m/c - 1 (|r| <= 2^-8), degree-3 minimax
//             polynomial for -2 log1p(r)/r                                                               6
//   sqrt      rsqrt.approx.f64 seed (it only sees the high word: rel ~2^-20) + one cubic step           5
//   N/D       even/odd split in w2 = t^2 (2 + 3), rcp.approx.f64 seed + one cubic step, t - N r1       9
// Stated tolerance of the whole map against the REFERENCE evaluation: |dz| <= 5e-13 absolute.
// The MUFU seeds only define the high word of their result; PTX zero-fills the low word with an extra move.
// The cubic corrections below absorb a relative seed error of 2^-18, so the low word may be anything: borrow
// the low word of a value that is dead by then, which lets the register allocator write the seed's high
// word next to it (no move).  SDE_SEED_GARBAGE_LOW=0 restores the zero low word.
#ifndef SDE_SEED_GARBAGE_LOW
#define SDE_SEED_GARBAGE_LOW 1
#endif
#if SDE_SEED_GARBAGE_LOW == 2
// the low word is a register that is never written: ptxas is free to pair the seed's high word with any dead register
__device__ __forceinline__ double sde_seed_junk_low(double seed) {
    double r;
    asm("{ .reg .b32 lo, hi, junk; mov.b64 {lo, hi}, %1; mov.b64 %0, {junk, hi}; }" : "=d"(r) : "d"(seed));
    return r;
}
#define SDE_SEED_LOW(seed, donor) sde_seed_junk_low(seed)
#elif SDE_SEED_GARBAGE_LOW
#define SDE_SEED_LOW(seed, donor) __hiloint2double(__double2hiint(seed), __double2loint(donor))
#else
#define SDE_SEED_LOW(seed, donor) (seed)
#endif
// everything after the logarithm: w2 = -2 ln w  in [1.386, 73.5]  ->  x = t - N(t)/D(t), t = sqrt(w2).
// `d1`, `d2` are dead values whose low words seed the MUFU results (see SDE_SEED_LOW).
__device__ __forceinline__ double sde_icdf_as_tail(const double w2, const double d1, const double d2);
__device__ __forceinline__ double sde_icdf_as_tail_f32seed(const double w16);
// `base` = e (-2 ln 2) - 2 ln c from the tables of sde_icdf_table_load (times 16 with SDE_ICDF_F32SEED)
__device__ __forceinline__ double sde_icdf_as_core_b(const double m, const double2 tc, const double base) {
    const double r = fma(m, tc.x, -1.0);
#if SDE_ICDF_F32SEED
    double q = fma(r, SDE_KC(21), SDE_KC(22));               // 16 x the degree-3 minimax polynomial of -2 log1p(r) / r
    q = fma(q, r, 16.0);
    q = fma(q, r, -32.0);
    return sde_icdf_as_tail_f32seed(fma(q, r, base));
#else
    double q = fma(r, SDE_KC(0), SDE_KC(1));
    q = fma(q, r, 1.0);
    q = fma(q, r, -2.0);
    return sde_icdf_as_tail(fma(q, r, base), tc.x, base);
#endif
}
__device__ __forceinline__ double sde_icdf_as_tail(const double w2, const double d1, const double d2) {
    // t = sqrt(w2): y0 ~ w2^-1/2, g = w2 y0, e2 = 1 - w2 y0^2, t = g (1 + e2/2 + 3/8 e2^2)
    const double y0 = SDE_SEED_LOW(sde_rsqrt_approx(w2), d1);
    const double g = w2 * y0;
    const double e2 = fma(-g, y0, 1.0);
    const double ps = fma(e2, SDE_KC(8), 0.5);
#ifdef SDE_SEED_DONOR_LOCAL
    const double ep = e2 * ps;
    const double t = fma(g, ep, g);
#define SDE_R0_DONOR ep
#else
    const double t = fma(g, e2 * ps, g);
#define SDE_R0_DONOR d2
#endif
    // N(t) = (c0 + c2 w2) + c1 t and D(t) = (1 + d2 w2) + t (d1 + d3 w2): t^2 = w2 is known before the square root
    // is, so only one FMA of each polynomial waits for t (same operation count as Horner, shorter critical path)
    const double ne = fma(SDE_KU(2), w2, SDE_KC(4));
    const double de = fma(SDE_KC(6), w2, 1.0);
    const double dd = fma(SDE_KU(5), w2, SDE_KC(7));
    const double num = fma(SDE_KU(3), t, ne);
    const double den = fma(t, dd, de);
    // x = t - N/D with 1/D = r0 (1 + ed + ed^2), ed = 1 - D r0 (cubic step on the MUFU seed), folded into one final FMA
    const double r0 = SDE_SEED_LOW(sde_rcp_approx(den), SDE_R0_DONOR);
    const double ed = fma(-den, r0, 1.0);
    const double r1 = fma(r0, fma(ed, ed, ed), r0);
    return fma(-num, r1, t);
}
__device__ __forceinline__ double sde_icdf_as_core_v(const double m, const double2 tc, const double eterm) {
    return sde_icdf_as_core_b(m, tc, eterm + tc.y);         // base = e * (-2 ln 2) - 2 ln c
}
__device__ __forceinline__ double sde_icdf_as_core(sde_u32 mb_hi, sde_u32 mb_lo, const double* eterm, const double* s_table_lane) {
    const double m = __hiloint2double((int)(mb_hi | 0x3ff00000u), (int)mb_lo);       // in [1, 2)
    // s_table_lane = s_table + (lane & 7) * 2: the copy of the log table this lane's bank group owns
    const double2 tc = *reinterpret_cast<const double2*>(s_table_lane + (mb_hi >> 13) * (2 * SDE_ICDF_TABLE_REPL));
    return sde_icdf_as_core_v(m, tc, *eterm);
}

// p = j * 2^-53, j a 53-bit integer (rand's f64 is (u64 >> 11) * 2^-53).  min(p, 1-p), the
// exponent/mantissa split and the sign all happen in the integer pipe.
__device__ __forceinline__ double sde_icdf_normal_fast_j53(sde_u64 j, const double* s_table, int lane) {
    const bool upper = (j >> 52) != 0;                       // p >= 0.5  (increment.rs:165-169 uses 1 - p there)
    const sde_u64 jw = upper ? (0x20000000000000ull - j) : j;   // exact 1 - p
    const int h = 63 - __clzll((long long)jw);               // jw in [1, 2^52]
    const sde_u64 mb = (jw << (52 - h)) & 0xfffffffffffffull;
    double x = sde_icdf_as_core((sde_u32)(mb >> 32), (sde_u32)mb, s_table + SDE_ICDF_LOG_DOUBLES + h,
                                s_table + 2 * (lane & (SDE_ICDF_TABLE_REPL - 1)));
    int xhi = __double2hiint(x);
    xhi ^= upper ? 0 : 0x80000000;                           // p < 0.5 -> -x
    xhi = (j == 0) ? 0x7ff80000 : xhi;                       // ln(0) path of the reference -> NaN
    return __hiloint2double(xhi, __double2loint(x));
}

// Same map from the two raw 32-bit words of a u64 draw (rand: f64 = (u64 >> 11) * 2^-53), bit-identical to the j53 entry
// whenever min(p, 1-p) >= 2^-32 — all but 2^-32 of the draws.  X = u64 with its low 11 bits cleared is j 2^11, so
// p = X 2^-64 and min(p, 1-p) 2^64 is X or its two's complement -X; with a non-zero high word the leading one, the 52
// mantissa bits and the exponent come from one FLO and three funnel shifts on 32-bit words instead of 64-bit compare /
// select / shift sequences: ~16 integer instructions instead of ~26 (the ChaCha-driven modes are bound by the integer ALU
// pipe).  Branch-free: `*xh_out` receives the high word of the folded value; when it is zero (w < 2^-32, or p = 0) the
// returned value is meaningless and the caller must redo the draw through sde_icdf_normal_fast_j53 — the fused kernel
// does that once per step group, outside the straight-line draw code.
__device__ __forceinline__ double sde_icdf_normal_fast_w64(sde_u32 lo, sde_u32 hi, const double* s_table, int lane, sde_u32* xh_out) {
    const int sgn = (int)hi >> 31;                           // all ones when p >= 0.5
    const sde_u64 xs = ((sde_u64)(hi ^ (sde_u32)sgn) << 32) | (sde_u64)((lo & 0xfffff800u) ^ (sde_u32)sgn);
    const sde_u64 xf = xs + (sde_u64)(sde_u32)(-sgn);        // (X ^ S) - S: conditional 64-bit negate
    const sde_u32 xl = (sde_u32)xf, xh = (sde_u32)(xf >> 32);
    *xh_out = xh;
    int pos;                                                  // leading one of the high word: w = 1.m * 2^(pos - 32)
    asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(xh));
    const int sh = 31 - pos;
    const sde_u32 nh = __funnelshift_l(xl, xh, sh);          // normalised: leading one at bit 31 of nh
    const sde_u32 nl = xl << (sh & 31);
    const sde_u32 mb_hi = (nh & 0x7fffffffu) >> 11;          // 20 + 32 mantissa bits below the leading one (exact: X is a
    const sde_u32 mb_lo = __funnelshift_l(nl, nh, 21);       // multiple of 2^11, nothing is shifted out)
    double x = sde_icdf_as_core(mb_hi, mb_lo, s_table + SDE_ICDF_LOG_DOUBLES + 21 + (pos & 31),   // h - 53 = pos - 32
                                s_table + 2 * (lane & (SDE_ICDF_TABLE_REPL - 1)));
    const int xhi = __double2hiint(x) ^ (~sgn & 0x80000000);                          // p < 0.5 -> -x
    return __hiloint2double(xhi, __double2loint(x));
}
// high word of the folded value for the raw words of a draw (the rare-path test of the caller)
__device__ __forceinline__ sde_u32 sde_icdf_w64_folded_high(sde_u32 lo, sde_u32 hi) {
    const int sgn = (int)hi >> 31;
    const sde_u64 xs = ((sde_u64)(hi ^ (sde_u32)sgn) << 32) | (sde_u64)((lo & 0xfffff800u) ^ (sde_u32)sgn);
    return (sde_u32)((xs + (sde_u64)(sde_u32)(-sgn)) >> 32);
}

// p = (k + 1/2) * 2^-32, k = 32-bit digitally shifted Sobol integer: p = (2k+1) * 2^-33.
// 1 - p = (2 ~k + 1) * 2^-33, so min(p, 1-p) is a conditional bit flip of k.
__device__ __forceinline__ double sde_icdf_normal_fast_k32(sde_u32 k, const double* s_table, int lane) {
    const int sgn = (int)k >> 31;                            // all ones when p >= 0.5
    const sde_u32 j = ((k ^ (sde_u32)sgn) << 1) | 1u;        // w = min(p, 1-p) = j * 2^-33, j odd, 1 <= j < 2^32
    int pos;                                                  // leading one of j (FLO): w = 1.m * 2^(pos - 33)
    asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(j));
    const sde_u32 mh = __funnelshift_r(0u, j, pos);          // bits below the leading one, left aligned (pos = 0 -> 0)
    double x = sde_icdf_as_core(mh >> 12, mh << 20, s_table + SDE_ICDF_LOG_DOUBLES + 20 + pos,     // h - 53 = pos - 33
                                s_table + 2 * (lane & (SDE_ICDF_TABLE_REPL - 1)));
    const int xhi = __double2hiint(x) ^ (~sgn & 0x80000000);                         // p < 0.5 -> -x
    return __hiloint2double(xhi, __double2loint(x));
}

// Same map with explicit 32-bit shared-window addresses (persistent kernel): `tab_lane` = address of this lane's
// replica of log-table entry 0 of a table loaded with y_offset = SDE_ICDF_Y_OFFSET_K32.  Two integer instructions
// per table address, and the compiler cannot rematerialise the lane-dependent part inside the step loop.
// One shared-memory access per draw (LDS.128): the load/store data pipe is this kernel's busiest unit.
__device__ __forceinline__ double sde_icdf_fast_j32s(sde_u32 j, sde_u32 neg, sde_u32 tab_lane) {
    int pos;
    asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(j));
    const sde_u32 mh = __funnelshift_r(0u, j, pos);          // bits below the leading one, left aligned
    sde_u32 ta;                                              // tab_lane + (mh >> 25) * 128: shift + one multiply-add
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ta) : "r"(mh >> 25), "r"(16u * SDE_ICDF_TABLE_REPL), "r"(tab_lane));
    double2 tc;                                              // {1/c, -2 ln c + 194 ln 2}
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tc.x), "=d"(tc.y) : "r"(ta));
    // exponent term without a table: base = pos (-2 ln 2) + tc.y = -2 ln 2 (pos - 33) - 2 ln c, rounded once.
    // The int -> f64 conversion runs on the XU pipe (like FLO and the MUFU seeds), which has room; the variant
    // SDE_ICDF_EXP_MAGIC builds D = 32 + pos/2 with an integer multiply-add on the high word instead (two integer
    // instructions: the multiply-add and a zero low word).
#if SDE_ICDF_F32SEED
    const double base = fma((double)pos, SDE_KU(13), tc.y);  // the table is loaded times 16 (sde_icdf_table_load)
#elif SDE_ICDF_EXP_MAGIC
    sde_u32 dh;
    asm("mad.lo.u32 %0, %1, 16384, 0x40400000;" : "=r"(dh) : "r"((sde_u32)pos));
    const double base = fma(__hiloint2double((int)dh, 0), SDE_KC(9), tc.y);
#else
    const double base = fma((double)pos, SDE_KU(10), tc.y);
#endif
    const double m = __hiloint2double((int)((mh >> 12) | 0x3ff00000u), (int)(mh << 20));
    const double x = sde_icdf_as_core_b(m, tc, base);
    const int xhi = __double2hiint(x) ^ (int)(neg & 0x80000000u);
    return __hiloint2double(xhi, __double2loint(x));
}
__device__ __forceinline__ double sde_icdf_normal_fast_k32s(sde_u32 k, sde_u32 tab_lane) {
    const int sgn = (int)k >> 31;                            // all ones when p >= 0.5
    sde_u32 j;                                               // w = min(p, 1-p) = j * 2^-33, j = 2 (k ^ sgn) + 1
    asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(j) : "r"(k ^ (sde_u32)sgn));
    return sde_icdf_fast_j32s(j, ~(sde_u32)sgn, tab_lane);   // p < 0.5 -> -x
}
__device__ __forceinline__ double sde_icdf_normal_fast_y32s(sde_u32 y, sde_u32 tab_lane) {   // sign-folded entry, see y32w
    sde_u32 j;
    asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(j) : "r"(y));
    return sde_icdf_fast_j32s(j, ~y, tab_lane);
}

// ---- wide log table (persistent kernel, when shared memory allows): 1024 entries x 8 replicas = 128 KB.
// |r| <= 2^-11 makes the cubic Taylor polynomial of -2 log1p(r) exact to r^4/2 <= 2.8e-14, one FP64 instruction less
// than the 128-entry table's degree-3 minimax fit.  Built in the CTA prologue: invc = rn(1 / c_mid) and
// -2 ln c = 2 ln(invc) for exactly that invc (CUDA's f64 log: <= 1 ulp), so no constant array is needed.
#define SDE_ICDF_WIDE_BITS 10
#define SDE_ICDF_WIDE_DOUBLES ((1 << SDE_ICDF_WIDE_BITS) * 2 * SDE_ICDF_TABLE_REPL)
__device__ __forceinline__ void sde_icdf_wide_table_build(double* s_table, int tid, int nthreads, double y_offset,
                                                          double y_scale = (SDE_ICDF_F32SEED ? 16.0 : 1.0)) {
    for (int i = tid; i < (1 << SDE_ICDF_WIDE_BITS) * SDE_ICDF_TABLE_REPL; i += nthreads) {
        const int idx = i / SDE_ICDF_TABLE_REPL;
        const double c_mid = 1.0 + ((double)idx + 0.5) * (1.0 / (double)(1 << SDE_ICDF_WIDE_BITS));
        const double invc = __ddiv_rn(1.0, c_mid);
        s_table[2 * i] = invc;
        s_table[2 * i + 1] = __dadd_rn(__dmul_rn(2.0, log(invc)), y_offset) * y_scale;   // 1 or 16: exact scaling
    }
}
// FP32-unit seeds (SDE_ICDF_F32SEED): everything after the logarithm with MUFU.RSQ / MUFU.RCP (FP32: relative error 2^-22.4
// on the operand truncated to 24 bits, measured in tools/ubench/mufu_seed.cu; MUFU.RSQ64H / RCP64H only reach 2^-20) and
// ONE quadratic step each — 6 FP64 instructions for sqrt and 1/D instead of 9.  The operands travel between the f64 and
// f32 formats as bit patterns (funnel shift in, shift + add out), which only works when their exponents sit where the
// low 8 bits of the f64 exponent field are a valid f32 exponent: the map is therefore evaluated on w16 = 16 x (-2 ln w)
// and 16 D(t) (all coefficients scaled by powers of two: bit-identical products), whose f64 exponent fields 0x403..0x409
// read as f32 values 2^-128 times smaller.
//   sqrt:  s = (1/2) w16^-1/2 (1 + e),  P = w16 s,  E = -P s - 1/4 = -1/2 + e2/4  (e2 = 1 - (1+e)^2),
//          t = P + P E = (1/4) w16^1/2 (1 - 3/8 e2^2 ...) = sqrt(-2 ln w): relative error <= 3.8e-14, centred by the constant
//   1/D:   r0 = 1/(16 D) (1 + e),  ed = 1 - 16 D r0,  r1 = r0 + r0 ed   (relative error ed^2 <= 3e-14)
// Stated tolerance of the map unchanged: |dz| <= 5e-13 absolute vs the REFERENCE evaluation (tests/test_gpu_blocks.py).
__device__ __forceinline__ double sde_icdf_as_tail_f32seed(const double w16) {
    float ys, rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ys) : "f"(__uint_as_float(__funnelshift_l((sde_u32)__double2loint(w16), (sde_u32)__double2hiint(w16), 3))));
    const sde_u32 yb = __float_as_uint(ys);                  // ys = w16^-1/2 2^64
    const double s = __hiloint2double((int)((yb >> 3) + 0x33F00000u), (int)(yb << 29));   // ys 2^-65
    const double P = w16 * s;
    const double E = fma(-P, s, SDE_KU(12));
    const double t = fma(P, E, P);
#if SDE_ICDF_HORNER
    // 16 N(t) and 16 D(t) by Horner in t: every instruction reads two register pairs (t and the running value) plus a
    // uniform-register or immediate coefficient — no three-pair DFMA (the even/odd split below ends in den = t dd + de)
    const double num = fma(fma(t, SDE_KU(18), SDE_KC(16)), t, SDE_KU(15));
    const double den = fma(fma(fma(t, SDE_KU(19), SDE_KC(20)), t, SDE_KU(17)), t, 16.0);
#else
    // 16 N(t) = (16 c0 + c2 w16) + 16 c1 t,  16 D(t) = (16 + d2 w16) + t (16 d1 + d3 w16)
    const double ne = fma(SDE_KU(2), w16, SDE_KC(15));
    const double de = fma(SDE_KC(6), w16, 16.0);
    const double dd = fma(SDE_KU(5), w16, SDE_KC(17));
    const double num = fma(SDE_KU(16), t, ne);
    const double den = fma(t, dd, de);
#endif
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(__uint_as_float(__funnelshift_l((sde_u32)__double2loint(den), (sde_u32)__double2hiint(den), 3))));
    const sde_u32 rb = __float_as_uint(rs);                  // rs = 2^128 / (16 D)
    const double r0 = __hiloint2double((int)((rb >> 3) + 0x30000000u), (int)(rb << 29));   // rs 2^-128
    const double ed = fma(-den, r0, 1.0);
    const double r1 = fma(r0, ed, r0);
    return fma(-num, r1, t);
}
// core of the wide-table map: j = 2 v + 1 with v = the 31 folded bits of min(p, 1-p) 2^32 - 1/2; `neg` has bit 31 set when
// the result is to be negated (p < 1/2)
template <int F32SEED>
__device__ __forceinline__ double sde_icdf_fast_j32w_t(sde_u32 j, sde_u32 neg, sde_u32 tab_lane) {
    int pos;
    asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(j));
    const sde_u32 mh = __funnelshift_r(0u, j, pos);          // bits below the leading one, left aligned
    sde_u32 ta;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ta) : "r"(mh >> (32 - SDE_ICDF_WIDE_BITS)), "r"(16u * SDE_ICDF_TABLE_REPL), "r"(tab_lane));
    double2 tc;                                              // {1/c, -2 ln c + 66 ln 2}
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tc.x), "=d"(tc.y) : "r"(ta));
    const double m = __hiloint2double((int)((mh >> 12) | 0x3ff00000u), (int)(mh << 20));
    const double r = fma(m, tc.x, -1.0);
    double x;
    if (F32SEED) {
        const double base = fma((double)pos, SDE_KU(13), tc.y);  // 16 x (-2 ln 2 (pos - 33) - 2 ln c): table built with y_scale = 16
        double q = fma(r, SDE_KC(14), 16.0);
        q = fma(q, r, -32.0);
        x = sde_icdf_as_tail_f32seed(fma(q, r, base));
    } else {
        const double base = fma((double)pos, SDE_KU(10), tc.y);
        double q = fma(r, SDE_KC(11), 1.0);                  // -2 log1p(r) = r (-2 + r (1 - 2/3 r)) + O(r^4 / 2)
        q = fma(q, r, -2.0);
        x = sde_icdf_as_tail(fma(q, r, base), tc.x, base);
    }
    const int xhi = __double2hiint(x) ^ (int)(neg & 0x80000000u);
    return __hiloint2double(xhi, __double2loint(x));
}
__device__ __forceinline__ double sde_icdf_fast_j32w(sde_u32 j, sde_u32 neg, sde_u32 tab_lane) {
    return sde_icdf_fast_j32w_t<SDE_ICDF_F32SEED>(j, neg, tab_lane);
}
__device__ __forceinline__ double sde_icdf_normal_fast_k32w(sde_u32 k, sde_u32 tab_lane) {
    const int sgn = (int)k >> 31;                            // all ones when p >= 0.5
    sde_u32 j;                                               // w = min(p, 1-p) = j * 2^-33, j = 2 (k ^ sgn) + 1
    asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(j) : "r"(k ^ (sde_u32)sgn));
    return sde_icdf_fast_j32w(j, ~(sde_u32)sgn, tab_lane);   // p < 0.5 -> -x
}
// Sign-folded entry (persistent kernel, SDE_RES_FOLD): y = k ^ ((k >>s 31) & 0x7fffffff) arrives ready made — bit 31 is
// the sign bit of k, bits 30..0 are those of k ^ sgn — so 2 y + 1 (mod 2^32) is j and ~y carries the negation flag.
__device__ __forceinline__ double sde_icdf_normal_fast_y32w(sde_u32 y, sde_u32 tab_lane) {
    sde_u32 j;
    asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(j) : "r"(y));
    return sde_icdf_fast_j32w(j, ~y, tab_lane);
}

// General f64 entry (Cranley–Patterson compat mode, stand-alone kernel): p in [0, 1).
__device__ __forceinline__ double sde_icdf_normal_fast(double p, const double* s_table, int lane) {
    const bool lower = p < 0.5;
    const double w = lower ? p : 1.0 - p;                   // exact for every p the generators produce
    const int hi = __double2hiint(w);
    const int e = ((hi >> 20) & 0x7ff) - 1023;              // w = 1.m * 2^e, e in [-53, -1]
    double x = sde_icdf_as_core((sde_u32)hi & 0x000fffffu, (sde_u32)__double2loint(w), s_table + SDE_ICDF_LOG_DOUBLES + min(max(e + 53, 0), 63),
                                s_table + 2 * (lane & (SDE_ICDF_TABLE_REPL - 1)));
    if (!(w > 0.0)) x = __longlong_as_double(0x7ff8000000000000ll);   // p = 0 -> NaN like ln(0) in the reference
    return lower ? -x : x;
}

// ---- SINGLE: the same A&S map evaluated in FP32 (icdf = "single").
// MUFU.LG2 / MUFU.RSQ / MUFU.RCP and FP32 Horner: ~12 FP32-pipe + 3 MUFU instructions and no FP64-pipe work until the
// conversion of the result.  Stated tolerance against the REFERENCE evaluation: |dz| <= 4e-6 absolute over the whole
// range (<= 1e-6 for |z| <= 3; measured in tests/test_gpu_blocks.py) — two orders below A&S 26.2.23's own 4.5e-4
// approximation error, but far above the 1e-12 of the f64 tiers: a precision tier of its own, never a default.
__device__ __forceinline__ float sde_icdf_as_single_core(float wf) {
    const float w2 = -1.3862943611198906f * __log2f(wf);     // -2 ln w  (w <= 1/2: |log2 w| >= 1, <= 2 ulp)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(w2));
    const float t = w2 * y;
    const float num = fmaf(fmaf((float)SDE_AS_C2, t, (float)SDE_AS_C1), t, (float)SDE_AS_C0);
    const float den = fmaf(fmaf(fmaf((float)SDE_AS_D3, t, (float)SDE_AS_D2), t, (float)SDE_AS_D1), t, 1.0f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
    return fmaf(-num, r, t);
}
// p = (k + 1/2) 2^-32, k the 32-bit digitally shifted Sobol integer
__device__ __forceinline__ float sde_icdf_normal_single_k32(sde_u32 k) {   // float: f64 plans widen it, f32 plans use it as is
    const int sgn = (int)k >> 31;                            // all ones when p >= 0.5
    const sde_u32 v = k ^ (sde_u32)sgn;                      // min(p, 1-p) = (v + 1/2) 2^-32
    const float wf = fmaf((float)v, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float x = sde_icdf_as_single_core(wf);
    return __int_as_float(__float_as_int(x) ^ (~sgn & 0x80000000));                  // p < 0.5 -> -x
}
// general entry: p in [0, 1); p = 0 gives NaN like the reference's ln(0) path
__device__ __forceinline__ float sde_icdf_normal_single(double p) {
    const bool lower = p < 0.5;
    const float wf = (float)(lower ? p : 1.0 - p);
    const float x = sde_icdf_as_single_core(wf);
    return lower ? -x : x;
}

// increment.rs:182-200 verbatim.
__device__ __forceinline__ double sde_icdf_poisson(double u, double lambda) {
    if (lambda <= 0.0) return 0.0;
    double p = exp(-lambda), f = p;
    int k = 0;
    while (u > f && k < 200) {
        k += 1;
        p = __dmul_rn(p, __ddiv_rn(lambda, (double)k));
        f = __dadd_rn(f, p);
    }
    return (double)k;
}
