// sde_expr_helpers.cuh — device helpers referenced by generated coefficient code
// (csrc/host/expr.cpp).  Semantics follow the fasteval subset reachable from
// src/func.rs:18-42 ([3P-unverified]).
#pragma once

// sde_real: the type of the model state, the draws and the output rows — double, or float for dtype = f32 plans
// (the f32 variant: state, arithmetic and stored rows in single precision; time-grid quantities stay f64 on the host
// and are rounded once when a step reads them).
#ifndef SDE_F32
#define SDE_F32 0
#endif
#if SDE_F32
typedef float sde_real;
#else
typedef double sde_real;
#endif

// sde_u0_t: how the kernels hand the Runge-Kutta probe uniform u0 (runge_kutta.rs:18-22) to the step.  The step only asks
// u0 > 1/2.  Where the uniform is u0 = (w + 1/2) 2^-32 of a 32-bit word w (Sobol xor / none, Philox) that is exactly the top
// bit of w, so under arithmetic = fast the lowering defines SDE_U0_BITS and the kernels pass w itself: no conversion.
#ifndef SDE_U0_BITS
#define SDE_U0_BITS 0
#endif
#if SDE_U0_BITS
typedef unsigned int sde_u0_t;
#else
typedef sde_real sde_u0_t;
#endif

// sde_uc(v): the f64 constant v as a value of the UNIFORM datapath.
// On sm_100a a DFMA reads its register operands at one 64-bit pair per cycle (tools/ubench/dfma_operands.cu: three
// distinct register pairs = 3 cycles per warp instruction and sub-partition, two pairs or a uniform-register /
// immediate operand = 2), and the step loop is bound by exactly that operand traffic.  ptxas keeps loop-invariant f64
// constants in vector registers whatever their source (constant bank, literal, kernel parameter), so every
// fma(K, x, y) pays for three pairs.  Here the two words of the constant are formed by integer adds on a value that
// is zero at run time but opaque at compile time and warp-uniform by construction — gridDim.z - 1: every grid this
// library launches is one-dimensional — which ptxas evaluates on the uniform datapath and feeds to the DFMA as a
// uniform-register operand.  Same bits as the literal; no instruction inside the loop.
#ifndef SDE_UC
#define SDE_UC 1
#endif
__device__ __forceinline__ double sde_uc(const double v) {
#if SDE_UC
    const int uz = (int)gridDim.z - 1;
    const long long b = __double_as_longlong(v);
    return __hiloint2double((int)(unsigned)((unsigned long long)b >> 32) + uz, (int)(unsigned)b + uz);
#else
    return v;
#endif
}

#define SDE_F_EPS8 1.7763568394002505e-15   /* 8 * f64::EPSILON: fasteval's f64_eq! tolerance */
#define SDE_F_EPS8F 9.5367431640625e-7f     /* 8 * f32::EPSILON: the same rule for the f32 variant */

__device__ __forceinline__ double sde_f_sq(double x) { return __dmul_rn(x, x); }
__device__ __forceinline__ double sde_f_nan() { return __longlong_as_double(0x7ff8000000000000ll); }
__device__ __forceinline__ bool sde_f_is0(double x) { return fabs(x) <= SDE_F_EPS8; }
__device__ __forceinline__ double sde_f_not(double x) { return sde_f_is0(x) ? 1.0 : 0.0; }
__device__ __forceinline__ double sde_f_eq(double a, double b) { return fabs(a - b) <= SDE_F_EPS8 ? 1.0 : 0.0; }
__device__ __forceinline__ double sde_f_ne(double a, double b) { return fabs(a - b) <= SDE_F_EPS8 ? 0.0 : 1.0; }
// `and` / `or` return an operand, not a boolean; operands have no side effects so both are evaluated.
__device__ __forceinline__ double sde_f_and(double a, double b) { return sde_f_is0(a) ? a : b; }
__device__ __forceinline__ double sde_f_or(double a, double b) { return sde_f_is0(a) ? b : a; }
__device__ __forceinline__ double sde_f_sign(double x) { return (x != x) ? x : (signbit(x) ? -1.0 : 1.0); }   // f64::signum
// NaN-propagating min/max folded left to right
__device__ __forceinline__ double sde_f_min(double a, double b) { return (a != a || b != b) ? sde_f_nan() : (b < a ? b : a); }
__device__ __forceinline__ double sde_f_max(double a, double b) { return (a != a || b != b) ? sde_f_nan() : (b > a ? b : a); }

// x^0.5 under arithmetic = fast: rsqrt.approx.f64 seed (MUFU.RSQ64H, relative 2^-20) and one cubic step, 5 FP64 instructions
// instead of the ~10 + fix-up branch of the IEEE sequence; <= 1 ulp from the correctly rounded root.  Branch-free (so that
// repeated sub-expressions still fold): +inf and NaN pass through, negative arguments give NaN, +-0 gives +-0 and
// subnormal arguments are flushed to zero (the one stated deviation from sqrt, 1e-154 absolute).
__device__ __forceinline__ double sde_f_sqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y, e = fma(-g, y, 1.0);
    const double r = fma(g, e * fma(e, 0.375, 0.5), g);
    const double t = (x <= 1.7976931348623157e308) ? r : x;
    return (fabs(x) >= 2.2250738585072014e-308) ? t : x * 0.0;
}
__device__ __forceinline__ float sde_f_sqrt_fast(float x) { return sqrtf(x); }
// max(x, 0)^0.5 under arithmetic = fast (the full-truncation root of square-root diffusions), same value as
// sde_f_sqrt_fast(sde_f_max(x, 0.0)): NaN stays NaN, x <= 0 and subnormals give 0, +inf passes through.  The four FP64 compares
// of the two separate guards become one integer test of the high word and one FP64 compare: 6 FP64 instructions.
__device__ __forceinline__ double sde_f_sqrt_max0_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y, e = fma(-g, y, 1.0);
    const double r = fma(g, e * fma(e, 0.375, 0.5), g);
    const bool normal = (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u;    // positive, finite, not subnormal
    const bool pass = !(x < __longlong_as_double(0x7ff0000000000000ll));             // +inf and NaN (either sign): returned as they are
    const double z = pass ? x : 0.0;                                                 // x <= 0 (also -inf), subnormals
    return normal ? r : z;
}
// v with its sign flipped where the mask's top bit is set (mask = 0 or 0x80000000): products with sk = +-1 of the
// Runge-Kutta probe (runge_kutta.rs:18-22) without an FP64 instruction
__device__ __forceinline__ double sde_f_xorsign(double v, unsigned int m) { return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v)); }
__device__ __forceinline__ float sde_f_xorsign(float v, unsigned int m) { return __int_as_float(__float_as_int(v) ^ (int)m); }

// ---- f32 overloads (dtype = f32 plans; arithmetic = fast only, so no *_rn intrinsics are needed)
__device__ __forceinline__ float sde_f_sq(float x) { return x * x; }
__device__ __forceinline__ float sde_f_nanf() { return __int_as_float(0x7fc00000); }
__device__ __forceinline__ bool sde_f_is0(float x) { return fabsf(x) <= SDE_F_EPS8F; }
__device__ __forceinline__ float sde_f_not(float x) { return sde_f_is0(x) ? 1.0f : 0.0f; }
__device__ __forceinline__ float sde_f_eq(float a, float b) { return fabsf(a - b) <= SDE_F_EPS8F ? 1.0f : 0.0f; }
__device__ __forceinline__ float sde_f_ne(float a, float b) { return fabsf(a - b) <= SDE_F_EPS8F ? 0.0f : 1.0f; }
__device__ __forceinline__ float sde_f_and(float a, float b) { return sde_f_is0(a) ? a : b; }
__device__ __forceinline__ float sde_f_or(float a, float b) { return sde_f_is0(a) ? b : a; }
__device__ __forceinline__ float sde_f_sign(float x) { return (x != x) ? x : (signbit(x) ? -1.0f : 1.0f); }
__device__ __forceinline__ float sde_f_min(float a, float b) { return (a != a || b != b) ? sde_f_nanf() : (b < a ? b : a); }
__device__ __forceinline__ float sde_f_max(float a, float b) { return (a != a || b != b) ? sde_f_nanf() : (b > a ? b : a); }
