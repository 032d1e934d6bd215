// sde_expr_helpers.cuh — device helpers referenced by generated coefficient code
// (csrc/host/expr.cpp).  Semantics follow the fasteval subset reachable from
// src/func.rs:18-42 ([3P-unverified]).
#pragma once

#define SDE_F_EPS8 1.7763568394002505e-15   /* 8 * f64::EPSILON: fasteval's f64_eq! tolerance */

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
