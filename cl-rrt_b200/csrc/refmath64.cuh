// refmath64.cuh — double sin / cos / tan with bit-for-bit the results of the reference platform.
//
// PROVENANCE / LICENCE NOTE.  The GNU C Library is distributed under the LGPL v2.1 or later; its sinf / cosf / exp come from
// ARM Optimized Routines (MIT OR Apache-2.0 WITH LLVM-exception), atanf / acosf / asinf from FreeBSD msun (Sun Microsystems'
// permissive notice), sin / cos / tan from the IBM Accurate Mathematical Library (LGPL).  Nothing was copied from the
// reference repository (it contains no libm).  The routines below were written from the published algorithms against the
// behaviour of the installed library (tests/test_refmath.py sweeps them against it); the numeric tables are constants
// extracted from the installed libm by scripts/gen_refmath64_tables.py.  A redistributor who treats restated algorithms
// and extracted tables as derived work of glibc should ship these three files under the LGPL.
//
// The reference integrates its vehicle model in double through the C library: sin, cos (heading) and tan (steering angle)
// of glibc 2.39 on x86-64 (rrt/src/simulation.cpp:11-25, rrt/src/controller.cpp:56-57).  CUDA's own double routines agree
// with glibc's to the last ulp in most calls, not in all; the states then drift apart by ~1e-13, which is harmless except
// where the reference itself amplifies the last bit: a goal-biased reference repeats its junction point, and while the
// controller's three-point window straddles the pair the Lagrange interpolation divides by ~1e-15 —
// the saturated steer command takes a sign that hangs on the last bit of sin/cos/tan.  To follow the reference through
// those steps this header restates glibc's routines operation by operation:
//   * sin / cos: sysdeps/ieee754/dbl-64/s_sin.c (IBM Accurate Mathematical Library as simplified in glibc 2.28+): do_sin /
//     do_cos around the 440-entry table __sincostab (x_k = k/128), TAYLOR_SIN below 0.126, reduce_sincos (Cody-Waite in
//     four parts) up to |x| < 105414350.
//   * tan: sysdeps/ieee754/dbl-64/s_tan.c after the removal of its slow paths (glibc 2.35+): |x| <= 1.259e-8 -> x; <= 0.0608
//     -> odd polynomial; <= 0.787 -> table xfg (utan.tbl) + one division.  (The steering angle is saturated to +-0.52.)
// On x86-64 CPUs with FMA, glibc's ifunc dispatch selects the build of these sources compiled with -mfma -mavx2, in which
// the compiler contracted many a*b+c into fused multiply-adds; WHICH ones is not in the source, so the operation sequence
// below follows the machine code of that build (Ubuntu glibc 2.39-0ubuntu8, __sin_fma / __cos_fma / __tan_fma) — every
// fma() here is a vfmadd/vfnmadd/vfmsub there, every separate product or sum a vmulsd/vaddsd/vsubsd.  Constants are the
// published ones (usncs.h, utan.h), the two tables are generated from the installed libm (scripts/gen_refmath64_tables.py).
// tests/test_refmath.py compiles this header for the host (REFMATH_HOST) and compares it with the C library bit for bit
// over the ranges the rollout uses.  Outside the restated ranges (|x| >= 105414350 for sin/cos, |x| > 0.787 for tan,
// non-finite arguments) the functions fall back to the platform's own routine.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef REFMATH_HOST
#define RM64_FN static inline
#define RM64_CORE static inline
#define RM64_TABLE static const
static inline int64_t rm64_bits(double d) { int64_t u; memcpy(&u, &d, 8); return u; }
static inline double rm64_fallback_sin(double x) { return sin(x); }
static inline double rm64_fallback_cos(double x) { return cos(x); }
static inline double rm64_fallback_tan(double x) { return tan(x); }
#else
#define RM64_FN __device__ __forceinline__
#ifdef RM64_NOINLINE_CORE
#define RM64_CORE __device__ __noinline__
#else
#define RM64_CORE __device__ __forceinline__
#endif
#define RM64_TABLE static __device__ const
__device__ __forceinline__ int64_t rm64_bits(double d) { return (int64_t)__double_as_longlong(d); }
__device__ __noinline__ double rm64_fallback_sin(double x) { return sin(x); }
__device__ __noinline__ double rm64_fallback_cos(double x) { return cos(x); }
__device__ __noinline__ double rm64_fallback_tan(double x) { return tan(x); }
#endif

#include "refmath64_tables.inc"

// usncs.h
#define RM64_BIG 0x1.8p45
#define RM64_SN3 -0x1.5555555555515p-3
#define RM64_SN5 0x1.11110e829872fp-7
#define RM64_CS2 0x1p-1
#define RM64_CS4 -0x1.5555555555535p-5
#define RM64_CS6 0x1.6c16bedd9e239p-10
#define RM64_S1 -0x1.5555555555555p-3
#define RM64_S2 0x1.1111111110ecep-7
#define RM64_S3 -0x1.a01a019db08b8p-13
#define RM64_S4 0x1.71de27b9a7ed9p-19
#define RM64_S5 -0x1.addffc2fcdf59p-26
#define RM64_HP0 0x1.921fb54442d18p+0
#define RM64_HP1 0x1.1a62633145c07p-54
#define RM64_TOINT 0x1.8p52
#define RM64_HPINV 0x1.45f306dc9c883p-1
#define RM64_MP1 0x1.921fb58000000p+0
#define RM64_MP2 -0x1.dde973c000000p-27
#define RM64_PP3 -0x1.cb3b398000000p-55
#define RM64_PP4 -0x1.d747f23e32ed7p-83
// utan.h
#define RM64_G1 0x1.b096cp-27
#define RM64_G2 0x1.f212dp-5
#define RM64_G3 0x1.92f1ap-1
#define RM64_D3 0x1.5555555555555p-2
#define RM64_D5 0x1.11111111107c6p-3
#define RM64_D7 0x1.ba1ba1cdb8745p-5
#define RM64_D9 0x1.664ed49cfc666p-6
#define RM64_D11 0x1.2385a3cf2e4eap-7
#define RM64_E0 0x1.5555555554dbdp-2
#define RM64_E1 0x1.11112e0a6b45fp-3

// TAYLOR_SIN (s_sin.c): x + ((POLYNOMIAL(xx) * x - 0.5 * dx) * xx + dx)
RM64_FN double rm64_taylor_sin(double xx, double x, double dx) {
  double p = fma(RM64_S5, xx, RM64_S4);
  p = fma(p, xx, RM64_S3);
  p = fma(p, xx, RM64_S2);
  p = fma(p, xx, RM64_S1);
  const double h = dx * 0.5;
  const double t = fma(xx, fma(p, x, -h), dx);
  return x + t;
}

// do_sin (s_sin.c): sin(x + dx) for |x| < ~0.86, |dx| tiny.  `tab` = __sincostab (any address space)
RM64_CORE double rm64_do_sin(double x, double dx, const double* tab) {
  const double xold = x;
  if (fabs(x) < 0.126) return rm64_taylor_sin(x * x, x, dx);
  if (x <= 0) dx = -dx;
  const double u = RM64_BIG + fabs(x);
  x = fabs(x) - (u - RM64_BIG);
  const double xx = x * x;
  const double s = x + fma(x * xx, fma(xx, RM64_SN5, RM64_SN3), dx);
  const double c = fma(x, dx, xx * fma(xx, fma(xx, RM64_CS6, RM64_CS4), RM64_CS2));
  const int k = (int)((uint32_t)rm64_bits(u) << 2);
  const double sn = tab[k], ssn = tab[k + 1], cs = tab[k + 2], ccs = tab[k + 3];
  const double cor = fma(s, cs, fma(-c, sn, fma(s, ccs, ssn)));
  return copysign(sn + cor, xold);
}

// do_cos (s_sin.c): cos(x + dx)
RM64_CORE double rm64_do_cos(double x, double dx, const double* tab) {
  if (x < 0) dx = -dx;
  const double u = RM64_BIG + fabs(x);
  x = fabs(x) - (u - RM64_BIG);
  x = x + dx;
  const double xx = x * x;
  const double s = fma(x * xx, fma(xx, RM64_SN5, RM64_SN3), x);
  const double c = xx * fma(xx, fma(xx, RM64_CS6, RM64_CS4), RM64_CS2);
  const int k = (int)((uint32_t)rm64_bits(u) << 2);
  const double sn = tab[k], ssn = tab[k + 1], cs = tab[k + 2], ccs = tab[k + 3];
  const double cor = fma(-s, sn, fma(-c, cs, fma(-s, ssn, ccs)));
  return cs + cor;
}

// reduce_sincos (s_sin.c): x = n*pi/2 + a + da, |x| < 105414350; returns n mod 4
RM64_FN int rm64_reduce(double x, double* a, double* da) {
  const double t = fma(x, RM64_HPINV, RM64_TOINT);
  const double xn = t - RM64_TOINT;
  const int n = (int)(rm64_bits(t) & 3);
  const double y = fma(-xn, RM64_MP2, fma(-xn, RM64_MP1, x));
  const double t2 = fma(-xn, RM64_PP3, y);
  const double db = fma(-xn, RM64_PP3, y - t2);
  const double b = fma(-xn, RM64_PP4, t2);
  const double db2 = fma(-xn, RM64_PP4, t2 - b);
  *a = b;
  *da = db + db2;
  return n;
}

// __sin / __cos (s_sin.c).  k = high word of |x|.  (The tables are read through L1: staged in shared memory the rollout
// kernel was 6-20 % slower.)
RM64_FN double ref_sin(double x) {
  const double* tab = rm64_sincostab;
  const uint32_t k = (uint32_t)(rm64_bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;                            // |x| < 2^-26
  if (k < 0x3feb6000u) return rm64_do_sin(x, 0.0, tab);     // |x| < 0.855469
  if (k < 0x400368fdu) {                                    // |x| < 2.426265
    const double t = RM64_HP0 - fabs(x);
    return copysign(rm64_do_cos(t, RM64_HP1, tab), x);
  }
  if (k < 0x419921FBu) {                                    // |x| < 105414350
    double a, da;
    const int n = rm64_reduce(x, &a, &da);
    const double r = (n & 1) ? rm64_do_cos(a, da, tab) : rm64_do_sin(a, da, tab);
    return (n & 2) ? -r : r;
  }
  return rm64_fallback_sin(x);
}
RM64_FN double ref_cos(double x) {
  const double* tab = rm64_sincostab;
  const uint32_t k = (uint32_t)(rm64_bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;                          // |x| < 2^-27
  if (k < 0x3feb6000u) return rm64_do_cos(x, 0.0, tab);
  if (k < 0x400368fdu) {
    const double y = RM64_HP0 - fabs(x);
    const double a = y + RM64_HP1;
    const double da = (y - a) + RM64_HP1;
    return rm64_do_sin(a, da, tab);
  }
  if (k < 0x419921FBu) {
    double a, da;
    const int n = rm64_reduce(x, &a, &da) + 1;
    const double r = (n & 1) ? rm64_do_cos(a, da, tab) : rm64_do_sin(a, da, tab);
    return (n & 2) ? -r : r;
  }
  return rm64_fallback_cos(x);
}
// the reference calls std::sin and std::cos on the same heading
#ifdef RM64_PLAIN
RM64_FN void ref_sincos(double x, double* s, double* c) { *s = ref_sin(x); *c = ref_cos(x); }
#else
// Every range of __sin / __cos is one do_sin and one do_cos call whose results are swapped / negated: the arguments are
// selected first and the two routines are evaluated once for both results (one copy of each in the instruction stream, no
// divergence between lanes whose headings lie in different ranges); every operation sequence is the one glibc runs for that
// argument.  (Rollout kernel at 8 warps per SM: equal on C3, 3 % faster on latency-bound rounds than the two plain calls.)
RM64_FN void ref_sincos(double x, double* s, double* c) {
  const double* tab = rm64_sincostab;
  const uint32_t k = (uint32_t)(rm64_bits(x) >> 32) & 0x7fffffffu;
  if (!(k < 0x419921FBu)) { *s = rm64_fallback_sin(x); *c = rm64_fallback_cos(x); return; }
  double p, dp, q, dq;
  bool swap = false, neg_s = false, neg_c = false, sign_from_x = false;
  if (k < 0x3feb6000u) { p = x; dp = 0.0; q = x; dq = 0.0; }
  else if (k < 0x400368fdu) {
    const double y = RM64_HP0 - fabs(x);
    q = y; dq = RM64_HP1;
    p = y + RM64_HP1; dp = (y - p) + RM64_HP1;
    swap = true; sign_from_x = true;
  } else {
    double a, da;
    const int n = rm64_reduce(x, &a, &da);
    p = a; dp = da; q = a; dq = da;
    swap = (n & 1) != 0; neg_s = (n & 2) != 0; neg_c = ((n + 1) & 2) != 0;
  }
  const double ds = rm64_do_sin(p, dp, tab), dc = rm64_do_cos(q, dq, tab);
  double sv = swap ? dc : ds, cv = swap ? ds : dc;
  if (sign_from_x) sv = copysign(sv, x);
  if (neg_s) sv = -sv;
  if (neg_c) cv = -cv;
  if (k < 0x3e500000u) sv = x;
  if (k < 0x3e400000u) cv = 1.0;
  *s = sv; *c = cv;
}
#endif

// __tan (s_tan.c)
RM64_FN double ref_tan(double x) {
  const double* xfg = rm64_xfg;
  const double w = fabs(x);
  if (w <= RM64_G1) return x;
  if (w <= RM64_G2) {
    const double x2 = x * x;
    double t = fma(RM64_D11, x2, RM64_D9);
    t = fma(t, x2, RM64_D7);
    t = fma(t, x2, RM64_D5);
    t = fma(t, x2, RM64_D3);
    return fma(x * x2, t, x);
  }
  if (w <= RM64_G3) {
    const int i = (int)fma(256.0, w, -15.5);
    const double z = w - xfg[4 * i];
    const double z2 = z * z;
    const double pz = fma(z * z2, fma(z2, RM64_E1, RM64_E0), z);
    const double fi = xfg[4 * i + 1], gi = xfg[4 * i + 2];
    const double t2 = ((fi + gi) * pz) / (gi - pz);
    const double y = fi + t2;
    return (x < 0) ? -y : y;
  }
  return rm64_fallback_tan(x);
}

// __exp (sysdeps/ieee754/dbl-64/e_exp.c: the ARM Optimized Routines exp glibc adopted in 2.28, N = 128, degree-5 polynomial),
// again in the operation order of the FMA build.  Used by the exact-distance cost W2 * exp(-W3 * Dobs) (simulation.cpp:91).
#define RM64_EXP_INVLN2N 0x1.71547652b82fep+7
#define RM64_EXP_SHIFT 0x1.8p52
#define RM64_EXP_NEGLN2HIN -0x1.62e42fefa0000p-8
#define RM64_EXP_NEGLN2LON -0x1.cf79abc9e3b3ap-47
#define RM64_EXP_C2 0x1.ffffffffffdbdp-2
#define RM64_EXP_C3 0x1.555555555543cp-3
#define RM64_EXP_C4 0x1.55555cf172b91p-5
#define RM64_EXP_C5 0x1.1111167a4d017p-7
#ifdef REFMATH_HOST
static inline double rm64_from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline double rm64_fallback_exp(double x) { return exp(x); }
#else
__device__ __forceinline__ double rm64_from_bits(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __noinline__ double rm64_fallback_exp(double x) { return exp(x); }
#endif
RM64_FN double ref_exp(double x) {
  const uint64_t ix = (uint64_t)rm64_bits(x);
  uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {                     // |x| < 2^-54 or |x| >= 512 or not finite
    if (abstop < 0x3c9u) return 1.0 + x;             // tiny
    if (abstop >= 0x409u) {                          // |x| >= 1024, inf, NaN
      if (ix == 0xfff0000000000000ull) return 0.0;   // -inf
      if (abstop >= 0x7ffu) return 1.0 + x;          // +inf, NaN
      if (ix >> 63) return 0.0;                      // underflow (upstream: __math_uflow)
      return rm64_fallback_exp(x);                   // overflow
    }
    abstop = 0;                                      // 512 <= |x| < 1024: the result may be subnormal
  }
  const double kd0 = fma(x, RM64_EXP_INVLN2N, RM64_EXP_SHIFT);
  const uint64_t ki = (uint64_t)rm64_bits(kd0);
  const double kd = kd0 - RM64_EXP_SHIFT;
  double r = fma(kd, RM64_EXP_NEGLN2HIN, x);
  r = fma(kd, RM64_EXP_NEGLN2LON, r);
  const int idx = 2 * (int)(ki & 127u);
  const uint64_t top = ki << 45;
  const double tail = rm64_from_bits(rm64_exp_tab[idx]);
  uint64_t sbits = rm64_exp_tab[idx + 1] + top;
  const double r2 = r * r;
  const double p23 = fma(r, RM64_EXP_C3, RM64_EXP_C2);
  const double p45 = fma(r, RM64_EXP_C5, RM64_EXP_C4);
  const double tmp = fma(r2 * r2, p45, fma(p23, r2, r + tail));
  if (abstop == 0) {
    // specialcase (e_exp.c): only the k < 0 branch can occur for the negative arguments of the cost term
    if ((ki & 0x80000000u) == 0) return rm64_fallback_exp(x);
    sbits += 1022ull << 52;
    const double scale = rm64_from_bits(sbits);
    const double st = tmp * scale;
    double y = scale + st;
    if (y < 1.0) {
      const double hi = y + 1.0;
      double lo = (scale - y) + st;
      lo = ((1.0 - hi) + y) + lo;
      y = (lo + hi) - 1.0;
      if (y == 0.0) y = 0.0;
    }
    return 0x1p-1022 * y;
  }
  const double scale = rm64_from_bits(sbits);
  return fma(scale, tmp, scale);
}
