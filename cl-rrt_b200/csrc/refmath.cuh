// refmath.cuh — float transcendental functions with bit-for-bit the results of the reference platform.
//
// PROVENANCE / LICENCE NOTE.  The GNU C Library is distributed under the LGPL v2.1 or later; its sinf / cosf / exp come from
// ARM Optimized Routines (MIT OR Apache-2.0 WITH LLVM-exception), atanf / acosf / asinf from FreeBSD msun (Sun Microsystems'
// permissive notice), sin / cos / tan from the IBM Accurate Mathematical Library (LGPL).  Nothing was copied from the
// reference repository (it contains no libm).  The routines below were written from the published algorithms against the
// behaviour of the installed library (tests/test_refmath.py sweeps them against it); the numeric tables are constants
// extracted from the installed libm by scripts/gen_refmath64_tables.py.  A redistributor who treats restated algorithms
// and extracted tables as derived work of glibc should ship these three files under the LGPL.
//
// The reference computes its Dubins metric (rrt/src/rrtplanner.cpp:371-406) and OBB geometry
// (rrt/src/old_collisioncheck.cpp:56-65) in float through the C library: cosf, sinf, atan2f, acosf, asinf of
// glibc 2.39 libm on x86-64.  Candidate order and separating-axis signs hang on the last bit of
// those results, and CUDA's own float routines (1-2 ulp) differ from glibc's in 1-16 % of calls.  This header
// therefore restates glibc's published algorithms operation by operation:
//   * sinf / cosf: the ARM Optimized Routines implementation glibc adopted in 2.28 (sysdeps/ieee754/flt-32/
//     s_sinf.c, s_cosf.c, sincosf.h, sincosf_table.c): double-precision reduction by pi/2 and two degree-7/8
//     polynomials.  On x86-64 CPUs with FMA glibc's ifunc dispatch selects the build of the same source compiled
//     with -mfma, in which every `a + b*c` of the polynomial is one fused multiply-add; that variant is what runs
//     on every host a B200 can sit in, and is what is restated here (explicit fma()).
//   * atanf / atan2f / acosf / asinf: the fdlibm-derived single-precision routines (sysdeps/ieee754/flt-32/
//     s_atanf.c, e_atan2f.c, e_acosf.c, e_asinf.c): plain float arithmetic, no ifunc variants, no contraction.
// Constants are the published ones (checked against the installed libm).  tests/test_refmath.py compiles this
// header for the host (REFMATH_HOST) and sweeps it against the C library: every float in the domains used here.
// Arguments outside the restated ranges (|x| >= 120 for sinf/cosf) fall back to double evaluation, rounded once.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef REFMATH_HOST
#define RM_FN static inline
static inline uint32_t rm_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float rm_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float rm_sqrtf(float x) { return sqrtf(x); }
static inline float rm_divf(float a, float b) { return a / b; }
static inline int rm_d2i_rz(double d) { return (int)d; }
#else
#define RM_FN __device__ __forceinline__
__device__ __forceinline__ uint32_t rm_f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float rm_u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ float rm_sqrtf(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ float rm_divf(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ int rm_d2i_rz(double d) { return __double2int_rz(d); }
#endif

// ---- sinf / cosf (ARM Optimized Routines, FMA build) -------------------------------------------------------
// polynomial coefficients of __sincosf_table[0]; table[1] (quadrants 2,3) holds the negated cosine set
#define RM_HPI_INV 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define RM_HPI 0x1.921FB54442D18p0
#define RM_C0 0x1p0
#define RM_C1 -0x1.ffffffd0c621cp-2
#define RM_C2 0x1.55553e1068f19p-5
#define RM_C3 -0x1.6c087e89a359dp-10
#define RM_C4 0x1.99343027bf8c3p-16
#define RM_S1 -0x1.555545995a603p-3
#define RM_S2 0x1.1107605230bc4p-7
#define RM_S3 -0x1.994eb3774cf24p-13

// sinf_poly (sincosf.h): n even -> sine polynomial of x, n odd -> cosine polynomial; neg selects table[1]
RM_FN float rm_sinf_poly(double x, double x2, int n, bool neg) {
  if ((n & 1) == 0) {
    const double x3 = x * x2;
    const double s1 = fma(x2, RM_S3, RM_S2);
    const double x7 = x3 * x2;
    const double s = fma(x3, RM_S1, x);
    return (float)fma(x7, s1, s);
  } else {
    const double sg = neg ? -1.0 : 1.0;  // table[1]: c0..c4 negated (exact sign flips)
    const double x4 = x2 * x2;
    const double c2 = fma(x2, sg * RM_C4, sg * RM_C3);
    const double c1 = fma(x2, sg * RM_C1, sg * RM_C0);
    const double x6 = x4 * x2;
    const double c = fma(x4, sg * RM_C2, c1);
    return (float)fma(x6, c2, c);
  }
}
// reduce_fast (sincosf.h): quadrant from a scaled truncating conversion, one fused subtraction
RM_FN double rm_reduce_fast(double x, int* np) {
  const double r = x * RM_HPI_INV;
  const int n = (rm_d2i_rz(r) + 0x800000) >> 24;
  *np = n;
  return fma(-(double)n, RM_HPI, x);
}
RM_FN float ref_sinf(float y) {
  const uint32_t top = (rm_f2u(y) >> 20) & 0x7ff;
  const double x = (double)y;
  if (top < 0x3f4) {  // |y| < pi/4
    if (top < 0x398) return y;  // |y| < 2^-12
    return rm_sinf_poly(x, x * x, 0, false);
  } else if (top < 0x42f) {  // |y| < 120
    int n;
    const double xr = rm_reduce_fast(x, &n);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign[n & 3] = {1,-1,-1,1}
    // table[1] differs from table[0] only in the cosine coefficients; the sine set is shared
    return rm_sinf_poly(xr * s, xr * xr, n, (n & 2) != 0);
  }
  return (float)sin(x);
}
RM_FN float ref_cosf(float y) {
  const uint32_t top = (rm_f2u(y) >> 20) & 0x7ff;
  const double x = (double)y;
  if (top < 0x3f4) {
    if (top < 0x398) return 1.0f;
    return rm_sinf_poly(x, x * x, 1, false);
  } else if (top < 0x42f) {
    int n;
    const double xr = rm_reduce_fast(x, &n);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return rm_sinf_poly(xr * s, xr * xr, n ^ 1, (n & 2) != 0);
  }
  return (float)cos(x);
}
RM_FN void ref_sincosf(float a, float* s, float* c) {
  *s = ref_sinf(a);
  *c = ref_cosf(a);
}

// ---- atanf (fdlibm s_atanf.c) ----------------------------------------------------------------------------------
RM_FN float ref_atanf(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  const uint32_t hx = rm_f2u(x), ix = hx & 0x7fffffffu;
  int id;
  if (ix >= 0x4c000000u) {  // |x| >= 2^25
    if (ix > 0x7f800000u) return x + x;
    return ((int32_t)hx > 0) ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000u) {  // |x| < 0.4375
    if (ix < 0x31000000u) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000u) {
      if (ix < 0x3f300000u) { id = 0; x = rm_divf(2.0f * x - 1.0f, 2.0f + x); }
      else { id = 1; x = rm_divf(x - 1.0f, x + 1.0f); }
    } else {
      if (ix < 0x401c0000u) { id = 2; x = rm_divf(x - 1.5f, 1.0f + 1.5f * x); }
      else { id = 3; x = rm_divf(-1.0f, x); }
    }
  }
  const float z = x * x;
  const float w = z * z;
  const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) return x - x * (s1 + s2);
  const float zz = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return ((int32_t)hx < 0) ? -zz : zz;
}

// ---- atan2f (fdlibm e_atan2f.c) ---------------------------------------------------------------------------------
RM_FN float ref_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = (int32_t)rm_f2u(x), hy = (int32_t)rm_f2u(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return ref_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0:
      case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = ref_atanf(fabsf(rm_divf(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return rm_u2f(rm_f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

// ---- acosf (fdlibm e_acosf.c) -----------------------------------------------------------------------------------
RM_FN float ref_acosf(float x) {
  const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
              pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
              pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
              qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
  const int32_t hx = (int32_t)rm_f2u(x), ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) {
    if (hx > 0) return 0.0f;
    return pi + 2.0f * pio2_lo;
  } else if (ix > 0x3f800000) {
    return rm_divf(x - x, x - x);
  }
  if (ix < 0x3f000000) {  // |x| < 0.5
    if (ix <= 0x32800000) return pio2_hi + pio2_lo;
    const float z = x * x;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = rm_divf(p, q);
    return pio2_hi - (x - (pio2_lo - r * x));
  } else if (hx < 0) {  // x < -0.5
    const float z = (one + x) * 0.5f;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float s = rm_sqrtf(z);
    const float r = rm_divf(p, q);
    const float w = r * s - pio2_lo;
    return pi - 2.0f * (s + w);
  } else {  // x > 0.5
    const float z = (one - x) * 0.5f;
    const float s = rm_sqrtf(z);
    const float df = rm_u2f(rm_f2u(s) & 0xfffff000u);
    const float c = rm_divf(z - df * df, s + df);
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = rm_divf(p, q);
    const float w = r * s + c;
    return 2.0f * (df + w);
  }
}

// ---- asinf (glibc e_asinf.c: fdlibm structure with a degree-4 polynomial) ---------------------------------------
RM_FN float ref_asinf(float x) {
  // bit patterns of: pio2_hi 1.57079637050628662109375, pio2_lo -4.37113900018624283e-8, pio4_hi 0.785398185253143310546875,
  // p0 1.666675248e-1, p1 7.495297643e-2, p2 4.547037598e-2, p3 2.417951451e-2, p4 4.216630880e-2
  const float one = 1.0f, pio2_hi = rm_u2f(0x3fc90fdbu), pio2_lo = rm_u2f(0xb33bbd2eu), pio4_hi = rm_u2f(0x3f490fdbu),
              p0 = rm_u2f(0x3e2aaae4u), p1 = rm_u2f(0x3d9980f2u), p2 = rm_u2f(0x3d3a3f25u), p3 = rm_u2f(0x3cc6141eu),
              p4 = rm_u2f(0x3d2cb694u);
  const int32_t hx = (int32_t)rm_f2u(x), ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return x * pio2_hi + x * pio2_lo;
  else if (ix > 0x3f800000) return rm_divf(x - x, x - x);
  else if (ix < 0x3f000000) {  // |x| < 0.5
    if (ix < 0x32000000) return x;  // |x| < 2^-27
    const float t = x * x;
    const float w = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
    return x + x * w;
  }
  float w = one - fabsf(x);
  float t = w * 0.5f;
  float p = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
  const float s = rm_sqrtf(t);
  if (ix >= 0x3f79999a) {  // |x| > 0.975
    t = pio2_hi - (2.0f * (s + s * p) - pio2_lo);
  } else {
    w = rm_u2f(rm_f2u(s) & 0xfffff000u);
    const float c = rm_divf(t - w * w, s + w);
    const float r = p;
    p = 2.0f * s * r - (pio2_lo - 2.0f * c);
    const float q = pio4_hi - 2.0f * w;
    t = pio4_hi - (p - q);
  }
  return (hx > 0) ? t : -t;
}
