// refmath.cuh — float transcendental functions with the reference platform's results.
//
// The reference computes its Dubins metric and OBB geometry in float through glibc 2.39 libm (cosf, sinf,
// atan2f, acosf, asinf).  CUDA's float functions are 1-2 ulp routines with different results,
// and discrete decisions (candidate order, separating-axis sign) hang on the last bit.  The functions here are
// therefore evaluated in double and rounded once ("correctly rounded" up to double-rounding cases of
// probability ~2^-29 per call).  glibc's own float routines are faithful but not always correctly rounded;
// tests/test_nearest_gpu.py audits every disagreement and requires it to be a last-ulp key tie.
#pragma once
#include <math.h>

__device__ __forceinline__ void ref_sincosf(float a, float* s, float* c) {
  double sd, cd;
  sincos((double)a, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}
__device__ __forceinline__ float ref_sinf(float a) { return (float)sin((double)a); }
__device__ __forceinline__ float ref_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float ref_acosf(float x) { return (float)acos((double)x); }
__device__ __forceinline__ float ref_asinf(float x) { return (float)asin((double)x); }
