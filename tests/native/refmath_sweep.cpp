// Host sweep of cl-rrt_b200/csrc/refmath.cuh against the C library (glibc libm), bit for bit.
// usage: refmath_sweep <stride>   (stride 1 = every float of each domain)
// Built by tests/test_refmath.py with: g++ -O2 -mfma -ffp-contract=off -DREFMATH_HOST
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include "../../cl-rrt_b200/csrc/refmath.cuh"

static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static bool same(float a, float b) { return f2u(a) == f2u(b) || (a != a && b != b); }

int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
  long bad[6] = {0, 0, 0, 0, 0, 0}, n[6] = {0, 0, 0, 0, 0, 0};
  // sinf / cosf: every float with |x| < 120 (both signs)
  for (uint32_t sign = 0; sign < 2; sign++)
    for (uint32_t u = 0; u < 0x42f00000u; u += stride) {
      const float x = u2f(u | (sign << 31));
      n[0]++; if (!same(ref_sinf(x), sinf(x))) { if (bad[0]++ < 3) printf("sinf(%a): %a vs %a\n", x, ref_sinf(x), sinf(x)); }
      n[1]++; if (!same(ref_cosf(x), cosf(x))) { if (bad[1]++ < 3) printf("cosf(%a): %a vs %a\n", x, ref_cosf(x), cosf(x)); }
    }
  // acosf / asinf: every float in [-1, 1] plus a few outside
  for (uint32_t sign = 0; sign < 2; sign++)
    for (uint32_t u = 0; u <= 0x3f800100u; u += stride) {
      const float x = u2f(u | (sign << 31));
      n[2]++; if (!same(ref_acosf(x), acosf(x))) { if (bad[2]++ < 3) printf("acosf(%a): %a vs %a\n", x, ref_acosf(x), acosf(x)); }
      n[3]++; if (!same(ref_asinf(x), asinf(x))) { if (bad[3]++ < 3) printf("asinf(%a): %a vs %a\n", x, ref_asinf(x), asinf(x)); }
    }
  // atanf: every finite float
  for (uint32_t sign = 0; sign < 2; sign++)
    for (uint32_t u = 0; u < 0x7f800000u; u += stride) {
      const float x = u2f(u | (sign << 31));
      n[4]++; if (!same(ref_atanf(x), atanf(x))) { if (bad[4]++ < 3) printf("atanf(%a): %a vs %a\n", x, ref_atanf(x), atanf(x)); }
    }
  // atan2f: pseudo-random pairs over the magnitudes the planner sees, plus special values
  uint64_t st = 88172645463325252ull;
  const long pairs = 400000000l / stride + 1000;
  for (long i = 0; i < pairs; i++) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const float y = ((int32_t)(st & 0xffffffff)) * (200.0f / 2147483648.0f);
    const float x = ((int32_t)(st >> 32)) * (200.0f / 2147483648.0f);
    n[5]++; if (!same(ref_atan2f(y, x), atan2f(y, x))) { if (bad[5]++ < 3) printf("atan2f(%a,%a): %a vs %a\n", y, x, ref_atan2f(y, x), atan2f(y, x)); }
  }
  const float sp[] = {0.0f, -0.0f, 1.0f, -1.0f, INFINITY, -INFINITY, NAN, 1e-40f, 4.77f, 1e30f, -1e30f, 1e-30f};
  for (float y : sp) for (float x : sp) { n[5]++; if (!same(ref_atan2f(y, x), atan2f(y, x))) { if (bad[5]++ < 10) printf("atan2f(%a,%a): %a vs %a\n", y, x, ref_atan2f(y, x), atan2f(y, x)); } }
  const char* names[6] = {"sinf", "cosf", "acosf", "asinf", "atanf", "atan2f"};
  long total = 0;
  for (int i = 0; i < 6; i++) { printf("%s: %ld mismatches of %ld\n", names[i], bad[i], n[i]); total += bad[i]; }
  return total ? 1 : 0;
}
