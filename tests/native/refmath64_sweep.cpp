// refmath64.cuh (the double sin/cos/tan restatement the rollout kernels use) against the C library, bit for bit.
// TEST INFRASTRUCTURE.  usage: refmath64_sweep [millions of samples per range]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include "../../cl-rrt_b200/csrc/refmath64.cuh"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rng() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static inline double uni(double lo, double hi) { return lo + (hi - lo) * ((rng() >> 11) * (1.0 / 9007199254740992.0)); }
static inline bool same(double a, double b) { return rm64_bits(a) == rm64_bits(b) || (a != a && b != b); }

int main(int argc, char** argv) {
  const long n = (argc > 1 ? atol(argv[1]) : 4) * 1000000L;
  struct Range { double lo, hi; bool logscale; };
  const Range trig[] = {{1e-12, 1e-6, true}, {1e-6, 0.126, true}, {0.1, 0.86, false}, {0.85, 2.43, false}, {2.4, 10.0, false},
                        {10.0, 1000.0, false}, {1000.0, 1.0e8, true}};
  long bad_sin = 0, bad_cos = 0, bad_sc = 0, tot = 0;
  for (const Range& r : trig)
    for (long i = 0; i < n; i++) {
      double x = r.logscale ? exp(uni(log(r.lo), log(r.hi))) : uni(r.lo, r.hi);
      if (rng() & 1) x = -x;
      const double s = ref_sin(x), c = ref_cos(x);
      double s2, c2;
      ref_sincos(x, &s2, &c2);
      if (!same(s, sin(x))) { if (bad_sin++ < 5) printf("  sin(%a): %a vs libm %a\n", x, s, sin(x)); }
      if (!same(c, cos(x))) { if (bad_cos++ < 5) printf("  cos(%a): %a vs libm %a\n", x, c, cos(x)); }
      if (!same(s2, s) || !same(c2, c)) bad_sc++;
      tot++;
    }
  printf("sin: %ld mismatches of %ld\ncos: %ld mismatches of %ld\nsincos: %ld mismatches of %ld\n", bad_sin, tot, bad_cos, tot, bad_sc, tot);
  const Range tr[] = {{1e-12, 1.3e-8, true}, {1e-8, 0.0608, true}, {0.05, 0.07, false}, {0.06, 0.787, false}, {0.5, 0.53, false}};
  long bad_tan = 0, tt = 0;
  for (const Range& r : tr)
    for (long i = 0; i < n; i++) {
      double x = r.logscale ? exp(uni(log(r.lo), log(r.hi))) : uni(r.lo, r.hi);
      if (rng() & 1) x = -x;
      const double t = ref_tan(x);
      if (!same(t, tan(x))) { if (bad_tan++ < 5) printf("  tan(%a): %a vs libm %a\n", x, t, tan(x)); }
      tt++;
    }
  printf("tan: %ld mismatches of %ld\n", bad_tan, tt);
  // special values
  const double sp[] = {0.0, -0.0, 0.126, -0.126, 0.855469, 2.426265, 0x1.921fb54442d18p+0, 3.141592653589793, 0.52, -0.52, 0.0608, 0.787};
  long bad_sp = 0;
  for (double x : sp) bad_sp += !same(ref_sin(x), sin(x)) + !same(ref_cos(x), cos(x)) + (fabs(x) <= 0.787 ? !same(ref_tan(x), tan(x)) : 0);
  printf("special: %ld mismatches\n", bad_sp);
  // exp over the cost term's arguments (-W3 * Dobs <= 0) and a band of positive ones
  const Range er[] = {{1e-20, 1e-15, true}, {1e-15, 1e-3, true}, {1e-3, 1.0, true}, {1.0, 700.0, false}, {700.0, 1100.0, false}};
  long bad_exp = 0, te = 0;
  for (const Range& r : er)
    for (long i = 0; i < n; i++) {
      double x = -(r.logscale ? exp(uni(log(r.lo), log(r.hi))) : uni(r.lo, r.hi));
      if ((rng() & 7) == 0 && x > -500.0) x = -x;
      const double e = ref_exp(x);
      if (!same(e, exp(x))) { if (bad_exp++ < 5) printf("  exp(%a): %a vs libm %a\n", x, e, exp(x)); }
      te++;
    }
  const double se[] = {0.0, -0.0, -745.0, -745.2, -744.9, -708.4, -1023.9, -1024.0, -1e9, -INFINITY, 1e-300, -1e-300, -511.9999, -512.0};
  for (double x : se) bad_exp += !same(ref_exp(x), exp(x));
  printf("exp: %ld mismatches of %ld\n", bad_exp, te);
  return (bad_sin || bad_cos || bad_sc || bad_tan || bad_sp || bad_exp) ? 1 : 0;
}
