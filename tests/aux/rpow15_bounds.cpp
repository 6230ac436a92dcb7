#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
/* odx_rpow15_fast's device branch (csrc/odx_model_prelude.cuh) with the hardware seed replaced by a perturbed exact one:
 * y0 = x^(-1/2) (1 + d), |d| <= 2^-bits (rsqrt.approx.ftz.f64 delivers more than 20 good bits) */
static double rpow15_fast(double x, double y0) {
  const double t = y0 * y0;
  const double e = fma(-x, t, 1.0);
  const double c = t * y0;
  const double p = fma(e, 1.875, 1.5);
  return fma(p, c * e, c);
}
int main(int argc, char **argv) {
  const int bits = argc > 1 ? atoi(argv[1]) : 20;
  const long nsamp = argc > 2 ? atol(argv[2]) : 2000000;
  uint64_t st = 0x9E3779B97F4A7C15ULL;
  double emax = 0;
  for (long i = 0; i < nsamp; i++) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const double u = (double)(st >> 11) * (1.0 / 9007199254740992.0);
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const double v = (double)(st >> 11) * (1.0 / 9007199254740992.0);
    const double x = exp2(120.0 * u - 60.0);                              /* r^2 over 36 decades */
    const double d = (i % 4 == 0 ? (v < 0.5 ? -1.0 : 1.0) : 2.0 * v - 1.0) * exp2(-bits);
    const double y0 = (double)((1.0L / sqrtl((long double)x)) * (1.0L + (long double)d));
    const long double ref = 1.0L / ((long double)x * sqrtl((long double)x));
    const double got = rpow15_fast(x, y0);
    const double rel = fabs((double)(((long double)got - ref) / ref));
    if (rel > emax) emax = rel;
  }
  printf("%.3e\n", emax);
  return 0;
}
