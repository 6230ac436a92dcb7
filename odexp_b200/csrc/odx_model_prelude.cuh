/* odx_model_prelude.cuh — helpers visible to generated model code (host + device). */
#pragma once
#include <math.h>
#include "odx_rng.h"

/* pow(x, c) with literal exponents (emitted by the generator's peephole; <= 1 ulp from pow) */
ODX_HD double odx_pow05(double x) { return sqrt(x); }
ODX_HD double odx_pow15(double x) { return x * sqrt(x); }
ODX_HD double odx_pow2(double x) { return x * x; }
ODX_HD double odx_pow3(double x) { return x * x * x; }

/* sin and cos of one argument with a single Cody-Waite range reduction and constant-bank polynomial coefficients
 * (fdlibm __kernel_sin / __kernel_cos minimax sets on [-pi/4, pi/4], error < 2^-58 before the final rounding).
 * nvcc does not merge sin(x) and cos(x) of the same x, and each libdevice call re-reduces the argument and loads its
 * coefficient table through LDG; the emitter therefore hoists every sin(A) / cos(A) pair into one odx_sincos(A).
 * Outside |x| < 105615 (where three-term reduction stops being exact) it defers to the library. */
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static void odx_sincos_library(double x, double *sp, double *cp) { sincos(x, sp, cp); }
/* coefficients live in the constant bank on the device: a 64-bit literal would be rebuilt from two 32-bit uniform
 * moves in front of every DFMA (three issue slots per term instead of one) */
#define ODX_SC_TABLE { 0.63661977236758134308, 0x1.921fb54442d18p+0, 0x1.1a62633145c00p-54, 0x1.b839a252049c0p-104, \
  1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04, \
  8.33333333332248946124e-03, -1.66666666666666324348e-01, \
  -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05, \
  -1.38888888888741095749e-03, 4.16666666666666019037e-02, \
  /* [16..] Taylor coefficients of odx_sincos_inc: sin d / d - 1 and cos d - 1 in z = d^2 */ \
  2.75573192239858906526e-06, -1.98412698412698412698e-04, 8.33333333333333333333e-03, -1.66666666666666666667e-01, \
  -2.75573192239858906526e-07, 2.48015873015873015873e-05, -1.38888888888888888889e-03, 4.16666666666666666667e-02, \
  /* [24..] table-assisted odx_sincos: 1024 / (2 pi), head and tail of 2 pi / 1024 */ \
  162.97466172610082, 0x1.921fb54442d18p-8, 0x1.1a62633145c07p-62 }
#if defined(__CUDACC__)
__constant__ double odx_sc_dev[27] = ODX_SC_TABLE;
#endif
static const double odx_sc_host[27] = ODX_SC_TABLE;
/* Table-assisted form for the persistent step kernels (contexts tagged by odx_ctx_sctab).  The argument is reduced to the nearest
 * multiple of 2 pi / 1024 with the same shifted-sum rounding and a two-term Cody-Waite product (the third term is < 6e-27 here),
 * sin / cos of that node come from a 16 KB shared-memory table the kernel fills once per launch (sincospi: the nodes are exact
 * binary fractions of pi), and the residual |r| <= pi / 1024 needs only r - r^3/6 + r^5/120 and 1 - r^2/2 + r^4/24 (truncation
 * < 1e-18): 14 FP64 operations and one 16-byte shared load instead of 26 FP64 + a dozen integer ones, and 6 instead of 16
 * polynomial coefficients to keep in registers.  Error <= ~2 ulp of 1 (table 0.5 ulp + combination), far inside the 1e-10 parity
 * bound; the quadrant logic is the table index. */
template <class C> struct odx_ctx_sctab { static constexpr bool value = false; };
#define ODX_SCT_N 1024
#if defined(__CUDACC__)
__shared__ double2 odx_sct_smem[ODX_SCT_N];          /* allocated only in kernels that reference it */
__device__ __forceinline__ void odx_sct_fill() {     /* whole CTA; the caller synchronises afterwards */
  for (int k = threadIdx.x; k < ODX_SCT_N; k += blockDim.x) {
    double sv, cv;
    sincospi((double)k * (2.0 / ODX_SCT_N), &sv, &cv);
    odx_sct_smem[k] = make_double2(sv, cv);
  }
}
#endif
/* EXACT = false: straight-line fast path for every argument; out-of-range arguments only raise c.slow, and the caller redoes
 * the particle with EXACT = true.  (A branch to the library inside the fast path splits the basic block and keeps the compiler
 * from interleaving the two particles a thread carries: measured 135 -> ~100 issued instructions per particle-stage.) */
template <bool EXACT, class C>
ODX_HD void odx_sincos(C &ctx, double x, double *sp, double *cp) {
#if defined(__CUDA_ARCH__)
  const double *T = odx_sc_dev;
#else
  const double *T = odx_sc_host;
#endif
  if (EXACT) { odx_sincos_library(x, sp, cp); return; }
#if defined(__CUDA_ARCH__)
  ctx.slow |= !((__double2hiint(x) & 0x7fffffff) < 0x40f9c8f0);      /* !(|x| < 105615.0) on the integer pipe: the bound's low word is 0 */
  if (odx_ctx_sctab<C>::value) {
    const double qs = fma(x, T[24], 6755399441055744.0);
    const int n = __double2loint(qs);
    const double q = qs - 6755399441055744.0;
    double r = fma(-q, T[25], x);                         /* 2 pi / 1024, head (exact difference) and tail */
    r = fma(-q, T[26], r);
    const double z = r * r;
    const double sr = fma(r * z, fma(z, T[18], T[19]), r);      /* sin r = r - r^3/6 + r^5/120 */
    const double cm = z * fma(z, T[23], -0.5);                  /* cos r - 1 = -r^2/2 + r^4/24 */
    const double2 nd = odx_sct_smem[n & (ODX_SCT_N - 1)];
    *sp = fma(nd.x, cm, fma(nd.y, sr, nd.x));
    *cp = fma(nd.y, cm, fma(-nd.x, sr, nd.y));
    return;
  }
#else
  ctx.slow |= !(fabs(x) < 105615.0);
#endif
  /* round-to-nearest by the 1.5 * 2^52 shift: the quadrant is the low word of the shifted sum (no FRND / F2I, whose FP64 forms
   * run at conversion-unit rate), valid far beyond the range bound above */
#if defined(__CUDA_ARCH__)
  const double qs = fma(x, T[0], 6755399441055744.0);
  const int n = __double2loint(qs);
  const double q = qs - 6755399441055744.0;
#else
  const double q = rint(x * T[0]);
  const int n = (int)q;
#endif
  double r = fma(-q, T[1], x);
  r = fma(-q, T[2], r);
  r = fma(-q, T[3], r);
  const double z = r * r;
  double ps = fma(z, T[4], T[5]);
  ps = fma(z, ps, T[6]);
  ps = fma(z, ps, T[7]);
  ps = fma(z, ps, T[8]);
  ps = fma(z, ps, T[9]);
  const double s = fma(r * z, ps, r);
  double pc = fma(z, T[10], T[11]);
  pc = fma(z, pc, T[12]);
  pc = fma(z, pc, T[13]);
  pc = fma(z, pc, T[14]);
  pc = fma(z, pc, T[15]);
  const double c = fma(z, fma(z, pc, -0.5), 1.0);
  /* quadrant: swap for odd n, flip signs through the sign bit (integer pipe, not DADD on the FP64 pipe) */
  const double a = (n & 1) ? c : s, b = (n & 1) ? s : c;
#if defined(__CUDA_ARCH__)
  *sp = __hiloint2double(__double2hiint(a) ^ ((n & 2) << 30), __double2loint(a));
  *cp = __hiloint2double(__double2hiint(b) ^ (((n + 1) & 2) << 30), __double2loint(b));
#else
  *sp = (n & 2) ? -a : a;
  *cp = ((n + 1) & 2) ? -b : b;
#endif
}

/* sin and cos of x from those of a nearby x0 (the previous stage state of the same particle): angle addition with short
 * Taylor polynomials in d = x - x0, |d| < 1/8 (truncation < 3e-18 relative), written so that the known values are only
 * corrected (s = s0 + c0 sin d + s0 (cos d - 1)): no range reduction, no quadrant logic, 17 FP64 operations instead of 26 plus
 * a dozen integer ones.  A larger step raises ctx.slow and the engine redoes the particle with the library path. */
template <bool EXACT, class C>
ODX_HD void odx_sincos_inc(C &ctx, double x, double x0, double s0, double c0, double *sp, double *cp) {
  if (EXACT) { odx_sincos_library(x, sp, cp); return; }
#if defined(__CUDA_ARCH__)
  const double *T = odx_sc_dev;
#else
  const double *T = odx_sc_host;
#endif
  const double d = x - x0;
#if defined(__CUDA_ARCH__)
  ctx.slow |= !((__double2hiint(d) & 0x7fffffff) < 0x3fc00000);      /* !(|d| < 0.125), integer pipe */
#else
  ctx.slow |= !(fabs(d) < 0.125);
#endif
  const double z = d * d;
  double ps = fma(z, T[16], T[17]);
  ps = fma(z, ps, T[18]);
  ps = fma(z, ps, T[19]);
  const double sd = fma(d * z, ps, d);                       /* sin d */
  double pc = fma(z, T[20], T[21]);
  pc = fma(z, pc, T[22]);
  pc = fma(z, pc, T[23]);
  pc = fma(z, pc, -0.5);
  const double cm = z * pc;                                  /* cos d - 1 */
  *sp = fma(s0, cm, fma(c0, sd, s0));
  *cp = fma(c0, cm, fma(-s0, sd, c0));
}

/* x^(-3/2) and x^(-1/2) emitted for 1 / pow(x, 1.5) and 1 / sqrt(x) */
ODX_HD double odx_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
ODX_HD double odx_rpow15(double x) { const double r = odx_rsqrt(x); return r * r * r; }

/* x^(-3/2) for the guarded product below: one MUFU.RSQ64H seed y0 (>= 20 good bits), then x^(-3/2) = y0^3 (1 - e)^(-3/2) with
 * e = 1 - x y0^2 expanded to second order (truncation 2.2 e^3 < 2e-18): 6 FP64 operations instead of the 7 of rsqrt(x)^3, no
 * special-case branch.  x = 0, subnormal or huge gives NaN / inf / 0; every non-finite value sends the j-tile through the exact
 * forms (see odx_zmul), which is why this form is only emitted as odx_zmul's third argument. */
template <bool EXACT>
ODX_HD double odx_rpow15_fast(double x) {
#if defined(__CUDA_ARCH__)
  if (EXACT) return odx_rpow15(x);
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double t = y0 * y0;
  const double e = fma(-x, t, 1.0);
  const double c = t * y0;
  const double p = fma(e, 1.875, 1.5);
  return fma(p, c * e, c);
#else
  return odx_rpow15(x);
#endif
}

/* (a == b) ? 0.0 : (a - b) * r, emitted for that pattern inside %C terms (the per-component guard of a 1/r^n kernel).
 * Fast path: for a finite r the product is already +-0 when a == b, so neither the FP64 compare nor the select is needed and
 * the product fuses into the accumulation.  A non-finite r (the self term: r = 1/0) makes the product inf or NaN, which
 * poisons the running sum it is added to: the pair kernel checks its sums after every j-tile, and on a non-finite sum
 * restores the values from before the tile and repeats the tile with the exact form (OdxModel::CPL_POISON; once per
 * i-particle for a self term; coincident or overflowing particles, where the reference itself produces inf / NaN, take the
 * same route).  With a finite r, a = b = +-inf gives NaN instead of the reference's 0. */
template <bool EXACT, class C>
ODX_HD double odx_zmul(C &ctx, double a, double b, double r) {
  (void)ctx;
#if defined(__CUDA_ARCH__)
  if (!EXACT) return (a - b) * r;
#endif
  return (a == b) ? 0.0 : (a - b) * r;
}

/* reciprocal emitted for `x / f(...)` (a pure call the compiler can share between terms) */
ODX_HD double odx_rcp(double x) { return 1.0 / x; }

/* branch-free conditional emitted for `c ? a : b` in model functions (both arms are pure) */
ODX_HD double odx_sel(bool c, double a, double b) { return c ? a : b; }

/* STOPIF(cond): raise the stop flag, value is the condition (src/odexp.h:400-406) */
template <class C> ODX_HD int odx_stopif(C &c, int cond) { if (cond) c.stop = 1; return cond; }

/* interp(x, y, xi, nx, start) on user arrays (src/utils_odexp.c:173-203): piecewise-linear interpolation at xi in the table (x, y) whose
 * abscissae increase circularly from `start`; dichotomy search with the reference's index arithmetic (the same code serves delay() through
 * DevCtx::delay_lookup).  The reference prints a message when xi lies in front of the table and goes on: so does this, without the message. */
ODX_HD double interp(const double *x, const double *y, double xi, int nx, int start) {
  int i1 = 0, i9 = nx - 1;
  while (i9 > i1 + 1) {
    const int i5 = (i1 + i9 + 1) / 2;
    if (x[(start + i5) % nx] > xi) i9 = i5; else i1 = i5;
  }
  if (i1 == i9) return y[(i1 + start) % nx];
  return y[(i1 + start) % nx] + (y[(i9 + start) % nx] - y[(i1 + start) % nx]) * (xi - x[(i1 + start) % nx]) / (x[(i9 + start) % nx] - x[(i1 + start) % nx]);
}

/* device twins of the pure helpers a model expression may call (src/utils_odexp.c:11-107) */
ODX_HD double sum(const double *array, long len) { double s = 0.0; for (long i = 0; i < len; i++) s += array[i]; return s; }
ODX_HD double sumsub(const double *array, const long *ind, long len)      /* const: a `decl const long idx[]` array is the usual argument */ { double s = 0.0; for (long i = 0; i < len; i++) s += array[ind[i]]; return s; }
ODX_HD double sumstep(const double *array, long len, long step) { double s = 0.0; for (long i = 0; i < len; i += step) s += array[i]; return s; }
ODX_HD double prod(const double *array, long len) { double s = 1.0; for (long i = 0; i < len; i++) s *= array[i]; return s; }
ODX_HD double dotprod(const double *x, const double *y, long len) { double s = 0.0; for (long i = 0; i < len; i++) s += x[i] * y[i]; return s; }
ODX_HD double minus(double x, double y) { return x - y; }
ODX_HD double plus(double x, double y) { return x + y; }
/* one link of a linear chain delay (src/utils_odexp.c:162-167); LAG lines expand into an array of these */
ODX_HD double linchaindelay(const double root, const double *chain, const int link, const double delay, const int len) {
  const double beta = (double)len / delay;
  return beta * ((link == 0 ? root : chain[link - 1]) - chain[link]);
}
ODX_HD double identity(double x) { return x; }
/* sum over all but one dimension of a 3-index array, as coded (src/utils_odexp.c:38-60: offsets i + d2*j + d3*k) */
ODX_HD double sum3(const double *array, long d1, long d2, long d3, long dim) {
  double s = 0.0; long n[3] = {d1, d2, d3};
  n[dim - 1] = 1;
  for (long i = 0; i < n[0]; i++) for (long j = 0; j < n[1]; j++) for (long k = 0; k < n[2]; k++) s += array[i + d2 * j + d3 * k];
  return s;
}
ODX_HD double conv(const double *u, const double *v, long len) { double s = 0.0; for (long i = 0; i < len; i++) s += u[len - 1 - i] * v[i]; return s; }
/* sumxy / kern take model functions as arguments (src/utils_odexp.c:122-152): templates, so that the calls inline on the device */
template <class F, class G> ODX_HD double sumxy(long len, F f, G g, const double *x, const double yi) {
  double s = 0.0; for (long j = 0; j < len; j++) s += f(g(x[j], yi)); return s;
}
template <class F> ODX_HD double kern(const double *Wi, F f, double xi, const double *x, double *p, long len) {
  double s = 0.0; for (long j = 0; j < len; j++) s += Wi[j] * f(xi, x[j], p); return s;
}
