/* odx_rng.h — random number sources of the population integrator (host + device).
 *
 * Two sources, replacing the reference's libc-rand() based src/rand_gen.c:9-72:
 *
 *  1. Philox4x32-10, counter based (Salmon et al., SC'11).  key = (seed_lo, seed_hi); counter =
 *     (particle id, draw index within the call site, stream index, domain tag).  Every draw of every
 *     particle is addressable, so the result does not depend on how particles are sharded or ordered.
 *     This is the production source (DWDT noise, tau-leap / replication draws, device ICs).
 *
 *  2. HostRand: the serial libc stream the reference consumes (glibc TYPE_3 additive feedback
 *     generator, r[i] = r[i-3] + r[i-31], output >> 1, srand() = Park-Miller fill + 310 discards),
 *     with rand01 / randN / poisson / exprand written against it exactly as the reference composes
 *     them.  Host only; used to draw bit-identical initial conditions in parity runs ("replay").
 */
#ifndef ODX_RNG_H
#define ODX_RNG_H
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define ODX_HD __host__ __device__ __forceinline__
#else
#define ODX_HD inline
#endif

#ifndef ODX_FAST_NORMAL
#define ODX_FAST_NORMAL 1
#endif

namespace odx {

enum : uint32_t { DOM_RHS = 0x52485300u, DOM_IC = 0x49430000u, DOM_LEAP = 0x4c454150u };

ODX_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

/* the same block function with the ten round keys given (rk[2r] = k0 + r W0, rk[2r + 1] = k1 + r W1): when they come from the kernel
 * arguments they are constant-bank operands of the XORs, and the twenty per-thread additions of the key schedule leave the draw */
ODX_HD void philox4x32_10_rk(uint32_t c[4], const uint32_t (&rk)[20]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ rk[2 * r], n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ rk[2 * r + 1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
}
ODX_HD void philox_round_keys(uint32_t k0, uint32_t k1, uint32_t (&rk)[20]) {
  for (int r = 0; r < 10; r++) { rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u; rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u; }
}

/* 53-bit uniform in the open interval (0,1) from two 32-bit words */
ODX_HD double u01_from_bits(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

struct PhiloxStream {
  uint32_t k0, k1;      /* key   = seed */
  uint32_t id;          /* c[0]  = particle id */
  uint32_t stream;      /* c[2]  = RHS-evaluation / leap / init index */
  uint32_t dom;         /* c[3]  = domain tag */
  uint32_t slot;        /* c[1]  = running draw index at this call site */
  uint32_t rk[20];      /* round keys, valid when has_rk (set from the kernel arguments by the device contexts) */
  int has_rk = 0;
  ODX_HD void draw4(uint32_t out[4]) {
    out[0] = id; out[1] = slot++; out[2] = stream; out[3] = dom;
    if (has_rk) philox4x32_10_rk(out, rk); else philox4x32_10(out, k0, k1);
  }
  ODX_HD double rand01() { uint32_t r[4]; draw4(r); return u01_from_bits(r[0], r[1]); }
  /* Box-Muller (cosine branch); one Philox block per normal keeps draws independently addressable */
  ODX_HD double randN(double mu, double sigma) {
    uint32_t r[4]; draw4(r);
    const double u1 = u01_from_bits(r[0], r[1]), u2 = u01_from_bits(r[2], r[3]);
#if defined(__CUDA_ARCH__) && ODX_FAST_NORMAL
    /* the same two uniforms; cospi needs no range reduction of 2 pi u2 (cos of it costs a Cody-Waite reduction on top) */
    return mu + sigma * (sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
#else
    return mu + sigma * (sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2));
#endif
  }
  ODX_HD double exprand(double r) { const double u = rand01(); return log(1.0 / (1.0 - u)) / r; }
  /* lambda < 30: Knuth's product method, as the reference composes it (src/rand_gen.c:37-46).  Above that the reference's form
   * is O(lambda) serial work and, from lambda ~ 745 on, silently wrong (exp(-lambda) underflows to 0 and the loop only ends when
   * the product itself underflows: the count saturates near 745).  The keyed stream has no bit-parity constraint with libc rand(),
   * so large means use Hoermann's transformed rejection (PTRS, 1993): exact in distribution, ~1.1 iterations of two uniforms. */
  ODX_HD long poisson(double lambda) {
    if (!(lambda >= 30.0)) {
      long k = 0; const double L = exp(-lambda); double p = 1.0;
      do { ++k; p *= rand01(); } while (p > L);
      return k - 1;
    }
    const double slam = sqrt(lambda), loglam = log(lambda);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
      const double U = rand01() - 0.5, V = rand01();
      const double us = 0.5 - fabs(U);
      const double kf = floor((2.0 * a / us + b) * U + lambda + 0.43);
      if (us >= 0.07 && V <= vr) return (long)kf;
      if (kf < 0.0 || (us < 0.013 && V > us)) continue;
      if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lambda + kf * loglam - lgamma(kf + 1.0)) return (long)kf;
    }
  }
};

/* Serial libc stream restated (used on the host only). */
struct HostRand {
  int32_t ring[31]; int f, r;
  ODX_HD void seed(uint32_t s) {
    if (s == 0) s = 1;
    int32_t word = (int32_t)s; ring[0] = word;
    for (int i = 1; i < 31; i++) {
      const long hi = word / 127773, lo = word % 127773;
      word = (int32_t)(16807 * lo - 2836 * hi);
      if (word < 0) word += 2147483647;
      ring[i] = word;
    }
    f = 3; r = 0;
    for (int i = 0; i < 310; i++) next();
  }
  ODX_HD int next() {
    const uint32_t v = (uint32_t)ring[f] + (uint32_t)ring[r];
    ring[f] = (int32_t)v;
    if (++f >= 31) f = 0;
    if (++r >= 31) r = 0;
    return (int)(v >> 1);
  }
  /* src/rand_gen.c:9-12 */
  ODX_HD double rand01() { return (double)next() / 2147483647; }
  /* src/rand_gen.c:14-26 (ratio of uniforms with quick accept / reject bounds) */
  ODX_HD double randN(double mu, double sigma) {
    double u, v, x, y, q;
    do {
      u = rand01();
      v = 1.7156 * (rand01() - 0.5);
      x = u - 0.449871;
      y = fabs(v) + 0.386595;
      q = x * x + y * (0.19600 * y - 0.25472 * x);
    } while (q > 0.27597 && (q > 0.27846 || v * v > -4.0 * log(u) * u * u));
    return mu + sigma * v / u;
  }
  /* src/rand_gen.c:37-46 */
  ODX_HD long poisson(double lambda) {
    long k = 0; const double L = exp(-lambda); double p = 1.0;
    do { ++k; p *= next() / (double)2147483647; } while (p > L);
    return k - 1;
  }
  /* src/rand_gen.c:68-72 */
  ODX_HD double exprand(double rr) { const double mu = rand01(); return log(1 / (1 - mu)) / rr; }
};

}  // namespace odx
#endif
