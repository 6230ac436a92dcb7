/* odx_engine.cuh — B200 (sm_100a) population engine, instantiated once per generated model.
 *
 * Included at the end of a generated model.cu, after `struct OdxModel` (per-particle RHS, %M terms,
 * %C terms, rates, initial conditions as inlinable __host__ __device__ functions with compile-time
 * dimensions).  Everything here is templated on that struct, so the model code is inlined into the
 * stage kernels and all column indices are compile-time constants (the reference resolves every
 * MY()/MU()/... by a run-time strncmp search, src/odexp.h:218-364).
 *
 * Data layout (HBM): structure of arrays, FP64, one column per variable, column stride = capacity:
 *   Y[nvar][cap]   state at the start of the step          (reference: y[] AoS + p->y, methods_odexp.c:290)
 *   YS[nvar][cap]  current stage state                      (reference: k[] in dde_apply, ytmp in GSL)
 *   ACC[nvar][cap] running sum of weighted stage slopes     (chain methods: fe, dde)
 *   KS[7][nvar][cap] stage slopes k1..k6 + dydt_out         (embedded pairs rkf45 / rkck)
 *   E[nexpr][cap], AUX[naux][cap], PSI[npsi][cap], DEATH[cap], REPLI[cap], ID[cap] (int32)
 *
 * One RHS evaluation of the reference (generated pop_rhs, src/odexp:801-858) is
 *   pass 1 (aux with zeroed mean fields) -> one list walk per %M -> pass 2 (N^2 %C walk + ode_rhs).
 * Here a mean-field barrier is one grid-wide reduction, so a Runge-Kutta stage s is split at the
 * barrier: kernel s finishes stage s-1 (k_{s-1} = ode_rhs(y_{s-1}, mfd_{s-1})), forms y_s and
 * accumulates the %M terms of y_s (warp shuffle + block partials + last-block fixed-order sum).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <string>
#include <vector>
#include <algorithm>
#include "odx.h"
#include "odx_rng.h"

#ifdef ODX_WITH_NCCL
#include <nccl.h>
#endif

namespace odx {

typedef OdxModel M;
constexpr int NVAR = M::NVAR, NEXPR = M::NEXPR, NAUX = M::NAUX, NPSI = M::NPSI, NMFD = M::NMFD, NPAR = M::NPAR;
constexpr int A1(int n) { return n > 0 ? n : 1; }      /* array extent that is never zero */
constexpr int NRED = NMFD + 2;                          /* [%M sums..., %BIRTH sum, sum(death+repli)] */
constexpr int RED_BIRTH = NMFD, RED_RATES = NMFD + 1;
constexpr int BLOCK = 256;
constexpr int MAXSTAGE = 7;
#ifndef ODX_NO_SCTAB
#define ODX_NO_SCTAB 0
#endif
constexpr bool SCTAB_MODEL = (M::NHINT > 0) && !ODX_NO_SCTAB;      /* table-assisted sincos in the persistent step kernels (odx_model_prelude.cuh) */
#ifndef STAGE_LOAD_CS
#define STAGE_LOAD_CS 0
#endif
#ifndef STAGE_STORE_CS
#define STAGE_STORE_CS 1   /* streaming stores: neutral at N=2^24, +5% at N=2^22 (measured) */
#endif
#ifndef CHAIN_MIN_BLOCKS
#define CHAIN_MIN_BLOCKS 4
#endif

/* ------------------------------------------------------------------------------------------------
 * evaluation context handed to the generated model code
 * ---------------------------------------------------------------------------------------------- */
/* history stacks of one delay() call site (src/odexp.h:151-162: DELAY_SSTACK samples of (t, x) and the write position), one per
 * particle that the site indexes.  Sample s of stack q lives at [s * K + q]: the particles of a warp advance in lock step, so their
 * pushes and most probes of their searches touch consecutive addresses. */
constexpr int DELAY_SSTACK = 1000;
struct DelayView { double *T, *X; int *pos; long K; };

struct DevCtx {
  const DelayView *dl;      /* [NDELAY] */
  /* interp() on the circular stack (src/utils_odexp.c:173-203), index arithmetic as coded: start = pos + 1, the samples are read at
   * (start + i) % DELAY_SSTACK.  idx outside the declared array: the reference reads foreign memory; here NaN. */
  ODX_HD double delay_lookup(int k, double xi, long idx) const {
#if defined(__CUDA_ARCH__)
    const DelayView v = dl[k];
    if (idx < 0 || idx >= v.K) return __longlong_as_double(0x7ff8000000000000LL);
    const int nx = DELAY_SSTACK, start = v.pos[idx] + 1;
    int i1 = 0, i9 = nx - 1;
    while (i9 > i1 + 1) {
      const int i5 = (i1 + i9 + 1) / 2;
      if (v.T[(long)((start + i5) % nx) * v.K + idx] > xi) i9 = i5; else i1 = i5;
    }
    const double x1 = v.T[(long)((i1 + start) % nx) * v.K + idx], y1 = v.X[(long)((i1 + start) % nx) * v.K + idx];
    if (i1 == i9) return y1;
    const double x9 = v.T[(long)((i9 + start) % nx) * v.K + idx], y9 = v.X[(long)((i9 + start) % nx) * v.K + idx];
    return y1 + (y9 - y1) * (xi - x1) / (x9 - x1);
#else
    (void)k; (void)xi; (void)idx; return 0.0;
#endif
  }
  const double *mu;
  double pop_size;          /* POP_SIZE */
  double h;                 /* *SIM->h (DWDT = randN(0,1)/sqrt(h), src/odexp.h:384) */
  PhiloxStream rng;
  int32_t id, oid;          /* ID / OID */
  int atbirth, atrepli;     /* ATBIRTH / ATREPLI (src/odexp.h:366-367) */
  int ist0;                 /* IST0: initial conditions of the very first population (EVENT_TYPE == EVENT_T0) */
  const double *sis_y, *sis_e, *sis_aux;   /* sister particle (valid only during replication ICs) */
  int has_sister;
  int stop;                 /* STOPIF */
  int slow;                 /* a fast-path sincos saw an out-of-range argument: redo this particle with the library path */
  /* KY(name, degree) (src/odexp.h:334-357): variable `name` of the degree-th neighbour in list order, the list closed to a ring;
   * the neighbour's value is the state the current RHS evaluation is made at (p->y after pass 1 of pop_rhs) */
  const double *ky_base; int64_t ky_i, ky_n, ky_stride;
  ODX_HD double ky(int v, long d) const {
    int64_t j = (ky_i + d) % ky_n;
    if (j < 0) j += ky_n;
    return ky_base[(int64_t)v * ky_stride + j];
  }
  ODX_HD double rand01() { return rng.rand01(); }
  ODX_HD double randN(double m, double s) { return rng.randN(m, s); }
  ODX_HD double exprand(double r) { return rng.exprand(r); }
  ODX_HD long poisson(double l) { return rng.poisson(l); }
#if ODX_FAST_NORMAL
  /* DWDT = randN(0,1)/sqrt(*SIM->h) (src/odexp.h:384) as a product with 1/sqrt(h): h is uniform over the launch, so the square root and
   * the reciprocal leave the particle loop and no FP64 division is left per draw (<= 1 ulp from the quotient) */
  ODX_HD double dwdt() { return rng.randN(0.0, 1.0) * (1.0 / sqrt(h)); }
#else
  ODX_HD double dwdt() { return rng.randN(0.0, 1.0) / sqrt(h); }
#endif
};

/* the same context inside the persistent step kernels, which hold the sincos node table in shared memory (odx_model_prelude.cuh):
 * the tag selects the table-assisted odx_sincos at compile time */
struct DevCtxT : DevCtx {};
}  // namespace odx
template <> struct odx_ctx_sctab<odx::DevCtxT> { static constexpr bool value = true; };
namespace odx {
template <bool TAB> struct CtxSel { typedef DevCtx type; };
template <> struct CtxSel<true> { typedef DevCtxT type; };

/* host context for replayed initial conditions: draws through the serial libc stream */
struct HostCtx {
  const double *mu; double pop_size; double h; HostRand *hr; int32_t id, oid; int atbirth, atrepli, ist0;
  const double *sis_y, *sis_e, *sis_aux; int has_sister; int stop; int slow;
  ODX_HD double ky(int, long) const { return 0.0; }      /* not reachable: KY is rejected inside initial conditions */
  ODX_HD double delay_lookup(int, double, long) const { return 0.0; }      /* not reachable: delay() is rejected there too */
  ODX_HD double rand01() { return hr->rand01(); }
  ODX_HD double randN(double m, double s) { return hr->randN(m, s); }
  ODX_HD double exprand(double r) { return hr->exprand(r); }
  ODX_HD long poisson(double l) { return hr->poisson(l); }
  ODX_HD double dwdt() { return hr->randN(0.0, 1.0) / sqrt(h); }
};

/* device view of the population */
constexpr int MUV_MAX = 48;                                   /* parameters carried by value in the kernel arguments */
constexpr int MUV_N = (NPAR >= 1 && NPAR <= MUV_MAX) ? NPAR : 1;
struct Pop {
  double *Y, *YS, *ACC, *E, *AUX, *PSI, *DEATH, *REPLI, *KS;
  double *Y_OUT;            /* where the last stage stores the new state: Y (the other buffer for Euler / iteration with KY()) */
  double *YS_OUT;           /* where a stage stores the stage state it forms: YS, except for models with KY(), whose stages
                               alternate between two buffers because neighbours still read the previous stage state */
  int32_t *ID;
  int64_t stride;           /* column stride = capacity */
  int64_t n;                /* particles on this GPU */
  double n_global;          /* POP_SIZE across all GPUs */
  const double *mu;
  double muv[MUV_N];        /* the parameters by value: kernel-argument (constant bank) operands instead of a load per use */
  uint32_t k0, k1;          /* Philox key */
  uint32_t rk[20];          /* its ten round keys (philox_round_keys): constant-bank operands of every draw */
  int *stop;                /* sticky STOPIF flag of the whole population (models with STOPIF only) */
  DelayView dl[A1(M::NDELAY)];      /* history stacks of the delay() call sites */
};

struct Red {                /* grid reduction scratch */
  double *partial;          /* [grid][NRED] */
  unsigned int *counter;
  double *sums;             /* [NRED] raw sums (divided by POP_SIZE by the consumer) */
};

/* ------------------------------------------------------------------------------------------------
 * helpers
 * ---------------------------------------------------------------------------------------------- */
template <int NC>
__device__ __forceinline__ void load_cols(const double *__restrict__ base, int64_t stride, int64_t i, double (&out)[A1(NC)]) {
#pragma unroll
  for (int c = 0; c < NC; c++) out[c] = base[(int64_t)c * stride + i];
}
template <int NC>
__device__ __forceinline__ void store_cols(double *__restrict__ base, int64_t stride, int64_t i, const double (&v)[A1(NC)]) {
#pragma unroll
  for (int c = 0; c < NC; c++) base[(int64_t)c * stride + i] = v[c];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

/* Block-level sum of red[0..NRED) followed by a deterministic grid-level sum: every block writes its
 * partials, the last block to arrive adds them in a fixed order (independent of arrival order). */
template <int NR>
__device__ __forceinline__ void grid_reduce(double (&red)[NR], const Red &r) {
  __shared__ double s_part[BLOCK / 32][NR];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NR; k++) {
    const double v = warp_sum(red[k]);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NR) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) v += s_part[w][threadIdx.x];
    r.partial[(size_t)blockIdx.x * NR + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(r.counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double acc[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) acc[k] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLOCK) {
#pragma unroll
      for (int k = 0; k < NR; k++) acc[k] += __ldcg(&r.partial[(size_t)b * NR + k]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NR; k++) {
      const double v = warp_sum(acc[k]);
      if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < BLOCK / 32; w++) v += s_part[w][threadIdx.x];
      r.sums[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *r.counter = 0u;
  }
}

/* mean fields from raw sums: SIM->meanfield[k] /= POP_SIZE (src/odexp:897) */
__device__ __forceinline__ void load_mfd(const Pop &p, const double *__restrict__ sums, double n_global, double (&mfd)[A1(NMFD)]) {
#pragma unroll
  for (int k = 0; k < NMFD; k++) mfd[k] = (sums != nullptr) ? sums[k] / n_global : 0.0;
  if (M::HAS_VECT_MFD && sums != nullptr) M::mfd_vect((NPAR >= 1 && NPAR <= MUV_MAX) ? p.muv : p.mu, n_global, mfd);      /* %M ... {vect} */
}

__device__ __forceinline__ DevCtx make_ctx(const Pop &p, double h, uint32_t stream, uint32_t dom, int32_t id) {
  DevCtx c;
  c.mu = (NPAR >= 1 && NPAR <= MUV_MAX) ? p.muv : p.mu; c.pop_size = p.n_global; c.h = h;
  /* round keys from the kernel arguments only where the RHS draws (DWDT, rand01 ... in AUX / D/DT / rates): in a model without them the
   * twenty words would only add to the register pressure of the step kernels (Kuramoto ring kernel: 132 -> 332 bytes of spill loads) */
  if (M::RHS_RNG) {
#pragma unroll
    for (int r = 0; r < 20; r++) c.rng.rk[r] = p.rk[r];
    c.rng.has_rk = 1;
  } else {
    c.rng.has_rk = 0;
  }
  c.rng.k0 = p.k0; c.rng.k1 = p.k1; c.rng.id = (uint32_t)id; c.rng.stream = stream; c.rng.dom = dom; c.rng.slot = 0;
  c.id = id; c.oid = id; c.atbirth = 0; c.atrepli = 0; c.ist0 = 0; c.sis_y = c.sis_e = c.sis_aux = nullptr; c.has_sister = 0; c.stop = 0; c.slow = 0;
  c.ky_base = nullptr; c.ky_i = 0; c.ky_n = 1; c.ky_stride = 0;
  c.dl = p.dl;
  return c;
}
template <bool TAB>
__device__ __forceinline__ typename CtxSel<TAB>::type make_ctx_t(const Pop &p, double h, uint32_t stream, uint32_t dom, int32_t id) {
  typename CtxSel<TAB>::type c;
  static_cast<DevCtx &>(c) = make_ctx(p, h, stream, dom, id);
  return c;
}
__device__ __forceinline__ void ctx_neighbours(DevCtx &c, const Pop &p, const double *ycur, int64_t i) {
  if (M::USES_KY) { c.ky_base = ycur; c.ky_i = i; c.ky_n = p.n; c.ky_stride = p.stride; }
}

/* STOPIF (src/odexp.h:400-406): the reference sets the global EVENT_TYPE from inside ode_rhs; here the full (second-pass)
 * evaluation of a particle raises one sticky device flag that the time loop reads after the step.  Compiled out for models
 * without STOPIF.  A fast-path evaluation that is about to be redone (c.slow) does not count. */
__device__ __forceinline__ void note_stop(const Pop &p, const DevCtx &c) {
  if (M::HAS_STOPIF) { if (c.stop && !c.slow) atomicOr(p.stop, 1); }
}

/* %M terms of one particle state.  Pass-1 semantics of the reference: when a %M expression reads an
 * auxiliary variable (MA), that value comes from ode_rhs evaluated with zeroed mean fields and the
 * previous psi (src/odexp:820-833). */
constexpr bool USE_HINTS = (M::NHINT > 0) && !M::TERMS_USE_AUX;     /* stage-to-stage sincos increments (odx_sincos_inc) */
constexpr bool USE_SCTAB = SCTAB_MODEL;                          /* models whose RHS and %M terms take sin / cos: table-assisted sincos in the persistent kernels */
template <bool EXACT, class CT>
__device__ __forceinline__ void particle_terms_x(CT &c, double t, const double (&y)[A1(NVAR)], const double (&e)[A1(NEXPR)],
                                                 const double (&psi)[A1(NPSI)], double (&tm)[NRED], const OdxHint *hint = nullptr) {
  double aux[A1(NAUX)];
  if (M::TERMS_USE_AUX) {
    double f[A1(NVAR)], mz[A1(NMFD)], d = 0.0, r = 0.0;
#pragma unroll
    for (int k = 0; k < A1(NMFD); k++) mz[k] = 0.0;
    CT c1 = c;
    M::template rhs<EXACT>(c1, t, y, e, psi, mz, f, aux, d, r);
    c.slow |= c1.slow;
  }
  double m[A1(NMFD)];
  if (NMFD > 0) {
    if (USE_HINTS && hint != nullptr) M::template mfd_terms_inc<EXACT>(c, t, y, e, aux, m, *hint);
    else M::template mfd_terms<EXACT>(c, t, y, e, aux, m);
#pragma unroll
    for (int k = 0; k < NMFD; k++) tm[k] = m[k];
  }
  if (M::HAS_BIRTH) tm[RED_BIRTH] = M::template birth_term<EXACT>(c, t, y, e, aux);
}
/* accumulating form with the rare redo (kernels that carry one particle per thread) */
__device__ __forceinline__ void particle_terms(DevCtx &c, double t, const double (&y)[A1(NVAR)], const double (&e)[A1(NEXPR)],
                                               const double (&psi)[A1(NPSI)], double (&red)[NRED]) {
  double tm[NRED];
#pragma unroll
  for (int k = 0; k < NRED; k++) tm[k] = 0.0;
  DevCtx c0 = c;
  particle_terms_x<false>(c, t, y, e, psi, tm);
  if (c.slow) { c = c0; particle_terms_x<true>(c, t, y, e, psi, tm); }
#pragma unroll
  for (int k = 0; k < NRED; k++) red[k] += tm[k];
}
/* ode_rhs with the rare redo */
__device__ __forceinline__ void rhs_checked(DevCtx &c, double t, const double (&y)[A1(NVAR)], const double (&e)[A1(NEXPR)],
                                            const double (&psi)[A1(NPSI)], const double (&mfd)[A1(NMFD)], double (&f)[A1(NVAR)],
                                            double (&aux)[A1(NAUX)], double &death, double &repli) {
  DevCtx c0 = c;
  M::template rhs<false>(c, t, y, e, psi, mfd, f, aux, death, repli);
  if (c.slow) { c = c0; M::template rhs<true>(c, t, y, e, psi, mfd, f, aux, death, repli); }
}
__device__ __forceinline__ void rhs_checked(const Pop &p, DevCtx &c, double t, const double (&y)[A1(NVAR)], const double (&e)[A1(NEXPR)],
                                            const double (&psi)[A1(NPSI)], const double (&mfd)[A1(NMFD)], double (&f)[A1(NVAR)],
                                            double (&aux)[A1(NAUX)], double &death, double &repli) {
  rhs_checked(c, t, y, e, psi, mfd, f, aux, death, repli);
  note_stop(p, c);
}

/* ------------------------------------------------------------------------------------------------
 * kernels
 * ---------------------------------------------------------------------------------------------- */
/* moments of a state (start of a step when the previous kernel did not provide them) */
__global__ void __launch_bounds__(BLOCK) k_moments(Pop p, const double *__restrict__ ysrc, double t, Red r) {
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; k++) red[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
    double y[A1(NVAR)], e[A1(NEXPR)], psi[A1(NPSI)];
    load_cols<NVAR>(ysrc, p.stride, i, y);
    load_cols<NEXPR>(p.E, p.stride, i, e);
    if (M::TERMS_USE_AUX) load_cols<NPSI>(p.PSI, p.stride, i, psi);
    DevCtx c = make_ctx(p, 0.0, 0, DOM_RHS, M::NEEDS_ID ? p.ID[i] : 0);
    particle_terms(c, t, y, e, psi, red);
  }
  grid_reduce<NRED>(red, r);
}

/* Bounded wait of the in-kernel exchanges: a peer (CTA or GPU) that never arrives must not hang the box.  Returns true when the
 * wait is to be abandoned: after SPIN_LIMIT polls (tens of seconds; a legitimate wait lasts micro- to milliseconds), or as soon
 * as any earlier wait on this GPU has given up -- so that after the first time-out the remaining stages drain in ~1024 polls each
 * and the host reports the error (odx_sync) instead of waiting SPIN_LIMIT polls per stage. */
constexpr unsigned long long SPIN_LIMIT = 1ull << 26;      /* ~ 20 - 45 s of polling; a rank may legitimately be late by seconds (host-side output of a large frame) */
__device__ __forceinline__ bool spin_give_up(unsigned long long &spins, int *abort_flag) {
  if ((++spins & 1023ull) != 0) return false;
  if (spins > SPIN_LIMIT || (abort_flag != nullptr && *(volatile int *)abort_flag != 0)) {
    if (abort_flag != nullptr) *abort_flag = 1;
    return true;
  }
  return false;
}

struct ChainArgs {
  const double *sums_in;    /* raw %M sums of the state the RHS is evaluated at */
  double t_eval;            /* time of that state */
  double t_next;            /* time of the state formed by this kernel (for its %M terms) */
  double h;                 /* step size (DWDT) */
  double a;                 /* mid: y_s = y0 + a * k         | last: y_new = y0 + a * (acc + w * k) */
  double w;                 /* weight of k in the accumulator */
  uint32_t eval_index;      /* RHS evaluation counter (Philox stream) */
  int want_out;             /* write aux / rates (last step before an output or event) */
  int reverse;              /* sweep direction of this stage */
  double recon;             /* != 0: this is the stage after FIRST and the accumulator w1*k1 is rebuilt from the stage state,
                               acc = (ys - y0) * recon with recon = w1 / a1, instead of being stored and re-read (16 B per
                               particle-step less traffic; the rebuilt slope is exact to <= 1 ulp of y0 / (a1 h)) */
  int skip_acc_store;       /* FIRST stage: the next stage reconstructs the accumulator */
  int iterate;              /* iteration_apply (src/methods_odexp.c:1816-1823): y <- f(y) instead of y <- y + h f(y) */
};

/* chain methods (Euler, classic RK4): stage state depends on the previous slope only.
 * FIRST: the RHS is evaluated at Y itself (k1).  LAST: apply the step to Y and produce the %M sums of the
 * new state, which are exactly what the first stage of the next step needs.
 * Each thread carries two adjacent particles per iteration: every column is moved with 128-bit accesses and the two
 * independent dependency chains hide the FP64 latency of the per-particle model code. */
template <bool FIRST, bool LAST, bool EXACT, bool TAB>
__device__ __forceinline__ int chain_lane(const Pop &p, const ChainArgs &a, const double (&mfd)[A1(NMFD)], int32_t id, int64_t gi,
                                          const double (&y0)[A1(NVAR)], const double (&yc)[A1(NVAR)], const double (&e)[A1(NEXPR)],
                                          const double (&psi)[A1(NPSI)], const double (&acc_in)[A1(NVAR)], double (&acc)[A1(NVAR)],
                                          double (&yn)[A1(NVAR)], double (&aux)[A1(NAUX)], double &death, double &repli, double (&tm)[NRED]) {
  double f[A1(NVAR)];
  typename CtxSel<TAB>::type c = make_ctx_t<TAB>(p, a.h, a.eval_index, DOM_RHS, id);
  ctx_neighbours(c, p, FIRST ? p.Y : p.YS, gi);
  OdxHint hint;
  M::template rhs<EXACT>(c, a.t_eval, yc, e, psi, mfd, f, aux, death, repli, USE_HINTS ? &hint : nullptr);
  note_stop(p, c);
  int slow = c.slow;
#pragma unroll
  for (int k = 0; k < NRED; k++) tm[k] = 0.0;
  if (LAST) {
#pragma unroll
    for (int v = 0; v < NVAR; v++) yn[v] = a.iterate ? f[v] : y0[v] + a.a * ((FIRST ? 0.0 : acc_in[v]) + a.w * f[v]);
    tm[RED_RATES] = death + repli;
  } else {
#pragma unroll
    for (int v = 0; v < NVAR; v++) {
      acc[v] = FIRST ? a.w * f[v] : acc_in[v] + a.w * f[v];
      yn[v] = y0[v] + a.a * f[v];
    }
  }
  if (NMFD > 0 || M::HAS_BIRTH) {
    typename CtxSel<TAB>::type c2 = make_ctx_t<TAB>(p, a.h, a.eval_index + 1, DOM_RHS, id);
    particle_terms_x<EXACT>(c2, a.t_next, yn, e, psi, tm, USE_HINTS ? &hint : nullptr);
    slow |= c2.slow;
  }
  return slow;
}

template <int NC>
__device__ __forceinline__ void load_cols2(const double *__restrict__ base, int64_t stride, int64_t i, bool two,
                                           double (&o0)[A1(NC)], double (&o1)[A1(NC)]) {
#pragma unroll
  for (int c = 0; c < NC; c++) {
#if STAGE_LOAD_CS
    if (two) { const double2 v = __ldcs(reinterpret_cast<const double2 *>(base + (int64_t)c * stride + i)); o0[c] = v.x; o1[c] = v.y; }
#else
    if (two) { const double2 v = *reinterpret_cast<const double2 *>(base + (int64_t)c * stride + i); o0[c] = v.x; o1[c] = v.y; }
#endif
    else { o0[c] = base[(int64_t)c * stride + i]; o1[c] = o0[c]; }
  }
}
template <int NC>
__device__ __forceinline__ void store_cols2(double *__restrict__ base, int64_t stride, int64_t i, bool two,
                                            const double (&v0)[A1(NC)], const double (&v1)[A1(NC)]) {
#pragma unroll
  for (int c = 0; c < NC; c++) {
#if STAGE_STORE_CS
    if (two) __stcs(reinterpret_cast<double2 *>(base + (int64_t)c * stride + i), make_double2(v0[c], v1[c]));
#else
    if (two) *reinterpret_cast<double2 *>(base + (int64_t)c * stride + i) = make_double2(v0[c], v1[c]);
#endif
    else base[(int64_t)c * stride + i] = v0[c];
  }
}

/* ------------------------------------------------------------------------------------------------
 * TMA staging for the persistent kernel.  A CTA consumes one contiguous tile of 2 x 736 particles per iteration, i.e. one
 * contiguous 11.5 KB run per column: one elected thread fetches the runs of the next tiles with cp.async.bulk (the TMA unit, no
 * registers and no issue slots of the consumer warps) into a ring of shared-memory buffers, mbarrier complete_tx tells the
 * consumers when a tile has landed, and a second mbarrier per buffer (one arrival per warp) tells the producer when every
 * warp has copied its two particles out.  Compared with plain 128-bit loads this multiplies the bytes in flight per SM
 * (ring depth x tile instead of what the stalled warps hold), which is what a latency-bound HBM stream needs.
 * ---------------------------------------------------------------------------------------------- */
#ifndef PERSIST_TMA
#define PERSIST_TMA 1
#endif
#ifndef PERSIST_THREADS_TMA
#define PERSIST_THREADS_TMA 768    /* 23 consumer warps + 1 producer warp, 80 registers: measured best of 512 / 640 / 768 / 896 / 1024 */
#endif
#ifndef PERSIST_THREADS
/* without the ring: 32 warps x 64 registers hide the load latency best; with the sincos table the stage body needs ~80 registers
 * (at 64 it spills inside the sweep: 0.068 instead of 0.055 ms per step at 2^21 particles), so 24 warps x 85 */
#define PERSIST_THREADS (SCTAB_MODEL ? 768 : 1024)
#endif
#ifndef TMA_MAXBUF_N
#define TMA_MAXBUF_N 8   /* <= 16 (state word of tma_produce) */
#endif
constexpr int TMA_MAXBUF = TMA_MAXBUF_N;
constexpr int TMA_NCOL_MAX = 3 * NVAR + NEXPR + NPSI;                       /* y0, e, psi, ys, acc */
#ifndef TMA_SMEM_KB
#define TMA_SMEM_KB 192      /* ring size per CTA: 192 KB with one CTA per SM */
#endif
constexpr size_t TMA_SMEM = (size_t)TMA_SMEM_KB * 1024;
constexpr int PT_TMA = PERSIST_THREADS_TMA, PT_LD = PERSIST_THREADS;       /* CTA size of the two kernel variants */
constexpr int TMA_CONS = PT_TMA - 32;                                      /* consumer threads (the last warp produces) */
/* pairs of particles per consumer thread and ring tile: two when two tiles of the widest stage still fit (the per-tile work of a
 * consumer -- barrier wait, slot release, index arithmetic, ~45 issue slots -- is then shared by four particles instead of two) */
#ifndef TMA_PPT_MAX
#define TMA_PPT_MAX 2
#endif
constexpr int TMA_PPT = (TMA_PPT_MAX >= 2 && (size_t)TMA_NCOL_MAX * 4 * TMA_CONS * sizeof(double) * 2 <= TMA_SMEM) ? 2 : 1;
constexpr int TMA_TILE = 2 * TMA_PPT * TMA_CONS;                           /* particles per ring tile */
constexpr size_t TMA_COL_BYTES = (size_t)TMA_TILE * sizeof(double);        /* one column of one tile */
/* the ring is used when at least two tiles of the widest stage fit */
constexpr bool TMA_ON = PERSIST_TMA && ((size_t)TMA_NCOL_MAX * TMA_COL_BYTES * 2 <= TMA_SMEM);
/* barrier over the NT computing threads of the CTA: all of it (NT = PT_LD), or all but the TMA producer warp, which never joins */
template <int NT>
__device__ __forceinline__ void cta_sync() {
  if (NT != TMA_CONS) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}
/* ring geometry of one stage: columns in ring order Y, E, PSI, [YS], [ACC] */
__device__ __forceinline__ void tma_geometry(bool first, bool use_acc, int &ncol, uint32_t &tile_bytes, int &nbuf) {
  ncol = NVAR + NEXPR + NPSI + (first ? 0 : NVAR) + (use_acc ? NVAR : 0);
  tile_bytes = (uint32_t)ncol * (uint32_t)TMA_COL_BYTES;
  nbuf = (int)(TMA_SMEM / tile_bytes);
  if (nbuf > TMA_MAXBUF) nbuf = TMA_MAXBUF;
}
struct TmaRing {
  unsigned char *buf;            /* TMA_SMEM bytes, 128-byte aligned */
  uint32_t base;                 /* the same as a 32-bit shared-window address */
  uint64_t *full, *empty;        /* [TMA_MAXBUF] each */
  uint32_t full_a, empty_a;      /* their 32-bit shared-window addresses (no generic-to-shared conversion per tile) */
  uint32_t full_par;             /* bit b: parity of the next phase of full[b] this thread will wait for */
};
__device__ __forceinline__ uint32_t smem_u32(const void *q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { mbar_arrive_a(smem_u32(bar)); }
__device__ __forceinline__ uint32_t mbar_try_wait_a(uint32_t bar_a, uint32_t parity) {
  uint32_t ok;
  /* the suspend-time hint lets the hardware park the warp until the phase completes instead of returning after the default
   * (short) limit: fewer re-issues of the try, fewer issue slots taken from the computing warps */
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar_a), "r"(parity), "r"(0x989680u) : "memory");
  return ok;
}
/* one thread waits */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  while (!mbar_try_wait_a(a, parity)) {}
}
/* a whole (converged) warp waits: the vote keeps the loop warp-uniform, so the compiler may go on using the uniform datapath
 * (constant-bank coefficients in uniform registers) in the code that follows.  Address and parity are formed before the loop:
 * a waiting warp issues three instructions per try, and every issue slot it takes is one the computing warps of its scheduler
 * do not get. */
__device__ __forceinline__ void mbar_wait_warp_a(uint32_t bar_a, uint32_t parity) {
  while (!__all_sync(0xffffffffu, mbar_try_wait_a(bar_a, parity))) {}
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* this thread's pair out of column c of a ring slot (slot = 32-bit shared address of the thread's pair in column 0) */
template <int NC>
__device__ __forceinline__ void unstage_cols(uint32_t slot, double (&o0)[A1(NC)], double (&o1)[A1(NC)]) {
#pragma unroll
  for (int c = 0; c < NC; c++) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o0[c]), "=d"(o1[c]) : "r"(slot + (uint32_t)(c * TMA_COL_BYTES)) : "memory");
  }
}

/* the producer lane of the TMA ring: fetches the column runs of every tile of one sweep, never more than nbuf tiles ahead of
 * the slowest consumer warp.  Called from ONE place, the producer branch at the top of the persistent kernel: instantiated
 * inside every stage sweep (or called out of line from there) it made ptxas park the sincos coefficient table on the stack of
 * every consumer thread.
 * state: bits 0-15 parity of the number of tiles issued into slot b, bits 16-31 slot b used at all. */
__device__ __forceinline__ uint32_t tma_produce(const double *Y, const double *E, const double *PSI, const double *YS, const double *ACC,
                                             int64_t stride, int64_t n, int64_t per_iter, int64_t cta_off, int niter, int reverse,
                                             int nbuf, uint32_t tile_bytes, unsigned char *buf, uint64_t *full, uint64_t *empty,
                                             uint32_t state) {
  const int ncol = NVAR + NEXPR + NPSI + (YS ? NVAR : 0) + (ACC ? NVAR : 0);
  int b = 0;
  for (int k = 0; k < niter; k++) {
    const int kk = reverse ? (niter - 1 - k) : k;
    const int64_t start = (int64_t)kk * per_iter + cta_off;
    int64_t cnt = n - start;
    if (cnt > (int64_t)TMA_TILE) cnt = (int64_t)TMA_TILE;
    if (cnt < 0) cnt = 0;
    cnt = (cnt + 1) & ~(int64_t)1;                      /* whole pairs: the partner of a trailing single particle lies inside the
                                                           (32-particle padded) column and is ignored by the consumer */
    const uint32_t bytes = (uint32_t)(cnt * sizeof(double));
    if ((state >> (16 + b)) & 1u) mbar_wait(empty + b, ((state >> b) & 1u) ^ 1u);
    state |= 1u << (16 + b); state ^= 1u << b;
    mbar_expect_tx(full + b, bytes * (uint32_t)ncol);
    if (bytes) {
      unsigned char *dst = buf + (size_t)b * tile_bytes;
      for (int c = 0; c < NVAR; c++, dst += TMA_COL_BYTES) bulk_g2s(dst, Y + (int64_t)c * stride + start, bytes, full + b);
      for (int c = 0; c < NEXPR; c++, dst += TMA_COL_BYTES) bulk_g2s(dst, E + (int64_t)c * stride + start, bytes, full + b);
      for (int c = 0; c < NPSI; c++, dst += TMA_COL_BYTES) bulk_g2s(dst, PSI + (int64_t)c * stride + start, bytes, full + b);
      if (YS) for (int c = 0; c < NVAR; c++, dst += TMA_COL_BYTES) bulk_g2s(dst, YS + (int64_t)c * stride + start, bytes, full + b);
      if (ACC) for (int c = 0; c < NVAR; c++, dst += TMA_COL_BYTES) bulk_g2s(dst, ACC + (int64_t)c * stride + start, bytes, full + b);
    }
    if (++b == nbuf) b = 0;
  }
  return state;
}

/* the arithmetic and the stores of one thread's pair of particles.  FULL: the pair lies inside the population (known per
 * tile, so the tail predicates vanish from the hot path) */
template <bool FIRST, bool LAST, bool FULL, bool TAB>
__device__ __forceinline__ void chain_pair(const Pop &p, const ChainArgs &a, const double (&mfd)[A1(NMFD)], int64_t i, bool two_rt,
                                           double (&y0)[2][A1(NVAR)], double (&yc)[2][A1(NVAR)], double (&e)[2][A1(NEXPR)],
                                           double (&psi)[2][A1(NPSI)], double (&acc)[2][A1(NVAR)], double (&red)[NRED]) {
  const bool two = FULL ? true : two_rt;
  double yn[2][A1(NVAR)], aux[2][A1(NAUX)];
  double death[2] = {0.0, 0.0}, repli[2] = {0.0, 0.0};
  if (FIRST) {
#pragma unroll
    for (int v = 0; v < NVAR; v++) { yc[0][v] = y0[0][v]; yc[1][v] = y0[1][v]; }
  } else if (a.recon != 0.0) {
#pragma unroll
    for (int v = 0; v < NVAR; v++) { acc[0][v] = (yc[0][v] - y0[0][v]) * a.recon; acc[1][v] = (yc[1][v] - y0[1][v]) * a.recon; }
  }
  int32_t id0 = 0, id1 = 0;
  if (M::NEEDS_ID) { id0 = p.ID[i]; id1 = two ? p.ID[i + 1] : id0; }
  double acco[2][A1(NVAR)], tm[2][NRED];
  const int s0 = chain_lane<FIRST, LAST, false, TAB>(p, a, mfd, id0, i, y0[0], yc[0], e[0], psi[0], acc[0], acco[0], yn[0], aux[0], death[0], repli[0], tm[0]);
  const int s1 = chain_lane<FIRST, LAST, false, TAB>(p, a, mfd, id1, i + 1, y0[1], yc[1], e[1], psi[1], acc[1], acco[1], yn[1], aux[1], death[1], repli[1], tm[1]);
  if (s0 | s1) {          /* rare: an argument outside the fast sincos range -> library path for this pair */
    chain_lane<FIRST, LAST, true, TAB>(p, a, mfd, id0, i, y0[0], yc[0], e[0], psi[0], acc[0], acco[0], yn[0], aux[0], death[0], repli[0], tm[0]);
    chain_lane<FIRST, LAST, true, TAB>(p, a, mfd, id1, i + 1, y0[1], yc[1], e[1], psi[1], acc[1], acco[1], yn[1], aux[1], death[1], repli[1], tm[1]);
  }
#pragma unroll
  for (int k2 = 0; k2 < NRED; k2++) red[k2] += tm[0][k2] + (two ? tm[1][k2] : 0.0);
  if (LAST) {
    store_cols2<NVAR>(p.Y_OUT, p.stride, i, two, yn[0], yn[1]);
    if (a.want_out) {
      store_cols2<NAUX>(p.AUX, p.stride, i, two, aux[0], aux[1]);
      if (M::HAS_DEATH) { p.DEATH[i] = death[0]; if (two) p.DEATH[i + 1] = death[1]; }
      if (M::HAS_REPLI) { p.REPLI[i] = repli[0]; if (two) p.REPLI[i + 1] = repli[1]; }
    }
  } else {
    if (!(FIRST && a.skip_acc_store)) store_cols2<NVAR>(p.ACC, p.stride, i, two, acco[0], acco[1]);
    store_cols2<NVAR>(p.YS_OUT, p.stride, i, two, yn[0], yn[1]);
  }
}

/* One pass over this GPU's particles, two particles per thread and load (128-bit accesses).  body(i, two, full, y0, e, psi, ys,
 * acc) does the arithmetic and the stores of one pair; full (a compile-time tag) says that the pair lies inside the population.
 * The grid sweeps the columns as a moving window of gridDim.x tiles; alternate stages sweep in opposite directions so that the
 * tail of one stage (still resident in the 126 MB L2) is the head of the next.
 * TMA: warp-specialised.  The last warp of the CTA is the producer (one lane issues the bulk copies and runs ahead as far as
 * the ring allows); the other warps consume TMA_PPT pairs per thread from every tile and release the slot as soon as the last
 * pair has been copied out.  The producer's registers do not add to the consumers' live state. */
struct FullTile { static constexpr bool value = true; };
struct RaggedTile { static constexpr bool value = false; };
template <bool TMA, class F>
__device__ __forceinline__ void tile_sweep(const Pop &p, int reverse, bool has_ys, bool has_acc, TmaRing &ring, F &&body) {
  constexpr int PPT = TMA ? TMA_PPT : 1;
  const int ncons = TMA ? TMA_CONS : (int)blockDim.x;
  const int tile = 2 * PPT * ncons;                                   /* particles per CTA and iteration */
  const int64_t per_iter = (int64_t)gridDim.x * tile;
  const int niter = (int)((p.n + per_iter - 1) / per_iter);
  /* iterations kk <= kfull find this CTA's whole tile inside the population (CTA-uniform) */
  const int64_t need = (int64_t)tile * (blockIdx.x + 1);
  const int kfull = (TMA && p.n >= need) ? (int)((p.n - need) / per_iter) : -1;
  int ncol, nbuf; uint32_t tile_bytes;
  tma_geometry(!has_ys, has_acc, ncol, tile_bytes, nbuf);
  const int64_t toff = (int64_t)tile * blockIdx.x + 2 * threadIdx.x;
  int bcur = 0;
  /* shared-window addresses formed once per sweep and made opaque: left to itself the compiler re-derives them from the special
   * registers (S2R CgaCtaId / SWINHI / TID.X + alignment arithmetic, ~10 instructions) in front of every pair of every tile */
  uint32_t slot0 = 0, full_a = 0, empty_a = 0;
  if (TMA) {
    slot0 = ring.base + threadIdx.x * (uint32_t)sizeof(double2); full_a = ring.full_a; empty_a = ring.empty_a;
    asm volatile("" : "+r"(slot0), "+r"(full_a), "+r"(empty_a));
  }
  for (int k = 0; k < niter; k++) {
    const int kk = reverse ? (niter - 1 - k) : k;
    const int64_t i0 = (int64_t)kk * per_iter + toff;
    int full_tile = (kk <= kfull);                      /* kept as a register: re-deriving it costs six instructions per pair */
    if (TMA) asm volatile("" : "+r"(full_tile));
    uint32_t slot = 0;
    if (TMA) {
      mbar_wait_warp_a(full_a + 8u * bcur, (ring.full_par >> bcur) & 1u);
      ring.full_par ^= 1u << bcur;
      slot = slot0 + (uint32_t)bcur * tile_bytes;
    }
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      const int64_t i = i0 + 2 * q * ncons;
      const bool two = (i + 1 < p.n);
      double y0[2][A1(NVAR)], ys[2][A1(NVAR)], e[2][A1(NEXPR)], psi[2][A1(NPSI)], acc[2][A1(NVAR)];
      if (TMA) {
        const uint32_t sq = slot + (uint32_t)(q * TMA_CONS * sizeof(double2));
        unstage_cols<NVAR>(sq, y0[0], y0[1]);
        unstage_cols<NEXPR>(sq + (uint32_t)(NVAR * TMA_COL_BYTES), e[0], e[1]);
        if (NPSI > 0) unstage_cols<NPSI>(sq + (uint32_t)((NVAR + NEXPR) * TMA_COL_BYTES), psi[0], psi[1]);
        if (has_ys) unstage_cols<NVAR>(sq + (uint32_t)((NVAR + NEXPR + NPSI) * TMA_COL_BYTES), ys[0], ys[1]);
        if (has_acc) unstage_cols<NVAR>(sq + (uint32_t)((2 * NVAR + NEXPR + NPSI) * TMA_COL_BYTES), acc[0], acc[1]);
        if (q == PPT - 1) {
          /* the slot is refilled through the async proxy: order this thread's generic-proxy reads of it before the release
           * (without this fence a refill was observed to overtake the reads once in a few 10^5 tiles) */
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if ((threadIdx.x & 31) == 0) mbar_arrive_a(empty_a + 8u * bcur);
          if (++bcur == nbuf) bcur = 0;
        }
        if (full_tile) { body(i, true, FullTile(), y0, e, psi, ys, acc); continue; }
      } else {
        if (i >= p.n) continue;
        load_cols2<NVAR>(p.Y, p.stride, i, two, y0[0], y0[1]);
        load_cols2<NEXPR>(p.E, p.stride, i, two, e[0], e[1]);
        if (NPSI > 0) load_cols2<NPSI>(p.PSI, p.stride, i, two, psi[0], psi[1]);
        if (has_ys) load_cols2<NVAR>(p.YS, p.stride, i, two, ys[0], ys[1]);
        if (has_acc) load_cols2<NVAR>(p.ACC, p.stride, i, two, acc[0], acc[1]);
      }
      if (i < p.n) body(i, two, RaggedTile(), y0, e, psi, ys, acc);
    }
  }
  /* the next stage fetches what this one stored through the async proxy (TMA), which reads L2 directly: the stores must have
   * reached gpu scope (a CTA-scope release could leave them in flight between the SM and L2), then cross the proxy boundary */
  if (TMA) {
    __threadfence();
    asm volatile("fence.proxy.async.global;" ::: "memory");
  }
}

/* one pass of a chain stage (stored scheme) */
template <bool FIRST, bool LAST, bool TMA, bool TAB = false>
__device__ __forceinline__ void chain_sweep(const Pop &p, const ChainArgs &a, const double (&mfd)[A1(NMFD)], double (&red)[NRED],
                                            TmaRing &ring) {
  const bool use_acc = !FIRST && a.recon == 0.0;
  tile_sweep<TMA>(p, a.reverse, !FIRST, use_acc, ring,
                  [&](int64_t i, bool two, auto full, double (&y0)[2][A1(NVAR)], double (&e)[2][A1(NEXPR)], double (&psi)[2][A1(NPSI)],
                      double (&ys)[2][A1(NVAR)], double (&acc)[2][A1(NVAR)]) {
                    chain_pair<FIRST, LAST, decltype(full)::value, TAB>(p, a, mfd, i, two, y0, ys, e, psi, acc, red);
                  });
}

/* ------------------------------------------------------------------------------------------------
 * Lean classic RK4 for HBM-resident mean-field populations (persistent TMA kernel only).
 *
 * The stored scheme above moves 152 B per particle-step for a one-variable model (stage states and the slope accumulator go
 * to HBM and come back); the compulsory traffic with one mean-field barrier per stage is 72 B.  Here stages are recomputed or
 * recovered instead of stored wherever that costs fewer FP64 instructions than the bytes cost time:
 *   stage 1: read y0, e          k1 = f(y0; M0), y2 = y0 + h/2 k1                        -> sums of y2      (nothing stored)
 *   stage 2: read y0, e          k1 again (M0 kept in shared memory), y2, k2 = f(y2; M1), y3 = y0 + h/2 k2  -> store y3 (YS)
 *   stage 3: read y0, e, y3      k3 = f(y3; M2), y4 = y0 + h k3                                             -> store y4 (ACC)
 *   stage 4: read y0, e, y3, y4  k1 again, k2 = (y3 - y0) / (h/2), k3 = (y4 - y0) / h (recovered: exact to <= 1 ulp of
 *                                y0 / (a h), like `recon` above), k4 = f(y4; M3), y += h/6 (k1 + 2 k2 + 2 k3 + k4)  -> store y
 * = 16 + 24 + 32 + 40 = 112 B per particle-step for nvar = nexpr = 1, six RHS evaluations instead of four.  The recomputed
 * evaluations reuse the sincos triple of y0 through angle addition (rhs_inc), so a step costs ~245 FP64 instructions
 * instead of ~200.  Requires: no %C columns (psi of earlier stages would have to be kept), no random numbers or STOPIF in the
 * RHS.  While the population sits in L2 the stored scheme is faster (FP64-bound there) and stays in use.
 * ---------------------------------------------------------------------------------------------- */
constexpr bool LEAN_OK = (NPSI == 0) && !M::RHS_RNG && !M::HAS_STOPIF && !M::HAS_PAIR && !M::HAS_LOWRANK && !M::USES_KY;
/* ODX_LEAN_RECOMP_Y4 = 1: stage 3 does not store y4 either; stage 4 evaluates k3 = f(y3; M2) itself (M2 kept like M0) and forms y4
 * from it: 16 + 24 + 24 + 32 = 96 B per particle-step, seven RHS evaluations.  Measured and NOT used (profiles/r2w_kbench_lean_recompute_y4.txt):
 * 0.347 instead of 0.323 ms per step at 2^24 particles, 0.183 instead of 0.172 at 2^23 -- the stage-by-stage model of DESIGN 3.1 said
 * -6 %, but the seventh evaluation takes stage 4 from 132 to 428 bytes of spill loads at 80 registers and makes it issue-bound. */
#ifndef ODX_LEAN_RECOMP_Y4
#define ODX_LEAN_RECOMP_Y4 0
#endif
constexpr bool LEAN_RECOMP_Y4 = ODX_LEAN_RECOMP_Y4 != 0;
template <int LS, bool EXACT>
__device__ __forceinline__ int lean_lane(const Pop &p, const ChainArgs &a, const double (&mfd0)[A1(NMFD)], const double (&mfd2)[A1(NMFD)],
                                         const double (&mfd)[A1(NMFD)], int32_t id, const double (&y0)[A1(NVAR)], const double (&e)[A1(NEXPR)],
                                         const double (&y3)[A1(NVAR)], const double (&y4)[A1(NVAR)], double (&yout)[A1(NVAR)],
                                         double (&aux)[A1(NAUX)], double &death, double &repli, double (&tm)[NRED]) {
  double psi[A1(NPSI)], f[A1(NVAR)], f2[A1(NVAR)];
#pragma unroll
  for (int k = 0; k < A1(NPSI); k++) psi[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NRED; k++) tm[k] = 0.0;
  const double t = a.t_eval, h = a.h, half = a.h / 2.0;
  typename CtxSel<USE_SCTAB>::type c = make_ctx_t<USE_SCTAB>(p, h, a.eval_index, DOM_RHS, id);
  OdxHint h0, h1;
  OdxHint *hl = &h0;                    /* triple the %M terms of the new state are incremented from */
  double t_next;
  if (LS == 0) {
    M::template rhs<EXACT>(c, t, y0, e, psi, mfd, f, aux, death, repli, USE_HINTS ? &h0 : nullptr);
#pragma unroll
    for (int v = 0; v < NVAR; v++) yout[v] = y0[v] + half * f[v];
    t_next = t + half;
  } else if (LS == 1) {
    double y2[A1(NVAR)];
    M::template rhs<EXACT>(c, t, y0, e, psi, mfd0, f, aux, death, repli, USE_HINTS ? &h0 : nullptr);
#pragma unroll
    for (int v = 0; v < NVAR; v++) y2[v] = y0[v] + half * f[v];
    M::template rhs_inc<EXACT>(c, t + half, y2, e, psi, mfd, f2, aux, death, repli, h0, USE_HINTS ? &h1 : nullptr);
#pragma unroll
    for (int v = 0; v < NVAR; v++) yout[v] = y0[v] + half * f2[v];
    hl = &h1; t_next = t + half;
  } else if (LS == 2) {
    M::template rhs<EXACT>(c, t + half, y3, e, psi, mfd, f, aux, death, repli, USE_HINTS ? &h0 : nullptr);
#pragma unroll
    for (int v = 0; v < NVAR; v++) yout[v] = y0[v] + h * f[v];
    t_next = t + h;
  } else {
    M::template rhs<EXACT>(c, t, y0, e, psi, mfd0, f, aux, death, repli, USE_HINTS ? &h0 : nullptr);
    const double r2 = 1.0 / half, r3 = 1.0 / h;
    if (LEAN_RECOMP_Y4) {
      double f3[A1(NVAR)], y4r[A1(NVAR)];
      M::template rhs<EXACT>(c, t + half, y3, e, psi, mfd2, f3, aux, death, repli, nullptr);      /* k3 again: the same instructions as in stage 3 */
#pragma unroll
      for (int v = 0; v < NVAR; v++) y4r[v] = y0[v] + h * f3[v];
      M::template rhs_inc<EXACT>(c, t + h, y4r, e, psi, mfd, f2, aux, death, repli, h0, USE_HINTS ? &h1 : nullptr);
#pragma unroll
      for (int v = 0; v < NVAR; v++) {
        double acc = f[v];
        acc = acc + 2.0 * ((y3[v] - y0[v]) * r2);
        acc = acc + 2.0 * f3[v];
        yout[v] = y0[v] + (h / 6.0) * (acc + f2[v]);
      }
    } else {
      M::template rhs_inc<EXACT>(c, t + h, y4, e, psi, mfd, f2, aux, death, repli, h0, USE_HINTS ? &h1 : nullptr);
#pragma unroll
      for (int v = 0; v < NVAR; v++) {
        double acc = f[v];                                  /* k1 + 2 k2 + 2 k3, then y += h/6 (acc + k4): the order of the stored scheme */
        acc = acc + 2.0 * ((y3[v] - y0[v]) * r2);
        acc = acc + 2.0 * ((y4[v] - y0[v]) * r3);
        yout[v] = y0[v] + (h / 6.0) * (acc + f2[v]);
      }
    }
    tm[RED_RATES] = death + repli;
    hl = &h1; t_next = t + h;
  }
  int slow = c.slow;
  if (NMFD > 0 || M::HAS_BIRTH) {
    typename CtxSel<USE_SCTAB>::type c2 = make_ctx_t<USE_SCTAB>(p, h, a.eval_index + 1, DOM_RHS, id);
    particle_terms_x<EXACT>(c2, t_next, yout, e, psi, tm, USE_HINTS ? hl : nullptr);
    slow |= c2.slow;
  }
  return slow;
}

template <int LS, bool FULL>
__device__ __forceinline__ void lean_pair(const Pop &p, const ChainArgs &a, const double (&mfd0)[A1(NMFD)], const double (&mfd2)[A1(NMFD)],
                                          const double (&mfd)[A1(NMFD)], int64_t i, bool two_rt, double (&y0)[2][A1(NVAR)], double (&e)[2][A1(NEXPR)],
                                          double (&y3)[2][A1(NVAR)], double (&y4)[2][A1(NVAR)], double (&red)[NRED]) {
  const bool two = FULL ? true : two_rt;
  double yn[2][A1(NVAR)], aux[2][A1(NAUX)], tm[2][NRED];
  double death[2] = {0.0, 0.0}, repli[2] = {0.0, 0.0};
  int32_t id0 = 0, id1 = 0;
  if (M::NEEDS_ID) { id0 = p.ID[i]; id1 = two ? p.ID[i + 1] : id0; }
  const int s0 = lean_lane<LS, false>(p, a, mfd0, mfd2, mfd, id0, y0[0], e[0], y3[0], y4[0], yn[0], aux[0], death[0], repli[0], tm[0]);
  const int s1 = lean_lane<LS, false>(p, a, mfd0, mfd2, mfd, id1, y0[1], e[1], y3[1], y4[1], yn[1], aux[1], death[1], repli[1], tm[1]);
  if (s0 | s1) {          /* rare: a sincos argument outside the fast range, or a stage increment too large for angle addition */
    lean_lane<LS, true>(p, a, mfd0, mfd2, mfd, id0, y0[0], e[0], y3[0], y4[0], yn[0], aux[0], death[0], repli[0], tm[0]);
    lean_lane<LS, true>(p, a, mfd0, mfd2, mfd, id1, y0[1], e[1], y3[1], y4[1], yn[1], aux[1], death[1], repli[1], tm[1]);
  }
#pragma unroll
  for (int k2 = 0; k2 < NRED; k2++) red[k2] += tm[0][k2] + (two ? tm[1][k2] : 0.0);
  if (LS == 1) store_cols2<NVAR>(p.YS, p.stride, i, two, yn[0], yn[1]);
  if (LS == 2 && !LEAN_RECOMP_Y4) store_cols2<NVAR>(p.ACC, p.stride, i, two, yn[0], yn[1]);
  if (LS == 3) {
    store_cols2<NVAR>(p.Y, p.stride, i, two, yn[0], yn[1]);
    if (a.want_out) {
      store_cols2<NAUX>(p.AUX, p.stride, i, two, aux[0], aux[1]);
      if (M::HAS_DEATH) { p.DEATH[i] = death[0]; if (two) p.DEATH[i + 1] = death[1]; }
      if (M::HAS_REPLI) { p.REPLI[i] = repli[0]; if (two) p.REPLI[i + 1] = repli[1]; }
    }
  }
}

/* ring columns of lean stage LS: Y, E, [YS = y3], [ACC = y4] */
__device__ __forceinline__ bool lean_has_ys(int ls) { return ls >= 2; }
__device__ __forceinline__ bool lean_has_acc(int ls) { return ls == 3 && !LEAN_RECOMP_Y4; }
template <int LS, bool TMA>
__device__ __forceinline__ void lean_sweep(const Pop &p, const ChainArgs &a, const double (&mfd0)[A1(NMFD)], const double (&mfd2)[A1(NMFD)],
                                           const double (&mfd)[A1(NMFD)], double (&red)[NRED], TmaRing &ring) {
  tile_sweep<TMA>(p, a.reverse, LS >= 2, LS == 3 && !LEAN_RECOMP_Y4, ring,
                  [&](int64_t i, bool two, auto full, double (&y0)[2][A1(NVAR)], double (&e)[2][A1(NEXPR)], double (&psi)[2][A1(NPSI)],
                      double (&y3)[2][A1(NVAR)], double (&y4)[2][A1(NVAR)]) {
                    lean_pair<LS, decltype(full)::value>(p, a, mfd0, mfd2, mfd, i, two, y0, e, y3, y4, red);
                  });
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(BLOCK, CHAIN_MIN_BLOCKS) k_chain(Pop p, ChainArgs a, Red r) {
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; k++) red[k] = 0.0;
  double mfd[A1(NMFD)];
  load_mfd(p, a.sums_in, p.n_global, mfd);
  TmaRing none{};
  chain_sweep<FIRST, LAST, false>(p, a, mfd, red, none);
  if (NMFD > 0 || M::HAS_BIRTH || (LAST && (M::HAS_DEATH || M::HAS_REPLI))) grid_reduce<NRED>(red, r);
}

/* ------------------------------------------------------------------------------------------------
 * Persistent step kernel: `nsteps` whole steps of a chain method in ONE cooperative launch.
 *
 * The mean-field barrier between stages is a grid barrier inside the kernel (arrive counter + generation word in
 * global memory, all CTAs co-resident) instead of a kernel boundary, and across GPUs it is an exchange of the NRED
 * partial sums through peer memory: after the local barrier CTA 0 adds the block partials in a fixed order and
 * stores the result, tagged with the epoch, straight into every peer's mailbox over NVLink (plain st.global on
 * IPC-mapped pointers, __threadfence_system, then the flag); every CTA then waits for the flags of all ranks in its
 * own GPU's mailbox and adds the rank sums in rank order.  No NCCL call, no host round trip per stage: the fused
 * compute + collective the latency-bound per-stage allreduce (24-40 bytes) needs.
 * ---------------------------------------------------------------------------------------------- */
constexpr int MAX_RANKS = 8;
struct PersistArgs {
  int nst; long long nsteps; double t0, h; int want_out_last; uint32_t eval_index0; int recon; int iterate;
  int lean;                        /* classic RK4 by the lean scheme (lean_sweep) */
  const double *sums_y;            /* global raw sums of the %M terms of Y at entry */
  double *partial;                 /* [2][grid][NRED] */
  unsigned int *bar;               /* [2][PERSIST_MAX_CTAS] epoch tags of the published block partials, then [2] global-sum flags */
  double *totals;                  /* [2][NRED] global sums published by CTA 0 (multi-GPU) */
  double *sums_new, *sums_report;  /* out: sums of the new Y (+ rates) / sums used by the last RHS evaluation */
  int rank, nranks; uint32_t epoch0;
  double *mailbox; uint32_t *mailflag;                 /* local: [2][MAX_RANKS][NRED], [2][MAX_RANKS] (k_peer_allreduce) */
  double *peer_box[MAX_RANKS]; uint32_t *peer_flag[MAX_RANKS];
  unsigned long long *ll_part;     /* [2][PERSIST_MAX_CTAS][LLW] block partials as LL words */
  unsigned long long *ll_gtot;     /* [2][LLW] global sums republished by CTA 0 (multi-GPU) */
  unsigned long long *mail_ll;     /* local mailbox [2][MAX_RANKS][LLW] written by the peers */
  unsigned long long *peer_ll[MAX_RANKS];
  int *abort_flag;
};

/* acquire / release accesses used by the in-kernel mean-field exchange (gpu scope inside one GPU, sys scope across NVLink) */
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
/* block partial of red[] -> partial[blockIdx][*] */
template <int NT>
__device__ __forceinline__ void block_partial(const double (&red)[NRED], double *__restrict__ partial) {
  __shared__ double s_part[NT / 32][NRED];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NRED; k++) {
    const double v = warp_sum(red[k]);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) v += s_part[w][threadIdx.x];
    partial[(size_t)blockIdx.x * NRED + threadIdx.x] = v;
  }
  __syncthreads();
}
/* low-latency words: {tag:32 | data:32}; a double is two words (low half, high half) */
constexpr int LLW = 2 * NRED;
__device__ __forceinline__ unsigned long long ll_pack(double v, int half, uint32_t tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const uint32_t d = half ? (uint32_t)(bits >> 32) : (uint32_t)bits;
  return ((unsigned long long)tag << 32) | d;
}
__device__ __forceinline__ double ll_unpack(uint32_t lo, uint32_t hi) {
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
__device__ __forceinline__ void ll_store_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
/* both words of a double in one 16-byte request (each word still carries its own tag: no 16-byte atomicity is assumed) */
__device__ __forceinline__ void ll_load2_gpu(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void ll_load2_sys(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void ll_store_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#ifndef PERSIST_MAX_CTAS_N
#define PERSIST_MAX_CTAS_N 256
#endif
constexpr int PERSIST_MAX_CTAS = PERSIST_MAX_CTAS_N;
constexpr size_t PF_BYTES = TMA_SMEM + 128;
/* ---- mean-field barrier without a barrier, low-latency words.  Every exchanged double travels as two 8-byte words
 * {32 data bits, 32-bit epoch tag}: an 8-byte store is atomic, so data and "ready" flag arrive together and no fence
 * or separate flag round trip is needed (the LL idea of NCCL's small-message protocol).  Every CTA publishes its
 * block partials as such words, every CTA polls the words of all CTAs and adds the partials itself in a fixed order
 * (identical bits everywhere, no atomics).  Buffers alternate with the epoch parity: a CTA is one epoch ahead at most.
 * Three CTA barriers per stage: [A] warp partials in shared memory (warp 0 alone then forms and publishes the block partial),
 * [B] all words of all CTAs in shared memory, [C] the totals in s_tot.  Warp k adds value k over the CTAs (lane-strided, then a
 * shuffle tree: one fixed order) and, across GPUs, exchanges it with the peers itself: CTA 0's warp k stores the rank
 * total as LL words straight into every peer's mailbox over NVLink, warp k of EVERY CTA polls the peers' words in this GPU's
 * mailbox and adds the rank totals in rank order -- no further barrier, only the NVLink latency.
 * Used by the plain (L2-resident, latency-sensitive) variant: 4.4 instead of 5.2 us per stage with one tile per CTA, 50.8 instead of
 * 55.5 us per step at 2^21 particles.  The ring variant keeps the older eight-barrier form below: inlined there, the shuffle trees
 * and poll loops of this one raised the register pressure of its 80-register stage sweeps (432 bytes of spills, 0.326 -> 0.347 ms
 * per step at 2^24), and out of line the call itself cost more (0.40 ms). */
/* block sum of red[] into shared s_bp[NRED] */
template <int NT>
__device__ __forceinline__ void block_partial_smem(const double (&red)[NRED], double *s_bp) {
  __shared__ double s_part4[NT / 32][NRED];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cta_sync<NT>();
#pragma unroll
  for (int k = 0; k < NRED; k++) {
    const double v = warp_sum(red[k]);
    if (lane == 0) s_part4[warp][k] = v;
  }
  cta_sync<NT>();
  if (threadIdx.x < NRED) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) v += s_part4[w][threadIdx.x];
    s_bp[threadIdx.x] = v;
  }
  cta_sync<NT>();
}
/* fixed-order block sum of per-thread accumulators into s_tot[NRED] */
template <int NT>
__device__ __forceinline__ void block_total(const double (&acc)[NRED], double *s_tot) {
  __shared__ double s_part3[NT / 32][NRED];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cta_sync<NT>();
#pragma unroll
  for (int k = 0; k < NRED; k++) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_part3[warp][k] = v;
  }
  cta_sync<NT>();
  if (threadIdx.x < NRED) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) v += s_part3[w][threadIdx.x];
    s_tot[threadIdx.x] = v;
  }
  cta_sync<NT>();
}
struct ExchangeArgs {          /* the few fields of PersistArgs the exchange needs, by value (a reference to the kernel parameter block
                                  would be copied to the stack) */
  unsigned long long *ll_part, *mail_ll; unsigned long long *const *peer_ll;      /* peer_ll: this CTA's shared-memory copy of the peer mailboxes */
  int *abort_flag; int rank, nranks;
};
template <int NTC, bool PF_ON>
__device__ __forceinline__ void stage_exchange(const double (&red)[NRED], const ExchangeArgs &g, uint32_t epoch, double *s_tot,
                                            double (*s_wp)[NRED], uint32_t *s_words, uint64_t *s_go) {
  const size_t par = (size_t)(epoch & 1u);
  const uint32_t tag = epoch + 1u;
  unsigned long long *ll_part = g.ll_part + par * (size_t)PERSIST_MAX_CTAS * LLW;
  constexpr int NW = NTC / 32;
  /* the birth and rate slots are structurally zero in a model without %BIRTH / %DEATH / %REPLI: not reduced, not exchanged */
  constexpr int NX = (M::HAS_BIRTH || M::HAS_DEATH || M::HAS_REPLI) ? NRED : A1(NMFD), LLX = 2 * NX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NX; k++) {
    const double v = warp_sum(red[k]);
    if (lane == 0) s_wp[warp][k] = v;
  }
  cta_sync<NTC>();                                         /* [A] */
  if (PF_ON && threadIdx.x == 0) mbar_arrive(s_go);     /* every consumer has stored (and fenced) this stage */
  if (warp == 0) {
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < NX; k++) {
      double v = (lane < NW) ? s_wp[lane][k] : 0.0;
      v = warp_sum(v);
      v = __shfl_sync(0xffffffffu, v, 0);
      if ((lane >> 1) == k) mine = v;
    }
    if (lane < LLX) ll_store_gpu(ll_part + (size_t)blockIdx.x * LLX + lane, ll_pack(mine, lane & 1, tag));
  }
#ifndef ODX_LL_POLL16
#define ODX_LL_POLL16 1
#endif
  if (ODX_LL_POLL16) {
    /* one 16-byte request per double: when the CTAs arrive together, the grid x grid x LLX polls of the same few cache lines are what
     * the barrier costs (every L2 slice serialises the requests to its line) */
    for (unsigned int idx = threadIdx.x; idx < gridDim.x * NX; idx += NTC) {
      unsigned long long w0, w1; unsigned long long spins = 0;
      for (;;) {
        ll_load2_gpu(ll_part + 2 * idx, w0, w1);
        if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag) break;
        if (spin_give_up(spins, g.abort_flag)) break;
      }
      s_words[2 * idx] = (uint32_t)w0; s_words[2 * idx + 1] = (uint32_t)w1;
    }
  } else {
    for (unsigned int idx = threadIdx.x; idx < gridDim.x * LLX; idx += NTC) {
      unsigned long long w; unsigned long long spins = 0;
      while ((uint32_t)((w = ll_load_gpu(ll_part + idx)) >> 32) != tag) { if (spin_give_up(spins, g.abort_flag)) break; }
      s_words[idx] = (uint32_t)w;
    }
  }
  cta_sync<NTC>();                                         /* [B] */
  for (int k = warp; k < NX; k += NW) {
    double a = 0.0;
    for (unsigned int b2 = lane; b2 < gridDim.x; b2 += 32) a += ll_unpack(s_words[b2 * LLX + 2 * k], s_words[b2 * LLX + 2 * k + 1]);
    a = warp_sum(a);
    if (g.nranks > 1) {
      a = __shfl_sync(0xffffffffu, a, 0);
      /* lane r < nranks handles peer r: CTA 0 stores this rank's total of value k as two LL words into the peer's mailbox, every CTA
       * polls the peer's two words in its own GPU's mailbox with one 16-byte request (148 CTAs x nranks pollers share a few cache
       * lines of the mailbox: half the requests of a poll per word) */
      uint32_t wlo = 0, whi = 0;
      if (lane < g.nranks && lane != g.rank) {
        const int r = lane;
        if (blockIdx.x == 0) {
          unsigned long long *dst = g.peer_ll[r] + (par * MAX_RANKS + g.rank) * LLW + 2 * k;
          ll_store_sys(dst, ll_pack(a, 0, tag)); ll_store_sys(dst + 1, ll_pack(a, 1, tag));
        }
        unsigned long long w0, w1; unsigned long long spins = 0;
        for (;;) {
          ll_load2_sys(g.mail_ll + (par * MAX_RANKS + r) * LLW + 2 * k, w0, w1);
          if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag) break;
          if (spin_give_up(spins, g.abort_flag)) break;
        }
        wlo = (uint32_t)w0; whi = (uint32_t)w1;
      }
      double tot = 0.0;
      for (int q = 0; q < g.nranks; q++) {
        const uint32_t lo = __shfl_sync(0xffffffffu, wlo, q), hi = __shfl_sync(0xffffffffu, whi, q);
        tot += (q == g.rank) ? a : ll_unpack(lo, hi);
      }
      a = tot;
    }
    if (lane == 0) s_tot[k] = a;
  }
  cta_sync<NTC>();                                         /* [C] */
}


/* PF_ON: the TMA-ring variant (768 threads, HBM-resident populations); otherwise plain 128-bit loads (1024 threads), which is
 * faster while the working set sits in L2 (measured crossover: 5-6 M Kuramoto particles per GPU) */
#ifndef PERSIST_TMA_CTAS_PER_SM
#define PERSIST_TMA_CTAS_PER_SM 1
#endif
template <bool PF_ON>
__global__ void __launch_bounds__(PF_ON ? PT_TMA : PT_LD, PF_ON ? PERSIST_TMA_CTAS_PER_SM : ((1024 / PT_LD < 1) ? 1 : 1024 / PT_LD)) k_chain_persist(Pop p, PersistArgs g) {
  __shared__ double s_tot[NRED], s_bp[NRED], s_wp[(PF_ON ? TMA_CONS : PT_LD) / 32][NRED];
  __shared__ uint32_t s_words[PERSIST_MAX_CTAS * LLW];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ __align__(8) uint64_t s_full[TMA_MAXBUF], s_empty[TMA_MAXBUF], s_go;
  constexpr int NTC = PF_ON ? TMA_CONS : PT_LD;    /* threads that compute (all but the producer warp) */
  TmaRing ring_;
  ring_.buf = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(s_dyn) + 127) & ~(uintptr_t)127);
  ring_.base = smem_u32(ring_.buf); ring_.full = s_full; ring_.empty = s_empty; ring_.full_par = 0;
  ring_.full_a = smem_u32(s_full); ring_.empty_a = smem_u32(s_empty);
  if (PF_ON && threadIdx.x == 0) {
    for (int b = 0; b < TMA_MAXBUF; b++) { mbar_init(s_full + b, 1); mbar_init(s_empty + b, TMA_CONS / 32); }
    mbar_init(&s_go, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  double mfd[A1(NMFD)], mfd0[A1(NMFD)], mfd2[A1(NMFD)];
#pragma unroll
  for (int k = 0; k < A1(NMFD); k++) { mfd0[k] = 0.0; mfd2[k] = 0.0; }
  if (threadIdx.x < NRED) s_tot[threadIdx.x] = (g.sums_y != nullptr) ? g.sums_y[threadIdx.x] : 0.0;
  __shared__ unsigned long long *s_peer[MAX_RANKS];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < MAX_RANKS; r++) s_peer[r] = g.peer_ll[r];
  }
  ExchangeArgs xa;
  xa.ll_part = g.ll_part; xa.mail_ll = g.mail_ll; xa.peer_ll = s_peer; xa.abort_flag = g.abort_flag; xa.rank = g.rank; xa.nranks = g.nranks;
  if (USE_SCTAB) odx_sct_fill();
  __syncthreads();
  /* role by warp index, broadcast from lane 0 so that the compiler knows the branch is warp-uniform */
  const int warp_id = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (PF_ON && warp_id >= NTC / 32) {
    /* ---- producer warp: walks the same (step, stage) sequence as the consumers and keeps the ring full.  A stage may only be
     * fetched once this CTA's consumers have stored the previous one (same CTA, same particle ranges in every stage): s_go. */
    if (threadIdx.x == NTC) {
      const int64_t per_iter = (int64_t)gridDim.x * TMA_TILE;
      const int niter = (int)((p.n + per_iter - 1) / per_iter);
      uint32_t state = 0, go_par = 0, ep = g.epoch0;
      for (long long step = 0; step < g.nsteps; step++) {
        for (int s = 0; s < g.nst; s++) {
          const bool first = g.lean ? !lean_has_ys(s) : (s == 0);
          const bool use_acc = g.lean ? lean_has_acc(s) : (!first && !(s == 1 && g.recon));
          int ncol, nbuf; uint32_t tile_bytes;
          tma_geometry(first, use_acc, ncol, tile_bytes, nbuf);
          if (step != 0 || s != 0) { mbar_wait(&s_go, go_par); go_par ^= 1u; }
          asm volatile("fence.proxy.async;" ::: "memory");

          state = tma_produce(p.Y, p.E, p.PSI, first ? nullptr : p.YS, use_acc ? p.ACC : nullptr, p.stride, p.n, per_iter,
                              (int64_t)blockIdx.x * TMA_TILE, niter, (int)(ep & 1u), nbuf, tile_bytes, ring_.buf, s_full, s_empty, state);
          ep++;
        }
      }
    }
    return;
  }
  uint32_t epoch = g.epoch0;
  uint32_t eval = g.eval_index0;
  /* sharded population: k_peer_agree (same stream, just before this launch) has compared this rank's request with its peers' */
  if (g.nranks > 1 && g.abort_flag != nullptr && *(volatile int *)g.abort_flag != 0) return;
  for (long long step = 0; step < g.nsteps; step++) {
    const double t = g.t0 + (double)step * g.h;
    for (int s = 0; s < g.nst; s++) {
#pragma unroll
      for (int k = 0; k < NMFD; k++) mfd[k] = s_tot[k] / p.n_global;
      if (M::HAS_VECT_MFD) M::mfd_vect((NPAR >= 1 && NPAR <= MUV_MAX) ? p.muv : p.mu, p.n_global, mfd);
      const bool last_stage = (s == g.nst - 1);
      if (last_stage && step == g.nsteps - 1 && blockIdx.x == 0 && threadIdx.x < NRED) g.sums_report[threadIdx.x] = s_tot[threadIdx.x];
      /* no barrier here: s_tot is rewritten only behind barriers [A] and [B] of this stage's exchange */
      ChainArgs a;
      a.sums_in = nullptr; a.h = g.h; a.eval_index = eval; eval += 2; a.reverse = (int)(epoch & 1u);
      a.recon = 0.0; a.skip_acc_store = 0; a.iterate = g.iterate;
      a.want_out = (last_stage && step == g.nsteps - 1) ? g.want_out_last : 0;
      double red[NRED];
#pragma unroll
      for (int k = 0; k < NRED; k++) red[k] = 0.0;
      if (LEAN_OK && PF_ON && g.lean) {                          /* lean classic RK4: a.t_eval is the start of the step in every stage */
        a.t_eval = t; a.t_next = t; a.a = 0.0; a.w = 0.0;
        if (s == 0) {
#pragma unroll
          for (int k = 0; k < A1(NMFD); k++) mfd0[k] = mfd[k];
          lean_sweep<0, PF_ON>(p, a, mfd0, mfd2, mfd, red, ring_);
        } else if (s == 1) lean_sweep<1, PF_ON>(p, a, mfd0, mfd2, mfd, red, ring_);
        else if (s == 2) {
#pragma unroll
          for (int k = 0; k < A1(NMFD); k++) mfd2[k] = mfd[k];
          lean_sweep<2, PF_ON>(p, a, mfd0, mfd2, mfd, red, ring_);
        } else lean_sweep<3, PF_ON>(p, a, mfd0, mfd2, mfd, red, ring_);
      } else if (g.nst == 1) {                            /* Euler */
        a.t_eval = t; a.t_next = t + g.h; a.a = g.h; a.w = 1.0;
        chain_sweep<true, true, PF_ON, USE_SCTAB>(p, a, mfd, red, ring_);
      } else if (s == 0) {                                /* RK4 (methods_odexp.c:1846-1863) */
        a.t_eval = t; a.t_next = t + g.h / 2.0; a.a = g.h / 2.0; a.w = 1.0; a.skip_acc_store = g.recon ? 1 : 0;
        chain_sweep<true, false, PF_ON, USE_SCTAB>(p, a, mfd, red, ring_);
      } else if (s == 1) {
        a.t_eval = t + g.h / 2.0; a.t_next = t + g.h / 2.0; a.a = g.h / 2.0; a.w = 2.0; a.recon = g.recon ? 1.0 / (g.h / 2.0) : 0.0;
        chain_sweep<false, false, PF_ON, USE_SCTAB>(p, a, mfd, red, ring_);
      } else if (s == 2) {
        a.t_eval = t + g.h / 2.0; a.t_next = t + g.h; a.a = g.h; a.w = 2.0;
        chain_sweep<false, false, PF_ON, USE_SCTAB>(p, a, mfd, red, ring_);
      } else {
        a.t_eval = t + g.h; a.t_next = t + g.h; a.a = g.h / 6.0; a.w = 1.0;
        chain_sweep<false, true, PF_ON, USE_SCTAB>(p, a, mfd, red, ring_);
      }
#ifndef ODX_RING_NEW_XCHG
#define ODX_RING_NEW_XCHG 0
#endif
      if (!PF_ON || ODX_RING_NEW_XCHG) {
        stage_exchange<NTC, PF_ON>(red, xa, epoch, s_tot, s_wp, s_words, &s_go);
      } else {
        /* ---- mean-field barrier without a barrier, low-latency words.  Every exchanged double travels as two 8-byte words
         * {32 data bits, 32-bit epoch tag}: an 8-byte store is atomic, so data and "ready" flag arrive together and no fence
         * or separate flag round trip is needed (the LL idea of NCCL's small-message protocol).  Every CTA publishes its
         * block partials as such words, every CTA polls the words of all CTAs and adds the partials itself in a fixed order
         * (identical bits everywhere, no atomics).  Buffers alternate with the epoch parity: a CTA is one epoch ahead at most. */
        const size_t par = (size_t)(epoch & 1u);
        const uint32_t tag = epoch + 1u;
        unsigned long long *ll_part = g.ll_part + par * (size_t)PERSIST_MAX_CTAS * LLW;
        block_partial_smem<NTC>(red, s_bp);
        if (PF_ON && threadIdx.x == 0) mbar_arrive(&s_go);     /* every consumer has stored (and fenced) this stage */
        if (threadIdx.x < LLW) ll_store_gpu(ll_part + (size_t)blockIdx.x * LLW + threadIdx.x, ll_pack(s_bp[threadIdx.x >> 1], threadIdx.x & 1, tag));
        if (ODX_LL_POLL16) {
          for (unsigned int idx = threadIdx.x; idx < gridDim.x * NRED; idx += NTC) {
            unsigned long long w0, w1; unsigned long long spins = 0;
            for (;;) {
              ll_load2_gpu(ll_part + 2 * idx, w0, w1);
              if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag) break;
              if (spin_give_up(spins, g.abort_flag)) break;
            }
            s_words[2 * idx] = (uint32_t)w0; s_words[2 * idx + 1] = (uint32_t)w1;
          }
        } else {
          for (unsigned int idx = threadIdx.x; idx < gridDim.x * LLW; idx += NTC) {
            unsigned long long w; unsigned long long spins = 0;
            while ((uint32_t)((w = ll_load_gpu(ll_part + idx)) >> 32) != tag) { if (spin_give_up(spins, g.abort_flag)) break; }
            s_words[idx] = (uint32_t)w;
          }
        }
        cta_sync<NTC>();
        {
          double acc[NRED];
  #pragma unroll
          for (int k = 0; k < NRED; k++) acc[k] = 0.0;
          for (unsigned int b2 = threadIdx.x; b2 < gridDim.x; b2 += NTC) {
  #pragma unroll
            for (int k = 0; k < NRED; k++) acc[k] += ll_unpack(s_words[b2 * LLW + 2 * k], s_words[b2 * LLW + 2 * k + 1]);
          }
          block_total<NTC>(acc, s_tot);
        }
        if (g.nranks > 1) {
          /* cross-GPU: CTA 0 stores this rank's sums as LL words straight into every peer's mailbox over NVLink (8-byte
           * st.relaxed.sys on IPC-mapped pointers), polls the words of the other ranks in its own GPU's mailbox, adds the rank
           * sums in rank order and republishes the global sums to the other CTAs of this GPU, again as LL words */
          /* every CTA polls the peers' words itself (no local republish hop) */
          if (blockIdx.x == 0 && threadIdx.x < LLW) {
            const unsigned long long w = ll_pack(s_tot[threadIdx.x >> 1], threadIdx.x & 1, tag);
            for (int r = 0; r < g.nranks; r++)
              if (r != g.rank) ll_store_sys(g.peer_ll[r] + (par * MAX_RANKS + g.rank) * LLW + threadIdx.x, w);
          }
          if (threadIdx.x < (unsigned)(g.nranks * LLW)) {
            const int r = threadIdx.x / LLW, k = threadIdx.x % LLW;
            if (r != g.rank) {
              unsigned long long w; unsigned long long spins = 0;
              while ((uint32_t)((w = ll_load_sys(g.mail_ll + (par * MAX_RANKS + r) * LLW + k)) >> 32) != tag) {
                if (spin_give_up(spins, g.abort_flag)) break;
              }
              s_words[threadIdx.x] = (uint32_t)w;
            }
          }
          cta_sync<NTC>();
          if (threadIdx.x < NRED) {
            double v = 0.0;
            for (int r = 0; r < g.nranks; r++)
              v += (r == g.rank) ? s_tot[threadIdx.x] : ll_unpack(s_words[r * LLW + 2 * threadIdx.x], s_words[r * LLW + 2 * threadIdx.x + 1]);
            s_bp[threadIdx.x] = v;
          }
          cta_sync<NTC>();
          if (threadIdx.x < NRED) s_tot[threadIdx.x] = s_bp[threadIdx.x];
          cta_sync<NTC>();
        }
      }
      epoch++;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < NRED) g.sums_new[threadIdx.x] = s_tot[threadIdx.x];
}

/* full RHS evaluation with outputs: aux, rates and (optionally) f into fout[nvar][cap] */
struct RhsArgs { const double *sums_in; double t; double h; uint32_t eval_index; double *fout; };
__global__ void __launch_bounds__(BLOCK) k_rhs_out(Pop p, const double *__restrict__ ysrc, RhsArgs a, Red r) {
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; k++) red[k] = 0.0;
  double mfd[A1(NMFD)];
  load_mfd(p, a.sums_in, p.n_global, mfd);
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
    double y[A1(NVAR)], e[A1(NEXPR)], psi[A1(NPSI)], f[A1(NVAR)], aux[A1(NAUX)];
    double death = 0.0, repli = 0.0;
    load_cols<NVAR>(ysrc, p.stride, i, y);
    load_cols<NEXPR>(p.E, p.stride, i, e);
    if (NPSI > 0) load_cols<NPSI>(p.PSI, p.stride, i, psi);
    DevCtx c = make_ctx(p, a.h, a.eval_index, DOM_RHS, M::NEEDS_ID ? p.ID[i] : 0);
    ctx_neighbours(c, p, ysrc, i);
    rhs_checked(p, c, a.t, y, e, psi, mfd, f, aux, death, repli);
    store_cols<NAUX>(p.AUX, p.stride, i, aux);
    if (M::HAS_DEATH) p.DEATH[i] = death;
    if (M::HAS_REPLI) p.REPLI[i] = repli;
    if (a.fout) store_cols<NVAR>(a.fout, p.stride, i, f);
    red[RED_RATES] += death + repli;
  }
  if (M::HAS_DEATH || M::HAS_REPLI) grid_reduce<NRED>(red, r);
}

/* O(N^2) %C coupling: psi_k(i) = (1/POP_SIZE) sum_j expr_k(i,j) (src/odexp:922-950).
 * Grid = (i-tiles of 256 particles) x (j-segments).  Two i-particles per thread live in registers (two independent FP64
 * dependency chains), j-particles are staged through shared memory in tiles of PAIR_TJ and broadcast to the CTA.  Every
 * CTA writes the raw partial sums of its (i-tile, j-segment); k_pair_finish adds the segments in ascending j order and
 * divides by POP_SIZE, so the result is deterministic and the j-segments of other GPUs (allgathered columns) are just
 * more segments.  Arithmetic is FP64 FMA on the CUDA cores. */
#ifndef PAIR_THREADS_N
#define PAIR_THREADS_N 128
#endif
#ifndef PAIR_TJ_N
#define PAIR_TJ_N 128
#endif
#ifndef PAIR_UNROLL_N
#define PAIR_UNROLL_N 4
#endif
#ifndef PAIR_WAVES
#define PAIR_WAVES 16      /* CTAs per resident slot: the tail of the last wave costs <= 1 / PAIR_WAVES of the launch */
#endif
constexpr int PAIR_THREADS = PAIR_THREADS_N, PAIR_ITILE = 2 * PAIR_THREADS_N, PAIR_TJ = PAIR_TJ_N, PAIR_UNROLL = PAIR_UNROLL_N;
constexpr int PAIR_NCOL = NVAR + (M::CPL_USES_EXPR ? NEXPR : 0);       /* columns of the other particles read from memory (and allgathered) */
/* OA() inside %C: the other particle's auxiliary values are evaluated once per staged j-particle and kept next to its columns in
 * the shared-memory tile (1 / PAIR_ITILE of the pair work); the emitter admits only auxiliaries that read neither mean fields
 * nor couplings, for which the reference's pass-1 and pass-2 values coincide */
constexpr int PAIR_SCOL = PAIR_NCOL + (M::CPL_USES_OAUX ? NAUX : 0);
static_assert(sizeof(double) * (PAIR_SCOL > 0 ? PAIR_SCOL : 1) * PAIR_TJ <= 40 * 1024, "j-tile of the %C kernel exceeds static shared memory");
/* Fast path + tile-local redo: the straight-line fast forms of the model code only signal when they cannot represent a term
 * (range-limited sincos: the context's slow flag; odx_zmul: a non-finite sum); the thread then discards what this j-tile added
 * to its sums and repeats the tile with the exact forms.  The self term of a 1/r^n kernel does this once per i-particle (1 tile
 * in N / PAIR_TJ). */
__device__ __forceinline__ void pair_rows(const Pop &p, const double *__restrict__ jy, const double *__restrict__ je,
                                          const int32_t *__restrict__ jid, int64_t jstride, int64_t jbeg, int64_t jend, double t,
                                          const double (&ya)[A1(NVAR)], const double (&ea)[A1(NEXPR)], const double (&yb)[A1(NVAR)],
                                          const double (&eb)[A1(NEXPR)], int32_t ida, int32_t idb, double (&acca)[A1(NPSI)],
                                          double (&accb)[A1(NPSI)], double (*s_col)[PAIR_TJ], int32_t *s_id,
                                          const double (&auxa)[A1(NAUX)], const double (&auxb)[A1(NAUX)]) {
  double aux[A1(NAUX)];      /* the other particle's auxiliaries (OA()), filled from the tile when the model reads them */
#pragma unroll
  for (int k = 0; k < A1(NPSI); k++) { acca[k] = 0.0; accb[k] = 0.0; }
  DevCtx ca = make_ctx(p, 0.0, 0, DOM_RHS, ida);
  DevCtx cb = make_ctx(p, 0.0, 0, DOM_RHS, idb);
  for (int64_t j0 = jbeg; j0 < jend; j0 += PAIR_TJ) {
    const int tj = (int)((jend - j0) < PAIR_TJ ? (jend - j0) : PAIR_TJ);
    __syncthreads();
    for (int jj = threadIdx.x; jj < tj; jj += PAIR_THREADS) {
#pragma unroll
      for (int v = 0; v < NVAR; v++) s_col[v][jj] = jy[(int64_t)v * jstride + j0 + jj];
      if (M::CPL_USES_EXPR) {
#pragma unroll
        for (int v = 0; v < NEXPR; v++) s_col[NVAR + v][jj] = je[(int64_t)v * jstride + j0 + jj];
      }
      if (M::NEEDS_ID) s_id[jj] = jid[j0 + jj];
      if (M::CPL_USES_OAUX) {
        double oyv[A1(NVAR)], oev[A1(NEXPR)], zp[A1(NPSI)], zm[A1(NMFD)], fo[A1(NVAR)], oax[A1(NAUX)], d = 0.0, r = 0.0;
#pragma unroll
        for (int v = 0; v < NVAR; v++) oyv[v] = s_col[v][jj];
#pragma unroll
        for (int v = 0; v < NEXPR; v++) oev[v] = s_col[NVAR + v][jj];          /* CPL_USES_EXPR is set together with CPL_USES_OAUX */
#pragma unroll
        for (int k = 0; k < A1(NPSI); k++) zp[k] = 0.0;
#pragma unroll
        for (int k = 0; k < A1(NMFD); k++) zm[k] = 0.0;
        DevCtx co = make_ctx(p, 0.0, 0, DOM_RHS, M::NEEDS_ID ? jid[j0 + jj] : 0);
        M::template rhs<true>(co, t, oyv, oev, zp, zm, fo, oax, d, r);
#pragma unroll
        for (int v = 0; v < NAUX; v++) s_col[PAIR_NCOL + v][jj] = oax[v];
      }
    }
    __syncthreads();
    double sava[A1(NPSI)], savb[A1(NPSI)];
#pragma unroll
    for (int k = 0; k < A1(NPSI); k++) { sava[k] = acca[k]; savb[k] = accb[k]; }
    ca.slow = 0; cb.slow = 0;
#pragma unroll PAIR_UNROLL
    for (int jj = 0; jj < tj; jj++) {
      double oy[A1(NVAR)], oe[A1(NEXPR)];
#pragma unroll
      for (int v = 0; v < NVAR; v++) oy[v] = s_col[v][jj];
      if (M::CPL_USES_EXPR) {
#pragma unroll
        for (int v = 0; v < NEXPR; v++) oe[v] = s_col[NVAR + v][jj];
      }
      if (M::NEEDS_ID) { ca.oid = s_id[jj]; cb.oid = s_id[jj]; }
      if (M::CPL_USES_OAUX) {
#pragma unroll
        for (int v = 0; v < NAUX; v++) aux[v] = s_col[PAIR_NCOL + v][jj];
      }
      M::template cpl_terms<false>(ca, t, ya, ea, auxa, oy, oe, aux, acca);
      M::template cpl_terms<false>(cb, t, yb, eb, auxb, oy, oe, aux, accb);
    }
    int redo = ca.slow | cb.slow;
    if (M::CPL_POISON) {
#pragma unroll
      for (int k = 0; k < NPSI; k++)
        redo |= ((__double2hiint(acca[k]) & 0x7ff00000) == 0x7ff00000) | ((__double2hiint(accb[k]) & 0x7ff00000) == 0x7ff00000);
    }
    if (redo) {
#pragma unroll
      for (int k = 0; k < A1(NPSI); k++) { acca[k] = sava[k]; accb[k] = savb[k]; }
#pragma unroll 1
      for (int jj = 0; jj < tj; jj++) {
        double oy[A1(NVAR)], oe[A1(NEXPR)];
#pragma unroll
        for (int v = 0; v < NVAR; v++) oy[v] = s_col[v][jj];
        if (M::CPL_USES_EXPR) {
#pragma unroll
          for (int v = 0; v < NEXPR; v++) oe[v] = s_col[NVAR + v][jj];
        }
        if (M::NEEDS_ID) { ca.oid = s_id[jj]; cb.oid = s_id[jj]; }
        if (M::CPL_USES_OAUX) {
#pragma unroll
          for (int v = 0; v < NAUX; v++) aux[v] = s_col[PAIR_NCOL + v][jj];
        }
        M::template cpl_terms<true>(ca, t, ya, ea, auxa, oy, oe, aux, acca);
        M::template cpl_terms<true>(cb, t, yb, eb, auxb, oy, oe, aux, accb);
      }
    }
  }
}
__global__ void __launch_bounds__(PAIR_THREADS) k_pair(Pop p, const double *__restrict__ ysrc, const double *__restrict__ jy,
                                                       const double *__restrict__ je, const int32_t *__restrict__ jid, int64_t jstride,
                                                       int64_t nj, double t, double *__restrict__ part, int seg_base,
                                                       const int *__restrict__ only_if) {
  __shared__ double s_col[A1(PAIR_SCOL)][PAIR_TJ];
  __shared__ int32_t s_id[PAIR_TJ];
  if (only_if != nullptr && *only_if == 0) return;      /* ordered exact pass after k_pair_sym: runs only when that kernel met a non-finite term */
  /* this CTA's j-range: segment blockIdx.y of gridDim.y, boundaries aligned to the j-tile */
  const int64_t tiles = (nj + PAIR_TJ - 1) / PAIR_TJ;
  const int64_t tps = (tiles + gridDim.y - 1) / gridDim.y;
  const int64_t jbeg = (int64_t)blockIdx.y * tps * PAIR_TJ;
  int64_t jend = jbeg + tps * PAIR_TJ; if (jend > nj) jend = nj;
  const int64_t ia = (int64_t)blockIdx.x * PAIR_ITILE + threadIdx.x, ib = ia + PAIR_THREADS;
  const bool la = ia < p.n, lb = ib < p.n;
  const int64_t ja = la ? ia : p.n - 1, jb = lb ? ib : p.n - 1;
  double ya[A1(NVAR)], ea[A1(NEXPR)], yb[A1(NVAR)], eb[A1(NEXPR)], acca[A1(NPSI)], accb[A1(NPSI)];
  load_cols<NVAR>(ysrc, p.stride, ja, ya); load_cols<NEXPR>(p.E, p.stride, ja, ea);
  load_cols<NVAR>(ysrc, p.stride, jb, yb); load_cols<NEXPR>(p.E, p.stride, jb, eb);
  const int32_t ida = M::NEEDS_ID ? p.ID[ja] : 0, idb = M::NEEDS_ID ? p.ID[jb] : 0;
  double auxa[A1(NAUX)], auxb[A1(NAUX)];
  if (M::CPL_USES_AUX) {    /* MA() inside %C: what pass 1 of pop_rhs left in myself_->aux = ode_rhs with zeroed mean fields and the psi
                             * of the previous evaluation (still in the PSI columns: k_pair_finish runs after this kernel) */
    double psia[A1(NPSI)], psib[A1(NPSI)], mz[A1(NMFD)], f[A1(NVAR)], d = 0.0, r = 0.0;
    load_cols<NPSI>(p.PSI, p.stride, ja, psia); load_cols<NPSI>(p.PSI, p.stride, jb, psib);
#pragma unroll
    for (int k = 0; k < A1(NMFD); k++) mz[k] = 0.0;
    DevCtx ca = make_ctx(p, 0.0, 0, DOM_RHS, ida), cb = make_ctx(p, 0.0, 0, DOM_RHS, idb);
    M::template rhs<true>(ca, t, ya, ea, psia, mz, f, auxa, d, r);
    M::template rhs<true>(cb, t, yb, eb, psib, mz, f, auxb, d, r);
  }
  pair_rows(p, jy, je, jid, jstride, jbeg, jend, t, ya, ea, yb, eb, ida, idb, acca, accb, s_col, s_id, auxa, auxb);
  if (M::CPL_SCALED) {      /* factors that do not depend on the other particle, hoisted out of the sum by the emitter */
    double sa[A1(NPSI)], sb[A1(NPSI)];
    DevCtx ca = make_ctx(p, 0.0, 0, DOM_RHS, ida), cb = make_ctx(p, 0.0, 0, DOM_RHS, idb);
    M::template cpl_scale<true>(ca, t, ya, ea, auxa, sa);
    M::template cpl_scale<true>(cb, t, yb, eb, auxb, sb);
#pragma unroll
    for (int k = 0; k < NPSI; k++) { acca[k] *= sa[k]; accb[k] *= sb[k]; }
  }
  double *dst = part + (size_t)(seg_base + blockIdx.y) * A1(NPSI) * p.stride;
#pragma unroll
  for (int k = 0; k < NPSI; k++) {
    if (la) dst[(int64_t)k * p.stride + ia] = acca[k];
    if (lb) dst[(int64_t)k * p.stride + ib] = accb[k];
  }
}
/* psi = (sum over segments, ascending j) / POP_SIZE; only the pairwise couplings are touched (a low-rank psi keeps its column) */
__global__ void __launch_bounds__(BLOCK) k_pair_finish(Pop p, const double *__restrict__ part, int nseg, double *__restrict__ psi,
                                                       const int *__restrict__ only_if) {
  if (only_if != nullptr && *only_if == 0) return;
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
#pragma unroll
    for (int k = 0; k < NPSI; k++) {
      if (!M::psi_is_pairwise(k)) continue;
      double a = 0.0;
      for (int s2 = 0; s2 < nseg; s2++) a += part[((size_t)s2 * A1(NPSI) + k) * p.stride + i];
      psi[(int64_t)k * p.stride + i] = a / p.n_global;
    }
  }
}

#ifdef ODX_MODEL_CPL_SYM
/* ------------------------------------------------------------------------------------------------
 * Symmetric all-pairs %C (Newton's third law) for couplings the emitter has proven antisymmetric bit for bit
 * (OdxModel::CPL_SYM: term_k(i,j) = (y_i - y_j) R with R even in every difference, so term_k(j,i) = -term_k(i,j) exactly):
 * every unordered pair is evaluated once and added to both particles -- for the N-body coupling 18 FP64 instructions per
 * unordered pair (3 differences, 3 for r^2, 6 for r^-3, 3 + 3 FMAs) = 9 per ordered pair instead of the 15 of k_pair.
 *
 * Work item = one pair of tiles (A <= B) of PSYM_TILE particles, one CTA.  Warp w owns i-superblock w of tile A (32 x PSYM_PI
 * particles, PSYM_PI per lane, in registers with their sums for the whole work item) and walks over the j-superblocks of tile B
 * (32 x PSYM_PJ particles) in a staggered order, so that at any time every j-superblock is held by one warp only.  Inside a
 * (i-superblock, j-superblock) block the lanes rotate over the j-groups: at rotation step r lane l reads the positions of
 * j-group (l + r) mod 32 from the warp's shared-memory copy (consecutive lanes, consecutive addresses: conflict-free) and holds
 * that group's running sums in registers; after the step the sums move to the next lane by one shuffle each, and after 32 steps
 * they are back in their home lane.  Register blocking PSYM_PI x PSYM_PJ is what makes the j-side cheap: 6 PSYM_PJ shuffles and
 * PSYM_NYC PSYM_PJ shared loads per PSYM_PI PSYM_PJ pairs.  The j-side sums of a block are then added into the tile's
 * shared-memory accumulators (exclusive owner, fixed order), the i-side sums stay in registers.  A tile pair writes two blocks
 * of raw partial sums, part[A][B] (i-side) and part[B][A] (j-side); the diagonal pair A = B evaluates all ordered pairs through
 * the i-side only, skipping i = j.  k_pair_sym_finish adds the partial blocks of a tile in ascending order of the other tile:
 * the result is deterministic (fixed schedule, no atomics) but the order of summation differs from k_pair's ascending-j order,
 * i.e. from the reference's by rounding (tests: <= 1e-12 relative on psi).  Blocks that contain the end of the population or the
 * diagonal take a masked form.  A non-finite term (coincident particles: the fast term is 0 x inf) poisons the sums; the finish
 * kernel raises a flag, and the ordered kernels (which redo such tiles with the exact forms) then run over the whole
 * population -- they are launched after every symmetric evaluation and return at once while the flag is down. */
/* measured at N = 2^17 (profiles/r2p_pair_sym_variants.txt), PI x PJ / CTAs per SM: 8x2/1 10.27 ms, 8x4/1 10.50, 8x1/2 10.59,
 * 4x2/2 10.65, 4x4/2 10.72, 4x4/1 10.97; ordered kernel 15.76 */
#ifndef PSYM_PI
#define PSYM_PI 8
#endif
#ifndef PSYM_PJ
#define PSYM_PJ 2
#endif
#ifndef PSYM_MINB
#define PSYM_MINB 1
#endif
constexpr int PSYM_TILE = 1024, PSYM_SI = 32 * PSYM_PI, PSYM_SJ = 32 * PSYM_PJ, PSYM_W = PSYM_TILE / PSYM_SI, PSYM_NJ = PSYM_TILE / PSYM_SJ;
constexpr int PSYM_THREADS = 32 * PSYM_W;
static_assert(PSYM_TILE % PSYM_SI == 0 && PSYM_TILE % PSYM_SJ == 0 && PSYM_NJ % PSYM_W == 0, "k_pair_sym: j-superblocks must not be larger than i-superblocks");
constexpr int psym_count_cols() { int c = 0; for (int v = 0; v < NVAR; v++) c += (M::CPL_SYM_YMASK >> v) & 1u; return c; }
constexpr int PSYM_NYC = psym_count_cols();                       /* dynamical variables the terms read */
constexpr int psym_col(int c) { int k = 0; for (int v = 0; v < NVAR; v++) if ((M::CPL_SYM_YMASK >> v) & 1u) { if (k == c) return v; k++; } return 0; }
constexpr size_t PSYM_SMEM = sizeof(double) * ((size_t)PSYM_W * PSYM_NYC * PSYM_SJ + (size_t)A1(NPSI) * PSYM_TILE);

/* Sharded populations: the tiles of all ranks form one list of slots (rank order, tile_base[q] = first slot of rank q); a rank
 * evaluates the pairs of its own tiles with each other as above, and HALF of the pairs (own tile a, tile b of rank q): those with
 * a + b even when its rank is the lower one, odd when it is the higher one -- the other rank evaluates the rest, with the roles of
 * the two tiles exchanged.  The j-side sums of such a pair belong to the other rank: they are written to `part_remote`, summed
 * over the own tiles (k_pair_sym_remote_sum, ascending tile order) and sent to their owners by grouped ncclSend / ncclRecv
 * (3 x N / G doubles per pair of ranks); k_pair_sym_finish adds, in ascending rank order, either the own partial blocks or what
 * the peers sent.  Each rank does N^2 / (2 G) pair evaluations instead of N^2 / G. */
struct PsymArgs {
  const double *jsrc;            /* column c of rank q's particles at jsrc + q * jrank_stride + c * jcol_stride */
  int64_t jcol_stride, jrank_stride;
  int rank, nranks, nslots;      /* nslots = tiles of all ranks */
  int tile_base[MAX_RANKS + 1];
  int64_t all_n[MAX_RANKS];
  double *part;                  /* [own tile][slot][NPSI][PSYM_TILE]: sums for the own tile from the tile in `slot` */
  double *part_remote;           /* [slot][own tile][NPSI][PSYM_TILE]: sums for the (remote) tile in `slot` from the own tile */
};
__device__ __forceinline__ bool psym_mine(int rank, int q, int a, int b) {      /* which of the two owner ranks evaluates the pair of tiles */
  return ((a + b) & 1) == (rank < q ? 0 : 1);
}
template <bool MASKED>
__device__ __forceinline__ void psym_block(DevCtx &c, double t, const double (&yi)[PSYM_PI][A1(NVAR)], double (&ai)[PSYM_PI][A1(NPSI)],
                                           const double (&yj)[PSYM_PJ][A1(NVAR)], double (&aj)[PSYM_PJ][A1(NPSI)], int64_t gi0, int64_t gj0,
                                           int64_t ni, int64_t nj, bool same_list) {
#pragma unroll
  for (int u = 0; u < PSYM_PI; u++) {
#pragma unroll
    for (int v = 0; v < PSYM_PJ; v++) {
      if (MASKED) {
        const int64_t gi = gi0 + 32 * u, gj = gj0 + 32 * v;
        if (gi < ni && gj < nj && !(same_list && gi == gj)) M::template cpl_terms_sym<false>(c, t, yi[u], yj[v], ai[u], aj[v]);
      } else {
        M::template cpl_terms_sym<false>(c, t, yi[u], yj[v], ai[u], aj[v]);
      }
    }
  }
}
template <bool MASKED>
__device__ __forceinline__ void psym_rotate(DevCtx &c, double t, const double (&yi)[PSYM_PI][A1(NVAR)], double (&ai)[PSYM_PI][A1(NPSI)],
                                            double (&aj)[PSYM_PJ][A1(NPSI)], const double *__restrict__ s_pos, int lane, int64_t gi0,
                                            int64_t jbase, int64_t ni, int64_t nj, bool same_list) {
#pragma unroll 1
  for (int r = 0; r < 32; r++) {
    const int src = (lane + r) & 31;
    double yj[PSYM_PJ][A1(NVAR)];
#pragma unroll
    for (int v = 0; v < PSYM_PJ; v++) {
#pragma unroll
      for (int cc = 0; cc < PSYM_NYC; cc++) yj[v][psym_col(cc)] = s_pos[cc * PSYM_SJ + src + 32 * v];
    }
    psym_block<MASKED>(c, t, yi, ai, yj, aj, gi0, jbase + src, ni, nj, same_list);
    const int from = (lane + 1) & 31;
#pragma unroll
    for (int v = 0; v < PSYM_PJ; v++) {
#pragma unroll
      for (int k = 0; k < NPSI; k++) aj[v][k] = __shfl_sync(0xffffffffu, aj[v][k], from);
    }
  }
}
__global__ void __launch_bounds__(PSYM_THREADS, PSYM_MINB) k_pair_sym(Pop p, const double *__restrict__ ysrc, double t, PsymArgs g) {
  const int A = blockIdx.y, slot = blockIdx.x;
  int q = 0;
  while (q + 1 < g.nranks && slot >= g.tile_base[q + 1]) q++;
  const int B = slot - g.tile_base[q];
  const bool local = (q == g.rank);
  if (local ? (B < A) : !psym_mine(g.rank, q, A, B)) return;
  extern __shared__ __align__(16) double s_psym[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *s_pos = s_psym + (size_t)w * PSYM_NYC * PSYM_SJ;                    /* this warp's copy of the current j-superblock */
  double *s_acc = s_psym + (size_t)PSYM_W * PSYM_NYC * PSYM_SJ;               /* [NPSI][PSYM_TILE] j-side sums of tile B */
  const int64_t ni = p.n, nj = g.all_n[q];
  const double *__restrict__ jsrc = g.jsrc + (int64_t)q * g.jrank_stride;
  const int64_t ibase = (int64_t)A * PSYM_TILE + (int64_t)w * PSYM_SI;
  double yi[PSYM_PI][A1(NVAR)], ai[PSYM_PI][A1(NPSI)];
#pragma unroll
  for (int u = 0; u < PSYM_PI; u++) {
    int64_t gi = ibase + lane + 32 * u; if (gi > ni - 1) gi = ni - 1;
#pragma unroll
    for (int cc = 0; cc < PSYM_NYC; cc++) yi[u][psym_col(cc)] = ysrc[(int64_t)psym_col(cc) * p.stride + gi];
#pragma unroll
    for (int k = 0; k < A1(NPSI); k++) ai[u][k] = 0.0;
  }
  for (int idx = threadIdx.x; idx < A1(NPSI) * PSYM_TILE; idx += PSYM_THREADS) s_acc[idx] = 0.0;
  __syncthreads();
  DevCtx c = make_ctx(p, 0.0, 0, DOM_RHS, 0);
  const bool diag = local && (A == B);
#pragma unroll 1
  for (int s = 0; s < PSYM_NJ; s++) {
    const int jb = (w * (PSYM_NJ / PSYM_W) + s) % PSYM_NJ;
    const int64_t jbase = (int64_t)B * PSYM_TILE + (int64_t)jb * PSYM_SJ;
#pragma unroll
    for (int v = 0; v < PSYM_PJ; v++) {
      int64_t gj = jbase + lane + 32 * v; if (gj > nj - 1) gj = nj - 1;
#pragma unroll
      for (int cc = 0; cc < PSYM_NYC; cc++) s_pos[cc * PSYM_SJ + lane + 32 * v] = jsrc[(int64_t)psym_col(cc) * g.jcol_stride + gj];
    }
    __syncwarp();
    double aj[PSYM_PJ][A1(NPSI)];
#pragma unroll
    for (int v = 0; v < PSYM_PJ; v++) {
#pragma unroll
      for (int k = 0; k < A1(NPSI); k++) aj[v][k] = 0.0;
    }
    const bool masked = (jbase + PSYM_SJ > nj) || (ibase + PSYM_SI > ni) || (diag && jbase < ibase + PSYM_SI && ibase < jbase + PSYM_SJ);
    if (jbase < nj && ibase < ni) {
      if (masked) psym_rotate<true>(c, t, yi, ai, aj, s_pos, lane, ibase + lane, jbase, ni, nj, local);
      else psym_rotate<false>(c, t, yi, ai, aj, s_pos, lane, ibase + lane, jbase, ni, nj, local);
    }
    if (!diag) {
#pragma unroll
      for (int v = 0; v < PSYM_PJ; v++) {
#pragma unroll
        for (int k = 0; k < NPSI; k++) s_acc[k * PSYM_TILE + jb * PSYM_SJ + lane + 32 * v] += aj[v][k];
      }
    }
    __syncthreads();
  }
  double *pa = g.part + ((size_t)A * g.nslots + slot) * A1(NPSI) * PSYM_TILE;
#pragma unroll
  for (int u = 0; u < PSYM_PI; u++) {
#pragma unroll
    for (int k = 0; k < NPSI; k++) pa[(size_t)k * PSYM_TILE + w * PSYM_SI + lane + 32 * u] = ai[u][k];
  }
  if (!diag) {
    double *pb = local ? g.part + ((size_t)B * g.nslots + g.tile_base[g.rank] + A) * A1(NPSI) * PSYM_TILE
                       : g.part_remote + ((size_t)slot * gridDim.y + A) * A1(NPSI) * PSYM_TILE;
    for (int idx = threadIdx.x; idx < NPSI * PSYM_TILE; idx += PSYM_THREADS) pb[idx] = s_acc[idx];
  }
}
/* sharded: out[slot] = sum over the own tiles a that evaluated the pair (a, tile in slot), ascending a; what is sent to the owner */
__global__ void __launch_bounds__(BLOCK) k_pair_sym_remote_sum(PsymArgs g, int nown, double *__restrict__ out) {
  const int slot = blockIdx.x;
  int q = 0;
  while (q + 1 < g.nranks && slot >= g.tile_base[q + 1]) q++;
  if (q == g.rank) return;
  const int B = slot - g.tile_base[q];
  for (int idx = threadIdx.x; idx < NPSI * PSYM_TILE; idx += BLOCK) {
    double a = 0.0;
    for (int A = 0; A < nown; A++)
      if (psym_mine(g.rank, q, A, B)) a += g.part_remote[((size_t)slot * nown + A) * A1(NPSI) * PSYM_TILE + idx];
    out[(size_t)slot * A1(NPSI) * PSYM_TILE + idx] = a;
  }
}
/* psi_k(i) = scale_k(i) * (own partial blocks and the peers' contributions, ascending rank and tile) / POP_SIZE; a non-finite
 * result raises *poison (the ordered kernels then redo the evaluation with the exact forms).  recv: [rank][own tile][NPSI][PSYM_TILE] */
__global__ void __launch_bounds__(BLOCK) k_pair_sym_finish(Pop p, const double *__restrict__ ysrc, double t, PsymArgs g, int nown,
                                                           const double *__restrict__ recv, double *__restrict__ psi, int *__restrict__ poison) {
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
    const int64_t T = i / PSYM_TILE, il = i % PSYM_TILE;
    double sc[A1(NPSI)];
#pragma unroll
    for (int k = 0; k < A1(NPSI); k++) sc[k] = 1.0;
    if (M::CPL_SCALED) {
      double y[A1(NVAR)], e[A1(NEXPR)], aux[A1(NAUX)];
      load_cols<NVAR>(ysrc, p.stride, i, y); load_cols<NEXPR>(p.E, p.stride, i, e);
      DevCtx c = make_ctx(p, 0.0, 0, DOM_RHS, M::NEEDS_ID ? p.ID[i] : 0);
      M::template cpl_scale<true>(c, t, y, e, aux, sc);
    }
    int bad = 0;
#pragma unroll
    for (int k = 0; k < NPSI; k++) {
      double a = 0.0;
      const double *src = g.part + ((size_t)T * g.nslots * A1(NPSI) + k) * PSYM_TILE + il;
      for (int q = 0; q < g.nranks; q++) {
        const int nb = g.tile_base[q + 1] - g.tile_base[q];
        for (int b = 0; b < nb; b++)
          if (q == g.rank || psym_mine(g.rank, q, (int)T, b)) a += src[(size_t)(g.tile_base[q] + b) * A1(NPSI) * PSYM_TILE];
        if (q != g.rank && nb > 0) a += recv[(((size_t)q * nown + T) * A1(NPSI) + k) * PSYM_TILE + il];
      }
      a = a * sc[k] / p.n_global;
      bad |= ((__double2hiint(a) & 0x7ff00000) == 0x7ff00000);
      psi[(int64_t)k * p.stride + i] = a;
    }
    if (bad) *poison = 1;
  }
}
#endif

/* embedded Runge-Kutta pairs (GSL rkf45 / rkck semantics), general tableau with stored slopes.
 * Kernel for stage s (2..6): k_{s-1} = rhs(y_{s-1}); y_s = y0 + h sum_{j<s} b_{s,j} k_j; %M terms of y_s.
 * Kernel s = 7 ("final"): k_6; y_new = y0 + h sum c_j k_j (into YS); yerr = h sum ec_j k_j;
 *   r_i = |yerr_i| / (eps_abs + eps_rel |y_new,i|), grid max -> rmax (gsl_odeiv2_control standard, a_y=1, a_dydt=0). */
struct TabArgs {
  const double *sums_in; double t_eval, t_next, h; int s; /* index of the slope computed here, 1-based */
  double b[13];         /* weights of k_1..k_s for the state formed here */
  double ec[13];        /* error weights (final only); err_diff: the weights of the comparison solution instead */
  int err_diff;         /* yerr = h (sum ec_j k_j - sum b_j k_j), GSL's arithmetic form for rk8pd (ksum7 - ksum8) */
  double eps_abs, eps_rel; int final_; uint32_t eval_index; unsigned long long *rmax_bits;
  const double *ycur;   /* state the RHS is evaluated at (Y for s == 1, else YS) */
  double *K1;           /* slope k_1 (= dydt_in of gsl_odeiv2_evolve_apply); k_2..k_13 live in Pop::KS */
  const double *y0src;  /* state at the start of the (sub)step (Y, or the half-step state of the step-doubling rk4) */
  int skip_rhs;         /* s == 1 only: k_1 is already in K1 (reused dydt_out of the previous step, or a retry) */
};
__device__ __forceinline__ double *slope_ptr(const Pop &p, const TabArgs &a, int j) {   /* j 0-based */
  return j == 0 ? a.K1 : p.KS + (int64_t)(j - 1) * NVAR * p.stride;
}
__global__ void __launch_bounds__(BLOCK) k_tab(Pop p, TabArgs a, Red r) {
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; k++) red[k] = 0.0;
  double mfd[A1(NMFD)];
  load_mfd(p, a.sums_in, p.n_global, mfd);
  double rloc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
    double y0[A1(NVAR)], yc[A1(NVAR)], e[A1(NEXPR)], psi[A1(NPSI)], f[A1(NVAR)], aux[A1(NAUX)];
    double death = 0.0, repli = 0.0;
    load_cols<NVAR>(a.y0src, p.stride, i, y0);
    load_cols<NEXPR>(p.E, p.stride, i, e);
    if (NPSI > 0) load_cols<NPSI>(p.PSI, p.stride, i, psi);
    load_cols<NVAR>(a.ycur, p.stride, i, yc);
    const int32_t id = M::NEEDS_ID ? p.ID[i] : 0;
    if (a.skip_rhs) {
      load_cols<NVAR>(slope_ptr(p, a, a.s - 1), p.stride, i, f);
    } else {
      DevCtx c = make_ctx(p, a.h, a.eval_index, DOM_RHS, id);
      ctx_neighbours(c, p, a.ycur, i);
      rhs_checked(p, c, a.t_eval, yc, e, psi, mfd, f, aux, death, repli);
      store_cols<NVAR>(slope_ptr(p, a, a.s - 1), p.stride, i, f);
    }
    double yn[A1(NVAR)];
#pragma unroll
    for (int v = 0; v < NVAR; v++) {
      double sacc = 0.0, eacc = 0.0;
      for (int j = 0; j < a.s - 1; j++) {
        const double kj = slope_ptr(p, a, j)[(int64_t)v * p.stride + i];
        if (a.b[j] != 0.0) sacc += a.b[j] * kj;
        if (a.final_ && a.ec[j] != 0.0) eacc += a.ec[j] * kj;
      }
      if (a.b[a.s - 1] != 0.0) sacc += a.b[a.s - 1] * f[v];
      yn[v] = y0[v] + a.h * sacc;
      if (a.final_) {
        if (a.ec[a.s - 1] != 0.0) eacc += a.ec[a.s - 1] * f[v];
        const double yerr = a.err_diff ? a.h * (eacc - sacc) : a.h * eacc;
        const double D0 = a.eps_rel * fabs(yn[v]) + a.eps_abs;
        const double rr = fabs(yerr) / fabs(D0);
        if (rr > rloc) rloc = rr;          /* GSL_MAX_DBL: a NaN ratio is dropped */
      }
    }
    store_cols<NVAR>(p.YS_OUT, p.stride, i, yn);
    if (NMFD > 0 || M::HAS_BIRTH) {
      DevCtx c2 = make_ctx(p, a.h, a.eval_index + 1, DOM_RHS, id);
      particle_terms(c2, a.t_next, yn, e, psi, red);
    }
  }
  if (a.final_) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double other = __shfl_down_sync(0xffffffffu, rloc, o); if (other > rloc) rloc = other; }
    if ((threadIdx.x & 31) == 0) atomicMax(a.rmax_bits, (unsigned long long)__double_as_longlong(rloc));
  }
  if (NMFD > 0 || M::HAS_BIRTH) grid_reduce<NRED>(red, r);
}

/* step-doubling error of GSL's rk4: yerr = 4 (y_two_halves - y_one_full) / 15, rmax = max |yerr| / (eps_abs + eps_rel |y|) */
__global__ void __launch_bounds__(BLOCK) k_err_doubling(Pop p, const double *__restrict__ ytwo, const double *__restrict__ yone,
                                                         double eps_abs, double eps_rel, unsigned long long *rmax_bits) {
  double rloc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
#pragma unroll
    for (int v = 0; v < NVAR; v++) {
      const double y = ytwo[(int64_t)v * p.stride + i], y1 = yone[(int64_t)v * p.stride + i];
      const double yerr = 4.0 * (y - y1) / 15.0;
      const double rr = fabs(yerr) / fabs(eps_rel * fabs(y) + eps_abs);
      if (rr > rloc) rloc = rr;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double other = __shfl_down_sync(0xffffffffu, rloc, o); if (other > rloc) rloc = other; }
  if ((threadIdx.x & 31) == 0) atomicMax(rmax_bits, (unsigned long long)__double_as_longlong(rloc));
}

/* delay(): the side effects of the call sites, carried out between evaluations (src/utils_odexp.c:205-240).
 * k_delay_fill: `if (EVENT_TYPE && EVENT_T0)` -- a logical AND, true while ANY event flag is set (t0, a pending stopping time or
 *   birth / death event): the whole history is rewritten from the history function, sample s at t - (DELAY_SSTACK - s) h, pos = DELAY_SSTACK.
 * k_delay_push: the first evaluation of a dde step (DELAY_FIRST_EVAL, src/methods_odexp.c:1842-1844) appends (t, x) -- once per ode_rhs
 *   call, and pop_rhs calls ode_rhs twice per particle: two identical samples. */
__global__ void __launch_bounds__(BLOCK) k_delay_fill(Pop p, const double *__restrict__ ysrc, double t, double hstep) {
  const int64_t total = p.n * DELAY_SSTACK;
  for (int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x; q < total; q += (int64_t)gridDim.x * BLOCK) {
    const int64_t i = q % p.n; const int sl = (int)(q / p.n);
    double y[A1(NVAR)], e[A1(NEXPR)];
    load_cols<NVAR>(ysrc, p.stride, i, y);
    load_cols<NEXPR>(p.E, p.stride, i, e);
    DevCtx c = make_ctx(p, hstep, 0, DOM_RHS, p.ID[i]);
#pragma unroll
    for (int k = 0; k < M::NDELAY; k++) {
      double tt, xx; long idx;
      M::delay_sample(c, k, t, y, e, tt, xx, idx);
      if (idx < 0 || idx >= p.dl[k].K) continue;
      const double ti = tt - (double)(DELAY_SSTACK - sl) * hstep;
      p.dl[k].T[(long)sl * p.dl[k].K + idx] = ti;
      p.dl[k].X[(long)sl * p.dl[k].K + idx] = M::delay_ic(k, ti);
      if (sl == 0) p.dl[k].pos[idx] = DELAY_SSTACK;
    }
  }
}
__global__ void __launch_bounds__(BLOCK) k_delay_push(Pop p, const double *__restrict__ ysrc, double t) {
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
    double y[A1(NVAR)], e[A1(NEXPR)];
    load_cols<NVAR>(ysrc, p.stride, i, y);
    load_cols<NEXPR>(p.E, p.stride, i, e);
    DevCtx c = make_ctx(p, 0.0, 0, DOM_RHS, p.ID[i]);
#pragma unroll
    for (int k = 0; k < M::NDELAY; k++) {
      double tt, xx; long idx;
      M::delay_sample(c, k, t, y, e, tt, xx, idx);
      if (idx < 0 || idx >= p.dl[k].K) continue;
      int pos = p.dl[k].pos[idx];
      for (int twice = 0; twice < 2; twice++) {
        pos = (pos + 1) % DELAY_SSTACK;
        p.dl[k].T[(long)pos * p.dl[k].K + idx] = tt;
        p.dl[k].X[(long)pos * p.dl[k].K + idx] = xx;
      }
      p.dl[k].pos[idx] = pos;
    }
  }
}

/* AoS <-> SoA transposes for upload / download (host arrays are particle-major like the reference's y[]) */
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, int ncol, int64_t n, int64_t stride, int64_t first) {
  const int64_t total = n * ncol;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / ncol; const int c = (int)(q % ncol);
    soa[(int64_t)c * stride + first + i] = aos[q];
  }
}
__global__ void k_soa_to_aos(const double *__restrict__ soa, double *__restrict__ aos, int ncol, int64_t n, int64_t stride, int64_t first) {
  const int64_t total = n * ncol;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / ncol; const int c = (int)(q % ncol);
    aos[q] = soa[(int64_t)c * stride + first + i];
  }
}
/* every stride-th particle of [first, first + m*stride) -> AoS rows (thinned trajectory output) */
__global__ void k_soa_to_aos_strided(const double *__restrict__ soa, double *__restrict__ aos, int ncol, int64_t m, int64_t stride_cols,
                                     int64_t first, int64_t step) {
  const int64_t total = m * ncol;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / ncol; const int c = (int)(q % ncol);
    aos[q] = soa[(int64_t)c * stride_cols + first + i * step];
  }
}
__global__ void k_gather_ids(const int32_t *__restrict__ id, int32_t *__restrict__ out, int64_t m, int64_t first, int64_t step) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) out[i] = id[first + i * step];
}
/* position of the particle with a given id in this GPU's part of the list (getpar(), src/datastruct.c): ids are unique */
__global__ void k_find_id(const int32_t *__restrict__ id, int64_t n, int32_t want, long long *__restrict__ pos) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (id[i] == want) *pos = (long long)i;
}
__global__ void k_iota_ids(int32_t *id, int64_t n, int32_t first) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) id[i] = first + (int32_t)i;
}
/* ode_init_conditions on the device (keyed Philox), particles [first, first+count) */
/* mode 1: the first population (ATBIRTH, IST0, coupling / aux / rate columns cleared); mode 0: re-initialisation at a stopping
 * time (pop_init_conditions on the living particles, src/methods_odexp.c:704-714): only Y and E are rewritten */
__global__ void __launch_bounds__(BLOCK) k_init(Pop p, int64_t first, int64_t count, double t, uint32_t stream, int first_pop) {
  for (int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x; q < count; q += (int64_t)gridDim.x * BLOCK) {
    const int64_t i = first + q;
    double y[A1(NVAR)], e[A1(NEXPR)];
    load_cols<NVAR>(p.Y, p.stride, i, y);
    load_cols<NEXPR>(p.E, p.stride, i, e);
    DevCtx c = make_ctx(p, 0.0, stream, first_pop ? DOM_IC : DOM_IC + 4, p.ID[i]);
    c.atbirth = first_pop; c.ist0 = first_pop;
    M::ic(c, t, y, e);
    store_cols<NVAR>(p.Y, p.stride, i, y);
    store_cols<NEXPR>(p.E, p.stride, i, e);
    if (!first_pop) continue;
    if (NPSI > 0) { for (int k = 0; k < NPSI; k++) p.PSI[(int64_t)k * p.stride + i] = 0.0; }
    if (NAUX > 0) { for (int k = 0; k < NAUX; k++) p.AUX[(int64_t)k * p.stride + i] = 0.0; }
    p.DEATH[i] = 0.0; p.REPLI[i] = 0.0;
  }
}

}  // namespace odx

#include "odx_runtime.cuh"
