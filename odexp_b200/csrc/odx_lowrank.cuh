/* odx_lowrank.cuh — low-rank %C coupling on the device (compiled only for models with `%C ... lrwexp {...}`).
 *
 *   psi_i = sum_j (U_i . V_j) f(x_j - x_i)     U, V: per-particle EXPR arrays of length r
 *
 * Reference: src/lrexp.c — lrexpars :82-101, lrexprank :106-173, lrexpwp :378-475, lrwpkern :528-539, and the
 * generated gather / scatter src/odexp:1565-1594.  Same algorithm (Chebyshev fit of f on [-range, range] converted to the
 * monomial basis), laid out for the GPU:
 *   k_lr_stats    min / max / sum of x over the population                      (HBM stream, 8 B per particle)
 *   k_lr_rank_plan  mean, range, adaptive order p (fit + sample sums on the whole grid), cc[], B[l], B[l]*C(l,k)
 *   k_lr_moments    A[m][k] = sum_j V_jm xi_j^k: block partials of an r x (p+1) reduction over particles, fixed-order sum
 *   k_lr_apply      psi_i = sum_k (sum_m U_im A[m][k]) (sum_{l>=k} B[l] C(l,k) (-xi_i)^(l-k))
 * The two contractions are skinny FP64 GEMMs (K = N, then K = r) that stream U and V once (16 N r bytes); with r = 64,
 * p + 1 <= 32 their arithmetic intensity is (p+1)/4 flop per byte, i.e. HBM-bound below p + 1 ~ 23 and FP64-bound above.
 * Rank 64 (BASELINE config 3) runs them on the FP64 tensor cores from TMA-staged tiles (k_lr_moments_tma, k_lr_apply_tma below):
 * same peak as DFMA, a quarter of the operand traffic; other ranks use the CUDA-core kernels k_lr_moments / k_lr_apply.
 */
#pragma once

namespace odx {

constexpr int LR_PMAX = 31;                 /* opt lrmax (src/options.c:33) */
constexpr int LR_P1 = LR_PMAX + 1;
constexpr int LR_R = M::LR0_RANK;
constexpr int LR_MTILE = 32;                /* particles per staged tile in k_lr_moments */
constexpr int LR_NACC = (LR_R * LR_P1 + BLOCK - 1) / BLOCK;

struct LrPlan {
  double meanx, range;
  int p;
  double cc[LR_P1], B[LR_P1], BC[LR_P1][LR_P1];
};

/* gsl_cheb_init restated: c_j = (2/(n+1)) sum_k f(cos(pi (k+1/2)/(n+1)) * range) cos(pi j (k+1/2)/(n+1)) on [-range, range].
 * The (order+1)^2 cosines are spread over the CTA (prod[j][k]); the sum over k then runs in the reference's order. */
__device__ void lr_cheb_init(int order, double range, double *fk, double *c, double (*prod)[LR_P1]) {
  for (int k = threadIdx.x; k <= order; k += blockDim.x) {
    const double y = cos(M_PI * (k + 0.5) / (order + 1));
    fk[k] = M::lr0_fun(y * range + 0.0);
  }
  __syncthreads();
  const int o1 = order + 1;
  for (int q = threadIdx.x; q < o1 * o1; q += blockDim.x) {
    const int j = q / o1, k = q % o1;
    prod[j][k] = fk[k] * cos(M_PI * j * (k + 0.5) / (order + 1));
  }
  __syncthreads();
  const double fac = 2.0 / (order + 1.0);
  for (int j = threadIdx.x; j <= order; j += blockDim.x) {
    double sum = 0.0;
    for (int k = 0; k <= order; k++) sum += prod[j][k];
    c[j] = fac * sum;
  }
  __syncthreads();
}

/* Chebyshev polynomials in the monomial basis, T_{n+1} = 2 x T_n - T_{n-1}: exact in 64-bit integers up to n = 50; built at
 * compile time (as a serial loop of one thread it was most of k_lr_plan's 29 us) */
struct LrTC { long long v[LR_P1][LR_P1]; };
constexpr LrTC lr_make_tc() {
  LrTC tc{};
  tc.v[0][0] = 1; tc.v[1][1] = 1;
  for (int n = 1; n + 1 < LR_P1; n++)
    for (int l = 0; l <= n + 1; l++) tc.v[n + 1][l] = (l > 0 ? 2 * tc.v[n][l - 1] : 0) - tc.v[n - 1][l];
  return tc;
}
__device__ const LrTC lr_tc = lr_make_tc();

__global__ void __launch_bounds__(BLOCK) k_lr_stats(const double *__restrict__ x, int64_t n, double *partial, unsigned int *counter,
                                                     double *stats) {
  __shared__ double s_min[BLOCK / 32], s_max[BLOCK / 32], s_sum[BLOCK / 32];
  __shared__ bool s_last;
  double mn = INFINITY, mx = -INFINITY, sm = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * BLOCK) {
    const double v = x[i];
    mn = v < mn ? v : mn; mx = v > mx ? v : mx; sm += v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_down_sync(0xffffffffu, mn, o), b = __shfl_down_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn; mx = b > mx ? b : mx; sm += __shfl_down_sync(0xffffffffu, sm, o);
  }
  if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; s_sum[warp] = sm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BLOCK / 32; w++) { mn = s_min[w] < mn ? s_min[w] : mn; mx = s_max[w] > mx ? s_max[w] : mx; sm += s_sum[w]; }
    partial[3 * blockIdx.x + 0] = mn; partial[3 * blockIdx.x + 1] = mx; partial[3 * blockIdx.x + 2] = sm;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {                       /* the last block folds the block partials with all its threads, in a fixed order */
    __threadfence();
    double a = INFINITY, b = -INFINITY, c = 0.0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += BLOCK) {
      const double p0 = __ldcg(&partial[3 * k]), p1 = __ldcg(&partial[3 * k + 1]), p2 = __ldcg(&partial[3 * k + 2]);
      a = p0 < a ? p0 : a; b = p1 > b ? p1 : b; c += p2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double a2 = __shfl_down_sync(0xffffffffu, a, o), b2 = __shfl_down_sync(0xffffffffu, b, o);
      a = a2 < a ? a2 : a; b = b2 > b ? b2 : b; c += __shfl_down_sync(0xffffffffu, c, o);
    }
    __syncthreads();
    if (lane == 0) { s_min[warp] = a; s_max[warp] = b; s_sum[warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < BLOCK / 32; w++) { a = s_min[w] < a ? s_min[w] : a; b = s_max[w] > b ? s_max[w] : b; c += s_sum[w]; }
      stats[0] = a; stats[1] = b; stats[2] = c; stats[3] = -a;     /* [3] = -min so that one MAX allreduce covers min and max */
      *counter = 0u;
    }
  }
}

constexpr int LR_NCAND = 12;
struct LrRank {
  double c[LR_P1];            /* fit at order pmax */
  int p0, pstep, ncand;       /* candidates p0 + pstep * (1..ncand), capped at pmax */
  double spe[LR_NCAND];       /* sum_i |series_p(x_i)| over the errn sample points, per candidate */
};
/* lrexprank + the coefficient work of lrexpwp in one launch (as three kernels -- fit on one CTA, sample sums on the grid, plan on one
 * CTA -- they cost ~29 us as a dependent chain): every CTA
 * repeats the fit at the maximal order (same bits everywhere), adds its share of the sample sums, and the last CTA to finish
 * selects the order and writes the plan.  rk->spe and the counter are left zeroed for the next call. */
__global__ void __launch_bounds__(BLOCK) k_lr_rank_plan(const double *__restrict__ stats, double n_global, int pmax, double abstol, double reltol,
                                                         LrRank *rk, unsigned int *counter, LrPlan *plan) {
  __shared__ double s_f[LR_P1], s_c[LR_P1], s_c2[LR_P1], s_prod[LR_P1][LR_P1], s_B[LR_P1];
  __shared__ double s_red[BLOCK / 32][LR_NCAND];
  __shared__ int s_p0, s_pstep, s_ncand, s_p;
  __shared__ bool s_last;
  const double meanx = stats[2] / n_global;
  double range = stats[1] - (-stats[3]);
  if (range == 0.0) range = 1.0;
  lr_cheb_init(pmax, range, s_f, s_c, s_prod);
  if (threadIdx.x == 0) {
    double sc = 0.0; int oddity = 2;
    for (int i = 0; i < pmax + 1; i += 2) sc += fabs(s_c[i]);
    if (sc < (double)pmax / 2 * 1e-8) oddity = 1;
    for (int i = 1; i < pmax + 1; i += 2) sc += fabs(s_c[i]);
    if (sc < (double)pmax / 2 * 1e-8) oddity = 0;
    int pstep = 4, p0;
    if (oddity == 0) p0 = 0; else if (oddity == 1) p0 = -1; else { p0 = -1; pstep = 3; }
    int nc = 0;
    for (int p = p0 + pstep; nc < LR_NCAND; p += pstep) { nc++; if (p >= pmax) break; }
    s_p0 = p0; s_pstep = pstep; s_ncand = nc;
  }
  __syncthreads();
  const long long ng = (long long)n_global;
  const long long errn = (ng / 10 > 10) ? ng / 10 : 10;
  const int ncand = s_ncand, p0 = s_p0, pstep = s_pstep;
  double spe[LR_NCAND];
#pragma unroll
  for (int q = 0; q < LR_NCAND; q++) spe[q] = 0.0;
  for (long long i = (long long)blockIdx.x * BLOCK + threadIdx.x; i < errn; i += (long long)gridDim.x * BLOCK) {
    const double x = -1.0 + (double)i / (errn - 1) * 2.0;
    const double y = (2.0 * x - (-range) - range) / (range - (-range)), y2 = 2.0 * y;
#pragma unroll
    for (int q = 0; q < LR_NCAND; q++) {
      if (q < ncand) {
        const int p = p0 + pstep * (q + 1);
        const int eo = p < pmax ? p : pmax;
        double d1 = 0.0, d2 = 0.0;
        for (int k = eo; k >= 1; k--) { const double tmp = d1; d1 = y2 * d1 - d2 + s_c[k]; d2 = tmp; }
        spe[q] += fabs(y * d1 - d2 + 0.5 * s_c[0]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < LR_NCAND; q++) {
    const double v = warp_sum(spe[q]);
    if (lane == 0) s_red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < LR_NCAND) {
    double v = 0.0;
    for (int w = 0; w < BLOCK / 32; w++) v += s_red[w][threadIdx.x];
    atomicAdd(&rk->spe[threadIdx.x], v);
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(counter, 1u) == gridDim.x - 1); }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    int p = p0;
    for (int q = 0; q < ncand; q++) {
      p = p0 + pstep * (q + 1);
      const int eo = p < pmax ? p : pmax;
      double absc = 0.0;
      for (int i = 0; i <= eo; i++) absc += fabs(s_c[i]);
      const double abserr = fabs(s_c[eo]) + absc * 2.2204460492503131e-16;
      const double err = abserr * (double)errn;               /* abserr does not depend on the sample point */
      if (!((err / errn > (abstol + reltol * __ldcg(&rk->spe[q]))) & (p < pmax))) break;
    }
    s_p = p;
    for (int q = 0; q < LR_NCAND; q++) rk->spe[q] = 0.0;
    *counter = 0u;
  }
  __syncthreads();
  const int p = s_p;
  lr_cheb_init(p, range, s_f, s_c2, s_prod);
  if (threadIdx.x == 0) {
    s_c2[0] /= 2;
    plan->meanx = meanx; plan->range = range; plan->p = p;
  }
  __syncthreads();
  for (int l = threadIdx.x; l <= p; l += blockDim.x) {          /* one order per thread, the sum over k in the reference's order */
    double b = 0;
    for (int k = l; k <= p; k++) b += s_c2[k] * (double)lr_tc.v[k][l];
    s_B[l] = b; plan->B[l] = b; plan->cc[l] = s_c2[l];
  }
  __syncthreads();
  for (int l = threadIdx.x; l <= p; l += blockDim.x) {
    double ch = 1.0;                                      /* C(l,0); C(l,k+1) = C(l,k) (l-k)/(k+1) exact in double here */
    for (int k = 0; k <= l; k++) { plan->BC[l][k] = s_B[l] * ch; ch = ch * (double)(l - k) / (double)(k + 1); }
  }
}

/* block partials of A[m][k] = sum_j V_jm xi_j^k over this block's particles: a skinny FP64 GEMM (K = particles).
 * Register tiling when the output fits 8 accumulators per thread (r = 64, p+1 <= 32): thread t owns the 4 x 2 tile
 * m0 = 4 (t % (r/4)), k0 = 2 (t / (r/4)); per particle it reads 4 V values and 2 powers from shared memory (LDS.128 / broadcast)
 * for 8 DFMA, i.e. 0.6 shared-memory wavefronts per FMA instead of 3 for the scalar mapping used for small ranks. */
constexpr bool LR_TILED = (LR_R % 4 == 0) && ((LR_R / 4) * (LR_P1 / 2) == BLOCK);
__global__ void __launch_bounds__(BLOCK, 5) k_lr_moments(const double *__restrict__ x, const double *__restrict__ V, int64_t vstride, int64_t n,
                                                       const LrPlan *__restrict__ plan, double *__restrict__ partial) {
  __shared__ __align__(16) double s_v[LR_MTILE][LR_TILED ? LR_R : LR_R + 1];
  __shared__ __align__(16) double s_pw[LR_MTILE][LR_P1];
  const int p1 = plan->p + 1;
  const double meanx = plan->meanx, range = plan->range;
  double acc[LR_TILED ? 8 : LR_NACC];
#pragma unroll
  for (int s = 0; s < (LR_TILED ? 8 : LR_NACC); s++) acc[s] = 0.0;
  const int m0 = LR_TILED ? 4 * (threadIdx.x % (LR_R / 4)) : 0, k0 = LR_TILED ? 2 * (threadIdx.x / (LR_R / 4)) : 0;
  for (int64_t base = (int64_t)blockIdx.x * LR_MTILE; base < n; base += (int64_t)gridDim.x * LR_MTILE) {
    const int cnt = (int)((n - base) < LR_MTILE ? (n - base) : LR_MTILE);
    __syncthreads();
    for (int q = threadIdx.x; q < LR_MTILE * LR_R; q += BLOCK) {
      const int jj = q % LR_MTILE, m = q / LR_MTILE;                     /* coalesced along particles */
      s_v[jj][m] = (jj < cnt) ? V[(int64_t)m * vstride + base + jj] : 0.0;
    }
    if (threadIdx.x < LR_MTILE) {
      const int jj = threadIdx.x;
      const double xi = (jj < cnt) ? (x[base + jj] - meanx) / range : 0.0;
      double pw = 1.0;
      for (int k = 0; k < LR_P1; k++) { s_pw[jj][k] = (jj < cnt && k < p1) ? pw : 0.0; pw *= xi; }
    }
    __syncthreads();
    if (LR_TILED) {
      if (k0 < p1) {
#pragma unroll 4
        for (int jj = 0; jj < LR_MTILE; jj++) {
          const double2 va = *reinterpret_cast<const double2 *>(&s_v[jj][m0]), vb = *reinterpret_cast<const double2 *>(&s_v[jj][m0 + 2]);
          const double2 pk = *reinterpret_cast<const double2 *>(&s_pw[jj][k0]);
          acc[0] = fma(va.x, pk.x, acc[0]); acc[1] = fma(va.y, pk.x, acc[1]); acc[2] = fma(vb.x, pk.x, acc[2]); acc[3] = fma(vb.y, pk.x, acc[3]);
          acc[4] = fma(va.x, pk.y, acc[4]); acc[5] = fma(va.y, pk.y, acc[5]); acc[6] = fma(vb.x, pk.y, acc[6]); acc[7] = fma(vb.y, pk.y, acc[7]);
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < LR_NACC; s++) {
        const int q = threadIdx.x + s * BLOCK, m = q % LR_R, k = q / LR_R;
        if (q < LR_R * LR_P1 && k < p1) {
          double a = acc[s];
          for (int jj = 0; jj < LR_MTILE; jj++) a = fma(s_v[jj][m], s_pw[jj][k], a);
          acc[s] = a;
        }
      }
    }
  }
  double *dst = partial + (size_t)blockIdx.x * LR_R * LR_P1;             /* layout q = k * LR_R + m */
  if (LR_TILED) {
#pragma unroll
    for (int kk = 0; kk < 2; kk++)
#pragma unroll
      for (int mm = 0; mm < 4; mm++) dst[(k0 + kk) * LR_R + m0 + mm] = acc[kk * 4 + mm];
  } else {
#pragma unroll
    for (int s = 0; s < LR_NACC; s++) {
      const int q = threadIdx.x + s * BLOCK;
      if (q < LR_R * LR_P1) dst[q] = acc[s];
    }
  }
}
constexpr bool LR_WSPLIT = (LR_R == 64);          /* rank 64: TMA-staged tensor-core path below */

/* ------------------------------------------------------------------------------------------------
 * rank-64 path (BASELINE config 3): TMA-staged tiles, FP64 tensor-core contractions.
 *
 * Both contractions are dense skinny GEMMs that stream a 64-column array once.  The warps that do the arithmetic never issue a
 * global load: a producer warp fetches the column runs of the next tiles with cp.async.bulk into a ring of shared-memory
 * buffers (mbarrier complete_tx = "tile landed", one arrival per consumer warp = "slot free").  The first version did the
 * arithmetic as DFMAs fed by 16-byte broadcast reads; ncu showed it bound by the shared-memory pipe (a broadcast LDS.128 costs
 * ~3 cycles and feeds two DFMAs: 25-34 % FP64 pipe, short-scoreboard stalls).  DMMA (mma.sync m8n8k4, f64) has the same peak rate
 * as DFMA on B200 but needs one 8-byte operand read per 8 FMAs of a lane, so the operand traffic drops fourfold; this is the one
 * place of the hot path that is a dense contraction, hence the one place that uses tensor cores.
 *
 * Fragment layout of mma.m8n8k4 (g = lane >> 2, t = lane & 3): A[g][t], B[t][g], C[g][2t], C[g][2t+1].  Row strides are = 4
 * (mod 16) doubles, which makes every fragment read conflict-free.
 *
 * k_lr_moments_tma: A[m][k] = sum_j V[m][j] xi_j^k  (M = 64 columns, N = P1T orders, K = particles).  The ring holds chunks of 16
 * column runs of V for a batch of 256 particles (2 KB per run: 512-byte runs kept the TMA unit, not HBM, busy); four chunks make
 * a batch.  Each of the 8 consumer warps takes 32 particles of the batch (eight K-steps) and keeps all 64 x P1T accumulators.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int LRT_PT = 256, LRT_SU = LRT_PT + 4, LRT_CH = 16, LRT_CW = LRT_PT / 32, LRT_THREADS = LRT_PT + 32, LRT_STAGES = 4;
constexpr size_t lrt_smem_bytes(int p1t) {
  return sizeof(double) * ((size_t)LRT_STAGES * LRT_CH * LRT_SU + (size_t)LRT_PT * (p1t + 4)) + 2 * LRT_STAGES * sizeof(uint64_t);
}
template <int P1T>
__device__ __forceinline__ void lr_moments_tma_body(const double *__restrict__ x, const double *__restrict__ V, int64_t vstride, int64_t n,
                                                    const LrPlan *__restrict__ plan, double *__restrict__ partial) {
  extern __shared__ __align__(16) unsigned char lrt_smem[];
  const int p1 = plan->p + 1;
  constexpr int SP = P1T + 4, NB = P1T / 8, NCH = 64 / LRT_CH;
  double *s_v = reinterpret_cast<double *>(lrt_smem);                          /* ring: [stage][16 runs][256 + 4] */
  double (*s_pw)[SP] = reinterpret_cast<double (*)[SP]>(s_v + (size_t)LRT_STAGES * LRT_CH * LRT_SU);      /* [particle of the batch][k] */
  uint64_t *s_full = reinterpret_cast<uint64_t *>(&s_pw[LRT_PT][0]), *s_empty = s_full + LRT_STAGES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (threadIdx.x == 0) {
    for (int b = 0; b < LRT_STAGES; b++) { mbar_init(s_full + b, 1); mbar_init(s_empty + b, LRT_CW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int64_t nbatch = (n + LRT_PT - 1) / LRT_PT;
  double acc[8][NB][2];
#pragma unroll
  for (int mb = 0; mb < 8; mb++)
#pragma unroll
    for (int nb = 0; nb < NB; nb++) { acc[mb][nb][0] = 0.0; acc[mb][nb][1] = 0.0; }
  if (warp == LRT_CW) {                                         /* producer: lanes 0..15 issue the 16 runs (2 KB each) of a chunk */
    int it = 0;
    for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
      const int64_t base = bt * LRT_PT;
      int64_t cnt = n - base; if (cnt > LRT_PT) cnt = LRT_PT;
      const uint32_t bytes = (uint32_t)(((cnt + 1) & ~(int64_t)1) * sizeof(double));     /* whole 16-byte units; columns are padded */
      for (int c = 0; c < NCH; c++, it++) {
        const int b = it % LRT_STAGES;
        if (it >= LRT_STAGES) mbar_wait(s_empty + b, (uint32_t)((it / LRT_STAGES - 1) & 1));
        if (lane == 0) mbar_expect_tx(s_full + b, bytes * (uint32_t)LRT_CH);
        __syncwarp();
        if (lane < LRT_CH)
          bulk_g2s(s_v + ((size_t)b * LRT_CH + lane) * LRT_SU, V + (int64_t)(c * LRT_CH + lane) * vstride + base, bytes, s_full + b);
      }
    }
  } else {
    const double meanx = plan->meanx, range = plan->range;
    int it = 0;
    const int64_t first = (int64_t)blockIdx.x * LRT_PT + threadIdx.x;
    double xnext = (first < n) ? x[first] : 0.0;
    for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
      const int64_t i = bt * LRT_PT + threadIdx.x;              /* powers of this lane's particle: the reference's product order */
      {
        const double xi = (xnext - meanx) / range;
        const bool valid = i < n;
        double pw = 1.0;
#pragma unroll
        for (int k = 0; k < P1T; k++) { s_pw[threadIdx.x][k] = (valid && k < p1) ? pw : 0.0; pw *= xi; }
        const int64_t inext = i + (int64_t)gridDim.x * LRT_PT;  /* x of the next batch: its latency hides behind this one */
        xnext = (inext < n) ? x[inext] : 0.0;
      }
      __syncwarp();
      const int64_t wleft = n - (bt * LRT_PT + warp * 32);      /* particles of this warp inside the population */
#pragma unroll
      for (int c = 0; c < NCH; c++, it++) {
        const int b = it % LRT_STAGES;
        mbar_wait_warp_a(smem_u32(s_full + b), (uint32_t)((it / LRT_STAGES) & 1));
        const double *vb = s_v + (size_t)b * LRT_CH * LRT_SU + g * LRT_SU + warp * 32 + t;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          const bool valid = (4 * ks + t) < wleft;              /* beyond the population the slot holds stale bytes */
          double bf[NB];
#pragma unroll
          for (int nb = 0; nb < NB; nb++) bf[nb] = s_pw[warp * 32 + 4 * ks + t][8 * nb + g];
#pragma unroll
          for (int ml = 0; ml < LRT_CH / 8; ml++) {
            double a = vb[ml * 8 * LRT_SU + 4 * ks];
            a = valid ? a : 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; nb++) dmma884(acc[c * (LRT_CH / 8) + ml][nb][0], acc[c * (LRT_CH / 8) + ml][nb][1], a, bf[nb]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + b);
      }
    }
  }
  /* the consumer warps add their accumulators in warp order through shared memory (the ring is dead by now), then the CTA
   * writes ONE partial, layout q = k * 64 + m; orders >= P1T stay zero (the buffer is cleared by the host) */
  __syncthreads();
  double *s_red = s_v;                                          /* P1T * 64 doubles <= one ring slot */
  for (int w = 0; w < LRT_CW; w++) {
    if (warp == w) {
#pragma unroll
      for (int mb = 0; mb < 8; mb++)
#pragma unroll
        for (int nb = 0; nb < NB; nb++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            double *slot = &s_red[(8 * nb + 2 * t + e) * 64 + 8 * mb + g];
            if (w == 0) *slot = acc[mb][nb][e]; else *slot += acc[mb][nb][e];
          }
    }
    __syncthreads();
  }
  double *dst = partial + (size_t)blockIdx.x * LR_R * LR_P1;
  for (int q = threadIdx.x; q < P1T * 64; q += LRT_THREADS) dst[q] = s_red[q];
}
/* one launch for every order: the order is chosen on the device (k_lr_plan), the kernel branches to the body specialised on
 * p + 1 rounded up to a multiple of 8.  One persistent CTA per SM either way, so the larger bodies cost the smaller ones nothing. */
__global__ void __launch_bounds__(LRT_THREADS, 1) k_lr_moments_tma(const double *__restrict__ x, const double *__restrict__ V, int64_t vstride,
                                                                 int64_t n, const LrPlan *__restrict__ plan, double *__restrict__ partial) {
  switch ((plan->p + 8) / 8) {
    case 1: lr_moments_tma_body<8>(x, V, vstride, n, plan, partial); break;
    case 2: lr_moments_tma_body<16>(x, V, vstride, n, plan, partial); break;
    case 3: lr_moments_tma_body<24>(x, V, vstride, n, plan, partial); break;
    default: lr_moments_tma_body<32>(x, V, vstride, n, plan, partial); break;
  }
}

/* k_lr_apply_tma: UA[i][k] = sum_m U[m][i] A[m][k] (M = particles, N = P1T orders, K = 64 columns), then
 * psi_i = sum_k UA[i][k] phi_k(-xi_i).  The ring holds chunks of 16 column runs of U for a batch of 256 particles; four chunks
 * make a batch, a warp owns 32 particles of it.  The accumulator fragments leave particle i = 8 ib + g and orders 2t, 2t+1 (per
 * block of 8) in lane (g, t); phi_k(-xi_i) is produced in the same layout by one more small contraction (see the tail). */
constexpr int LRA_PT = 256, LRA_SU = LRA_PT + 4, LRA_CH = 16, LRA_CW = LRA_PT / 32, LRA_THREADS = LRA_PT + 32, LRA_STAGES = 4;
constexpr size_t lra_smem_bytes(int p1t) {
  return sizeof(double) * ((size_t)LRA_STAGES * LRA_CH * LRA_SU + (size_t)(64 + p1t) * (p1t + 4)) + 2 * LRA_STAGES * sizeof(uint64_t);
}
template <int P1T>
__device__ __forceinline__ void lr_apply_tma_body(const double *__restrict__ x, const double *__restrict__ U, int64_t ustride, int64_t n,
                                                  const LrPlan *__restrict__ plan, const double *__restrict__ A, double *__restrict__ psi) {
  extern __shared__ __align__(16) unsigned char lra_smem[];
  const int p1 = plan->p + 1;
  constexpr int SA = P1T + 4, NB = P1T / 8, NCH = 64 / LRA_CH;
  double *s_u = reinterpret_cast<double *>(lra_smem);
  double (*s_A)[SA] = reinterpret_cast<double (*)[SA]>(s_u + (size_t)LRA_STAGES * LRA_CH * LRA_SU);       /* [m][k] */
  double (*s_BCt)[SA] = reinterpret_cast<double (*)[SA]>(&s_A[64][0]);       /* [d][k] = BC[k+d][k] */
  uint64_t *s_full = reinterpret_cast<uint64_t *>(&s_BCt[P1T][0]), *s_empty = s_full + LRA_STAGES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (threadIdx.x == 0) {
    for (int b = 0; b < LRA_STAGES; b++) { mbar_init(s_full + b, 1); mbar_init(s_empty + b, LRA_CW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int64_t nbatch = (n + LRA_PT - 1) / LRA_PT;
  if (warp == LRA_CW) {                                         /* producer: lanes 0..15 issue the 16 runs of a chunk */
    int it = 0;
    for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
      const int64_t base = bt * LRA_PT;
      int64_t cnt = n - base; if (cnt > LRA_PT) cnt = LRA_PT;
      const uint32_t bytes = (uint32_t)(((cnt + 1) & ~(int64_t)1) * sizeof(double));
      for (int c = 0; c < NCH; c++, it++) {
        const int b = it % LRA_STAGES;
        if (it >= LRA_STAGES) mbar_wait(s_empty + b, (uint32_t)((it / LRA_STAGES - 1) & 1));
        if (lane == 0) mbar_expect_tx(s_full + b, bytes * (uint32_t)LRA_CH);
        __syncwarp();
        if (lane < LRA_CH)
          bulk_g2s(s_u + ((size_t)b * LRA_CH + lane) * LRA_SU, U + (int64_t)(c * LRA_CH + lane) * ustride + base, bytes, s_full + b);
      }
    }
    return;
  }
  /* consumers (a named barrier keeps the producer warp out of their synchronisation) */
  for (int q = threadIdx.x; q < P1T * 64; q += LRA_PT) s_A[q % 64][q / 64] = A[q];
  for (int q = threadIdx.x; q < P1T * P1T; q += LRA_PT) {
    const int d = q / P1T, k = q % P1T;
    s_BCt[d][k] = (k + d < p1) ? plan->BC[k + d][k] : 0.0;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(LRA_PT) : "memory");
  const double meanx = plan->meanx, range = plan->range;
  int it = 0;
  double xnext[4];
#pragma unroll
  for (int ib = 0; ib < 4; ib++) {
    const int64_t i = (int64_t)blockIdx.x * LRA_PT + warp * 32 + ib * 8 + g;
    xnext[ib] = (i < n) ? x[i] : 0.0;
  }
  for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
    const int64_t i0 = bt * LRA_PT + warp * 32 + g;
    double nxi[4];
#pragma unroll
    for (int ib = 0; ib < 4; ib++) {
      nxi[ib] = -(xnext[ib] - meanx) / range;
      const int64_t inext = i0 + (int64_t)gridDim.x * LRA_PT + ib * 8;         /* x of the next batch: its latency hides behind this one */
      xnext[ib] = (inext < n) ? x[inext] : 0.0;
    }
    double ua[4][NB][2];
#pragma unroll
    for (int ib = 0; ib < 4; ib++)
#pragma unroll
      for (int nb = 0; nb < NB; nb++) { ua[ib][nb][0] = 0.0; ua[ib][nb][1] = 0.0; }
    for (int c = 0; c < NCH; c++, it++) {
      const int b = it % LRA_STAGES;
      mbar_wait_warp_a(smem_u32(s_full + b), (uint32_t)((it / LRA_STAGES) & 1));
      const double *ub = s_u + (size_t)b * LRA_CH * LRA_SU + warp * 32 + g;
#pragma unroll
      for (int ks = 0; ks < LRA_CH / 4; ks++) {
        double bf[NB];
#pragma unroll
        for (int nb = 0; nb < NB; nb++) bf[nb] = s_A[c * LRA_CH + 4 * ks + t][8 * nb + g];
#pragma unroll
        for (int ib = 0; ib < 4; ib++) {
          const double a = ub[(4 * ks + t) * LRA_SU + ib * 8];
#pragma unroll
          for (int nb = 0; nb < NB; nb++) dmma884(ua[ib][nb][0], ua[ib][nb][1], a, bf[nb]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty + b);
    }
    /* phi_k(nxi_i) = sum_d nxi_i^d BC[k+d][k] is one more small contraction (M = particles, K = d, N = orders) whose result has
     * the fragment layout of UA: psi_i = sum_k UA[i][k] phi_k is a dot product inside the lane and a sum over the quad.  (As
     * Horner loops per lane this tail was two thirds of the kernel's instructions.)  Blocks with 4 db + 8 nb >= P1T are zero. */
#pragma unroll
    for (int ib = 0; ib < 4; ib++) {
      const double z = nxi[ib], z2 = z * z, z3 = z2 * z, z4 = z2 * z2;
      double a = (t == 0) ? 1.0 : (t == 1) ? z : (t == 2) ? z2 : z3;             /* nxi^(4 db + t) */
      double phi[NB][2];
#pragma unroll
      for (int nb = 0; nb < NB; nb++) { phi[nb][0] = 0.0; phi[nb][1] = 0.0; }
#pragma unroll
      for (int db = 0; db < P1T / 4; db++) {
#pragma unroll
        for (int nb = 0; nb < NB; nb++)
          if (4 * db + 8 * nb < P1T) dmma884(phi[nb][0], phi[nb][1], a, s_BCt[4 * db + t][8 * nb + g]);
        a *= z4;
      }
      double y = 0.0;
#pragma unroll
      for (int nb = 0; nb < NB; nb++) { y = fma(phi[nb][0], ua[ib][nb][0], y); y = fma(phi[nb][1], ua[ib][nb][1], y); }
      y += __shfl_xor_sync(0xffffffffu, y, 1);
      y += __shfl_xor_sync(0xffffffffu, y, 2);
      const int64_t i = i0 + ib * 8;
      if (t == 0 && i < n) psi[i] = y;
    }
  }
}
__global__ void __launch_bounds__(LRA_THREADS, 1) k_lr_apply_tma(const double *__restrict__ x, const double *__restrict__ U, int64_t ustride,
                                                               int64_t n, const LrPlan *__restrict__ plan, const double *__restrict__ A,
                                                               double *__restrict__ psi) {
  switch ((plan->p + 8) / 8) {
    case 1: lr_apply_tma_body<8>(x, U, ustride, n, plan, A, psi); break;
    case 2: lr_apply_tma_body<16>(x, U, ustride, n, plan, A, psi); break;
    case 3: lr_apply_tma_body<24>(x, U, ustride, n, plan, A, psi); break;
    default: lr_apply_tma_body<32>(x, U, ustride, n, plan, A, psi); break;
  }
}

/* fixed-order sum of groups of partials: out[g][q] = sum_{b in group g} in[b][q]; called twice (groups of 32, then all) */
__global__ void __launch_bounds__(BLOCK) k_lr_moments_sum(const double *__restrict__ in, int nin, int group, double *__restrict__ out) {
  const int q = blockIdx.x * BLOCK + threadIdx.x, g = blockIdx.y;
  if (q >= LR_R * LR_P1) return;
  const int b0 = g * group, b1 = (b0 + group < nin) ? b0 + group : nin;
  double a = 0.0;
  for (int b = b0; b < b1; b++) a += in[(size_t)b * LR_R * LR_P1 + q];
  out[(size_t)g * LR_R * LR_P1 + q] = a;                                 /* q = k * LR_R + m */
}

/* psi_i for the particles of this GPU.  P1T = p + 1 rounded up to a multiple of 8: the host launches the four
 * instantiations and the three that do not match the order chosen on the device return at once (no host round trip). */
template <int P1T>
__global__ void __launch_bounds__(128) k_lr_apply(const double *__restrict__ x, const double *__restrict__ U, int64_t ustride, int64_t n,
                                                   const LrPlan *__restrict__ plan, const double *__restrict__ A, double *__restrict__ psi) {
  __shared__ __align__(16) double s_A[LR_R][P1T];          /* [m][k]: two orders per 16-byte broadcast read */
  __shared__ double s_BC[P1T][P1T];
  const int p1 = plan->p + 1;
  if ((p1 + 7) / 8 * 8 != P1T) return;
  for (int q = threadIdx.x; q < P1T * LR_R; q += 128) s_A[q % LR_R][q / LR_R] = A[q];
  for (int q = threadIdx.x; q < P1T * P1T; q += 128) s_BC[q / P1T][q % P1T] = plan->BC[q / P1T][q % P1T];
  __syncthreads();
  const double meanx = plan->meanx, range = plan->range;
  for (int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 128) {
    double ua[P1T];
    const double nxi = -(x[i] - meanx) / range;
#pragma unroll
    for (int k = 0; k < P1T; k++) ua[k] = 0.0;
    constexpr int UB = (LR_R % 8 == 0) ? 8 : 1;              /* U values fetched one batch ahead of their FMAs */
    double un[UB];
#pragma unroll
    for (int q = 0; q < UB; q++) un[q] = U[(int64_t)q * ustride + i];
    for (int m = 0; m < LR_R; m += UB) {
      double u[UB];
#pragma unroll
      for (int q = 0; q < UB; q++) u[q] = un[q];
      if (m + UB < LR_R) {
#pragma unroll
        for (int q = 0; q < UB; q++) un[q] = U[(int64_t)(m + UB + q) * ustride + i];
      }
#pragma unroll
      for (int q = 0; q < UB; q++) {
#pragma unroll
        for (int k = 0; k < P1T; k += 2) {
          const double2 a2 = *reinterpret_cast<const double2 *>(&s_A[m + q][k]);
          ua[k] = fma(u[q], a2.x, ua[k]); ua[k + 1] = fma(u[q], a2.y, ua[k + 1]);
        }
      }
    }
    /* phi_k(nxi) = sum_{l>=k} B[l] C(l,k) nxi^(l-k): Horner in nxi (no table of powers in registers) */
    double y = 0.0;
#pragma unroll
    for (int k = 0; k < P1T; k++) {
      if (k < p1) {
        double phi = 0.0;
        for (int l = p1 - 1; l >= k; l--) phi = fma(phi, nxi, s_BC[l][k]);
        y = fma(phi, ua[k], y);
      }
    }
    psi[i] = y;
  }
}

}  // namespace odx
