/* odx_lowrank.cuh — low-rank %C coupling on the device (compiled only for models with `%C ... lrwexp {...}`).
 *
 *   psi_i = sum_j (U_i . V_j) f(x_j - x_i)     U, V: per-particle EXPR arrays of length r
 *
 * Reference: src/lrexp.c — lrexpars :82-101, lrexprank :106-173, lrexpwp :378-475, lrwpkern :528-539, and the
 * generated gather / scatter src/odexp:1565-1594.  Same algorithm (Chebyshev fit of f on [-range, range] converted to the
 * monomial basis), laid out for the GPU:
 *   k_lr_stats    min / max / sum of x over the population                      (HBM stream, 8 B per particle)
 *   k_lr_plan     one CTA: mean, range, adaptive order p, cc[], B[l], B[l]*C(l,k) (tiny)
 *   k_lr_moments  A[m][k] = sum_j V_jm xi_j^k: block partials of an r x (p+1) reduction over particles, fixed-order sum
 *   k_lr_apply    psi_i = sum_k (sum_m U_im A[m][k]) (sum_{l>=k} B[l] C(l,k) (-xi_i)^(l-k))
 * The two contractions are skinny FP64 GEMMs (K = N, then K = r) that stream U and V once (16 N r bytes); with r = 64,
 * p + 1 <= 32 their arithmetic intensity is (p+1)/4 flop per byte, i.e. HBM-bound below p + 1 ~ 23 and FP64-bound above.
 * They run as FP64 FMA on the CUDA cores: B200's FP64 tensor rate equals its CUDA-core rate, so DMMA would buy nothing.
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

/* gsl_cheb_init restated: c_j = (2/(n+1)) sum_k f(cos(pi (k+1/2)/(n+1)) * range) cos(pi j (k+1/2)/(n+1)) on [-range, range] */
__device__ void lr_cheb_init(int order, double range, double *fk, double *c) {
  for (int k = threadIdx.x; k <= order; k += blockDim.x) {
    const double y = cos(M_PI * (k + 0.5) / (order + 1));
    fk[k] = M::lr0_fun(y * range + 0.0);
  }
  __syncthreads();
  const double fac = 2.0 / (order + 1.0);
  for (int j = threadIdx.x; j <= order; j += blockDim.x) {
    double sum = 0.0;
    for (int k = 0; k <= order; k++) sum += fk[k] * cos(M_PI * j * (k + 0.5) / (order + 1));
    c[j] = fac * sum;
  }
  __syncthreads();
}

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
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double a = INFINITY, b = -INFINITY, c = 0.0;
    for (unsigned int k = 0; k < gridDim.x; k++) {
      const double p0 = __ldcg(&partial[3 * k]), p1 = __ldcg(&partial[3 * k + 1]), p2 = __ldcg(&partial[3 * k + 2]);
      a = p0 < a ? p0 : a; b = p1 > b ? p1 : b; c += p2;
    }
    stats[0] = a; stats[1] = b; stats[2] = c; stats[3] = -a;     /* [3] = -min so that one MAX allreduce covers min and max */
    *counter = 0u;
  }
}

/* one CTA: lrexpars tail + lrexprank + the coefficient part of lrexpwp */
__global__ void __launch_bounds__(BLOCK) k_lr_plan(const double *__restrict__ stats, double n_global, int pmax, double abstol, double reltol,
                                                    LrPlan *plan) {
  __shared__ double s_f[LR_P1], s_c[LR_P1], s_red[2][BLOCK / 32];
  __shared__ long long s_tc[LR_P1][LR_P1];
  __shared__ int s_p, s_pstep, s_go;
  const double meanx = stats[2] / n_global;
  double range = stats[1] - (-stats[3]);
  if (range == 0.0) range = 1.0;
  /* Chebyshev polynomials in the monomial basis: T_{n+1} = 2 x T_n - T_{n-1} (exact in 64-bit integers up to n = 50) */
  if (threadIdx.x == 0) {
    for (int a = 0; a < LR_P1; a++) for (int b = 0; b < LR_P1; b++) s_tc[a][b] = 0;
    s_tc[0][0] = 1; s_tc[1][1] = 1;
    for (int n = 1; n + 1 < LR_P1; n++)
      for (int l = 0; l <= n + 1; l++) s_tc[n + 1][l] = (l > 0 ? 2 * s_tc[n][l - 1] : 0) - s_tc[n - 1][l];
  }
  lr_cheb_init(pmax, range, s_f, s_c);
  if (threadIdx.x == 0) {
    double sc = 0.0; int oddity = 2;
    for (int i = 0; i < pmax + 1; i += 2) sc += fabs(s_c[i]);
    if (sc < (double)pmax / 2 * 1e-8) oddity = 1;
    for (int i = 1; i < pmax + 1; i += 2) sc += fabs(s_c[i]);
    if (sc < (double)pmax / 2 * 1e-8) oddity = 0;
    s_pstep = 4;
    if (oddity == 0) s_p = 0; else if (oddity == 1) s_p = -1; else { s_p = -1; s_pstep = 3; }
  }
  __syncthreads();
  const long long ng = (long long)n_global;
  const long long errn = (ng / 10 > 10) ? ng / 10 : 10;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  do {
    if (threadIdx.x == 0) s_p += s_pstep;
    __syncthreads();
    const int p = s_p;
    const int eo = p < pmax ? p : pmax;
    double absc = 0.0;
    for (int i = 0; i <= eo; i++) absc += fabs(s_c[i]);
    const double abserr = fabs(s_c[eo]) + absc * 2.2204460492503131e-16;
    double spe = 0.0;
    for (long long i = threadIdx.x; i < errn; i += BLOCK) {
      /* gsl_cheb_eval_n_err at x = -1 + 2 i/(errn-1) on the series fitted on [-range, range] (as coded in the reference) */
      const double x = -1.0 + (double)i / (errn - 1) * 2.0;
      const double y = (2.0 * x - (-range) - range) / (range - (-range)), y2 = 2.0 * y;
      double d1 = 0.0, d2 = 0.0;
      for (int k = eo; k >= 1; k--) { const double tmp = d1; d1 = y2 * d1 - d2 + s_c[k]; d2 = tmp; }
      spe += fabs(y * d1 - d2 + 0.5 * s_c[0]);
    }
    spe = warp_sum(spe);
    if (lane == 0) s_red[0][warp] = spe;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < BLOCK / 32; w++) tot += s_red[0][w];
      const double err = abserr * (double)errn;            /* abserr does not depend on x */
      s_go = ((err / errn > (abstol + reltol * tot)) & (p < pmax)) ? 1 : 0;
    }
    __syncthreads();
  } while (s_go);
  const int p = s_p;
  /* lrexpwp: refit at order p, cc[0] /= 2, B[l] = sum_{k>=l} cc[k] TC[k][l], BC[l][k] = B[l] * C(l,k) */
  lr_cheb_init(p, range, s_f, s_c);
  if (threadIdx.x == 0) {
    s_c[0] /= 2;
    plan->meanx = meanx; plan->range = range; plan->p = p;
    for (int l = 0; l <= p; l++) {
      double b = 0;
      for (int k = l; k <= p; k++) b += s_c[k] * (double)s_tc[k][l];
      plan->B[l] = b; plan->cc[l] = s_c[l];
    }
    for (int l = 0; l <= p; l++) {
      double ch = 1.0;                                    /* C(l,0); C(l,k+1) = C(l,k) (l-k)/(k+1) exact in double here */
      for (int k = 0; k <= l; k++) { plan->BC[l][k] = plan->B[l] * ch; ch = ch * (double)(l - k) / (double)(k + 1); }
    }
  }
}

/* block partials of A[m][k] = sum_j V_jm xi_j^k over this block's particles */
__global__ void __launch_bounds__(BLOCK) k_lr_moments(const double *__restrict__ x, const double *__restrict__ V, int64_t vstride, int64_t n,
                                                       const LrPlan *__restrict__ plan, double *__restrict__ partial) {
  __shared__ double s_v[LR_MTILE][LR_R + 1];
  __shared__ double s_pw[LR_MTILE][LR_P1];
  const int p1 = plan->p + 1;
  const double meanx = plan->meanx, range = plan->range;
  double acc[LR_NACC];
#pragma unroll
  for (int s = 0; s < LR_NACC; s++) acc[s] = 0.0;
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
      for (int k = 0; k < p1; k++) { s_pw[jj][k] = (jj < cnt) ? pw : 0.0; pw *= xi; }
    }
    __syncthreads();
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
#pragma unroll
  for (int s = 0; s < LR_NACC; s++) {
    const int q = threadIdx.x + s * BLOCK;
    if (q < LR_R * LR_P1) partial[(size_t)blockIdx.x * LR_R * LR_P1 + q] = acc[s];
  }
}
__global__ void __launch_bounds__(BLOCK) k_lr_moments_sum(const double *__restrict__ partial, int nblocks, double *__restrict__ A) {
  const int q = blockIdx.x * BLOCK + threadIdx.x;
  if (q >= LR_R * LR_P1) return;
  double a = 0.0;
  for (int b = 0; b < nblocks; b++) a += partial[(size_t)b * LR_R * LR_P1 + q];
  A[q] = a;                                                             /* q = k * LR_R + m */
}

__global__ void __launch_bounds__(128) k_lr_apply(const double *__restrict__ x, const double *__restrict__ U, int64_t ustride, int64_t n,
                                                   const LrPlan *__restrict__ plan, const double *__restrict__ A, double *__restrict__ psi) {
  __shared__ double s_A[LR_P1][LR_R];
  __shared__ double s_BC[LR_P1][LR_P1];
  const int p1 = plan->p + 1;
  for (int q = threadIdx.x; q < LR_P1 * LR_R; q += 128) s_A[q / LR_R][q % LR_R] = A[q];
  for (int q = threadIdx.x; q < LR_P1 * LR_P1; q += 128) s_BC[q / LR_P1][q % LR_P1] = plan->BC[q / LR_P1][q % LR_P1];
  __syncthreads();
  const double meanx = plan->meanx, range = plan->range;
  for (int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 128) {
    double ua[LR_P1], pw[LR_P1];
    const double nxi = -(x[i] - meanx) / range;
    double w = 1.0;
#pragma unroll
    for (int k = 0; k < LR_P1; k++) { ua[k] = 0.0; pw[k] = w; w *= nxi; }
    for (int m = 0; m < LR_R; m++) {
      const double u = U[(int64_t)m * ustride + i];
#pragma unroll
      for (int k = 0; k < LR_P1; k++) ua[k] = fma(u, s_A[k][m], ua[k]);
    }
    double y = 0.0;
#pragma unroll
    for (int k = 0; k < LR_P1; k++) {
      if (k < p1) {
        double phi = 0.0;
#pragma unroll
        for (int l = 0; l < LR_P1; l++) if (l >= k && l < p1) phi = fma(s_BC[l][k], pw[l - k], phi);
        y = fma(phi, ua[k], y);
      }
    }
    psi[i] = y;
  }
}

}  // namespace odx
