/* TEST INFRASTRUCTURE — clean-room restatement of the parts of the GNU Scientific Library that the
 * reference's population path calls.  PARITY WITH REAL GSL IS UNPINNED: GSL (version unpinned by
 * the reference: CMakeLists.txt:35 `find_package(GSL REQUIRED)`) is not under /root/reference and is
 * not installed here, and the reference ships no golden outputs.  The algorithms below restate GSL's
 * published behaviour (gsl-2.x doc, "Ordinary Differential Equations" and "Chebyshev Approximations"):
 *
 *   gsl_odeiv2_step_{rk2,rk4,rkf45,rkck,rk8pd}  call sites  src/methods_odexp.c:154-183,445-447,698-701
 *   gsl_odeiv2_control_y_new          call site   src/methods_odexp.c:446
 *   gsl_odeiv2_evolve_apply           call site   src/methods_odexp.c:619
 *   gsl_cheb_{alloc,init,coeffs,eval_n_err,free}, gsl_sf_choose, gsl_pow_int
 *                                     call sites  src/lrexp.c:116-165,191-245,296-353,407-468
 *
 * What pins them here: order-of-convergence, the embedded-pair identities and closed forms
 * (tests/test_oracle_minigsl.py).  bsimp is not restated (abort if selected).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_odeiv2.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_chebyshev.h>
#include <gsl/gsl_sf.h>

/* ------------------------------------------------------------------ steppers */
enum { MG_RK2, MG_RK4, MG_RKF45, MG_RKCK, MG_RK8PD, MG_BSIMP };
struct gsl_odeiv2_step_type_struct { int kind; unsigned order; const char *name; };
static const gsl_odeiv2_step_type t_rk2 = {MG_RK2, 2, "rk2"}, t_rk4 = {MG_RK4, 4, "rk4"},
    t_rkf45 = {MG_RKF45, 5, "rkf45"}, t_rkck = {MG_RKCK, 5, "rkck"}, t_rk8pd = {MG_RK8PD, 8, "rk8pd"},
    t_bsimp = {MG_BSIMP, 2, "bsimp"};
const gsl_odeiv2_step_type *gsl_odeiv2_step_rk2 = &t_rk2;
const gsl_odeiv2_step_type *gsl_odeiv2_step_rk4 = &t_rk4;
const gsl_odeiv2_step_type *gsl_odeiv2_step_rkf45 = &t_rkf45;
const gsl_odeiv2_step_type *gsl_odeiv2_step_rkck = &t_rkck;
const gsl_odeiv2_step_type *gsl_odeiv2_step_rk8pd = &t_rk8pd;
const gsl_odeiv2_step_type *gsl_odeiv2_step_bsimp = &t_bsimp;

#define MG_MAXK 13
struct gsl_odeiv2_step_struct { const gsl_odeiv2_step_type *type; size_t dim; double *k[MG_MAXK], *y0, *ytmp, *yone; };
struct gsl_odeiv2_control_struct { double eps_abs, eps_rel, a_y, a_dydt; };
struct gsl_odeiv2_evolve_struct { size_t dim; double *y0, *yerr, *dydt_in, *dydt_out; double last_step; unsigned long count, failed_steps; };

gsl_odeiv2_step *gsl_odeiv2_step_alloc(const gsl_odeiv2_step_type *T, size_t dim) {
  gsl_odeiv2_step *s = calloc(1, sizeof *s);
  size_t n = dim ? dim : 1;
  s->type = T; s->dim = dim;
  for (int i = 0; i < MG_MAXK; i++) s->k[i] = calloc(n, sizeof(double));
  s->y0 = calloc(n, sizeof(double)); s->ytmp = calloc(n, sizeof(double)); s->yone = calloc(n, sizeof(double));
  return s;
}
void gsl_odeiv2_step_free(gsl_odeiv2_step *s) {
  if (!s) return;
  for (int i = 0; i < MG_MAXK; i++) free(s->k[i]);
  free(s->y0); free(s->ytmp); free(s->yone); free(s);
}
gsl_odeiv2_control *gsl_odeiv2_control_y_new(double eps_abs, double eps_rel) {
  gsl_odeiv2_control *c = calloc(1, sizeof *c);
  c->eps_abs = eps_abs; c->eps_rel = eps_rel; c->a_y = 1.0; c->a_dydt = 0.0; return c;
}
void gsl_odeiv2_control_free(gsl_odeiv2_control *c) { free(c); }
gsl_odeiv2_evolve *gsl_odeiv2_evolve_alloc(size_t dim) {
  gsl_odeiv2_evolve *e = calloc(1, sizeof *e);
  size_t n = dim ? dim : 1;
  e->dim = dim; e->y0 = calloc(n, sizeof(double)); e->yerr = calloc(n, sizeof(double));
  e->dydt_in = calloc(n, sizeof(double)); e->dydt_out = calloc(n, sizeof(double));
  return e;
}
void gsl_odeiv2_evolve_free(gsl_odeiv2_evolve *e) { if (!e) return; free(e->y0); free(e->yerr); free(e->dydt_in); free(e->dydt_out); free(e); }

#define FEVAL(sys, t, y, f) do { int s_ = (sys)->function((t), (y), (f), (sys)->params); if (s_ != GSL_SUCCESS) return s_; } while (0)

/* Embedded 4(5) pairs share one driver: 6 stages, k1 = dydt_in, propagate the 5th-order solution. */
typedef struct { double ah[5]; double b[5][5]; double c[6]; double ec[6]; } tableau;
static const tableau TAB_RKF45 = {
  {1.0/4.0, 3.0/8.0, 12.0/13.0, 1.0, 1.0/2.0},
  {{1.0/4.0},
   {3.0/32.0, 9.0/32.0},
   {1932.0/2197.0, -7200.0/2197.0, 7296.0/2197.0},
   {8341.0/4104.0, -32832.0/4104.0, 29440.0/4104.0, -845.0/4104.0},
   {-6080.0/20520.0, 41040.0/20520.0, -28352.0/20520.0, 9295.0/20520.0, -5643.0/20520.0}},
  {902880.0/7618050.0, 0.0, 3953664.0/7618050.0, 3855735.0/7618050.0, -1371249.0/7618050.0, 277020.0/7618050.0},
  {1.0/360.0, 0.0, -128.0/4275.0, -2197.0/75240.0, 1.0/50.0, 2.0/55.0}};
static const tableau TAB_RKCK = {
  {1.0/5.0, 3.0/10.0, 3.0/5.0, 1.0, 7.0/8.0},
  {{1.0/5.0},
   {3.0/40.0, 9.0/40.0},
   {0.3, -0.9, 1.2},
   {-11.0/54.0, 2.5, -70.0/27.0, 35.0/27.0},
   {1631.0/55296.0, 175.0/512.0, 575.0/13824.0, 44275.0/110592.0, 253.0/4096.0}},
  {37.0/378.0, 0.0, 250.0/621.0, 125.0/594.0, 0.0, 512.0/1771.0},
  {37.0/378.0 - 2825.0/27648.0, 0.0, 250.0/621.0 - 18575.0/48384.0, 125.0/594.0 - 13525.0/55296.0,
   -277.0/14336.0, 512.0/1771.0 - 0.25}};

/* GSL rk2 as a tableau (for the order-condition test only; rk2_apply below uses GSL's own arithmetic form) */
static const tableau TAB_RK2 = {
  {1.0/2.0, 1.0},
  {{1.0/2.0},
   {-1.0, 2.0}},
  {1.0/6.0, 4.0/6.0, 1.0/6.0},
  {-1.0/6.0, 1.0/3.0, -1.0/6.0}};

/* test hook: copy a tableau out (kind: 0 = rk2, 2 = rkf45, 3 = rkck); ah[5], b[25] row-major, c[6], ec[6] */
int minigsl_tableau(int kind, double *ah, double *b, double *c, double *ec) {
  const tableau *T = kind == 2 ? &TAB_RKF45 : (kind == 3 ? &TAB_RKCK : (kind == 0 ? &TAB_RK2 : NULL));
  int i, j;
  if (!T) return -1;
  for (i = 0; i < 5; i++) { ah[i] = T->ah[i]; for (j = 0; j < 5; j++) b[5 * i + j] = T->b[i][j]; }
  for (i = 0; i < 6; i++) { c[i] = T->c[i]; ec[i] = T->ec[i]; }
  return 0;
}

static int embedded_apply(gsl_odeiv2_step *s, const tableau *T, double t, double h, double y[], double yerr[],
                          const double dydt_in[], double dydt_out[], const gsl_odeiv2_system *sys) {
  const size_t n = s->dim; size_t i; int st, j;
  double **k = s->k, *y0 = s->y0, *ytmp = s->ytmp;
  memcpy(y0, y, n * sizeof(double));
  if (dydt_in) memcpy(k[0], dydt_in, n * sizeof(double)); else FEVAL(sys, t, y, k[0]);
  for (st = 1; st < 6; st++) {
    for (i = 0; i < n; i++) {
      double acc = 0.0;
      for (j = 0; j < st; j++) acc += T->b[st-1][j] * k[j][i];
      ytmp[i] = y[i] + h * acc;
    }
    FEVAL(sys, t + T->ah[st-1] * h, ytmp, k[st]);
  }
  for (i = 0; i < n; i++) {
    double d = 0.0;
    for (j = 0; j < 6; j++) if (T->c[j] != 0.0) d += T->c[j] * k[j][i];
    y[i] += h * d;
  }
  if (dydt_out) {
    int s_ = sys->function(t + h, y, dydt_out, sys->params);
    if (s_ != GSL_SUCCESS) { memcpy(y, y0, n * sizeof(double)); return s_; }
  }
  for (i = 0; i < n; i++) {
    double e = 0.0;
    for (j = 0; j < 6; j++) if (T->ec[j] != 0.0) e += T->ec[j] * k[j][i];
    yerr[i] = h * e;
  }
  return GSL_SUCCESS;
}

/* classic RK4 substep on (y, y0=s->y0, k=s->k[0] holding k1 on entry) */
static int rk4_substep(gsl_odeiv2_step *s, double y[], double h, double t, const gsl_odeiv2_system *sys) {
  const size_t n = s->dim; size_t i; double *k = s->k[0], *y0 = s->y0, *ytmp = s->ytmp;
  for (i = 0; i < n; i++) { y[i] += h / 6.0 * k[i]; ytmp[i] = y0[i] + 0.5 * h * k[i]; }
  FEVAL(sys, t + 0.5 * h, ytmp, k);
  for (i = 0; i < n; i++) { y[i] += h / 3.0 * k[i]; ytmp[i] = y0[i] + 0.5 * h * k[i]; }
  FEVAL(sys, t + 0.5 * h, ytmp, k);
  for (i = 0; i < n; i++) { y[i] += h / 3.0 * k[i]; ytmp[i] = y0[i] + h * k[i]; }
  FEVAL(sys, t + h, ytmp, k);
  for (i = 0; i < n; i++) y[i] += h / 6.0 * k[i];
  return GSL_SUCCESS;
}
/* GSL "rk4": step doubling, 11 RHS evaluations, yerr = 4 (y_two_halves - y_one) / 15 */
static int rk4_apply(gsl_odeiv2_step *s, double t, double h, double y[], double yerr[],
                     const double dydt_in[], double dydt_out[], const gsl_odeiv2_system *sys) {
  const size_t n = s->dim; size_t i; int st;
  double *k = s->k[0], *k1 = s->k[1], *y0 = s->y0, *yone = s->yone, *ysave = s->k[2];
  memcpy(y0, y, n * sizeof(double)); memcpy(ysave, y, n * sizeof(double));
  if (dydt_in) memcpy(k, dydt_in, n * sizeof(double)); else FEVAL(sys, t, y0, k);
  memcpy(k1, k, n * sizeof(double));
  memcpy(yone, y, n * sizeof(double));
  if ((st = rk4_substep(s, yone, h, t, sys)) != GSL_SUCCESS) return st;
  memcpy(k, k1, n * sizeof(double));
  if ((st = rk4_substep(s, y, h / 2.0, t, sys)) != GSL_SUCCESS) { memcpy(y, ysave, n * sizeof(double)); return st; }
  { int s_ = sys->function(t + h / 2.0, y, k, sys->params); if (s_ != GSL_SUCCESS) { memcpy(y, ysave, n * sizeof(double)); return s_; } }
  memcpy(y0, y, n * sizeof(double));
  if ((st = rk4_substep(s, y, h / 2.0, t + h / 2.0, sys)) != GSL_SUCCESS) { memcpy(y, ysave, n * sizeof(double)); return st; }
  if (dydt_out) { int s_ = sys->function(t + h, y, dydt_out, sys->params); if (s_ != GSL_SUCCESS) { memcpy(y, ysave, n * sizeof(double)); return s_; } }
  for (i = 0; i < n; i++) yerr[i] = 4.0 * (y[i] - yone[i]) / 15.0;
  return GSL_SUCCESS;
}

/* GSL "rk2": embedded Runge-Kutta (2, 3).  k2 at the midpoint, k3 at y + h (-k1 + 2 k2); the third-order combination
 * (k1 + 4 k2 + k3) / 6 is propagated, yerr = h (k2 - (k1 + 4 k2 + k3) / 6), order() = 2 */
static int rk2_apply(gsl_odeiv2_step *s, double t, double h, double y[], double yerr[],
                     const double dydt_in[], double dydt_out[], const gsl_odeiv2_system *sys) {
  const size_t n = s->dim; size_t i;
  double *k1 = s->k[0], *k2 = s->k[1], *k3 = s->k[2], *ytmp = s->ytmp, *y0 = s->y0;
  memcpy(y0, y, n * sizeof(double));
  if (dydt_in) memcpy(k1, dydt_in, n * sizeof(double)); else FEVAL(sys, t, y, k1);
  for (i = 0; i < n; i++) ytmp[i] = y[i] + 0.5 * h * k1[i];
  FEVAL(sys, t + 0.5 * h, ytmp, k2);
  for (i = 0; i < n; i++) ytmp[i] = y[i] + h * (-k1[i] + 2.0 * k2[i]);
  FEVAL(sys, t + h, ytmp, k3);
  for (i = 0; i < n; i++) {
    const double ksum3 = (k1[i] + 4.0 * k2[i] + k3[i]) / 6.0;
    y[i] += h * ksum3;
    yerr[i] = h * (k2[i] - ksum3);
  }
  if (dydt_out) {
    int s_ = sys->function(t + h, y, dydt_out, sys->params);
    if (s_ != GSL_SUCCESS) { memcpy(y, y0, n * sizeof(double)); return s_; }
  }
  return GSL_SUCCESS;
}

/* GSL "rk8pd": Prince-Dormand 8(7) with 13 stages (RK8(7)13M).  The 8th-order weights Abar are propagated,
 * yerr = h (ksum7 - ksum8) with the 7th-order weights A, order() = 8.  Sums run over the non-zero coefficients in stage order. */
static const double RK8_AH[12] = {
   1.0 / 18.0, 1.0 / 12.0, 1.0 / 8.0, 5.0 / 16.0, 3.0 / 8.0, 59.0 / 400.0, 93.0 / 200.0, 5490023248.0 / 9719169821.0, 13.0 / 20.0,
   1201146811.0 / 1299019798.0, 1.0, 1.0};
static const double RK8_B[12][12] = {
   {1.0 / 18.0},
   {1.0 / 48.0, 1.0 / 16.0},
   {1.0 / 32.0, 0.0, 3.0 / 32.0},
   {5.0 / 16.0, 0.0, -75.0 / 64.0, 75.0 / 64.0},
   {3.0 / 80.0, 0.0, 0.0, 3.0 / 16.0, 3.0 / 20.0},
   {29443841.0 / 614563906.0, 0.0, 0.0, 77736538.0 / 692538347.0, -28693883.0 / 1125000000.0, 23124283.0 / 1800000000.0},
   {16016141.0 / 946692911.0, 0.0, 0.0, 61564180.0 / 158732637.0, 22789713.0 / 633445777.0, 545815736.0 / 2771057229.0,
    -180193667.0 / 1043307555.0},
   {39632708.0 / 573591083.0, 0.0, 0.0, -433636366.0 / 683701615.0, -421739975.0 / 2616292301.0, 100302831.0 / 723423059.0,
    790204164.0 / 839813087.0, 800635310.0 / 3783071287.0},
   {246121993.0 / 1340847787.0, 0.0, 0.0, -37695042795.0 / 15268766246.0, -309121744.0 / 1061227803.0, -12992083.0 / 490766935.0,
    6005943493.0 / 2108947869.0, 393006217.0 / 1396673457.0, 123872331.0 / 1001029789.0},
   {-1028468189.0 / 846180014.0, 0.0, 0.0, 8478235783.0 / 508512852.0, 1311729495.0 / 1432422823.0, -10304129995.0 / 1701304382.0,
    -48777925059.0 / 3047939560.0, 15336726248.0 / 1032824649.0, -45442868181.0 / 3398467696.0, 3065993473.0 / 597172653.0},
   {185892177.0 / 718116043.0, 0.0, 0.0, -3185094517.0 / 667107341.0, -477755414.0 / 1098053517.0, -703635378.0 / 230739211.0,
    5731566787.0 / 1027545527.0, 5232866602.0 / 850066563.0, -4093664535.0 / 808688257.0, 3962137247.0 / 1805957418.0,
    65686358.0 / 487910083.0},
   {403863854.0 / 491063109.0, 0.0, 0.0, -5068492393.0 / 434740067.0, -411421997.0 / 543043805.0, 652783627.0 / 914296604.0,
    11173962825.0 / 925320556.0, -13158990841.0 / 6184727034.0, 3936647629.0 / 1978049680.0, -160528059.0 / 685178525.0,
    248638103.0 / 1413531060.0, 0.0}};
static const double RK8_ABAR[13] = {
   14005451.0 / 335480064.0, 0.0, 0.0, 0.0, 0.0, -59238493.0 / 1068277825.0, 181606767.0 / 758867731.0, 561292985.0 / 797845732.0,
   -1041891430.0 / 1371343529.0, 760417239.0 / 1151165299.0, 118820643.0 / 751138087.0, -528747749.0 / 2220607170.0, 1.0 / 4.0};
static const double RK8_A[13] = {
   13451932.0 / 455176623.0, 0.0, 0.0, 0.0, 0.0, -808719846.0 / 976000145.0, 1757004468.0 / 5645159321.0, 656045339.0 / 265891186.0,
   -3867574721.0 / 1518517206.0, 465885868.0 / 322736535.0, 53011238.0 / 667516719.0, 2.0 / 45.0, 0.0};
/* test hook: the tableau as dense arrays, b[144] row-major */
int minigsl_rk8pd_tableau(double *ah, double *b, double *abar, double *a7) {
  int i, j;
  for (i = 0; i < 12; i++) { ah[i] = RK8_AH[i]; for (j = 0; j < 12; j++) b[12 * i + j] = RK8_B[i][j]; }
  for (i = 0; i < 13; i++) { abar[i] = RK8_ABAR[i]; a7[i] = RK8_A[i]; }
  return 0;
}
static int rk8pd_apply(gsl_odeiv2_step *s, double t, double h, double y[], double yerr[],
                       const double dydt_in[], double dydt_out[], const gsl_odeiv2_system *sys) {
  const size_t n = s->dim; size_t i; int st, j;
  double **k = s->k, *y0 = s->y0, *ytmp = s->ytmp;
  memcpy(y0, y, n * sizeof(double));
  if (dydt_in) memcpy(k[0], dydt_in, n * sizeof(double)); else FEVAL(sys, t, y, k[0]);
  for (st = 1; st < 13; st++) {
    for (i = 0; i < n; i++) {
      double acc = 0.0;
      for (j = 0; j < st; j++) if (RK8_B[st-1][j] != 0.0) acc += RK8_B[st-1][j] * k[j][i];
      ytmp[i] = y[i] + h * acc;
    }
    FEVAL(sys, t + RK8_AH[st-1] * h, ytmp, k[st]);
  }
  for (i = 0; i < n; i++) {
    double ksum8 = 0.0;
    for (j = 0; j < 13; j++) if (RK8_ABAR[j] != 0.0) ksum8 += RK8_ABAR[j] * k[j][i];
    y[i] += h * ksum8;
  }
  if (dydt_out) {
    int s_ = sys->function(t + h, y, dydt_out, sys->params);
    if (s_ != GSL_SUCCESS) { memcpy(y, y0, n * sizeof(double)); return s_; }
  }
  for (i = 0; i < n; i++) {
    double ksum8 = 0.0, ksum7 = 0.0;
    for (j = 0; j < 13; j++) if (RK8_ABAR[j] != 0.0) ksum8 += RK8_ABAR[j] * k[j][i];
    for (j = 0; j < 13; j++) if (RK8_A[j] != 0.0) ksum7 += RK8_A[j] * k[j][i];
    yerr[i] = h * (ksum7 - ksum8);
  }
  return GSL_SUCCESS;
}

static int step_apply(gsl_odeiv2_step *s, double t, double h, double y[], double yerr[],
                      const double dydt_in[], double dydt_out[], const gsl_odeiv2_system *sys) {
  switch (s->type->kind) {
    case MG_RKF45: return embedded_apply(s, &TAB_RKF45, t, h, y, yerr, dydt_in, dydt_out, sys);
    case MG_RKCK:  return embedded_apply(s, &TAB_RKCK, t, h, y, yerr, dydt_in, dydt_out, sys);
    case MG_RK4:   return rk4_apply(s, t, h, y, yerr, dydt_in, dydt_out, sys);
    case MG_RK2:   return rk2_apply(s, t, h, y, yerr, dydt_in, dydt_out, sys);
    case MG_RK8PD: return rk8pd_apply(s, t, h, y, yerr, dydt_in, dydt_out, sys);
    default:
      fprintf(stderr, "minigsl: stepper '%s' is not restated (no GSL source offline)\n", s->type->name);
      abort();
  }
}

/* standard control, a_y = 1, a_dydt = 0 (gsl_odeiv2_control_y_new) */
enum { HADJ_DEC = -1, HADJ_NIL = 0, HADJ_INC = 1 };
static int hadjust(const gsl_odeiv2_control *c, size_t dim, unsigned ord, const double y[], const double yerr[],
                   const double yp[], double *h) {
  const double S = 0.9, h_old = *h; double rmax = DBL_MIN; size_t i;
  for (i = 0; i < dim; i++) {
    const double D0 = c->eps_rel * (c->a_y * fabs(y[i]) + c->a_dydt * fabs(h_old * yp[i])) + c->eps_abs;
    const double r = fabs(yerr[i]) / fabs(D0);
    rmax = r > rmax ? r : rmax;   /* GSL_MAX_DBL(r, rmax): a NaN r is dropped */
  }
  if (rmax > 1.1) {
    double r = S / pow(rmax, 1.0 / ord);
    if (r < 0.2) r = 0.2;
    *h = r * h_old; return HADJ_DEC;
  } else if (rmax < 0.5) {
    double r = S / pow(rmax, 1.0 / (ord + 1.0));
    if (r > 5.0) r = 5.0;
    if (r < 1.0) r = 1.0;
    *h = r * h_old; return HADJ_INC;
  }
  return HADJ_NIL;
}

int gsl_odeiv2_evolve_apply(gsl_odeiv2_evolve *e, gsl_odeiv2_control *con, gsl_odeiv2_step *step,
                            const gsl_odeiv2_system *dydt, double *t, double t1, double *h, double y[]) {
  const double t0 = *t; double h0 = *h; int step_status, final_step = 0; const double dt = t1 - t0;
  if (e->dim != step->dim) return GSL_EINVAL;
  if ((dt < 0.0 && h0 > 0.0) || (dt > 0.0 && h0 < 0.0)) return GSL_EINVAL;
  memcpy(e->y0, y, e->dim * sizeof(double));
  if (e->count == 0) { int s_ = dydt->function(t0, y, e->dydt_in, dydt->params); if (s_) return s_; }
  else memcpy(e->dydt_in, e->dydt_out, e->dim * sizeof(double));
try_step:
  if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) { h0 = dt; final_step = 1; } else final_step = 0;
  step_status = step_apply(step, t0, h0, y, e->yerr, e->dydt_in, e->dydt_out, dydt);
  if (step_status == GSL_EFAULT || step_status == GSL_EBADFUNC) return step_status;
  if (step_status != GSL_SUCCESS) {
    const double h_old = h0; h0 *= 0.5;
    { volatile double t_curr = *t, t_next = (*t) + h0;
      if (fabs(h0) < fabs(h_old) && t_next != t_curr) { memcpy(y, e->y0, dydt->dimension * sizeof(double)); e->failed_steps++; goto try_step; }
      else { *h = h0; *t = t0; return step_status; } }
  }
  e->count++; e->last_step = h0;
  *t = final_step ? t1 : t0 + h0;
  if (con != NULL) {
    const double h_old = h0;
    const int hs = hadjust(con, step->dim, step->type->order, y, e->yerr, e->dydt_out, &h0);
    if (hs == HADJ_DEC) {
      volatile double t_curr = *t, t_next = (*t) + h0;
      if (fabs(h0) < fabs(h_old) && t_next != t_curr) { memcpy(y, e->y0, dydt->dimension * sizeof(double)); e->failed_steps++; goto try_step; }
      else { *h = h0; return GSL_FAILURE; }
    }
  }
  if (final_step == 0) *h = h0;
  return step_status;
}

/* ------------------------------------------------------------------ cheb / sf / pow_int */
gsl_cheb_series *gsl_cheb_alloc(const size_t order) {
  gsl_cheb_series *cs = calloc(1, sizeof *cs);
  cs->order = order; cs->order_sp = order; cs->c = calloc(order + 1, sizeof(double)); cs->f = calloc(order + 1, sizeof(double));
  return cs;
}
void gsl_cheb_free(gsl_cheb_series *cs) { if (!cs) return; free(cs->c); free(cs->f); free(cs); }
double *gsl_cheb_coeffs(const gsl_cheb_series *cs) { return cs->c; }
int gsl_cheb_init(gsl_cheb_series *cs, const gsl_function *func, const double a, const double b) {
  size_t k, j; const double bma = 0.5 * (b - a), bpa = 0.5 * (b + a), fac = 2.0 / (cs->order + 1.0);
  if (a >= b) return GSL_EDOM;
  cs->a = a; cs->b = b;
  for (k = 0; k <= cs->order; k++) { double y = cos(M_PI * (k + 0.5) / (cs->order + 1)); cs->f[k] = GSL_FN_EVAL(func, (y * bma + bpa)); }
  for (j = 0; j <= cs->order; j++) {
    double sum = 0.0;
    for (k = 0; k <= cs->order; k++) sum += cs->f[k] * cos(M_PI * j * (k + 0.5) / (cs->order + 1));
    cs->c[j] = fac * sum;
  }
  return GSL_SUCCESS;
}
int gsl_cheb_eval_n_err(const gsl_cheb_series *cs, const size_t order, const double x, double *result, double *abserr) {
  size_t i; double d1 = 0.0, d2 = 0.0, absc = 0.0;
  const double y = (2.0 * x - cs->a - cs->b) / (cs->b - cs->a), y2 = 2.0 * y;
  const size_t eo = order < cs->order ? order : cs->order;
  for (i = eo; i >= 1; i--) { double temp = d1; d1 = y2 * d1 - d2 + cs->c[i]; d2 = temp; }
  *result = y * d1 - d2 + 0.5 * cs->c[0];
  for (i = 0; i <= eo; i++) absc += fabs(cs->c[i]);
  *abserr = fabs(cs->c[eo]) + absc * GSL_DBL_EPSILON;
  return GSL_SUCCESS;
}
double gsl_sf_choose(unsigned int n, unsigned int m) {
  double r = 1.0; unsigned int k;
  if (m > n) return 0.0;
  if (m > n - m) m = n - m;
  for (k = 1; k <= m; k++) r = r * (double)(n - m + k) / (double)k;   /* exact for the n <= 51 used by lrexp */
  return floor(r + 0.5);
}
double gsl_pow_int(double x, int n) {
  double value = 1.0; unsigned int u;
  if (n < 0) { x = 1.0 / x; u = (unsigned int)(-n); } else u = (unsigned int)n;
  do { if (u & 1) value *= x; u >>= 1; x *= x; } while (u);
  return value;
}
