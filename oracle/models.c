/* TEST INFRASTRUCTURE — the fixture models of models/ (.pop files) written out by hand as plain C, the way the
 * reference's generator lays out ode_rhs / %M / %C / ode_init_conditions (src/odexp:981-1069, 874-954; a concrete
 * instance is quoted in SURVEY.md Appendix B).  Hand-written on purpose: the oracle must not share the product's
 * code generator. */
#include <math.h>
#include <string.h>
#include "odx_oracle.h"

/* ---------------------------------------------------------------- kuramoto.pop */
#define KPI 3.141592
static const double kur_pars[] = {0.1};
static void kur_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                    double *aux, double *death, double *repli) {
  const double K = c->mu[0], omega = e[0], theta = y_[0], A1 = mfd[0], A2 = mfd[1];
  const double y = sin(theta);
  const double coupling = (A1 * cos(theta) - A2 * sin(theta));
  (void)t; (void)psi; (void)death; (void)repli;
  f[0] = omega + K * coupling;
  aux[0] = y; aux[1] = coupling;
}
static double kur_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)aux; (void)mfd;
  return k == 0 ? sin(y[0]) : (k == 1 ? cos(y[0]) : e[0]);
}
static void kur_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t;
  e[0] = orc_randN(c, 2 * KPI / 10.0, 0.1);
  y[0] = 2 * KPI * orc_rand01(c);
}

/* ---------------------------------------------------------------- nbody.pop */
#define NB_G 6.674e-11
static const double nb_pars[] = {1e+09};
static double normCube(double x, double y, double z) { return pow(x * x + y * y + z * z, 1.5); }
static double grav(double numx, double numxi, double x, double y, double z, double xi, double yi, double zi) {
  return (numx == numxi) ? 0.0 : (numx - numxi) / normCube(x - xi, y - yi, z - zi);
}
static void nb_rhs(orc_ctx *c, double t, const double *y, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  (void)c; (void)t; (void)e; (void)mfd; (void)aux; (void)death; (void)repli;
  f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
  f[3] = psi[0]; f[4] = psi[1]; f[5] = psi[2];
}
static void nb_cpl(orc_ctx *c, const double *y, const double *e, const double *oy, const double *oe, double *psi) {
  const double mass = c->mu[0];
  (void)e; (void)oe;
  psi[0] += -NB_G * mass * grav(y[0], oy[0], y[0], y[1], y[2], oy[0], oy[1], oy[2]);
  psi[1] += -NB_G * mass * grav(y[1], oy[1], y[0], y[1], y[2], oy[0], oy[1], oy[2]);
  psi[2] += -NB_G * mass * grav(y[2], oy[2], y[0], y[1], y[2], oy[0], oy[1], oy[2]);
}
static void nb_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t; (void)e;
  y[0] = orc_rand01(c); y[1] = orc_rand01(c); y[2] = orc_rand01(c);
  y[3] = 0.0; y[4] = 0.0; y[5] = 0.0;
}

/* ---------------------------------------------------------------- nbody_const.pop (constant arrays indexed by ID, as in examples/Nbody.pop) */
static const double nbc_pars[] = {1, 1e+09};
static const double nbc_x_ic[] = { 0.0, 1.0, 0.5 };
static const double nbc_y_ic[] = { 0.0, 0.0, 0.8660254037844386 };
static const double nbc_z_ic[] = { 0.0, 0.0, 0.0 };
static const double nbc_vx0[] = { 0.1, 0.0, -.1 };
static const double nbc_vy0[] = { 0.0, 0.1, 0.0 };
static const double nbc_vz0[] = { 0.0, 0.0, 0.0 };
static void nbc_cpl(orc_ctx *c, const double *y, const double *e, const double *oy, const double *oe, double *psi) {
  const double mass = c->mu[1];
  (void)e; (void)oe;
  psi[0] += -NB_G * mass * grav(y[0], oy[0], y[0], y[1], y[2], oy[0], oy[1], oy[2]);
  psi[1] += -NB_G * mass * grav(y[1], oy[1], y[0], y[1], y[2], oy[0], oy[1], oy[2]);
  psi[2] += -NB_G * mass * grav(y[2], oy[2], y[0], y[1], y[2], oy[0], oy[1], oy[2]);
}
static void nbc_ic(orc_ctx *c, double t, double *y, double *e) {
  const double v0 = c->mu[0]; const int ID = c->id;
  (void)t; (void)e;
  y[0] = nbc_x_ic[ID]; y[1] = nbc_y_ic[ID]; y[2] = nbc_z_ic[ID];
  y[3] = v0*nbc_vx0[ID]; y[4] = v0*nbc_vy0[ID]; y[5] = v0*nbc_vz0[ID];
}

/* ---------------------------------------------------------------- chain_array.pop (LET NC 4): array D/DT, AUX and EXPR entries */
static const double ca_pars[] = {1.5, 0.3};
static double ca_sum(const double *array, long len) { double s = 0.0; long i; for (i = 0; i < len; i++) s += array[i]; return s; }      /* src/utils_odexp.c:11-21 */
static double ca_dotprod(const double *x, const double *y, long len) { double s = 0.0; long i; for (i = 0; i < len; i++) s += x[i] * y[i]; return s; }
static void ca_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double k = c->mu[0], a = c->mu[1];
  const double *w = e, z = y_[0], *X = y_ + 1, u = y_[5], SBAR = mfd[0], ZBAR = mfd[1];
  double sq[4], nrm, dp; size_t i;
  (void)t; (void)psi; (void)death; (void)repli;
  for (i = 0; i < 4; ++i) sq[i] = X[i]*X[i];
  nrm = sqrt(ca_sum(sq,4));
  dp = ca_dotprod(X,w,4);
  f[0] = -z + 0.1*dp;
  for (i = 0; i < 4; ++i) f[1 + i] = (i == 0) ? (-k*X[0] + a*SBAR) : k*(X[i-1] - X[i]);
  f[5] = X[4-1] - a*u*ZBAR;
  for (i = 0; i < 4; ++i) aux[i] = sq[i];
  aux[4] = nrm; aux[5] = dp;
}
static double ca_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)e; (void)mfd;
  return k == 0 ? aux[4] : y[0];                        /* %M SBAR MA(nrm); %M ZBAR MY(z) */
}
static void ca_ic(orc_ctx *c, double t, double *y, double *e) {
  size_t i;
  (void)t;
  for (i = 0; i < 4; ++i) e[i] = 0.5 + orc_rand01(c);
  y[0] = 1.0;
  for (i = 0; i < 4; ++i) y[1 + i] = orc_rand01(c)/(i+1);
  y[5] = 0.0;
}

/* ---------------------------------------------------------------- fun_block.pop (LET NW 3): FUN ... end blocks, pointer arguments */
static const double fb_pars[] = {2.0, 0.5};
static double fb_mean(double *x) { return ca_sum(x,3)/3; }
static double fb_wmean(double *x, double *w) {
  double s = 0.0, sw = 0.0;
  int j;
  for (j = 0;j < 3;j++) { s += w[j]*x[j]; sw += w[j]; }
  return s/sw;
}
static double fb_clip(double x, double lo, double hi) {
  if (x < lo) { return lo; }
  if (x > hi) { return hi; }
  return x;
}
static double fb_fill(double *a, double x0, double n) {
  int j;
  for (j = 0;j < n;j++) { a[j] = x0*(j + 1); }
  return 0.0;
}
static void fb_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double a = c->mu[0], b = c->mu[1], MBAR = mfd[0];
  double w[3], c_[3], X[3], pw[3], m, wm; size_t i;
  (void)psi; (void)death; (void)repli;
  for (i = 0; i < 3; ++i) { w[i] = e[i]; c_[i] = e[3 + i]; X[i] = y_[i]; }
  m = fb_mean(X);
  wm = fb_wmean(X, w);
  fb_fill(pw, m, 3);                                     /* aux pw[i=0:NW] fill(pw, m, NW) {vect} */
  for (i = 0; i < 3; ++i) f[i] = -a*(X[i] - fb_clip(MBAR, -b, b)) + sin(t + i) + 0.1*pw[i] - c_[i];
  aux[0] = m; aux[1] = wm;
  for (i = 0; i < 3; ++i) aux[2 + i] = pw[i];
}
static double fb_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)y; (void)e; (void)mfd;
  return aux[1];                                         /* %M MBAR MA(wm) */
}
static void fb_ic(orc_ctx *c, double t, double *y, double *e) {
  size_t i;
  (void)t;
  for (i = 0; i < 3; ++i) e[i] = 0.5 + orc_rand01(c);
  fb_fill(e + 3, 0.25, 3);                               /* expr c[i=0:NW] fill(c, 0.25, NW) {vect} */
  for (i = 0; i < 3; ++i) y[i] = 2*orc_rand01(c) - 1;
}

/* ---------------------------------------------------------------- helper_calls.pop (LET NW 4): the helper functions of src/utils_odexp.c
 * restated next to their call sites: sumsub :23-33, sumstep :35-45, sum3 :47-69, prod :71-84, conv :99-111, minus / plus / identity
 * :113-126, sumxy :132-143, kern :149-160 */
static const double hc_pars[] = {1.5, 0.3};
static const long hc_pick[] = {0, 2, 3};
static double hc_sumsub(const double *array, const long *ind, long len) { double s = 0.0; long i; for (i = 0; i < len; i++) s += array[ind[i]]; return s; }
static double hc_sumstep(const double *array, long len, long step) { double s = 0.0; long i; for (i = 0; i < len; i += step) s += array[i]; return s; }
static double hc_sum3(const double *array, long d1, long d2, long d3, long dim) {
  double s = 0.0; long i, j, k; long n[3];
  n[0] = d1; n[1] = d2; n[2] = d3; n[dim - 1] = 1;
  for (i = 0; i < n[0]; i++) for (j = 0; j < n[1]; j++) for (k = 0; k < n[2]; k++) s += array[i + d2 * j + d3 * k];
  return s;
}
static double hc_prod(const double *array, long len) { double p = 1.0; long i; for (i = 0; i < len; i++) p *= array[i]; return p; }
static double hc_conv(const double *u, const double *v, long len) { double s = 0.0; long i; for (i = 0; i < len; i++) s += u[len - 1 - i] * v[i]; return s; }
static double hc_minus(double x, double y) { return x - y; }
static double hc_plus(double x, double y) { return x + y; }
static double hc_identity(double x) { return x; }
static double hc_sumxy(long len, double (*f)(double), double (*g)(double, double), const double *x, const double yi) {
  double s = 0.0; long j; for (j = 0; j < len; j++) s += f(g(x[j], yi)); return s;
}
static double hc_kern(const double *Wi, double (*f)(double, double, double *), double xi, const double *x, double *p, long len) {
  double s = 0.0; long j; for (j = 0; j < len; j++) s += Wi[j] * f(xi, x[j], p); return s;
}
static double hc_kf(double xi, double xj, double *p) { return p[0]*exp(-p[1]*(xi-xj)*(xi-xj)); }
static void hc_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double a = c->mu[0], b = c->mu[1], MB = mfd[0];
  double w[4], q[2], X[4], sxy[4], kn[4], s_sub, s_step, s_3, p_all, cv; size_t i;
  (void)t; (void)psi; (void)death; (void)repli;
  for (i = 0; i < 4; ++i) w[i] = e[i];
  for (i = 0; i < 2; ++i) q[i] = e[4 + i];
  for (i = 0; i < 4; ++i) X[i] = y_[i];
  s_sub = hc_sumsub(X, hc_pick, 3);
  s_step = hc_sumstep(X, 4, 2);
  s_3 = hc_sum3(X, 2, 2, 1, 1);
  p_all = hc_prod(w, 4);
  cv = hc_conv(X, w, 4);
  for (i = 0; i < 4; ++i) sxy[i] = hc_sumxy(4, hc_identity, hc_minus, X, X[i]);
  for (i = 0; i < 4; ++i) kn[i] = hc_kern(w, hc_kf, X[i], X, q, 4);
  for (i = 0; i < 4; ++i) f[i] = -a*X[i] + b*hc_plus(sxy[i], kn[i])/4 + 0.01*(s_sub + s_step + s_3)*p_all + 0.1*MB;
  aux[0] = s_sub; aux[1] = s_step; aux[2] = s_3; aux[3] = p_all; aux[4] = cv;
  for (i = 0; i < 4; ++i) aux[5 + i] = sxy[i];
  for (i = 0; i < 4; ++i) aux[9 + i] = kn[i];
}
static double hc_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)y; (void)e; (void)mfd;
  return aux[4];                                         /* %M MB MA(cv) */
}
static void hc_ic(orc_ctx *c, double t, double *y, double *e) {
  size_t i;
  (void)t;
  for (i = 0; i < 4; ++i) e[i] = 0.5 + orc_rand01(c);
  for (i = 0; i < 2; ++i) e[4 + i] = 0.2*(i+1);
  for (i = 0; i < 4; ++i) y[i] = 2*orc_rand01(c) - 1;
}

/* ---------------------------------------------------------------- delay_pop.pop (LET NP 16): delay() with one history stack per particle */
static const double dp_pars[] = {1.0, 2.0, 0.4};
static orc_stack dp_hist[16];
static double dp_x0(double s) { return 0.5 + 0.1*sin(s); }
static void dp_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double tau = c->mu[0], beta = c->mu[1], k = c->mu[2], gam = e[0], x = y_[0], XM = mfd[0];
  const double xd = orc_delay(c, t, x, tau, dp_x0, &dp_hist[c->id]);
  (void)psi; (void)death; (void)repli;
  f[0] = beta*xd/(1 + xd*xd*xd*xd) - gam*x + k*(XM - x);
  aux[0] = xd;
}
static double dp_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];
}
static void dp_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t;
  e[0] = 0.8 + 0.4*orc_rand01(c);
  y[0] = 0.5 + 0.2*orc_rand01(c);
}

/* ---------------------------------------------------------------- kuramoto_lrexp.pop (LET N 200): unweighted %C ... lrexp */
static const double kle_pars[] = {0.8};
static void kle_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                    double *aux, double *death, double *repli) {
  const double K = c->mu[0], omega = e[0], CPL = psi[0];
  (void)t; (void)y_; (void)mfd; (void)death; (void)repli;
  f[0] = omega + K/200*CPL;
  aux[0] = CPL;
}
static void kle_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t;
  e[0] = orc_randN(c, 2 * 3.141592653589793 / 10.0, 0.1);
  y[0] = 2 * 3.141592653589793 * orc_rand01(c);
}
static double kle_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)e; (void)aux; (void)mfd;
  return k == 0 ? sin(y[0]) : cos(y[0]);
}

/* ---------------------------------------------------------------- pair_aux.pop: MA() and OE() inside an all-pairs %C */
static const double pa_pars[] = {0.5, 0.1};
static void pa_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double g = c->mu[0], cc = c->mu[1], q = e[0], x = y_[0], v = y_[1], FRC = psi[0], DRG = psi[1], XM = mfd[0];
  const double s = q*(1 + x*x) + cc*XM;
  const double w = q/(1 + v*v);
  const double fo = FRC;
  (void)t; (void)death; (void)repli;
  f[0] = v;
  f[1] = -x + g*FRC - cc*XM + 0.2*DRG;
  aux[0] = s; aux[1] = w; aux[2] = fo;
}
static void pa_cpl(orc_ctx *c, const double *y, const double *e, const double *oy, const double *oe, double *psi) {
  (void)e;
  psi[0] += c->my_aux[0]*oe[0]*sin(oy[0] - y[0]);        /* MA(s)*OE(q)*sin(OY(x) - MY(x)) */
  psi[1] += c->other_aux[1]*(oy[1] - y[1]);              /* OA(w)*(OY(v) - MY(v)) */
}
static double pa_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];
}
static void pa_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t;
  e[0] = 0.5 + orc_rand01(c);
  y[0] = 2*orc_rand01(c) - 1;
  y[1] = 0.0;
}

/* ---------------------------------------------------------------- repel2d.pop: softened antisymmetric all-pairs %C + two %M */
static const double rp_pars[] = {0.5, 0.05, 0.3};
static double rp_soft(double dx, double dy, double e) { return pow(dx*dx + dy*dy + e*e, 1.5); }
static double rp_rep(double a, double b, double dx, double dy, double e) { return (a == b) ? 0.0 : (a - b)/rp_soft(dx, dy, e); }
static void rp_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double gam = c->mu[2], x = y_[0], y = y_[1], u = y_[2], v = y_[3], FX = psi[0], FY = psi[1], CX = mfd[0], CY = mfd[1];
  (void)t; (void)e; (void)aux; (void)death; (void)repli;
  f[0] = u;
  f[1] = v;
  f[2] = FX - gam*u - 0.1*(x - CX);
  f[3] = FY - gam*v - 0.1*(y - CY);
}
static void rp_cpl(orc_ctx *c, const double *y, const double *e, const double *oy, const double *oe, double *psi) {
  (void)e; (void)oe;
  psi[0] += c->mu[0]*rp_rep(y[0], oy[0], y[0] - oy[0], y[1] - oy[1], c->mu[1]);
  psi[1] += c->mu[0]*rp_rep(y[1], oy[1], y[0] - oy[0], y[1] - oy[1], c->mu[1]);
}
static double rp_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)e; (void)aux; (void)mfd;
  return k == 0 ? y[0] : y[1];
}
static void rp_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t; (void)e;
  y[0] = orc_rand01(c); y[1] = orc_rand01(c);
  y[2] = 0.0; y[3] = 0.0;
}

/* ---------------------------------------------------------------- lineage.pop: SE / SA / SY, exprand / poisson in expressions */
static const double ln_pars[] = {0.8, 0.05, 0.6};
static void ln_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double r = c->mu[0], b = c->mu[2], life = e[0], gen = e[2], x = y_[0];
  const double size = x*(1 + 0.1*gen);
  (void)t; (void)psi; (void)mfd;
  f[0] = r*x*(1 - x) - 0.01*life*x;
  aux[0] = size;
  *death = c->mu[1];                                   /* %DEATH MU(d) */
  *repli = b*(x > 0.5);
}
static double ln_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)aux; (void)mfd;
  return k == 0 ? y[0] : e[2];                         /* %M XM MY(x); %M GM ME(gen) */
}
static double ln_birth(orc_ctx *c, const double *y, const double *e, const double *aux) { (void)c; (void)y; (void)e; (void)aux; return 0.3; }
static void ln_ic(orc_ctx *c, double t, double *y, double *e) {
  const orc_par *sis = c->sister;
  double life, kids, gen, tag;
  (void)t;
  life = orc_exprand(c, 2.0);
  kids = orc_poisson(c, 3.0);
  gen = c->atbirth ? (0.0) : ((sis == NULL ? 0.0 : sis->expr[2]) + 1);                 /* BR(0.0, SE(gen) + 1) */
  tag = (sis != NULL)*10 + kids;                                                       /* ISDAUGHTER*10 + kids */
  e[0] = life; e[1] = kids; e[2] = gen; e[3] = tag;
  y[0] = c->atbirth ? (0.1 + 0.05*kids)
                    : (0.5*(sis == NULL ? 0.0 : sis->aux[0])/(1 + 0.1*gen) + 0.25*(sis == NULL ? 0.0 : sis->y[0]));
}

/* ---------------------------------------------------------------- table_drive.pop: FILE table indexed by ID */
static const double td_pars[] = {0.7};
static const double td_tab[4][3] = {
    {0.0,1.0,0.1},
    {0.5,0.8,0.2},
    {1.0,0.6,0.3},
    {1.5,0.4,0.4}};                                      /* table_drive.csv { nrow = 4, ncol = 3, sep = ",", skip = 1 } */
static void td_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double k = c->mu[0], x = y_[0], XM = mfd[0]; const int ID = c->id;
  const double amp = td_tab[ID % 4][1];
  (void)e; (void)psi; (void)death; (void)repli;
  f[0] = -k*(x - td_tab[ID % 4][2]) + amp*sin(t + td_tab[ID % 4][0]) - 0.1*XM;
  aux[0] = amp;
}
static double td_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];
}
static void td_ic(orc_ctx *c, double t, double *y, double *e) {
  const int ID = c->id;
  (void)t; (void)e;
  y[0] = td_tab[ID % 4][2] + 0.1*orc_rand01(c);
}

/* ---------------------------------------------------------------- interp_drive.pop: interp() on constant user arrays */
static const double id_pars[] = {0.8};
static const double id_tx[] = { 0.0, 0.5, 1.25, 2.0, 3.5, 4.0 }, id_ty[] = { 0.0, 1.0, 0.25, -0.5, 0.75, 0.0 };
static const double id_rx[] = { 2.0, 3.5, 4.0, 0.0, 0.5, 1.25 }, id_ry[] = { -0.5, 0.75, 0.0, 0.0, 1.0, 0.25 };
static double id_interp(const double *x, const double *y, double xi, int nx, int start) {      /* src/utils_odexp.c:173-203 */
  int i1 = 0, i9 = nx - 1, i5;
  while (i9 > i1 + 1) {
    i5 = (i1 + i9 + 1)/2;
    if (x[(start + i5) % nx] > xi) { i9 = i5; } else { i1 = i5; }
  }
  if (i1 == i9) return y[(i1 + start) % nx];
  return y[(i1 + start) % nx] + (y[(i9 + start) % nx] - y[(i1 + start) % nx])*(xi - x[(i1 + start) % nx])/(x[(i9 + start) % nx] - x[(i1 + start) % nx]);
}
static void id_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double k = c->mu[0], ph = e[0], x = y_[0], XM = mfd[0];
  const double drive = id_interp(id_tx, id_ty, fmod(t + ph, 4.0), 6, 0);
  const double rot = id_interp(id_rx, id_ry, fmod(t + ph, 4.0), 6, 3);
  (void)psi; (void)death; (void)repli;
  f[0] = -k*x + drive + 0.5*rot - 0.1*XM;
  aux[0] = drive; aux[1] = rot;
}
static double id_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];
}
static void id_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t;
  e[0] = 0.5*orc_rand01(c);
  y[0] = 0.1*orc_rand01(c);
}

/* ---------------------------------------------------------------- lag_chain.pop (LET L 5): LAG -> linear chain of hidden variables */
static const double lg_pars[] = {2.0, 1.0};
static double lg_linchaindelay(const double root, const double *chain, const int link, const double delay, const int len) {
  double beta = (double)len/delay;                       /* src/utils_odexp.c:162-167 */
  return beta*((link==0 ? root : chain[link-1]) - chain[link]);
}
static void lg_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double tau = c->mu[0], k = c->mu[1], *xd_chain = y_, x = y_[5], XM = mfd[0];
  const double xd = xd_chain[5-1];
  size_t i;
  (void)t; (void)e; (void)psi; (void)death; (void)repli;
  for (i = 0; i < 5; ++i) f[i] = lg_linchaindelay(x,xd_chain,i,tau,5);
  f[5] = k*(1 - xd) - 0.2*x*XM;
  aux[0] = xd;
}
static double lg_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[5];
}
static void lg_ic(orc_ctx *c, double t, double *y, double *e) {
  size_t i;
  (void)t; (void)e;
  for (i = 0; i < 5; ++i) y[i] = 0.5;
  y[5] = orc_rand01(c);
}

/* ---------------------------------------------------------------- vdp_sde.pop */
static const double vdp_pars[] = {1.0, 0.1, 0.01, 0.02};
static void vdp_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                    double *aux, double *death, double *repli) {
  const double mu = c->mu[0], sigma = c->mu[1], b = c->mu[3];      /* d is {impl}: reached through MU(d) only */
  const double y = y_[0], x = y_[1];
  (void)t; (void)e; (void)psi; (void)mfd; (void)aux;
  f[0] = mu * ((1 - x * x) * y - x) + sigma * orc_dwdt(c);
  f[1] = y;
  *death = c->mu[2];
  *repli = b * (x > 0);
}
static double vdp_birth(orc_ctx *c, const double *y, const double *e, const double *aux) { (void)c; (void)y; (void)e; (void)aux; return 0.5; }
static void vdp_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t; (void)e;
  y[0] = 0.0;
  y[1] = c->atbirth ? 2.0 : (c->sister == NULL ? 0.0 : c->sister->y[1]);      /* BR(2.0, SY(x)) */
}

/* ---------------------------------------------------------------- kuramoto_lowrank.pop (LET R 8, LET N 256) */
#define LPI 3.141592653589793
#define LR_R 8
static const double klr_pars[] = {0.5};
static void klr_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                    double *aux, double *death, double *repli) {
  const double K = c->mu[0], omega = e[0], theta = y_[0], CPL = psi[0];
  const double y = sin(theta);
  const double cpl = CPL;
  (void)t; (void)mfd; (void)death; (void)repli;
  f[0] = omega + K / 256 * CPL;
  aux[0] = y; aux[1] = cpl;
}
static double klr_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)e; (void)aux; (void)mfd;
  return k == 0 ? sin(y[0]) : cos(y[0]);
}
static void klr_ic(orc_ctx *c, double t, double *y, double *e) {
  int i;
  (void)t;
  e[0] = orc_randN(c, 2 * LPI / 10.0, 0.1);
  for (i = 0; i < LR_R; i++) e[1 + i] = orc_rand01(c);
  for (i = 0; i < LR_R; i++) e[1 + LR_R + i] = orc_rand01(c);
  y[0] = 2 * LPI * orc_rand01(c);
}

/* ---------------------------------------------------------------- kuramoto_lowrank64.pop (LET R 64, LET N 1048576) */
#define LR64_R 64
static void klr64_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                      double *aux, double *death, double *repli) {
  const double K = c->mu[0], omega = e[0], theta = y_[0], CPL = psi[0];
  const double y = sin(theta);
  const double cpl = CPL;
  (void)t; (void)mfd; (void)death; (void)repli;
  f[0] = omega + K / 1048576 * CPL;
  aux[0] = y; aux[1] = cpl;
}
static void klr64_ic(orc_ctx *c, double t, double *y, double *e) {
  int i;
  (void)t;
  e[0] = orc_randN(c, 2 * LPI / 10.0, 0.1);
  for (i = 0; i < LR64_R; i++) e[1 + i] = orc_rand01(c);
  for (i = 0; i < LR64_R; i++) e[1 + LR64_R + i] = orc_rand01(c);
  y[0] = 2 * LPI * orc_rand01(c);
}

/* ---------------------------------------------------------------- logistic_map.pop */
static const double lm_pars[] = {3.7, 0.1};
static void lm_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double eps = c->mu[1], r_i = e[0], x = y_[0], XBAR = mfd[0];
  const double local = r_i * x * (1 - x);
  (void)t; (void)psi; (void)death; (void)repli;
  f[0] = (1 - eps) * local + eps * XBAR;
  aux[0] = local;
}
static double lm_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];
}
static void lm_ic(orc_ctx *c, double t, double *y, double *e) {
  const double r = c->mu[0];
  (void)t;
  e[0] = r + 0.05 * (orc_rand01(c) - 0.5);
  y[0] = 0.1 + 0.8 * orc_rand01(c);
}

/* ---------------------------------------------------------------- reset_decay.pop */
static const double rd_pars[] = {0.5, 0.35};
static void rd_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double k = c->mu[0], xmin = c->mu[1], x = y_[0];
  const double low = orc_stopif(c, x < xmin);
  (void)t; (void)e; (void)psi; (void)mfd; (void)death; (void)repli;
  f[0] = -k * x;
  aux[0] = low;
}
static double rd_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];
}
static void rd_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t;
  e[0] = 0.5 + orc_rand01(c);
  y[0] = c->ist0 ? e[0] : 2.0 * e[0];          /* IST0(x0, 2.0*x0) */
}

/* ---------------------------------------------------------------- ring_diffusion.pop (KY list neighbours) */
static const double ring_pars[] = {0.8, 0.05, 0.3};
static void ring_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                     double *aux, double *death, double *repli) {
  const double D = c->mu[0], eps = c->mu[1], a = c->mu[2], w = e[0], x = y_[0], v = y_[1], XBAR = mfd[0];
  const double lap = orc_ky(c, 0, 1) + orc_ky(c, 0, -1) - 2 * x;
  (void)t; (void)psi; (void)death; (void)repli;
  f[0] = w * v + D * lap + eps * (orc_ky(c, 0, 2) - orc_ky(c, 0, -2));
  f[1] = -w * x - a * v + 0.1 * (XBAR - x);
  aux[0] = lap;
}
static double ring_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];
}
static void ring_ic(orc_ctx *c, double t, double *y, double *e) {
  (void)t;
  e[0] = orc_randN(c, 1.0, 0.1);
  y[0] = orc_rand01(c);
  y[1] = 0.0;
}

/* ---------------------------------------------------------------- kuramoto_lowrank_aux.pop (LET R 3, LET N 100) */
#define LRA_R 3
#define LRA_PI 3.141592
static const double klra_pars[] = {0.1};
static void klra_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                     double *aux, double *death, double *repli) {
  const double K = c->mu[0], omega = e[0], th = y_[0], CPL = psi[0];
  const double y = sin(th);
  const double theta = th - ( (int) (th/2/LRA_PI) * 2 * LRA_PI );      /* aux theta: the expansion variable of the coupling */
  (void)t; (void)mfd; (void)death; (void)repli;
  f[0] = omega + K/100*CPL;
  aux[0] = y; aux[1] = theta;
}
static double klra_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)y; (void)mfd;
  return k == 0 ? e[0] : aux[0];                       /* %M Omega ME(omega); %M MEANY MA(y) */
}
static void klra_ic(orc_ctx *c, double t, double *y, double *e) {
  int i;
  (void)t;
  e[0] = orc_randN(c, 2*LRA_PI/10.0, 0.1);
  for (i = 0; i < LRA_R; i++) e[1 + i] = orc_rand01(c);
  for (i = 0; i < LRA_R; i++) e[1 + LRA_R + i] = orc_rand01(c);
  y[0] = 2*LRA_PI*orc_rand01(c);
}

/* ---------------------------------------------------------------- vect_meanfield.pop */
static const double vm_pars[] = {0.5, 0.1};
static void vm_rhs(orc_ctx *c, double t, const double *y_, const double *e, const double *psi, const double *mfd, double *f,
                   double *aux, double *death, double *repli) {
  const double a = c->mu[0], x = y_[0], TOTAL = mfd[1], SCALED = mfd[2];
  const double w = x * TOTAL;
  (void)t; (void)e; (void)psi; (void)repli;
  f[0] = -a * x + 0.01 * SCALED;
  aux[0] = w;
  *death = c->mu[1];                                  /* %DEATH MU(d) */
}
static double vm_mfd(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd) {
  (void)c; (void)k; (void)e; (void)aux; (void)mfd;
  return y[0];                                        /* %M XBAR MY(x) */
}
static const int vm_vect[] = {0, 1, 1};
static double vm_vect_eval(int k, const double *mu, double pop_size, const double *mfd) {
  return k == 1 ? pop_size : mu[0] * mfd[0] * pop_size;     /* TOTAL = POP_SIZE; SCALED = MU(a)*MF(XBAR)*POP_SIZE */
}
static double vm_birth(orc_ctx *c, const double *y, const double *e, const double *aux) { (void)c; (void)y; (void)e; (void)aux; return 2.0; }
static void vm_ic(orc_ctx *c, double t, double *y, double *e) { (void)t; (void)e; y[0] = orc_rand01(c); }

static const orc_model MODELS[] = {
  {"lag_chain", 6, 0, 1, 0, 1, 2, lg_pars, lg_rhs, lg_mfd, NULL, NULL, lg_ic, 0, 0, 0, 0, 0, NULL},
  {"table_drive", 1, 0, 1, 0, 1, 1, td_pars, td_rhs, td_mfd, NULL, NULL, td_ic, 0, 0, 0, 0, 0, NULL},
  {"interp_drive", 1, 1, 2, 0, 1, 1, id_pars, id_rhs, id_mfd, NULL, NULL, id_ic, 0, 0, 0, 0, 0, NULL},
  {"lineage", 1, 4, 1, 0, 2, 3, ln_pars, ln_rhs, ln_mfd, ln_birth, NULL, ln_ic, 0, 0, 0, 0, 0, NULL},
  {"pair_aux", 2, 1, 3, 2, 1, 2, pa_pars, pa_rhs, pa_mfd, NULL, pa_cpl, pa_ic, 0, 0, 0, 0, 0, NULL},
  {"repel2d", 4, 0, 0, 2, 2, 3, rp_pars, rp_rhs, rp_mfd, NULL, rp_cpl, rp_ic, 0, 0, 0, 0, 0, NULL},
  {"kuramoto_lrexp", 1, 1, 1, 1, 2, 1, kle_pars, kle_rhs, kle_mfd, NULL, NULL, kle_ic, 1, 0, 0, 0, 0, sin, NULL, NULL, 0, 1},
  {"fun_block", 3, 6, 5, 0, 1, 2, fb_pars, fb_rhs, fb_mfd, NULL, NULL, fb_ic, 0, 0, 0, 0, 0, NULL},
  {"helper_calls", 4, 6, 13, 0, 1, 2, hc_pars, hc_rhs, hc_mfd, NULL, NULL, hc_ic, 0, 0, 0, 0, 0, NULL},
  {"delay_pop", 1, 1, 1, 0, 1, 3, dp_pars, dp_rhs, dp_mfd, NULL, NULL, dp_ic, 0, 0, 0, 0, 0, NULL},
  {"chain_array", 6, 4, 6, 0, 2, 2, ca_pars, ca_rhs, ca_mfd, NULL, NULL, ca_ic, 0, 0, 0, 0, 0, NULL},
  {"nbody_const", 6, 0, 0, 3, 0, 2, nbc_pars, nb_rhs, NULL, NULL, nbc_cpl, nbc_ic, 0, 0, 0, 0, 0, NULL},
  {"kuramoto_lowrank_aux", 1, 1 + 2 * LRA_R, 2, 1, 2, 1, klra_pars, klra_rhs, klra_mfd, NULL, NULL, klra_ic, LRA_R, 0, 1, 1, 1 + LRA_R, sin, NULL, NULL, 1},
  {"vect_meanfield", 1, 0, 1, 0, 3, 2, vm_pars, vm_rhs, vm_mfd, vm_birth, NULL, vm_ic, 0, 0, 0, 0, 0, NULL, vm_vect, vm_vect_eval},
  {"ring_diffusion", 2, 1, 1, 0, 1, 3, ring_pars, ring_rhs, ring_mfd, NULL, NULL, ring_ic, 0, 0, 0, 0, 0, NULL},
  {"reset_decay", 1, 1, 1, 0, 1, 2, rd_pars, rd_rhs, rd_mfd, NULL, NULL, rd_ic, 0, 0, 0, 0, 0, NULL},
  {"logistic_map", 1, 1, 1, 0, 1, 2, lm_pars, lm_rhs, lm_mfd, NULL, NULL, lm_ic, 0, 0, 0, 0, 0, NULL},
  {"kuramoto", 1, 1, 2, 0, 3, 1, kur_pars, kur_rhs, kur_mfd, NULL, NULL, kur_ic, 0, 0, 0, 0, 0, NULL},
  {"nbody", 6, 0, 0, 3, 0, 1, nb_pars, nb_rhs, NULL, NULL, nb_cpl, nb_ic, 0, 0, 0, 0, 0, NULL},
  {"vdp_sde", 2, 0, 0, 0, 0, 4, vdp_pars, vdp_rhs, NULL, vdp_birth, NULL, vdp_ic, 0, 0, 0, 0, 0, NULL},
  {"kuramoto_lowrank", 1, 1 + 2 * LR_R, 2, 1, 2, 1, klr_pars, klr_rhs, klr_mfd, NULL, NULL, klr_ic, LR_R, 0, 0, 1, 1 + LR_R, sin},
  {"kuramoto_lowrank64", 1, 1 + 2 * LR64_R, 2, 1, 2, 1, klr_pars, klr64_rhs, klr_mfd, NULL, NULL, klr64_ic, LR64_R, 0, 0, 1, 1 + LR64_R, sin},
};
/* a model registered from outside this file: tests/test_emitter_vs_reference.py compiles the PRODUCT's generated model code for the
 * host and drives it through this checker's loop, against the reference binary built from the same .pop file */
const orc_model *orc_extra_model = NULL;
const orc_model *orc_find_model(const char *name) {
  unsigned i;
  for (i = 0; i < sizeof MODELS / sizeof MODELS[0]; i++) if (strcmp(MODELS[i].name, name) == 0) return &MODELS[i];
  if (orc_extra_model != NULL && strcmp(orc_extra_model->name, name) == 0) return orc_extra_model;
  return NULL;
}
