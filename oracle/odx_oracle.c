/* TEST INFRASTRUCTURE — CPU restatement (plain C) of the reference's population simulation path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * build, link, import or execute anything under oracle/.  It is the checker, never the product.
 *
 * What is restated, and from where (paths relative to the reference tree):
 *   orc_pop_rhs        generated pop_rhs, src/odexp:801-858 (two ode_rhs passes, one list walk per %M
 *                      :874-903, N^2 %C walk :905-954, %BIRTH accumulation :860-872)
 *   orc_dde_apply      dde_apply,  src/methods_odexp.c:1825-1874 (classic RK4, h clipped to tnext, h reset to h0)
 *   orc_fe_apply       fe_apply,   src/methods_odexp.c:1790-1814
 *   adaptive steps     gsl_odeiv2_evolve_apply call site src/methods_odexp.c:619 -> oracle/minigsl.c (clean room)
 *   orc_ssa_timestep   SSA_timestep, src/methods_odexp.c:1528-1553
 *   orc_ssa_apply      SSA_apply_birthdeath + ncumsum, src/methods_odexp.c:1556-1660, 1764-1777
 *   orc_tauleap_apply  tau_leaping_apply_birthdeath, src/methods_odexp.c:1662-1762 (including the as-coded
 *                      behaviour: any event draws a third uniform; "replicate" = clone, then the mother dies)
 *   particle list      par_birth / par_repli / delete_el, src/datastruct.c:416-563
 *   random numbers     rand01 / randN / poisson / exprand over libc rand(), src/rand_gen.c:9-72
 *   orc_odesolver      odesolver, src/methods_odexp.c:83-763 (population mode, no stopping times)
 *   output records     fwrite_SIM / fwrite_all_particles, src/datastruct.c:643-674
 *
 * Pinning: rng_mode 0 consumes the serial libc stream exactly like the reference, so this restatement is
 * compared bit-for-bit / to 1e-13 with the UNCHANGED reference binary (oracle/_ref) and with the committed
 * golden vectors made from it (tests/test_oracle_port.py).  rng_mode 1 swaps only the random source for the
 * keyed Philox convention of the CUDA library (odexp_b200/csrc/odx_rng.h), which a serial stream cannot feed.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_odeiv2.h>
#include "odx_oracle.h"

/* ------------------------------------------------------------------ random numbers */
static void hr_seed(orc_hostrand *h, uint32_t s) {
  int i; int32_t word;
  if (s == 0) s = 1;
  word = (int32_t)s; h->ring[0] = word;
  for (i = 1; i < 31; i++) {
    long hi = word / 127773, lo = word % 127773;
    word = (int32_t)(16807 * lo - 2836 * hi);
    if (word < 0) word += 2147483647;
    h->ring[i] = word;
  }
  h->f = 3; h->r = 0;
  for (i = 0; i < 310; i++) {
    h->ring[h->f] = (int32_t)((uint32_t)h->ring[h->f] + (uint32_t)h->ring[h->r]);
    if (++h->f >= 31) h->f = 0;
    if (++h->r >= 31) h->r = 0;
  }
}
static int hr_next(orc_hostrand *h) {
  uint32_t v = (uint32_t)h->ring[h->f] + (uint32_t)h->ring[h->r];
  h->ring[h->f] = (int32_t)v;
  if (++h->f >= 31) h->f = 0;
  if (++h->r >= 31) h->r = 0;
  h->ndraws++;
  return (int)(v >> 1);
}
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  int r;
  for (r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
/* known-answer hook for tests/: the block function alone (Random123 philox4x32-10 vectors) */
void orc_philox_block(uint32_t c[4], uint32_t k0, uint32_t k1) { philox(c, k0, k1); }
/* draws through the keyed convention with an explicit key / counter prefix (tests/: distribution checks of orc_poisson) */
long orc_poisson_keyed(uint32_t k0, uint32_t k1, uint32_t id, uint32_t stream, double lambda);
static double u01(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
static void keyed_draw(orc_ctx *c, uint32_t out[4]) {
  out[0] = c->kid; out[1] = c->kslot++; out[2] = c->kstream; out[3] = c->kdom;
  philox(out, c->sim->k0, c->sim->k1);
}
double orc_rand01(orc_ctx *c) {
  if (c->sim->rng_mode == 0) return (double)hr_next(&c->sim->hr) / 2147483647;       /* rand()/RAND_MAX */
  { uint32_t r[4]; keyed_draw(c, r); return u01(r[0], r[1]); }
}
double orc_randN(orc_ctx *c, double mu, double sigma) {
  if (c->sim->rng_mode == 0) {                                                       /* ratio of uniforms */
    double u, v, x, y, q;
    do {
      u = orc_rand01(c);
      v = 1.7156 * (orc_rand01(c) - 0.5);
      x = u - 0.449871;
      y = fabs(v) + 0.386595;
      q = x * x + y * (0.19600 * y - 0.25472 * x);
    } while (q > 0.27597 && (q > 0.27846 || v * v > -4.0 * log(u) * u * u));
    return mu + sigma * v / u;
  }
  { uint32_t r[4]; double u1, u2; keyed_draw(c, r); u1 = u01(r[0], r[1]); u2 = u01(r[2], r[3]);
    return mu + sigma * (sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2)); }
}
double orc_exprand(orc_ctx *c, double r) { double mu = orc_rand01(c); return log(1 / (1 - mu)) / r; }
long orc_poisson(orc_ctx *c, double lambda) {
  long k = 0; double L = exp(-lambda), p = 1;
  if (c->sim->rng_mode != 0 && lambda >= 30.0) {      /* keyed convention of odx_rng.h: Hoermann's PTRS for large means */
    const double slam = sqrt(lambda), loglam = log(lambda);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
      const double U = orc_rand01(c) - 0.5, V = orc_rand01(c);
      const double us = 0.5 - fabs(U);
      const double kf = floor((2.0 * a / us + b) * U + lambda + 0.43);
      if (us >= 0.07 && V <= vr) return (long)kf;
      if (kf < 0.0 || (us < 0.013 && V > us)) continue;
      if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lambda + kf * loglam - lgamma(kf + 1.0)) return (long)kf;
    }
  }
  do {
    ++k;
    p *= (c->sim->rng_mode == 0) ? hr_next(&c->sim->hr) / (double)2147483647 : orc_rand01(c);
  } while (p > L);
  return --k;
}
long orc_poisson_keyed(uint32_t k0, uint32_t k1, uint32_t id, uint32_t stream, double lambda) {
  orc_sim s; orc_ctx c;
  memset(&s, 0, sizeof s); memset(&c, 0, sizeof c);
  s.rng_mode = 1; s.k0 = k0; s.k1 = k1; c.sim = &s; c.kid = id; c.kstream = stream; c.kdom = ORC_DOM_HOST; c.kslot = 0;
  return orc_poisson(&c, lambda);
}
double orc_dwdt(orc_ctx *c) { return orc_randN(c, 0, 1) / sqrt(c->h); }
int orc_stopif(orc_ctx *c, int cond) { if (cond) c->sim->stop_flag = ORC_STOP_COND; return cond; }

static void host_ctx(orc_sim *s, orc_ctx *c, uint32_t kid) {     /* draws made by the solver loop itself */
  memset(c, 0, sizeof *c);
  c->sim = s; c->kid = kid; c->kstream = s->host_draws++; c->kdom = ORC_DOM_HOST; c->kslot = 0;
}

/* ------------------------------------------------------------------ particle list */
static orc_par *par_new(orc_sim *s) {
  const orc_model *m = s->model;
  orc_par *p = calloc(1, sizeof *p);
  p->y = calloc(m->nvar > 0 ? m->nvar : 1, sizeof(double));
  p->expr = calloc(m->nexpr > 0 ? m->nexpr : 1, sizeof(double));
  p->aux = calloc(m->naux > 0 ? m->naux : 1, sizeof(double));
  p->psi = calloc(m->npsi > 0 ? m->npsi : 1, sizeof(double));
  p->id = s->max_id++;
  p->sister = NULL;
  return p;
}
static void par_free(orc_par *p) { free(p->y); free(p->expr); free(p->aux); free(p->psi); free(p); }
static void list_append(orc_sim *s, orc_par *p) {
  if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 1024; s->p = realloc(s->p, sizeof(orc_par *) * s->cap); }
  s->p[s->n++] = p;
  s->size++;
}
static orc_par *par_birth(orc_sim *s) { orc_par *p = par_new(s); list_append(s, p); return p; }
static orc_par *par_repli(orc_sim *s, orc_par *mother) {
  const orc_model *m = s->model; orc_par *p = par_new(s);
  memcpy(p->y, mother->y, sizeof(double) * m->nvar); memcpy(p->expr, mother->expr, sizeof(double) * m->nexpr);
  memcpy(p->aux, mother->aux, sizeof(double) * m->naux); memcpy(p->psi, mother->psi, sizeof(double) * m->npsi);
  p->death = mother->death; p->repli = mother->repli; p->sister = mother;
  list_append(s, p);
  return p;
}
static void par_delete(orc_sim *s, orc_par *p) { p->dead = 1; s->size--; }     /* unlinked now, storage reclaimed by compact() */
static void compact(orc_sim *s) {
  int i, k = 0;
  for (i = 0; i < s->n; i++) { if (s->p[i]->dead) par_free(s->p[i]); else s->p[k++] = s->p[i]; }
  s->n = k;
}

/* ------------------------------------------------------------------ one RHS evaluation of the population */
static void rhs_ctx(orc_sim *s, orc_ctx *c, const orc_par *p) {
  memset(c, 0, sizeof *c);
  c->sim = s; c->mu = s->mu; c->pop_size = (double)s->size; c->h = s->h ? *s->h : 0.0; c->id = p->id; c->oid = p->id;
  c->kid = (uint32_t)p->id; c->kstream = s->eval_index; c->kdom = ORC_DOM_RHS; c->kslot = 0;
}
/* KY(name, degree) (src/odexp.h:334-357): walk |degree| links forwards (nextel, NULL -> list start) or backwards (prevel,
 * NULL -> list end) from the particle and read that neighbour's y as stored in the list: in pass 1 of pop_rhs the neighbours
 * before the particle already hold the state of this evaluation and those after it the state of the previous evaluation; in
 * pass 2 (the values that reach f) all hold the current state. */
double orc_ky(orc_ctx *c, int var, int degree) {
  orc_sim *s = c->sim;
  int j = (c->pos + degree) % s->n;
  if (j < 0) j += s->n;
  return s->p[j]->y[var];
}
/* interp and delay, src/utils_odexp.c:173-240, index arithmetic as coded: after the fill at t0 (`EVENT_TYPE && EVENT_T0`, a logical
 * AND: true whenever any event flag is set) pos is DELAY_SSTACK, so the search starts one element late and the oldest sample
 * sits at the logical end; a first evaluation of dde_apply pushes (t, x) -- once per ode_rhs call, i.e. twice per pop_rhs */
double orc_interp(double *x, double *y, double xi, int nx, int start) {
  int i1 = 0, i9 = nx - 1, i5;
  while (i9 > i1 + 1) {
    i5 = (i1 + i9 + 1) / 2;
    if (x[(start + i5) % nx] > xi) i9 = i5; else i1 = i5;
  }
  if (i1 == i9) return y[(i1 + start) % nx];
  return y[(i1 + start) % nx] + (y[(i9 + start) % nx] - y[(i1 + start) % nx])
         * (xi - x[(i1 + start) % nx]) / (x[(i9 + start) % nx] - x[(i1 + start) % nx]);
}
double orc_delay(orc_ctx *c, const double t, const double x, const double tau, double (*ic)(double), orc_stack *xhist) {
  double ti; const double h = *(c->sim->h);
  if (c->sim->stop_flag && ORC_EVENT_T0) {
    xhist->pos = 0;
    while (xhist->pos < ORC_DELAY_SSTACK) {
      ti = t - (ORC_DELAY_SSTACK - xhist->pos) * h;
      xhist->t[xhist->pos] = ti;
      xhist->array[xhist->pos] = ic(ti);
      xhist->pos++;
    }
  } else if (c->sim->first_eval_flag) {
    xhist->pos++;
    xhist->pos %= ORC_DELAY_SSTACK;
    xhist->t[xhist->pos] = t;
    xhist->array[xhist->pos] = x;
  }
  return orc_interp(xhist->t, xhist->array, t - tau, ORC_DELAY_SSTACK, xhist->pos + 1);
}

int orc_pop_rhs(double t, const double *y_, double *f_, void *params) {
  orc_sim *s = params; const orc_model *m = s->model;
  int i, j, k; orc_ctx c;
  clock_t c0 = clock();
  s->pop_birth_rate = 0.0;
  for (k = 0; k < m->nmfd; k++) s->mfd[k] = 0.0;
  /* pass 1: aux / rates with zeroed mean fields and the psi of the previous call; copy the state into the particles */
  for (i = 0; i < s->n; i++) {
    orc_par *p = s->p[i];
    rhs_ctx(s, &c, p); c.pos = i;
    m->rhs(&c, t, y_ + (size_t)i * m->nvar, p->expr, p->psi, s->mfd, f_ + (size_t)i * m->nvar, p->aux, &p->death, &p->repli);
    memcpy(p->y, y_ + (size_t)i * m->nvar, sizeof(double) * m->nvar);
    if (m->birth_term) s->pop_birth_rate += m->birth_term(&c, p->y, p->expr, p->aux);
  }
  if (m->lr_rank > 0) orc_lowrank_apply(s);      /* low-rank coupling between pass 1 and the mean fields (src/odexp:841) */
  /* one walk per mean field, summed in list order, then divided by POP_SIZE */
  for (k = 0; k < m->nmfd; k++) {
    if (m->mfd_vect && m->mfd_vect[k]) { s->mfd[k] = m->mfd_vect_eval(k, s->mu, (double)s->size, s->mfd); continue; }      /* {vect} */
    for (i = 0; i < s->n; i++) { orc_par *p = s->p[i]; rhs_ctx(s, &c, p); s->mfd[k] += m->mfd_term(&c, k, p->y, p->expr, p->aux, s->mfd); }
    s->mfd[k] /= s->size;
  }
  s->pop_birth_rate /= s->size;
  /* pass 2: psi_i = (1/POP_SIZE) sum_j expr(i,j), then the final ode_rhs */
  for (i = 0; i < s->n; i++) {
    orc_par *p = s->p[i];
    if (m->npsi > 0 && m->cpl_terms) {
      for (k = 0; k < m->npsi; k++) p->psi[k] = 0.0;
      for (j = 0; j < s->n; j++) {
        orc_par *o = s->p[j];
        rhs_ctx(s, &c, p); c.oid = o->id; c.my_aux = p->aux; c.other_aux = o->aux;
        m->cpl_terms(&c, p->y, p->expr, o->y, o->expr, p->psi);
      }
      for (k = 0; k < m->npsi; k++) p->psi[k] /= s->size;
    }
    rhs_ctx(s, &c, p); c.pos = i;
    m->rhs(&c, t, y_ + (size_t)i * m->nvar, p->expr, p->psi, s->mfd, f_ + (size_t)i * m->nvar, p->aux, &p->death, &p->repli);
  }
  s->eval_index += 2;
  s->rhs_evals++;
  s->function_ms += (clock() - c0) * 1000.0 / CLOCKS_PER_SEC;
  return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ reference-owned steppers */
static int orc_fe_apply(orc_sim *s, double *t, double tnext, double *h, double *y, int dim) {
  double *f = malloc(sizeof(double) * dim), dt = tnext - *t; int i;
  if (*h > (dt - s->o.hmin)) *h = dt;
  orc_pop_rhs(*t, y, f, s);
  for (i = 0; i < dim; i++) y[i] = y[i] + (*h) * f[i];
  *t += *h;
  *h = s->o.h0;
  free(f);
  return GSL_SUCCESS;
}
static int orc_dde_apply(orc_sim *s, double *t, double tnext, double *h, double *y, int dim) {
  double *k = malloc(sizeof(double) * dim * 5), *k1 = k + dim, *k2 = k1 + dim, *k3 = k2 + dim, *k4 = k3 + dim, dt = tnext - *t;
  int i;
  if (*h > (dt - s->o.hmin)) *h = dt;
  s->first_eval_flag = 1;                 /* DELAY_FIRST_EVAL, src/methods_odexp.c:1842-1844 */
  orc_pop_rhs(*t, y, k1, s);
  s->first_eval_flag = 0;
  for (i = 0; i < dim; i++) k[i] = y[i] + (*h) / 2.0 * k1[i];
  orc_pop_rhs(*t + *h / 2.0, k, k2, s);
  for (i = 0; i < dim; i++) k[i] = y[i] + (*h) / 2.0 * k2[i];
  orc_pop_rhs(*t + *h / 2.0, k, k3, s);
  for (i = 0; i < dim; i++) k[i] = y[i] + (*h) * k3[i];
  orc_pop_rhs(*t + *h, k, k4, s);
  for (i = 0; i < dim; i++) y[i] += (*h) / 6.0 * (k1[i] + 2.0 * (k2[i] + k3[i]) + k4[i]);
  *t += *h;
  *h = s->o.h0;
  free(k);
  return GSL_SUCCESS;
}

/* iteration_apply, src/methods_odexp.c:1816-1823: y = f(y) written in place (the RHS reads y_ and writes f_ == y_) */
static int orc_iteration_apply(orc_sim *s, double *t, double *y) {
  orc_pop_rhs(*t, y, y, s);
  *t += 1;
  return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ initial conditions */
static void run_ic(orc_sim *s, orc_par *p, double t, uint32_t dom, uint32_t stream) {
  orc_ctx c;
  memset(&c, 0, sizeof c);
  c.sim = s; c.mu = s->mu; c.pop_size = (double)s->size; c.id = p->id; c.oid = p->id;
  c.atbirth = (s->event[0] == -1); c.atrepli = (s->event[0] >= 0 && s->event[1] == 1);
  c.ist0 = (s->stop_flag == ORC_EVENT_T0);
  c.sister = p->sister;
  c.kid = (uint32_t)p->id; c.kstream = stream; c.kdom = dom; c.kslot = 0;
  s->model->ic(&c, t, p->y, p->expr);
}

/* ------------------------------------------------------------------ outputs */
static void write_stats(orc_sim *s, double t) {
  fwrite(&t, sizeof(double), 1, s->fstats);
  fwrite(s->mfd, sizeof(double), s->model->nmfd, s->fstats);
  fwrite(&s->size, sizeof(int), 1, s->fstats);
  fwrite(&s->stop_flag, sizeof(int), 1, s->fstats);
  fwrite(s->event, sizeof(int), 3, s->fstats);
}
/* fwrite_quick (src/methods_odexp.c:1353-1436): one (x, y, z) row of current.plot; dynamical variables are read at
 * y[cp * tx + ng] with tx = nvar + naux + npsi as coded there (NaN where that lies past the end of y) */
static void write_quick(orc_sim *s, double t, const double *y) {
  const orc_model *m = s->model; const int tx = m->nvar + m->naux + m->npsi, ng[3] = {s->ngx, s->ngy, s->ngz};
  const orc_par *pars = NULL; double row[3]; int i, k;
  if (!s->fquick) return;
  for (i = 0; i < s->n; i++) if (!s->p[i]->dead && s->p[i]->id == s->cp) pars = s->p[i];          /* getpar(cp) */
  for (k = 0; k < 3; k++) {
    const int g = ng[k];
    if (g == -1) row[k] = t;
    else if (g < m->nvar && pars != NULL) row[k] = ((long)s->cp * tx + g < (long)s->n * m->nvar) ? y[s->cp * tx + g] : NAN;
    else if (g < m->nvar + m->naux && pars != NULL) row[k] = pars->aux[g - m->nvar];
    else if (g < tx && pars != NULL) row[k] = pars->psi[g - m->nvar - m->naux];
    else if (g >= tx) row[k] = (g - tx < m->nmfd) ? s->mfd[g - tx] : NAN;
    else row[k] = NAN;
  }
  fwrite(row, sizeof(double), 3, s->fquick);
}
static void write_traj(orc_sim *s, double t) {
  const orc_model *m = s->model; int i;
  if (s->ftraj) {
    for (i = 0; i < s->n; i++) {
      orc_par *p = s->p[i];
      if (p->dead) continue;
      fwrite(&s->nbrsteps, sizeof(unsigned int), 1, s->ftraj);
      fwrite(&p->id, sizeof(int), 1, s->ftraj);
      fwrite(&t, sizeof(double), 1, s->ftraj);
      fwrite(p->y, sizeof(double), m->nvar, s->ftraj);
      fwrite(p->aux, sizeof(double), m->naux, s->ftraj);
      fwrite(p->psi, sizeof(double), m->npsi, s->ftraj);
      fwrite(s->mfd, sizeof(double), m->nmfd, s->ftraj);
      fwrite(p->expr, sizeof(double), m->nexpr, s->ftraj);
    }
  }
  s->nbrsteps++;
}

/* ------------------------------------------------------------------ birth / death */
static double orc_ssa_timestep(orc_sim *s, double *sumr) {
  double r = s->pop_birth_rate; int i; orc_ctx c;
  for (i = 0; i < s->n; i++) { r += s->p[i]->death; r += s->p[i]->repli; }
  host_ctx(s, &c, 0xFFFFFFFFu);
  *sumr = r;
  return orc_exprand(&c, r);
}
static void orc_ssa_apply(orc_sim *s, double t) {
  const int N = s->size; int i, idx = 0; double *r = malloc(sizeof(double) * (2 * N + 1)), sumr, u; orc_ctx c;
  for (i = 0; i < N; i++) { r[i] = s->p[i]->death; r[i + N] = s->p[i]->repli; }
  r[2 * N] = s->pop_birth_rate;
  for (i = 1; i < 2 * N + 1; i++) r[i] += r[i - 1];               /* ncumsum */
  sumr = r[2 * N];
  for (i = 0; i < 2 * N + 1; i++) r[i] /= r[2 * N];                /* the last element becomes 1 once reached */
  if (sumr <= 0.0) { free(r); return; }
  host_ctx(s, &c, 0xFFFFFFFFu);
  u = orc_rand01(&c);
  while (u > r[idx]) idx++;
  if (idx < N) {                                                    /* death */
    orc_par *p = s->p[idx];
    s->stop_flag = ORC_STOP_DEATH; s->event[0] = p->id; s->event[1] = -1; s->event[2] = -1;
    par_delete(s, p);
  } else if (idx < 2 * N) {                                         /* replication */
    orc_par *p = s->p[idx - N], *d;
    s->stop_flag = ORC_STOP_REPLI; s->event[0] = p->id; s->event[1] = 1;
    d = par_repli(s, p);
    s->event[2] = d->id;
    run_ic(s, p, t, ORC_DOM_IC + 1, s->event_index);
    run_ic(s, d, t, ORC_DOM_IC + 2, s->event_index);
    d->sister = NULL;
  } else {                                                          /* birth */
    orc_par *d;
    s->stop_flag = ORC_STOP_BIRTH;
    d = par_birth(s);
    s->event[0] = -1; s->event[1] = 1; s->event[2] = d->id;
    run_ic(s, d, t, ORC_DOM_IC + 3, s->event_index);
  }
  s->event_index++;
  write_stats(s, t);
  compact(s);
  free(r);
}
static void orc_tauleap_apply(orc_sim *s, double t, double dt) {
  int i; long nnew; orc_ctx c;
  for (i = 0; i < s->n; i++) {                    /* daughters appended during the sweep are visited too */
    orc_par *p = s->p[i]; double u1, u2;
    memset(&c, 0, sizeof c);
    c.sim = s; c.kid = (uint32_t)p->id; c.kstream = s->event_index; c.kdom = ORC_DOM_LEAP; c.kslot = 0;
    s->stop_flag = ORC_EVENT_NONE;
    u1 = orc_rand01(&c);
    if (u1 < (p->death * dt)) s->stop_flag = ORC_STOP_DEATH;
    u2 = orc_rand01(&c);
    if (u2 < (p->repli * dt)) s->stop_flag |= ORC_STOP_REPLI;
    s->event[0] = p->id;
    if (s->stop_flag & (ORC_STOP_DEATH | ORC_STOP_REPLI)) {
      /* as coded in the reference this branch is taken whenever EITHER flag is set (methods_odexp.c:1685) */
      if (orc_rand01(&c) < p->death / (p->death + p->repli)) {
        s->stop_flag = ORC_STOP_DEATH; s->event[1] = -1; s->event[2] = -1;
        par_delete(s, p);
        write_stats(s, t);
      } else {
        orc_par *d;
        s->stop_flag = ORC_STOP_REPLI;
        d = par_repli(s, p);
        s->event[1] = 1; s->event[2] = d->id;
        write_stats(s, t);
        run_ic(s, p, t, ORC_DOM_IC + 1, s->event_index);
        run_ic(s, d, t, ORC_DOM_IC + 2, s->event_index);
        d->sister = NULL;
        s->stop_flag = ORC_STOP_DEATH; s->event[1] = -1; s->event[2] = -1;
        par_delete(s, p);
        write_stats(s, t);
      }
    }
  }
  host_ctx(s, &c, 0xFFFFFFFEu);
  nnew = orc_poisson(&c, s->pop_birth_rate * dt);
  for (i = 0; i < (int)nnew; i++) {
    orc_par *d;
    s->stop_flag = ORC_STOP_BIRTH;
    d = par_birth(s);
    s->event[0] = -1; s->event[1] = 1; s->event[2] = d->id;
    write_stats(s, t);
    run_ic(s, d, t, ORC_DOM_IC + 3, s->event_index);
  }
  s->event_index++;
  compact(s);
}

/* ------------------------------------------------------------------ odesolver */
static void fill_y(orc_sim *s, double *y) {
  int i; const int nv = s->model->nvar;
  for (i = 0; i < s->n; i++) memcpy(y + (size_t)i * nv, s->p[i]->y, sizeof(double) * nv);
}
static void update_from_y(orc_sim *s, const double *y) {
  int i; const int nv = s->model->nvar;
  for (i = 0; i < s->n; i++) memcpy(s->p[i]->y, y + (size_t)i * nv, sizeof(double) * nv);
}

int orc_odesolver(const char *model_name, const orc_opts *o, orc_report *rep) {
  const orc_model *m = orc_find_model(model_name);
  orc_sim S, *s = &S;
  double t, t1, dt, tnext, dt_ssa, dt_dyn, dt_next, sum_rates, expected, h, *y = NULL, *f = NULL, tot_solver = 0.0;
  const gsl_odeiv2_step_type *odeT = NULL; gsl_odeiv2_step *gs = NULL; gsl_odeiv2_control *gc = NULL; gsl_odeiv2_evolve *ge = NULL;
  int status = GSL_SUCCESS, i, dim, tau = 0, nstops = 0, idx_stop = 0;
  double *tstops = NULL, nextstop = INFINITY;
  char path[4096];
  clock_t start = clock(), c0;
  if (!m) return GSL_EINVAL;
  if (o->tspan_len < 2) return GSL_EINVAL;
  memset(s, 0, sizeof *s);
  s->model = m; s->o = *o;
  if (o->method == ORC_METH_ITERATION) {         /* src/methods_odexp.c:187-193 */
    s->o.h0 = 1.0; s->o.res = (int)(o->tspan[o->tspan_len - 1] - o->tspan[0]) + 1;
  }
  o = &s->o;
  if (o->res < 2) return GSL_EINVAL;
  s->rng_mode = o->rng_mode; s->k0 = o->seed; s->k1 = 0;
  s->mu = malloc(sizeof(double) * (m->npar > 0 ? m->npar : 1));
  for (i = 0; i < m->npar; i++) s->mu[i] = o->mu ? o->mu[i] : m->par_defaults[i];
  s->mfd = calloc(m->nmfd > 0 ? m->nmfd : 1, sizeof(double));
  hr_seed(&s->hr, o->seed);                                         /* srand(seed), main.c:403 */
  h = o->h0; s->h = &h;
  if (o->method == ORC_METH_RKF45) odeT = gsl_odeiv2_step_rkf45;
  else if (o->method == ORC_METH_RKCK) odeT = gsl_odeiv2_step_rkck;
  else if (o->method == ORC_METH_RK4) odeT = gsl_odeiv2_step_rk4;
  else if (o->method == ORC_METH_RK2) odeT = gsl_odeiv2_step_rk2;
  else if (o->method == ORC_METH_RK8PD) odeT = gsl_odeiv2_step_rk8pd;
  t = o->tspan[0]; t1 = o->tspan[o->tspan_len - 1];
  /* stopping times = interior TIMESPAN values, sorted (src/methods_odexp.c:222-236); past the last one the reference reads
   * beyond its array, restated as "no further stop" */
  if (o->tspan_len > 2) {
    nstops = o->tspan_len - 2;
    tstops = malloc(sizeof(double) * nstops);
    for (i = 0; i < nstops; i++) tstops[i] = o->tspan[i + 1];
    for (i = 1; i < nstops; i++) { double v = tstops[i]; int j2 = i - 1; while (j2 >= 0 && tstops[j2] > v) { tstops[j2 + 1] = tstops[j2]; j2--; } tstops[j2 + 1] = v; }
    nextstop = tstops[idx_stop++];
  }
  s->stop_flag = ORC_EVENT_T0;
  dt = (t1 - t) / (double)(o->res - 1);
  snprintf(path, sizeof path, "%s/stats.dat", o->outdir);
  if (!(s->fstats = fopen(path, "w"))) return GSL_EFAILED;
  snprintf(path, sizeof path, "%s/current.plot", o->outdir);
  if (!(s->fquick = fopen(path, "w"))) return GSL_EFAILED;
  s->ngx = o->plot_x; s->ngy = o->plot_y; s->ngz = o->plot_z; s->cp = o->particle;
  if (o->write_traj) { snprintf(path, sizeof path, "%s/traj.dat", o->outdir); if (!(s->ftraj = fopen(path, "w"))) return GSL_EFAILED; }

  for (i = 0; i < o->popsize; i++) par_birth(s);
  s->event[0] = -1; s->event[1] = 1;
  dim = m->nvar * s->n;
  y = malloc(sizeof(double) * (dim > 0 ? dim : 1)); f = malloc(sizeof(double) * (dim > 0 ? dim : 1));
  for (i = 0; i < s->n; i++) run_ic(s, s->p[i], t, ORC_DOM_IC, 0);   /* pop_init_conditions, list order */
  fill_y(s, y);
  orc_pop_rhs(t, y, f, s);
  for (i = 0; i < o->popsize; i++) { s->event[2] = i; write_stats(s, t); }
  memset(s->event, 0, sizeof s->event);
  s->stop_flag = ORC_EVENT_NONE;
  write_quick(s, t, y);
  write_traj(s, t);
  if (odeT) { gs = gsl_odeiv2_step_alloc(odeT, dim); gc = gsl_odeiv2_control_y_new(o->abstol, o->reltol); ge = gsl_odeiv2_evolve_alloc(dim); }

  while (t < t1 && s->size > 0) {
    dt_dyn = fmin(dt, t1 - t);
    if ((t <= nextstop) && (dt_dyn >= (nextstop - t))) { dt_dyn = nextstop - t; s->stop_flag = ORC_STOP_TSPAN; }
    dt_ssa = orc_ssa_timestep(s, &sum_rates);
    expected = 1.0 / sum_rates;
    tau = !(expected > (dt_dyn * o->ssahmin));
    if (!tau) {
      if (dt_ssa < dt_dyn) { dt_next = dt_ssa; s->stop_flag = ORC_POP_MASK; } else dt_next = dt_dyn;
    } else {
      s->stop_flag = ORC_POP_MASK;
      dt_next = fmin(fmax(1, o->aleap * s->size) * expected, dt_dyn);
    }
    tnext = t + dt_next;
    c0 = clock();
    {
      gsl_odeiv2_system sys = {orc_pop_rhs, NULL, (size_t)dim, s};
      while (t < tnext) {
        const int nstep = s->size;
        if (o->method == ORC_METH_ITERATION) status = orc_iteration_apply(s, &t, y);
        else if (o->method == ORC_METH_FE) status = orc_fe_apply(s, &t, tnext, &h, y, dim);
        else if (o->method == ORC_METH_DDE) status = orc_dde_apply(s, &t, tnext, &h, y, dim);
        else status = gsl_odeiv2_evolve_apply(ge, gc, gs, &sys, &t, tnext, &h, y);
        s->steps++; s->particle_steps += nstep;
        if (s->stop_flag & ORC_STOP_COND) break;
        if (status != GSL_SUCCESS) { s->stop_flag = ORC_EVENT_GSL; break; }
        if (h < o->hmin) h = o->hmin;
      }
    }
    if (t >= t1) s->stop_flag = ORC_EVENT_TFINAL;
    update_from_y(s, y);
    write_stats(s, t);
    if (status != GSL_SUCCESS && status != GSL_EBADFUNC) break;
    tot_solver += (clock() - c0) * 1000.0 / CLOCKS_PER_SEC;
    write_quick(s, t, y);
    write_traj(s, t);
    if (s->stop_flag & ORC_POP_MASK) {
      if (!tau) orc_ssa_apply(s, t); else orc_tauleap_apply(s, t, dt_next);
      dim = m->nvar * s->n;
      y = realloc(y, sizeof(double) * (dim > 0 ? dim : 1)); f = realloc(f, sizeof(double) * (dim > 0 ? dim : 1));
      fill_y(s, y);
      orc_pop_rhs(t, y, f, s);
      write_quick(s, t, y);
      write_traj(s, t);
      memset(s->event, 0, sizeof s->event);
      if (odeT) { gsl_odeiv2_evolve_free(ge); gsl_odeiv2_step_free(gs); gs = gsl_odeiv2_step_alloc(odeT, dim); ge = gsl_odeiv2_evolve_alloc(dim); }
    }
    if (s->stop_flag & (ORC_STOP_COND | ORC_STOP_TSPAN)) {
      /* reset: pop_init_conditions for every particle, then a full RHS evaluation and a trajectory frame (:704-714) */
      for (i = 0; i < s->n; i++) run_ic(s, s->p[i], t, ORC_DOM_IC + 4, s->reinit_index);
      s->reinit_index++;
      fill_y(s, y);
      orc_pop_rhs(t, y, f, s);
      write_quick(s, t, y);
      write_traj(s, t);
    }
    if (s->stop_flag & ORC_STOP_TSPAN) nextstop = (idx_stop < nstops) ? tstops[idx_stop++] : INFINITY;
    s->stop_flag = ORC_EVENT_NONE;
  }
  free(tstops);
  if (rep) {
    rep->total_ms = (clock() - start) * 1000.0 / CLOCKS_PER_SEC; rep->solver_ms = tot_solver; rep->function_ms = s->function_ms;
    rep->steps = s->steps; rep->particle_steps = s->particle_steps; rep->final_n = s->size; rep->status = status;
    rep->rand_draws = s->hr.ndraws; rep->rhs_evals = s->rhs_evals;
  }
  fclose(s->fstats); if (s->ftraj) fclose(s->ftraj); if (s->fquick) fclose(s->fquick);
  if (odeT) { gsl_odeiv2_evolve_free(ge); gsl_odeiv2_control_free(gc); gsl_odeiv2_step_free(gs); }
  for (i = 0; i < s->n; i++) par_free(s->p[i]);
  free(s->p); free(y); free(f); free(s->mu); free(s->mfd);
  return status;
}
