/* odexp_gpu_branch.c — the B200 branch of the reference's odesolver().
 *
 * Compiled WITH the reference's own headers (-I <reference>/src) and linked into its model.out next to the unchanged
 * sources; integration/apply_gpu_branch.py adds the one call to it inside odesolver() (src/methods_odexp.c:83, after the
 * step-method name has been parsed and before the output files are opened).  Everything the UI reads afterwards is left
 * exactly as the CPU path leaves it:
 *   .odexp/stats.dat, traj.dat, current.plot   written by odx_odesolver with the reference's record layouts
 *                                              (src/datastruct.c:643-674, src/methods_odexp.c:1353-1436)
 *   SIM->pop                                   rebuilt from the device population with the reference's own par_birth() /
 *                                              delete_el() (getpar, generate_particle_file, `opt init 1` keep working)
 *   SIM->max_id, SIM->meanfield, SIM->nbrsteps, SIM->event, EVENT_TYPE
 * The branch is taken in `popmode population` when the environment names a device model (ODEXP_CUDA_MODEL =
 * libodx_<model>.so, built from the same .pop file by odexp_b200.build) and the step method is one the library has
 * (everything but the implicit GSL methods); otherwise it returns 0 and the CPU path runs unchanged.  Once the branch is
 * taken nothing falls back: a library that does not load, was built from another model, or finds no usable GPU ends the
 * run with an error.
 *   ODEXP_CUDA_DEVICE      CUDA device ordinal (default 0)
 *   ODEXP_CUDA_TRAJ_STRIDE write every k-th particle to traj.dat (default 1 = reference behaviour; 0 = no traj.dat)
 *   ODEXP_CUDA_MIRROR      0 = do not rebuild SIM->pop (populations too large for a host list)
 *   ODEXP_CUDA_DUMP_POP    path: after the run, walk SIM->pop and write id, y, expr, aux, psi, rates per particle (tests)
 */
#include <dlfcn.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* the reference's headers carry no include guards: take them exactly the way src/methods_odexp.c does */
#include <gsl/gsl_errno.h>
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_multiroots.h>
#include <gsl/gsl_vector.h>
#include "methods_odexp.h"      /* enum solver, odeic, double_array, SIM, get_int / get_dou / get_str, par_birth, delete_el */
#include "odx.h"

typedef struct {
  void *so;
  const odx_model_info *(*model)(void);
  int  (*create)(odx_handle **, int, int64_t);
  void (*destroy)(odx_handle *);
  const char *(*last_error)(const odx_handle *);
  int  (*odesolver)(odx_handle *, const odx_solver_opts *, odx_solver_report *);
  int  (*set_params)(odx_handle *, const double *, int);
  int  (*pop_init)(odx_handle *, int64_t, double, uint32_t, int);
  int  (*pop_upload)(odx_handle *, int64_t, const double *, const double *, const int32_t *, int32_t);
  int  (*pop_download)(odx_handle *, int64_t, int64_t, double *, double *, double *, double *, int32_t *, double *, double *);
  int  (*pop_size)(odx_handle *, int64_t *, int64_t *, int32_t *);
  int  (*meanfields)(odx_handle *, double *);
  void (*request_abort)(odx_handle *);
} odx_api;

static odx_api api;
static int api_bound = 0;
static odx_handle *volatile running_handle = NULL;

#define ODX_BIND(field, name) do { *(void **)(&api.field) = dlsym(api.so, name); if (!api.field) return -1; } while (0)
static int bind_api(const char *path) {
  if (api_bound) return 0;
  if (!(api.so = dlopen(path, RTLD_NOW | RTLD_LOCAL))) { PRINTERR("  CUDA model %s: %s\n", path, dlerror()); return -1; }
  ODX_BIND(model, "odx_model"); ODX_BIND(create, "odx_create"); ODX_BIND(destroy, "odx_destroy");
  ODX_BIND(last_error, "odx_last_error"); ODX_BIND(odesolver, "odx_odesolver"); ODX_BIND(set_params, "odx_set_params");
  ODX_BIND(pop_init, "odx_pop_init"); ODX_BIND(pop_upload, "odx_pop_upload"); ODX_BIND(pop_download, "odx_pop_download");
  ODX_BIND(pop_size, "odx_pop_size"); ODX_BIND(meanfields, "odx_meanfields"); ODX_BIND(request_abort, "odx_request_abort");
  api_bound = 1;
  return 0;
}

/* Ctrl-C: the reference's handler sets abort_odesolver_flag and EVENT_TYPE = EVENT_ABORT (src/methods_odexp.c:23-30); here the
 * request goes to the library, whose loop polls it like `while (t < t1 && !abort_odesolver_flag && POP_SIZE > 0)` (:451) */
static void gpu_abort_handler(int sig) {
  (void)sig;
  if (running_handle) api.request_abort(running_handle);
  EVENT_TYPE = EVENT_ABORT;
}

static int method_of(int solver) {
  switch ((enum solver)solver) {
    case H_FE: return ODX_METH_FE;
    case H_DDE: return ODX_METH_DDE;
    case H_ITERATION: return ODX_METH_ITERATION;
    case GSL_RKF45: return ODX_METH_RKF45;
    case GSL_RKCK: return ODX_METH_RKCK;
    case GSL_RK8PD: return ODX_METH_RK8PD;
    case GSL_RK4: return ODX_METH_RK4;
    case GSL_RK2: return ODX_METH_RK2;
    default: return -1;           /* implicit GSL methods (bsimp, ...): CPU path */
  }
}

/* SIM->pop := the device population (the reference's own list primitives, src/datastruct.c:416-563) */
static int mirror_population(odx_handle *h) {
  int64_t n = 0, i; int32_t max_id = 0; int k, st;
  const int nv = SIM->nbr_var, ne = SIM->nbr_expr, na = SIM->nbr_aux, np = SIM->nbr_psi;
  double *y, *e, *a, *p, *dr, *rr; int32_t *ids;
  par *pars = SIM->pop->start, *next;
  api.pop_size(h, &n, NULL, &max_id);
  while (pars != NULL) { next = pars->nextel; delete_el(SIM->pop, pars); pars = next; }
  SIM->max_id = 0;
  y = malloc(sizeof(double) * (n * nv + 1)); e = malloc(sizeof(double) * (n * ne + 1)); a = malloc(sizeof(double) * (n * na + 1));
  p = malloc(sizeof(double) * (n * np + 1)); dr = malloc(sizeof(double) * (n + 1)); rr = malloc(sizeof(double) * (n + 1));
  ids = malloc(sizeof(int32_t) * (n + 1));
  st = (y && e && a && p && dr && rr && ids) ? api.pop_download(h, 0, n, y, e, a, p, ids, dr, rr) : ODX_ENOMEM;
  for (i = 0; st == 0 && i < n; i++) {
    par_birth();
    pars = SIM->pop->end;
    for (k = 0; k < nv; k++) pars->y[k] = y[i * nv + k];
    for (k = 0; k < ne; k++) pars->expr[k] = e[i * ne + k];
    for (k = 0; k < na; k++) pars->aux[k] = a[i * na + k];
    for (k = 0; k < np; k++) pars->psi[k] = p[i * np + k];
    pars->death_rate = dr[i]; pars->repli_rate = rr[i];
    pars->id = ids[i];
  }
  SIM->max_id = max_id;
  free(y); free(e); free(a); free(p); free(dr); free(rr); free(ids);
  return st;
}

static void dump_population(const char *path) {
  FILE *f = fopen(path, "w");
  par *pars;
  if (!f) return;
  for (pars = SIM->pop->start; pars != NULL; pars = pars->nextel) {
    fwrite(&pars->id, sizeof(int), 1, f);
    fwrite(pars->y, sizeof(double), pars->nbr_y, f);
    fwrite(pars->expr, sizeof(double), pars->nbr_expr, f);
    fwrite(pars->aux, sizeof(double), pars->nbr_aux, f);
    fwrite(pars->psi, sizeof(double), pars->nbr_psi, f);
    fwrite(&pars->death_rate, sizeof(double), 1, f);
    fwrite(&pars->repli_rate, sizeof(double), 1, f);
  }
  fwrite(&SIM->max_id, sizeof(int), 1, f);
  fclose(f);
}

/* returns 1 when the run was handled here (*status_out = odesolver's return value), 0 to continue on the CPU path */
int odx_gpu_odesolver(int solver, double_array *tspan, int *status_out) {
  const char *lib = getenv("ODEXP_CUDA_MODEL");
  const odx_model_info *mi;
  odx_handle *h = NULL;
  odx_solver_opts o;
  odx_solver_report rep;
  struct sigaction act, old_act;
  int status, init = get_int("init"), method = method_of(solver), device = 0, have_old = 0;
  int64_t cap;

  if (!lib || !*lib) return 0;
  if (strncmp(get_str("popmode"), "single", 3) == 0) return 0;          /* single mode and the phase-space tools stay on the CPU */
  if (method < 0) return 0;
  /* a device model was asked for: from here on a problem is an error, never a silent run on the CPU */
  if (bind_api(lib) != 0) { *status_out = GSL_FAILURE; return 1; }
  mi = api.model();
  if (mi->nvar != SIM->nbr_var || mi->nexpr != SIM->nbr_expr || mi->naux != SIM->nbr_aux || mi->npsi != SIM->nbr_psi ||
      mi->nmfd != SIM->nbr_mfd || mi->npar != SIM->nbr_par) {
    PRINTERR("  CUDA model %s was built from another model file (dimensions differ)\n", lib);
    *status_out = GSL_FAILURE;
    return 1;
  }
  if (getenv("ODEXP_CUDA_DEVICE")) device = atoi(getenv("ODEXP_CUDA_DEVICE"));

  if (method == ODX_METH_ITERATION) {       /* as the CPU path does while parsing the method (:196-199) */
    set_dou("h0", 1.0);
    set_int("res", (int)(tspan->array[tspan->length - 1] - tspan->array[0]) + 1);
  }
  if (POP_SIZE == 0) { set_int("init", 0); init = 0; }                    /* :284-287 */
  if (init == 1) set_int("popsize", POP_SIZE);                            /* :289-293 */

  memset(&o, 0, sizeof o);
  o.method = method;
  o.res = get_int("res"); o.hmin = get_dou("hmin"); o.h0 = get_dou("h0");
  o.abstol = get_dou("abstol"); o.reltol = get_dou("reltol");
  o.ssahmin = get_dou("ssahmin"); o.aleap = get_dou("aleap");
  o.popsize = get_int("popsize"); o.seed = (uint32_t)get_int("seed");
  o.replay_ic = 1;                          /* initial conditions through the same serial rand() stream as the CPU path */
  o.traj_stride = getenv("ODEXP_CUDA_TRAJ_STRIDE") ? atoi(getenv("ODEXP_CUDA_TRAJ_STRIDE")) : 1;
  o.outdir = ".odexp";
  o.tspan = tspan->array; o.tspan_len = tspan->length;
  o.mu = SIM->mu;
  o.quick = 1; o.plot_x = get_int("x"); o.plot_y = get_int("y"); o.plot_z = get_int("z"); o.particle = get_int("particle");
  o.init_mode = 0;

  remove_id_files();
  cap = 4 * (int64_t)o.popsize + 4096;      /* room for births and replications */
  status = api.create(&h, device, cap);
  if (status != 0) {
    PRINTERR("  CUDA device not usable: %s (unset ODEXP_CUDA_MODEL to integrate on the CPU)\n", h ? api.last_error(h) : "odx_create failed");
    if (h) api.destroy(h);
    *status_out = status;
    return 1;
  }
  if ((status = api.set_params(h, SIM->mu, SIM->nbr_par)) != 0) goto fail;

  if (init == 1) {
    /* `opt init 1`: continue from the state the previous run left in SIM->pop (y, expr, ids), population intact (:298-316) */
    const int nv = SIM->nbr_var, ne = SIM->nbr_expr;
    const int64_t n = POP_SIZE;
    double *y = malloc(sizeof(double) * (n * nv + 1)), *e = malloc(sizeof(double) * (n * ne + 1));
    int32_t *ids = malloc(sizeof(int32_t) * (n + 1));
    par *pars; int64_t j = 0; int k;
    for (pars = SIM->pop->start; pars != NULL; pars = pars->nextel, j++) {
      for (k = 0; k < nv; k++) y[j * nv + k] = pars->y[k];
      for (k = 0; k < ne; k++) e[j * ne + k] = pars->expr[k];
      ids[j] = pars->id;
    }
    status = api.pop_upload(h, n, y, e, ids, SIM->max_id);
    free(y); free(e); free(ids);
    if (status != 0) goto fail;
    o.init_mode = 1;
  } else if (init == 2) {
    /* numerically set initial conditions (:339-354): ode_ic first (it draws the expressions), then y is overwritten */
    const int nv = SIM->nbr_var, ne = SIM->nbr_expr;
    double_array arr = {NULL, 0};
    arpn2array(get_str("init"), &arr);
    if (arr.length == o.popsize * nv) {
      double *e = malloc(sizeof(double) * ((int64_t)o.popsize * ne + 1));
      if ((status = api.pop_init(h, o.popsize, tspan->array[0], o.seed, 1)) != 0) { free(e); goto fail; }
      if ((status = api.pop_download(h, 0, o.popsize, NULL, e, NULL, NULL, NULL, NULL, NULL)) != 0) { free(e); goto fail; }
      status = api.pop_upload(h, o.popsize, arr.array, e, NULL, o.popsize);
      free(e);
      if (status != 0) goto fail;
      o.init_mode = 2;
    }
    free_double_array(arr);
  }

  /* SIGINT while the GPU integrates */
  running_handle = h;
  act.sa_handler = &gpu_abort_handler; act.sa_flags = 0; sigemptyset(&act.sa_mask);
  have_old = sigaction(SIGINT, &act, &old_act) == 0;

  PRINTV("  integrating on [%s%g%s, %s%g%s], in %s mode, on the GPU (%s)...\n", T_VAL, tspan->array[0], T_NOR, T_VAL,
         tspan->array[tspan->length - 1], T_NOR, get_str("popmode"), mi->name);
  status = api.odesolver(h, &o, &rep);
  running_handle = NULL;
  if (have_old) sigaction(SIGINT, &old_act, NULL);
  if (status != 0) goto fail;

  /* what the CPU path leaves behind in SIM */
  if (!(getenv("ODEXP_CUDA_MIRROR") && atoi(getenv("ODEXP_CUDA_MIRROR")) == 0)) {
    if ((status = mirror_population(h)) != 0) goto fail;
  }
  api.meanfields(h, SIM->meanfield);
  memset(SIM->event, 0, sizeof(SIM->event));
  if (EVENT_TYPE != EVENT_ABORT) EVENT_TYPE = EVENT_NONE;
  else printf("\n  simulation aborted\n");
  if (getenv("ODEXP_CUDA_DUMP_POP")) dump_population(getenv("ODEXP_CUDA_DUMP_POP"));
  if (rep.hmin_reached)      /* the reference's one-time warning (:604-612) */
    ADD_MSG_TO_BUFF("  warning: hmin reached at t = %f (h=%f). Continuing with h = %e", rep.hmin_t, o.hmin, o.hmin);
  /* the reference's closing line (:736-741); `function` = time inside pop_rhs, which here is all device time */
  PRINTV("\n  total: %s%g msec%s,", T_VAL, rep.total_ms, T_NOR);
  PRINTV(" solver: %s%g msec%s,", T_VAL, rep.solver_ms, T_NOR);
  PRINTV(" function: %s%g msec%s\n", T_VAL, rep.solver_ms, T_NOR);
  api.destroy(h);
  *status_out = 0;
  return 1;

fail:
  running_handle = NULL;
  PRINTERR("  CUDA solver error %d: %s\n", status, api.last_error(h));
  api.destroy(h);
  *status_out = status;
  return 1;
}
