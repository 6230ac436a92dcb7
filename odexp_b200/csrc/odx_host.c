/* odx_host.c — host-side simulation driver in C: the GPU branch of the reference's odesolver().
 *
 * Mirrors src/methods_odexp.c:83-763 for `popmode population`: output grid dt = (t1-t0)/(res-1), one loop iteration per
 * output interval or birth/death event, SSA / tau-leaping choice, stepper dispatch, and the two binary outputs the
 * reference's UI reads afterwards (.odexp/stats.dat, .odexp/traj.dat; record layouts src/datastruct.c:643-674).
 * The population itself never leaves the GPU between outputs: this file only talks to the C-ABI of include/odx.h.
 *
 * Differences from the reference that are deliberate and documented in DESIGN.md:
 *   - stopping times inside TIMESPAN (STOP_TSPAN) and STOPIF (STOP_COND) end the current output interval and re-initialise
 *     the living particles (src/methods_odexp.c:222-236, 453-459, 704-720);
 *   - traj.dat can be thinned (traj_stride) or switched off: at N = 2^24 one frame is 1.2 GB;
 *   - current.plot (the 3-column quick buffer of the gnuplot UI, fwrite_quick) is written when odx_solver_opts.quick is set:
 *     one particle and the mean fields per row, never the population;
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "odx.h"

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

typedef struct {
  FILE *fstats, *ftraj, *fquick;
  int ngx, ngy, ngz, cp;          /* opt x / y / z as indices, opt particle */
  unsigned int nbrsteps;
  int nmfd, nvar, naux, npsi, nexpr;
  double *mfd;
  /* download buffers */
  double *y, *e, *aux, *psi;
  int32_t *ids;
  int64_t bufcap;
  /* frames in flight (odx_frame_begin / odx_frame_end), oldest first */
  int pend[2], npend;
} out_state;

/* fwrite_SIM (src/datastruct.c:643-653): t, meanfield[nmfd], N, stop_flag, event[3] */
static void write_stats_row(out_state *os, double t, int n, int stop_flag, const int event[3]) {
  fwrite(&t, sizeof(double), 1, os->fstats);
  fwrite(os->mfd, sizeof(double), os->nmfd, os->fstats);
  fwrite(&n, sizeof(int), 1, os->fstats);
  fwrite(&stop_flag, sizeof(int), 1, os->fstats);
  fwrite(event, sizeof(int), 3, os->fstats);
}

static int ensure_buffers(out_state *os, int64_t n) {
  if (n <= os->bufcap) return 0;
  int64_t cap = n + n / 4 + 16;
  free(os->y); free(os->e); free(os->aux); free(os->psi); free(os->ids);
  os->y = malloc(sizeof(double) * cap * (os->nvar > 0 ? os->nvar : 1));
  os->e = malloc(sizeof(double) * cap * (os->nexpr > 0 ? os->nexpr : 1));
  os->aux = malloc(sizeof(double) * cap * (os->naux > 0 ? os->naux : 1));
  os->psi = malloc(sizeof(double) * cap * (os->npsi > 0 ? os->npsi : 1));
  os->ids = malloc(sizeof(int32_t) * cap);
  os->bufcap = cap;
  return (os->y && os->e && os->aux && os->psi && os->ids) ? 0 : ODX_ENOMEM;
}

/* fwrite_quick (src/methods_odexp.c:1353-1436): one (x, y, z) row of current.plot.  An axis index is -1 for T, then counts the
 * dynamical variables, auxiliary variables and couplings of the current particle (getpar(opt particle): looked up by id), then
 * the mean fields.  As coded in the reference a dynamical variable is read from the state array at y[cp * tx + ng] with
 * tx = nvar + naux + npsi (not cp * nvar): reproduced, NaN where the reference would read past the end of y.  Multi-GPU: a
 * rank that does not own the particle writes NaN in its columns. */
static int write_quick_row(odx_handle *h, out_state *os, double t) {
  const int tx = os->nvar + os->naux + os->npsi, ng[3] = {os->ngx, os->ngy, os->ngz};
  double row[3];
  int64_t n = 0, pos = -1;
  int k, st, need_par = 0;
  if (!os->fquick) return 0;
  odx_pop_size(h, &n, NULL, NULL);
  for (k = 0; k < 3; k++) if (ng[k] >= 0 && ng[k] < tx) need_par = 1;
  if (need_par) {                                 /* pars = getpar(cp): NULL when the particle is dead (or on another GPU) */
    if ((st = odx_pop_find_id(h, os->cp, &pos)) != 0) return st;
    if ((st = ensure_buffers(os, 1)) != 0) return st;
    if (pos >= 0 && (st = odx_pop_download(h, pos, 1, NULL, NULL, os->aux, os->psi, NULL, NULL, NULL)) != 0) return st;
  }
  for (k = 0; k < 3; k++) {
    const int g = ng[k];
    if (g == -1) row[k] = t;
    else if (g < os->nvar && pos >= 0) {          /* y + cp * tx + ng: an element of the particle-major state array */
      const int64_t el = (int64_t)os->cp * tx + g, ip = el / os->nvar;
      row[k] = NAN;
      if (ip < n) {
        if ((st = odx_pop_download(h, ip, 1, os->y, NULL, NULL, NULL, NULL, NULL, NULL)) != 0) return st;
        row[k] = os->y[el % os->nvar];
      }
    }
    else if (g < os->nvar + os->naux && pos >= 0) row[k] = os->aux[g - os->nvar];
    else if (g < tx && pos >= 0) row[k] = os->psi[g - os->nvar - os->naux];
    else if (g >= tx) row[k] = (g - tx < os->nmfd) ? os->mfd[g - tx] : NAN;
    else row[k] = NAN;
  }
  fwrite(row, sizeof(double), 3, os->fquick);
  return 0;
}

/* the oldest frame in flight: its records (assembled on the device in the file's byte layout) and its current.plot row go to the files */
static int flush_frame(odx_handle *h, out_state *os) {
  const void *rec = NULL; int64_t nrec = 0, recb = 0; double q[3]; int st;
  if (os->npend == 0) return 0;
  if ((st = odx_frame_end(h, os->pend[0], &rec, &nrec, &recb, q)) != 0) return st;
  os->pend[0] = os->pend[1]; os->npend--;
  if (os->fquick) fwrite(q, sizeof(double), 3, os->fquick);
  if (os->ftraj && nrec > 0) fwrite(rec, (size_t)recb, (size_t)nrec, os->ftraj);
  return 0;
}
static int flush_frames(odx_handle *h, out_state *os) {
  int st;
  while (os->npend > 0) if ((st = flush_frame(h, os)) != 0) return st;
  return 0;
}
/* fwrite_quick + fwrite_all_particles (src/methods_odexp.c:1353-1436, src/datastruct.c:655-674) without stalling the integration:
 * the frame is snapshotted and copied by the device asynchronously, written when the next one is started or at the next flush */
static int write_traj_frame_sync(odx_handle *h, out_state *os, double t, int stride);
static int write_traj_frame(odx_handle *h, out_state *os, double t, int stride) {
  int st, ticket = 0;
  if (os->nmfd > 24) return write_traj_frame_sync(h, os, t, stride);
  if (os->npend == 2 && (st = flush_frame(h, os)) != 0) return st;
  if ((st = odx_frame_begin(h, os->nbrsteps, t, (os->ftraj && stride > 0) ? stride : 0, os->mfd, os->fquick != NULL, os->ngx, os->ngy, os->ngz,
                            os->cp, &ticket)) != 0) return st;
  os->pend[os->npend++] = ticket;
  os->nbrsteps++;
  return 0;
}

/* the synchronous form (models with more mean fields than the record kernel takes by value) */
static int write_traj_frame_sync(odx_handle *h, out_state *os, double t, int stride) {
  int64_t n = 0, m, i;
  int st;
  if ((st = write_quick_row(h, os, t)) != 0) return st;          /* fwrite_quick precedes fwrite_all_particles at every call site */
  if (!os->ftraj || stride <= 0) { os->nbrsteps++; return 0; }
  odx_pop_size(h, &n, NULL, NULL);
  m = (n + stride - 1) / stride;                  /* particles 0, stride, 2*stride, ... */
  if ((st = ensure_buffers(os, m)) != 0) return st;
  if (stride == 1) st = odx_pop_download(h, 0, n, os->y, os->e, os->aux, os->psi, os->ids, NULL, NULL);
  else st = odx_pop_download_strided(h, 0, stride, m, os->y, os->e, os->aux, os->psi, os->ids);
  if (st != 0) return st;
  for (i = 0; i < m; i++) {
    fwrite(&os->nbrsteps, sizeof(unsigned int), 1, os->ftraj);
    fwrite(&os->ids[i], sizeof(int), 1, os->ftraj);
    fwrite(&t, sizeof(double), 1, os->ftraj);
    fwrite(os->y + i * os->nvar, sizeof(double), os->nvar, os->ftraj);
    fwrite(os->aux + i * os->naux, sizeof(double), os->naux, os->ftraj);
    fwrite(os->psi + i * os->npsi, sizeof(double), os->npsi, os->ftraj);
    fwrite(os->mfd, sizeof(double), os->nmfd, os->ftraj);
    fwrite(os->e + i * os->nexpr, sizeof(double), os->nexpr, os->ftraj);
  }
  os->nbrsteps++;
  return 0;
}

static int cmp_double(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

int odx_odesolver(odx_handle *h, const odx_solver_opts *o, odx_solver_report *rep) {
  const odx_model_info *mi = odx_model();
  const double t_start = now_ms();
  double tot_solver = 0.0;
  int hmin_alert = 0; double hmin_t = 0.0;
  double t, t1, dt, tnext, dt_ssa, dt_leaping, dt_dyn, dt_next, sum_rates = 0.0, birth_rate = 0.0, expected_dt_ssa;
  double hstep = o->h0;
  int res = o->res;
  int status = ODX_SUCCESS, stop_flag, event[3] = {0, 0, 0};
  int64_t n = 0, steps = 0, i;
  double particle_steps = 0.0;
  enum { SSA, TAU_LEAPING } bd_meth = SSA;
  out_state os;
  char path[4096];
  const odx_event_row *rows = NULL;       /* pinned memory owned by the library */
  int64_t nrows = 0;
  uint64_t leap_index = 0;
  double *tstops = NULL, nextstop = INFINITY;
  int nstops = 0, idx_stop = 0;

  memset(&os, 0, sizeof os);
  if (rep) memset(rep, 0, sizeof *rep);
  if (o->tspan_len < 2) return ODX_EINVAL;
  if (o->method == ODX_METH_ITERATION) {          /* set_dou("h0", 1.0); set_int("res", (int)(t1 - t0) + 1) */
    hstep = 1.0;
    res = (int)(o->tspan[o->tspan_len - 1] - o->tspan[0]) + 1;
  }
  if (res < 2) return ODX_EINVAL;
  if (o->tspan_len > 2) {                         /* interior TIMESPAN values are stopping times, sorted (:222-236) */
    nstops = o->tspan_len - 2;
    if ((tstops = malloc(sizeof(double) * nstops)) == NULL) return ODX_ENOMEM;
    memcpy(tstops, o->tspan + 1, sizeof(double) * nstops);
    qsort(tstops, nstops, sizeof(double), cmp_double);
    nextstop = tstops[idx_stop++];
  }
  os.nmfd = mi->nmfd; os.nvar = mi->nvar; os.naux = mi->naux; os.npsi = mi->npsi; os.nexpr = mi->nexpr;
  os.mfd = calloc(os.nmfd > 0 ? os.nmfd : 1, sizeof(double));

  t = o->tspan[0];
  t1 = o->tspan[o->tspan_len - 1];
  stop_flag = ODX_EVENT_T0;
  dt = (t1 - t) / (double)(res - 1);

  snprintf(path, sizeof path, "%s/stats.dat", o->outdir ? o->outdir : ".odexp");
  if ((os.fstats = fopen(path, "w")) == NULL) { status = ODX_EFAILED; goto done; }
  if (o->traj_stride > 0) {
    snprintf(path, sizeof path, "%s/traj.dat", o->outdir ? o->outdir : ".odexp");
    if ((os.ftraj = fopen(path, "w")) == NULL) { status = ODX_EFAILED; goto done; }
  }
  if (o->quick) {                                 /* QUICK_BUFFER, src/macros.h:30 */
    snprintf(path, sizeof path, "%s/current.plot", o->outdir ? o->outdir : ".odexp");
    if ((os.fquick = fopen(path, "w")) == NULL) { status = ODX_EFAILED; goto done; }
    os.ngx = o->plot_x; os.ngy = o->plot_y; os.ngz = o->plot_z; os.cp = o->particle;
  }

  if (o->mu && (status = odx_set_params(h, o->mu, mi->npar)) != 0) goto done;
  if ((status = odx_set_tolerances(h, o->abstol, o->reltol)) != 0) goto done;
  if ((status = odx_set_seed(h, o->seed)) != 0) goto done;
  /* population initialisation: par_birth() x popsize, event = {-1, 1, .}, ode_ic (methods_odexp.c:315-337).
   * Multi-GPU: every rank runs this same loop on its contiguous shard of the ordered population. */
  {
    int rank = 0, nranks = 1;
    int64_t base, extra, local;
    odx_comm_rank(h, &rank, &nranks);
    base = o->popsize / nranks; extra = o->popsize % nranks;
    local = base + (rank < extra ? 1 : 0);
    if (o->init_mode == 0) {
      if ((status = odx_pop_init(h, local, t, o->seed, o->replay_ic)) != 0) goto done;
    }
    /* init_mode 1 / 2: the caller has uploaded the population (opt init 1: the state the previous run left in SIM->pop) */
  }
  /* first RHS evaluation fills aux / psi / mean fields / rates (methods_odexp.c:364); EVENT_TYPE is EVENT_T0 while it runs */
  odx_set_event_type(h, stop_flag, hstep);
  if ((status = odx_rhs(h, t)) != 0) goto done;
  odx_set_event_type(h, ODX_EVENT_NONE, hstep);
  if ((status = odx_meanfields(h, os.mfd)) != 0) goto done;
  if (mi->has_stopif) odx_stop_flag(h, NULL);      /* EVENT_TYPE = EVENT_NONE after the initial evaluation (:380) */
  odx_pop_size(h, NULL, &n, NULL);
  for (i = 0; i < (o->init_mode == 0 ? (int64_t)o->popsize : n); i++) {      /* one stats row per initial particle (methods_odexp.c:372-376) */
    /* SIM->event is {-1, 1, .} after par_birth; with `opt init 1` it still holds the zeros the previous run left (:339-354, 675) */
    event[0] = o->init_mode == 1 ? 0 : -1; event[1] = o->init_mode == 1 ? 0 : 1; event[2] = (int)i;
    write_stats_row(&os, t, (int)n, stop_flag, event);
  }
  memset(event, 0, sizeof event);
  stop_flag = ODX_EVENT_NONE;
  if ((status = write_traj_frame(h, &os, t, o->traj_stride)) != 0) goto done;
  if ((status = odx_evolve_reset(h)) != 0) goto done;

  odx_clear_abort(h);
  while (t < t1 && n > 0 && !odx_abort_requested(h)) {        /* while (t < t1 && !abort_odesolver_flag && POP_SIZE > 0), :451 */
    double u;
    dt_dyn = fmin(dt, t1 - t);
    if ((t <= nextstop) && (dt_dyn >= (nextstop - t))) { dt_dyn = nextstop - t; stop_flag = ODX_STOP_TSPAN; }
    /* SSA_timestep (methods_odexp.c:1528-1553): always consumes one uniform */
    if ((status = odx_rates(h, &sum_rates, &birth_rate)) != 0) break;
    u = odx_draw_uniform(h, o->replay_ic);
    dt_ssa = log(1 / (1 - u)) / sum_rates;
    expected_dt_ssa = 1.0 / sum_rates;
    bd_meth = (expected_dt_ssa > (dt_dyn * o->ssahmin)) ? SSA : TAU_LEAPING;
    if (bd_meth == SSA) {
      if (dt_ssa < dt_dyn) { dt_next = dt_ssa; stop_flag = ODX_STOP_BIRTH | ODX_STOP_DEATH | ODX_STOP_REPLI; }
      else dt_next = dt_dyn;
    } else {
      stop_flag = ODX_STOP_BIRTH | ODX_STOP_DEATH | ODX_STOP_REPLI;
      dt_leaping = fmax(1, o->aleap * (double)n) * expected_dt_ssa;
      dt_next = fmin(dt_leaping, dt_dyn);
    }
    tnext = t + dt_next;

    {
      const double c0 = now_ms();
      while (t < tnext) {
        const int64_t n_step = n;
        if (mi->ndelay) odx_set_event_type(h, stop_flag, hstep);      /* what delay() sees: a pending stopping time / event flag */
        if (o->method == ODX_METH_ITERATION) {
          status = odx_iteration_apply(h, &t, 1);
        } else if (o->method == ODX_METH_FE || o->method == ODX_METH_DDE) {
          /* this step lands on tnext: keep aux / rates (always when a STOPIF may end the interval at any step) */
          const int last = mi->has_stopif || hstep > ((tnext - t) - o->hmin);
          status = (o->method == ODX_METH_FE) ? odx_fe_apply(h, &t, tnext, &hstep, o->hmin, o->h0, last)
                                              : odx_dde_apply(h, &t, tnext, &hstep, o->hmin, o->h0, last);
        } else {
          status = odx_evolve_apply(h, o->method, &t, tnext, &hstep);
        }
        if (mi->has_stopif && status == ODX_SUCCESS) {            /* EVENT_TYPE = STOP_COND from inside the RHS (:453-459) */
          int raised = 0;
          if ((status = odx_stop_flag(h, &raised)) == ODX_SUCCESS && raised) {
            stop_flag = ODX_STOP_COND; steps++; particle_steps += (double)n_step;
            break;
          }
        }
        if (status != ODX_SUCCESS) { stop_flag = ODX_EVENT_GSL; break; }
        steps++; particle_steps += (double)n_step;
        if (hstep < o->hmin) {                 /* :604-612: continue with h = hmin, tell the user once */
          hstep = o->hmin;
          if (!hmin_alert && t < t1) { hmin_alert = 1; hmin_t = t; }
        }
      }
      /* the interval's steps are queued: write the frames that have landed while the GPU works */
      if (status == ODX_SUCCESS && (status = flush_frames(h, &os)) != 0) stop_flag = ODX_EVENT_GSL;
      if (odx_abort_requested(h)) stop_flag = ODX_EVENT_ABORT;   /* the SIGINT handler's EVENT_TYPE = EVENT_ABORT (:23-30) */
      if (t >= t1) stop_flag = ODX_EVENT_TFINAL;
      if (odx_meanfields(h, os.mfd) != 0 && status == 0) status = ODX_ECUDA;
      write_stats_row(&os, t, (int)n, stop_flag, event);
      if (status != ODX_SUCCESS && status != ODX_EBADFUNC) break;
      tot_solver += now_ms() - c0;
    }
    if ((status = write_traj_frame(h, &os, t, o->traj_stride)) != 0) break;

    if (stop_flag & (ODX_STOP_BIRTH | ODX_STOP_DEATH | ODX_STOP_REPLI)) {
      if (bd_meth == SSA) {
        odx_event_row row;
        const double uu = odx_draw_uniform(h, o->replay_ic);
        if ((status = odx_ssa_apply(h, t, uu, &row)) != 0) break;
        if (row.event[1] != 0) {      /* sum of rates > 0: one event happened, one stats row (methods_odexp.c:1654) */
          write_stats_row(&os, t, row.n, row.stop_flag, row.event);
        }
      } else {
        if ((status = odx_tauleap_apply_view(h, t, dt_next, leap_index++, &rows, &nrows)) != 0) break;
        for (i = 0; i < nrows; i++) write_stats_row(&os, t, rows[i].n, rows[i].stop_flag, rows[i].event);
      }
      odx_pop_size(h, NULL, &n, NULL);
      if (n > 0) {
        if (mi->ndelay) odx_set_event_type(h, stop_flag, hstep);
        if ((status = odx_rhs(h, t)) != 0) break;                 /* methods_odexp.c:687 */
        if ((status = odx_meanfields(h, os.mfd)) != 0) break;
      }
      if ((status = write_traj_frame(h, &os, t, o->traj_stride)) != 0) break;
      memset(event, 0, sizeof event);
      odx_evolve_reset(h);                                         /* gsl step / evolve re-allocated: count = 0 */
    }
    if (stop_flag & (ODX_STOP_COND | ODX_STOP_TSPAN)) {            /* reset the dynamical variables (:704-714) */
      if ((status = odx_pop_reinit(h, t, o->replay_ic)) != 0) break;
      if (mi->ndelay) odx_set_event_type(h, stop_flag, hstep);
      if ((status = odx_rhs(h, t)) != 0) break;
      if ((status = odx_meanfields(h, os.mfd)) != 0) break;
      if ((status = write_traj_frame(h, &os, t, o->traj_stride)) != 0) break;
    }
    if (stop_flag & ODX_STOP_TSPAN) nextstop = (idx_stop < nstops) ? tstops[idx_stop++] : INFINITY;
    if (mi->has_stopif) odx_stop_flag(h, NULL);                     /* EVENT_TYPE = EVENT_NONE */
    stop_flag = ODX_EVENT_NONE;
    if (mi->ndelay) odx_set_event_type(h, stop_flag, hstep);
  }
  if (status == ODX_SUCCESS) status = odx_sync(h);

done:
  { int st2 = flush_frames(h, &os); if (status == ODX_SUCCESS) status = st2; }
  if (os.fstats) fclose(os.fstats);
  if (os.ftraj) fclose(os.ftraj);
  if (os.fquick) fclose(os.fquick);
  free(tstops);
  free(os.mfd); free(os.y); free(os.e); free(os.aux); free(os.psi); free(os.ids);
  if (rep) {
    rep->total_ms = now_ms() - t_start; rep->solver_ms = tot_solver; rep->steps = steps; rep->particle_steps = particle_steps;
    rep->status = status; rep->final_n = n; rep->hmin_reached = hmin_alert; rep->hmin_t = hmin_t;
  }
  return status;
}
