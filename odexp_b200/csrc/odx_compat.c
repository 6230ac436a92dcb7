/* odx_compat.c — the reference's stepper entry points (src/methods_odexp.h:47-63) by name and signature, implemented on the C-ABI
 * of include/odx.h.  Plain C, part of every model library.  See include/odx_compat.h for the meaning of each argument. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "odx_compat.h"

static struct {
  odx_handle *h;
  double hmin, h0;
  int replay;
  FILE *stats;
  uint64_t leap_index;
  odx_event_row ssa_row;
  const odx_event_row *rows; int64_t nrows;
} B;

void odx_compat_bind(odx_handle *h, double hmin, double h0, int replay, FILE *stats) {
  memset(&B, 0, sizeof B);
  B.h = h; B.hmin = hmin; B.h0 = h0; B.replay = replay; B.stats = stats;
}
int odx_compat_last_rows(const odx_event_row **rows, int64_t *nrows) {
  if (rows) *rows = B.rows;
  if (nrows) *nrows = B.nrows;
  return ODX_SUCCESS;
}

static int state_back(odx_handle *h, double y[]) {            /* y <- the new state, like the reference's steppers leave it */
  int64_t n = 0;
  if (y == NULL) return ODX_SUCCESS;
  odx_pop_size(h, &n, NULL, NULL);
  return odx_pop_download(h, 0, n, y, NULL, NULL, NULL, NULL, NULL, NULL);
}

/* fe_apply (src/methods_odexp.c:1790-1814) */
int fe_apply(odx_gsl_system *sys, double *t, double tnext, double *h, double y[]) {
  odx_handle *hd = (odx_handle *)sys->params;
  int st = odx_fe_apply(hd, t, tnext, h, B.hmin, B.h0, 1);
  return st != ODX_SUCCESS ? st : state_back(hd, y);
}
/* iteration_apply (src/methods_odexp.c:1816-1823) */
int iteration_apply(odx_gsl_system *sys, double *t, double y[]) {
  odx_handle *hd = (odx_handle *)sys->params;
  int st = odx_iteration_apply(hd, t, 1);
  return st != ODX_SUCCESS ? st : state_back(hd, y);
}
/* dde_apply (src/methods_odexp.c:1825-1874) */
int dde_apply(odx_gsl_system *sys, double *t, double tnext, double *h, double y[]) {
  odx_handle *hd = (odx_handle *)sys->params;
  int st = odx_dde_apply(hd, t, tnext, h, B.hmin, B.h0, 1);
  return st != ODX_SUCCESS ? st : state_back(hd, y);
}

/* SSA_timestep (src/methods_odexp.c:1528-1553): r = pop_birth_rate + sum(death + repli); dt = exprand(r), one uniform consumed */
double SSA_timestep(double *sumr) {
  double r = 0.0, birth = 0.0, u;
  if (B.h == NULL || odx_rates(B.h, &r, &birth) != ODX_SUCCESS) { if (sumr) *sumr = 0.0; return INFINITY; }
  u = odx_draw_uniform(B.h, B.replay);
  if (sumr) *sumr = r;
  return log(1 / (1 - u)) / r;                                /* exprand, src/rand_gen.c:68-72 */
}

static void write_rows(double t, const odx_event_row *rows, int64_t n) {
  const odx_model_info *mi = odx_model();
  double mfd[64];
  int64_t i;
  if (B.stats == NULL || n <= 0) return;
  if (mi->nmfd > 64 || odx_meanfields(B.h, mfd) != ODX_SUCCESS) return;
  for (i = 0; i < n; i++) {                                   /* fwrite_SIM (src/datastruct.c:643-653) */
    fwrite(&t, sizeof(double), 1, B.stats);
    fwrite(mfd, sizeof(double), mi->nmfd, B.stats);
    fwrite(&rows[i].n, sizeof(int), 1, B.stats);
    fwrite(&rows[i].stop_flag, sizeof(int), 1, B.stats);
    fwrite(rows[i].event, sizeof(int), 3, B.stats);
  }
}

/* SSA_apply_birthdeath (src/methods_odexp.c:1556-1660): one uniform picks the event over [death.., repli.., birth] */
void SSA_apply_birthdeath(const double t, odx_odeic single_ic) {
  double u;
  (void)single_ic;
  B.rows = NULL; B.nrows = 0;
  if (B.h == NULL) return;
  u = odx_draw_uniform(B.h, B.replay);
  if (odx_ssa_apply(B.h, t, u, &B.ssa_row) != ODX_SUCCESS) return;
  if (B.ssa_row.event[1] != 0) { B.rows = &B.ssa_row; B.nrows = 1; write_rows(t, B.rows, 1); }
}

/* tau_leaping_apply_birthdeath (src/methods_odexp.c:1662-1762) */
void tau_leaping_apply_birthdeath(const double t, const double dt, odx_odeic single_ic) {
  (void)single_ic;
  B.rows = NULL; B.nrows = 0;
  if (B.h == NULL) return;
  if (odx_tauleap_apply_view(B.h, t, dt, B.leap_index++, &B.rows, &B.nrows) != ODX_SUCCESS) { B.rows = NULL; B.nrows = 0; return; }
  write_rows(t, B.rows, B.nrows);
}

/* ncumsum (src/methods_odexp.c:1764-1777): running sum in place, total to *sumx, then normalised by the total */
int ncumsum(double *x, int len, double *sumx) {
  int i;
  for (i = 1; i < len; i++) x[i] += x[i - 1];
  *sumx = x[len - 1];
  for (i = 0; i < len; i++) x[i] /= x[len - 1];
  return 0;
}
