/* odx_compat.h — the reference's stepper entry points by NAME and SIGNATURE, over the resident device population.
 *
 * src/methods_odexp.h:47-63 declares
 *     double SSA_timestep(double *r);
 *     void   SSA_apply_birthdeath(const double t, odeic single_ic);
 *     void   tau_leaping_apply_birthdeath(const double t, const double dt, odeic single_ic);
 *     int    ncumsum(double *x, int len, double *sumx);
 *     int    fe_apply(gsl_odeiv2_system *sys, double *t, double tnext, double *h, double y[]);
 *     int    iteration_apply(gsl_odeiv2_system *sys, double *t, double y[]);
 *     int    dde_apply(gsl_odeiv2_system *sys, double *t, double tnext, double *h, double y[]);
 * Every model library exports the same symbols (csrc/odx_compat.c).  Differences in meaning, all forced by where the state lives:
 *   sys->params   the odx_handle of the population (the reference passes the list head SIM->pop->start, :535); sys->function and
 *                 sys->jacobian are not called -- the generated device code is the right-hand side;
 *   y[]           may be NULL (state stays on the device).  When it is not NULL it receives the new state, particle-major like the
 *                 reference's y (one device-to-host copy per call);
 *   hmin, h0      read by the reference from GOPTS inside the steppers (:1796, 1810): here from odx_compat_bind();
 *   SSA_* / tau_* act on the handle bound by odx_compat_bind() (the reference acts on the global SIM); single_ic is not called: initial
 *                 conditions of new particles run on the device (keyed Philox) or through the replayed serial stream (replay = 1);
 *                 the stats.dat rows the reference writes from inside these functions (fwrite_SIM, src/datastruct.c:643-653) go to the
 *                 FILE given to odx_compat_bind(), when one is given, and are always available through odx_compat_last_rows().
 */
#ifndef ODX_COMPAT_H
#define ODX_COMPAT_H
#include <stdio.h>
#include "odx.h"
#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

/* layout of gsl_odeiv2_system (GSL >= 1.15: function, jacobian, dimension, params); a host that includes <gsl/gsl_odeiv2.h>
 * passes its gsl_odeiv2_system* unchanged */
typedef struct {
  int (*function)(double t, const double y[], double dydt[], void *params);
  int (*jacobian)(double t, const double y[], double *dfdy, double dfdt[], void *params);
  size_t dimension;
  void *params;
} odx_gsl_system;
typedef int (*odx_odeic)(double t, double *y, void *params);      /* odeic, src/odexp.h:20 */

/* the values the reference reads from its globals: opt hmin / h0 (GOPTS), the replay switch, SIM->fid (stats.dat, may be NULL) */
void odx_compat_bind(odx_handle *h, double hmin, double h0, int replay, FILE *stats);
int  odx_compat_last_rows(const odx_event_row **rows, int64_t *nrows);

int fe_apply(odx_gsl_system *sys, double *t, double tnext, double *h, double y[]);
int iteration_apply(odx_gsl_system *sys, double *t, double y[]);
int dde_apply(odx_gsl_system *sys, double *t, double tnext, double *h, double y[]);
double SSA_timestep(double *r);
void SSA_apply_birthdeath(const double t, odx_odeic single_ic);
void tau_leaping_apply_birthdeath(const double t, const double dt, odx_odeic single_ic);
int ncumsum(double *x, int len, double *sumx);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif
