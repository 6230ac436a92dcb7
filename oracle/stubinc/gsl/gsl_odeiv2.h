/* TEST INFRASTRUCTURE (oracle build only). Stand-in for <gsl/gsl_odeiv2.h>; the implementation
 * behind it is the clean-room oracle/minigsl.c (GSL itself is not available offline: parity with
 * real GSL for the adaptive steppers is UNPINNED, see oracle/README.md). */
#ifndef ODX_STUB_GSL_ODEIV2_H
#define ODX_STUB_GSL_ODEIV2_H
#include <stdlib.h>
typedef struct {
  int (*function)(double t, const double y[], double dydt[], void *params);
  int (*jacobian)(double t, const double y[], double *dfdy, double dfdt[], void *params);
  size_t dimension;
  void *params;
} gsl_odeiv2_system;
typedef struct gsl_odeiv2_step_type_struct gsl_odeiv2_step_type;
typedef struct gsl_odeiv2_step_struct gsl_odeiv2_step;
typedef struct gsl_odeiv2_control_struct gsl_odeiv2_control;
typedef struct gsl_odeiv2_evolve_struct gsl_odeiv2_evolve;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_rk2;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_rk4;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_rkf45;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_rkck;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_rk8pd;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_bsimp;
gsl_odeiv2_step *gsl_odeiv2_step_alloc(const gsl_odeiv2_step_type *T, size_t dim);
void gsl_odeiv2_step_free(gsl_odeiv2_step *s);
gsl_odeiv2_control *gsl_odeiv2_control_y_new(double eps_abs, double eps_rel);
void gsl_odeiv2_control_free(gsl_odeiv2_control *c);
gsl_odeiv2_evolve *gsl_odeiv2_evolve_alloc(size_t dim);
void gsl_odeiv2_evolve_free(gsl_odeiv2_evolve *e);
int gsl_odeiv2_evolve_apply(gsl_odeiv2_evolve *e, gsl_odeiv2_control *con, gsl_odeiv2_step *step,
                            const gsl_odeiv2_system *dydt, double *t, double t1, double *h, double y[]);
#endif
