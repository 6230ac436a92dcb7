/* TEST INFRASTRUCTURE (oracle build only). Implementation in oracle/minigsl.c (clean room). */
#ifndef ODX_STUB_GSL_CHEBYSHEV_H
#define ODX_STUB_GSL_CHEBYSHEV_H
#include <stdlib.h>
#include <gsl/gsl_math.h>
typedef struct { double *c; size_t order; double a; double b; size_t order_sp; double *f; } gsl_cheb_series;
gsl_cheb_series *gsl_cheb_alloc(const size_t order);
void gsl_cheb_free(gsl_cheb_series *cs);
int gsl_cheb_init(gsl_cheb_series *cs, const gsl_function *func, const double a, const double b);
double *gsl_cheb_coeffs(const gsl_cheb_series *cs);
int gsl_cheb_eval_n_err(const gsl_cheb_series *cs, const size_t order, const double x, double *result, double *abserr);
#endif
