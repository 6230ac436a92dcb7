/* TEST INFRASTRUCTURE (oracle build only). Declarations for the steady-state tools (never executed by the oracle). */
#ifndef ODX_STUB_GSL_MULTIROOTS_H
#define ODX_STUB_GSL_MULTIROOTS_H
#include <gsl/gsl_vector.h>
#include <gsl/gsl_matrix.h>
typedef struct { int (*f)(const gsl_vector *x, void *params, gsl_vector *f); size_t n; void *params; } gsl_multiroot_function;
typedef struct { const char *name; } gsl_multiroot_fsolver_type;
typedef struct { const gsl_multiroot_fsolver_type *type; gsl_multiroot_function *function;
                 gsl_vector *x; gsl_vector *f; gsl_vector *dx; void *state; } gsl_multiroot_fsolver;
extern const gsl_multiroot_fsolver_type *gsl_multiroot_fsolver_hybrids;
gsl_multiroot_fsolver *gsl_multiroot_fsolver_alloc(const gsl_multiroot_fsolver_type *T, size_t n);
void gsl_multiroot_fsolver_free(gsl_multiroot_fsolver *s);
int gsl_multiroot_fsolver_set(gsl_multiroot_fsolver *s, gsl_multiroot_function *f, const gsl_vector *x);
int gsl_multiroot_fsolver_iterate(gsl_multiroot_fsolver *s);
int gsl_multiroot_test_residual(const gsl_vector *f, double epsabs);
int gsl_multiroot_test_delta(const gsl_vector *dx, const gsl_vector *x, double epsabs, double epsrel);
int gsl_multiroot_fdjacobian(gsl_multiroot_function *F, const gsl_vector *x, const gsl_vector *f, double epsrel, gsl_matrix *jacobian);
#endif
