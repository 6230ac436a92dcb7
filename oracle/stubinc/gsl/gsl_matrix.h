/* TEST INFRASTRUCTURE (oracle build only). Minimal stand-in for <gsl/gsl_matrix.h>. */
#ifndef ODX_STUB_GSL_MATRIX_H
#define ODX_STUB_GSL_MATRIX_H
#include <stdlib.h>
#include <gsl/gsl_vector.h>
typedef struct { size_t size1; size_t size2; size_t tda; double *data; void *block; int owner; } gsl_matrix;
gsl_matrix *gsl_matrix_alloc(const size_t n1, const size_t n2);
void gsl_matrix_free(gsl_matrix *m);
void gsl_matrix_set(gsl_matrix *m, const size_t i, const size_t j, const double x);
double gsl_matrix_get(const gsl_matrix *m, const size_t i, const size_t j);
#endif
