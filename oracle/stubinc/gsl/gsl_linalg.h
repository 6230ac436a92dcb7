/* TEST INFRASTRUCTURE (oracle build only). */
#ifndef ODX_STUB_GSL_LINALG_H
#define ODX_STUB_GSL_LINALG_H
#include <gsl/gsl_vector.h>
#include <gsl/gsl_matrix.h>
typedef struct { size_t size; size_t *data; } gsl_permutation;
gsl_permutation *gsl_permutation_alloc(const size_t n);
void gsl_permutation_free(gsl_permutation *p);
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum);
int gsl_linalg_LU_solve(const gsl_matrix *LU, const gsl_permutation *p, const gsl_vector *b, gsl_vector *x);
#endif
