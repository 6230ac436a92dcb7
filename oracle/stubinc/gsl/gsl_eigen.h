/* TEST INFRASTRUCTURE (oracle build only). */
#ifndef ODX_STUB_GSL_EIGEN_H
#define ODX_STUB_GSL_EIGEN_H
#include <gsl/gsl_vector.h>
#include <gsl/gsl_matrix.h>
typedef struct { size_t size; } gsl_eigen_nonsymm_workspace;
gsl_eigen_nonsymm_workspace *gsl_eigen_nonsymm_alloc(const size_t n);
void gsl_eigen_nonsymm_free(gsl_eigen_nonsymm_workspace *w);
int gsl_eigen_nonsymm(gsl_matrix *A, gsl_vector_complex *eval, gsl_eigen_nonsymm_workspace *w);
#endif
