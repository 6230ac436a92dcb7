/* TEST INFRASTRUCTURE (oracle build only). */
#ifndef ODX_STUB_GSL_QRNG_H
#define ODX_STUB_GSL_QRNG_H
typedef struct { const char *name; } gsl_qrng_type;
typedef struct { const gsl_qrng_type *type; unsigned int dimension; } gsl_qrng;
extern const gsl_qrng_type *gsl_qrng_sobol;
gsl_qrng *gsl_qrng_alloc(const gsl_qrng_type *T, unsigned int dimension);
void gsl_qrng_free(gsl_qrng *q);
int gsl_qrng_get(const gsl_qrng *q, double x[]);
#endif
