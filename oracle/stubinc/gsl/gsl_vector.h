/* TEST INFRASTRUCTURE (oracle build only). Minimal stand-in for <gsl/gsl_vector.h>. */
#ifndef ODX_STUB_GSL_VECTOR_H
#define ODX_STUB_GSL_VECTOR_H
#include <stdio.h>
#include <stdlib.h>
typedef struct { size_t size; size_t stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { double dat[2]; } gsl_complex;
#define GSL_REAL(z) ((z).dat[0])
#define GSL_IMAG(z) ((z).dat[1])
typedef struct { size_t size; size_t stride; double *data; void *block; int owner; } gsl_vector_complex;
gsl_vector *gsl_vector_alloc(const size_t n);
void gsl_vector_free(gsl_vector *v);
double gsl_vector_get(const gsl_vector *v, const size_t i);
void gsl_vector_set(gsl_vector *v, const size_t i, double x);
int gsl_vector_memcpy(gsl_vector *dest, const gsl_vector *src);
int gsl_vector_fprintf(FILE *stream, const gsl_vector *v, const char *format);
gsl_vector_complex *gsl_vector_complex_alloc(const size_t n);
void gsl_vector_complex_free(gsl_vector_complex *v);
gsl_complex gsl_vector_complex_get(const gsl_vector_complex *v, const size_t i);
gsl_vector_view gsl_vector_complex_real(gsl_vector_complex *v);
gsl_vector_view gsl_vector_complex_imag(gsl_vector_complex *v);
int gsl_vector_complex_fprintf(FILE *stream, const gsl_vector_complex *v, const char *format);
#endif
