/* TEST INFRASTRUCTURE (oracle build only). */
#ifndef ODX_STUB_GSL_MATH_H
#define ODX_STUB_GSL_MATH_H
#include <math.h>
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F,x) (*((F)->function))(x,(F)->params)
#define GSL_DBL_EPSILON 2.2204460492503131e-16
double gsl_pow_int(double x, int n);
#endif
