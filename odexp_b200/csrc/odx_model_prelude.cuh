/* odx_model_prelude.cuh — helpers visible to generated model code (host + device). */
#pragma once
#include <math.h>
#include "odx_rng.h"

/* pow(x, c) with literal exponents (emitted by the generator's peephole; <= 1 ulp from pow) */
ODX_HD double odx_pow05(double x) { return sqrt(x); }
ODX_HD double odx_pow15(double x) { return x * sqrt(x); }
ODX_HD double odx_pow2(double x) { return x * x; }
ODX_HD double odx_pow3(double x) { return x * x * x; }

/* STOPIF(cond): raise the stop flag, value is the condition (src/odexp.h:400-406) */
template <class C> ODX_HD int odx_stopif(C &c, int cond) { if (cond) c.stop = 1; return cond; }

/* device twins of the pure helpers a model expression may call (src/utils_odexp.c:11-107) */
ODX_HD double sum(const double *array, long len) { double s = 0.0; for (long i = 0; i < len; i++) s += array[i]; return s; }
ODX_HD double sumsub(const double *array, long *ind, long len) { double s = 0.0; for (long i = 0; i < len; i++) s += array[ind[i]]; return s; }
ODX_HD double sumstep(const double *array, long len, long step) { double s = 0.0; for (long i = 0; i < len; i += step) s += array[i]; return s; }
ODX_HD double prod(const double *array, long len) { double s = 1.0; for (long i = 0; i < len; i++) s *= array[i]; return s; }
ODX_HD double dotprod(const double *x, const double *y, long len) { double s = 0.0; for (long i = 0; i < len; i++) s += x[i] * y[i]; return s; }
ODX_HD double minus(double x, double y) { return x - y; }
ODX_HD double plus(double x, double y) { return x + y; }
ODX_HD double identity(double x) { return x; }
