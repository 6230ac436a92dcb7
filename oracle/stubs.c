/* TEST INFRASTRUCTURE — link stubs for building the UNCHANGED reference sources headless.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use anything under oracle/.  Nothing here is shipped in the product library.
 *
 * The reference (samubernard/odexp) links against GSL, readline and two BSD libc functions that
 * this image lacks.  Its population path only needs gsl_odeiv2 (oracle/minigsl.c) and gsl_cheb
 * (ditto); everything else below is UI / steady-state tooling that the oracle never executes, so
 * those entry points abort loudly if reached.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_vector.h>
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_multiroots.h>
#include <gsl/gsl_eigen.h>
#include <gsl/gsl_qrng.h>
#include <gsl/gsl_linalg.h>
#include <readline/readline.h>
#include <readline/history.h>

#define UNREACHED(name) do { fprintf(stderr, "oracle stub: %s reached (not on the population path)\n", name); abort(); } while (0)

const char *gsl_strerror(const int e) { static char b[32]; snprintf(b, sizeof b, "gsl status %d", e); return b; }

/* gsl_vector: real, tiny */
gsl_vector *gsl_vector_alloc(const size_t n) {
  gsl_vector *v = malloc(sizeof *v); v->size = n; v->stride = 1; v->data = calloc(n ? n : 1, sizeof(double)); v->block = NULL; v->owner = 1; return v; }
void gsl_vector_free(gsl_vector *v) { if (v) { free(v->data); free(v); } }
double gsl_vector_get(const gsl_vector *v, const size_t i) { return v->data[i * v->stride]; }
void gsl_vector_set(gsl_vector *v, const size_t i, double x) { v->data[i * v->stride] = x; }
int gsl_vector_memcpy(gsl_vector *d, const gsl_vector *s) { for (size_t i = 0; i < s->size; i++) d->data[i * d->stride] = s->data[i * s->stride]; return 0; }
int gsl_vector_fprintf(FILE *f, const gsl_vector *v, const char *fmt) { for (size_t i = 0; i < v->size; i++) { fprintf(f, fmt, v->data[i * v->stride]); fputc('\n', f); } return 0; }
gsl_vector_complex *gsl_vector_complex_alloc(const size_t n) { (void)n; UNREACHED("gsl_vector_complex_alloc"); return NULL; }
void gsl_vector_complex_free(gsl_vector_complex *v) { (void)v; }
gsl_complex gsl_vector_complex_get(const gsl_vector_complex *v, const size_t i) { gsl_complex z = {{0, 0}}; (void)v; (void)i; UNREACHED("gsl_vector_complex_get"); return z; }
gsl_vector_view gsl_vector_complex_real(gsl_vector_complex *v) { gsl_vector_view w; memset(&w, 0, sizeof w); (void)v; UNREACHED("gsl_vector_complex_real"); return w; }
gsl_vector_view gsl_vector_complex_imag(gsl_vector_complex *v) { gsl_vector_view w; memset(&w, 0, sizeof w); (void)v; UNREACHED("gsl_vector_complex_imag"); return w; }
int gsl_vector_complex_fprintf(FILE *f, const gsl_vector_complex *v, const char *fmt) { (void)f; (void)v; (void)fmt; UNREACHED("gsl_vector_complex_fprintf"); return 0; }

gsl_matrix *gsl_matrix_alloc(const size_t a, const size_t b) { (void)a; (void)b; UNREACHED("gsl_matrix_alloc"); return NULL; }
void gsl_matrix_free(gsl_matrix *m) { (void)m; }
void gsl_matrix_set(gsl_matrix *m, const size_t i, const size_t j, const double x) { (void)m; (void)i; (void)j; (void)x; UNREACHED("gsl_matrix_set"); }

static const gsl_multiroot_fsolver_type hybrids_t = {"hybrids-stub"};
const gsl_multiroot_fsolver_type *gsl_multiroot_fsolver_hybrids = &hybrids_t;
gsl_multiroot_fsolver *gsl_multiroot_fsolver_alloc(const gsl_multiroot_fsolver_type *T, size_t n) { (void)T; (void)n; UNREACHED("gsl_multiroot_fsolver_alloc"); return NULL; }
void gsl_multiroot_fsolver_free(gsl_multiroot_fsolver *s) { (void)s; }
int gsl_multiroot_fsolver_set(gsl_multiroot_fsolver *s, gsl_multiroot_function *f, const gsl_vector *x) { (void)s; (void)f; (void)x; UNREACHED("gsl_multiroot_fsolver_set"); return 0; }
int gsl_multiroot_fsolver_iterate(gsl_multiroot_fsolver *s) { (void)s; UNREACHED("gsl_multiroot_fsolver_iterate"); return 0; }
int gsl_multiroot_test_residual(const gsl_vector *f, double e) { (void)f; (void)e; UNREACHED("gsl_multiroot_test_residual"); return 0; }
int gsl_multiroot_test_delta(const gsl_vector *dx, const gsl_vector *x, double a, double r) { (void)dx; (void)x; (void)a; (void)r; UNREACHED("gsl_multiroot_test_delta"); return 0; }

gsl_eigen_nonsymm_workspace *gsl_eigen_nonsymm_alloc(const size_t n) { (void)n; UNREACHED("gsl_eigen_nonsymm_alloc"); return NULL; }
int gsl_eigen_nonsymm(gsl_matrix *A, gsl_vector_complex *e, gsl_eigen_nonsymm_workspace *w) { (void)A; (void)e; (void)w; UNREACHED("gsl_eigen_nonsymm"); return 0; }

static const gsl_qrng_type sobol_t = {"sobol-stub"};
const gsl_qrng_type *gsl_qrng_sobol = &sobol_t;
gsl_qrng *gsl_qrng_alloc(const gsl_qrng_type *T, unsigned int d) { (void)T; (void)d; UNREACHED("gsl_qrng_alloc"); return NULL; }
void gsl_qrng_free(gsl_qrng *q) { (void)q; }
int gsl_qrng_get(const gsl_qrng *q, double x[]) { (void)q; (void)x; UNREACHED("gsl_qrng_get"); return 0; }

gsl_permutation *gsl_permutation_alloc(const size_t n) { (void)n; UNREACHED("gsl_permutation_alloc"); return NULL; }
void gsl_permutation_free(gsl_permutation *p) { (void)p; }
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *s) { (void)A; (void)p; (void)s; UNREACHED("gsl_linalg_LU_decomp"); return 0; }
int gsl_linalg_LU_solve(const gsl_matrix *LU, const gsl_permutation *p, const gsl_vector *b, gsl_vector *x) { (void)LU; (void)p; (void)b; (void)x; UNREACHED("gsl_linalg_LU_solve"); return 0; }

/* readline / history: headless no-ops */
const char *rl_library_version = "stub";
const char *rl_readline_name = "odexp";
int rl_attempted_completion_over = 0;
rl_completion_func_t *rl_attempted_completion_function = NULL;
char *readline(const char *prompt) { (void)prompt; return NULL; }
int rl_read_init_file(const char *f) { (void)f; return 0; }
char **rl_completion_matches(const char *t, rl_compentry_func_t *f) { (void)t; (void)f; return NULL; }
void add_history(const char *l) { (void)l; }
void using_history(void) {}
int read_history(const char *f) { (void)f; return 0; }
int write_history(const char *f) { (void)f; return 0; }
void stifle_history(int m) { (void)m; }

/* BSD libc names used by the (macOS-targeted) reference */
int mergesort(void *base, size_t nel, size_t width, int (*compar)(const void *, const void *)) { qsort(base, nel, width, compar); return 0; }
int isnumber(int c) { return isdigit(c); }
