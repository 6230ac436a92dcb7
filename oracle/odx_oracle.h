/* TEST INFRASTRUCTURE — interface of the CPU restatement (see odx_oracle.c for scope and citations).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may use it. */
#ifndef ODX_ORACLE_H
#define ODX_ORACLE_H
#include <stdint.h>
#include <stdio.h>

enum { ORC_METH_FE = 0, ORC_METH_DDE = 1, ORC_METH_RKF45 = 2, ORC_METH_RKCK = 3, ORC_METH_RK4 = 4, ORC_METH_ITERATION = 5, ORC_METH_RK2 = 6, ORC_METH_RK8PD = 7 };
enum { ORC_EVENT_NONE = 0x00, ORC_EVENT_T0 = 0x01, ORC_EVENT_TFINAL = 0x02, ORC_STOP_TSPAN = 0x04, ORC_STOP_COND = 0x08, ORC_STOP_BIRTH = 0x10, ORC_STOP_DEATH = 0x20,
       ORC_STOP_REPLI = 0x40, ORC_POP_MASK = 0x70, ORC_EVENT_GSL = 0x100 };
/* Philox domain tags, identical to odexp_b200/csrc/odx_rng.h (rng_mode 1 only) */
#define ORC_DOM_RHS  0x52485300u
#define ORC_DOM_IC   0x49430000u
#define ORC_DOM_LEAP 0x4c454150u
#define ORC_DOM_HOST 0x484f5354u

typedef struct { int32_t ring[31]; int f, r; uint64_t ndraws; } orc_hostrand;

typedef struct orc_par {
  double *y, *expr, *aux, *psi;
  double death, repli;
  int id, dead;
  struct orc_par *sister;
} orc_par;

struct orc_sim;
typedef struct orc_ctx {
  struct orc_sim *sim;
  const double *mu; double pop_size, h;
  int id, oid, atbirth, atrepli, ist0;
  const orc_par *sister;
  int pos;                                 /* position of the particle in the list (KY) */
  const double *other_aux;                 /* other_->aux (OA() inside %C) */
  const double *my_aux;                    /* myself_->aux while the %C terms are summed (MA() inside %C: pass-1 values) */
  uint32_t kid, kstream, kdom, kslot;      /* keyed-Philox address of the next draw (rng_mode 1) */
} orc_ctx;

typedef struct {
  const char *name;
  int nvar, nexpr, naux, npsi, nmfd, npar;
  const double *par_defaults;
  void (*rhs)(orc_ctx *c, double t, const double *y, const double *e, const double *psi, const double *mfd, double *f, double *aux,
              double *death, double *repli);
  double (*mfd_term)(orc_ctx *c, int k, const double *y, const double *e, const double *aux, const double *mfd);
  double (*birth_term)(orc_ctx *c, const double *y, const double *e, const double *aux);
  void (*cpl_terms)(orc_ctx *c, const double *y, const double *e, const double *oy, const double *oe, double *psi);
  void (*ic)(orc_ctx *c, double t, double *y, double *e);
  /* low-rank %C ... lrwexp {var, fun, U, V, rank}: lr_rank == 0 means none */
  int lr_rank, lr_psi, lr_var, lr_iu, lr_iv;
  double (*lr_fun)(double);
  /* %M ... {vect}: mean field k with mfd_vect[k] != 0 is assigned once (no list walk, src/odexp:886-889) from POP_SIZE, the
   * parameters and the mean fields before it */
  const int *mfd_vect;
  double (*mfd_vect_eval)(int k, const double *mu, double pop_size, const double *mfd);
  /* what lr_var indexes: 0 a dynamical variable, 1 an auxiliary variable (its pass-1 value), 2 a parametric expression --
   * the generated gather calls mvar(name, ...) which searches the three in this order (src/datastruct.c:824-855) */
  int lr_var_kind;
  int lr_unweighted;      /* `%C ... lrexp {var, fun}`: lrkern instead of lrwpkern (lr_rank is set to 1, lr_iu / lr_iv are unused) */
} orc_model;

typedef struct {
  int method, res;
  double hmin, h0, abstol, reltol, ssahmin, aleap;
  int popsize; uint32_t seed;
  int rng_mode;           /* 0: serial libc stream (reference behaviour); 1: keyed Philox (CUDA library convention) */
  int write_traj;
  const char *outdir;
  const double *tspan; int tspan_len;
  const double *mu;
  int plot_x, plot_y, plot_z, particle;      /* opt x / y / z as indices (-1 = T), opt particle: current.plot is always written */
} orc_opts;

typedef struct { double total_ms, solver_ms, function_ms; int64_t steps; double particle_steps; int64_t final_n; int status;
                 uint64_t rand_draws, rhs_evals; } orc_report;

typedef struct orc_sim {
  const orc_model *model; orc_opts o;
  orc_par **p; int n, cap, size;         /* n = array length (with tombstones during a sweep), size = POP_SIZE */
  int max_id;
  double *mu, *mfd, pop_birth_rate, *h;
  int stop_flag, event[3];
  unsigned int nbrsteps;
  FILE *fstats, *ftraj, *fquick; int ngx, ngy, ngz, cp;
  int rng_mode; orc_hostrand hr; uint32_t k0, k1, eval_index, host_draws, event_index;
  uint64_t rhs_evals; int64_t steps; double particle_steps, function_ms;
  int lr_last_p; uint32_t reinit_index;
  int first_eval_flag;                   /* DELAY_FIRST_EVAL = SIM->first_eval_flag: set by dde_apply around its first RHS evaluation */
} orc_sim;
void orc_lowrank_apply(orc_sim *s);

double orc_rand01(orc_ctx *c);
double orc_randN(orc_ctx *c, double mu, double sigma);
double orc_exprand(orc_ctx *c, double r);
long   orc_poisson(orc_ctx *c, double lambda);
double orc_dwdt(orc_ctx *c);
int    orc_stopif(orc_ctx *c, int cond);      /* STOPIF(cond), src/odexp.h:400-406 */
double orc_ky(orc_ctx *c, int var, int degree); /* KY(name, degree), src/odexp.h:334-357 */
/* delay(): one history stack of DELAY_SSTACK samples (src/odexp.h:151-162), src/utils_odexp.c:173-240 */
#define ORC_DELAY_SSTACK 1000
typedef struct { double t[ORC_DELAY_SSTACK]; double array[ORC_DELAY_SSTACK]; int pos; } orc_stack;
double orc_interp(double *x, double *y, double xi, int nx, int start);
double orc_delay(orc_ctx *c, const double t, const double x, const double tau, double (*ic)(double), orc_stack *xhist);

const orc_model *orc_find_model(const char *name);
extern const orc_model *orc_extra_model;
int orc_pop_rhs(double t, const double *y, double *f, void *sim);
int orc_odesolver(const char *model_name, const orc_opts *o, orc_report *rep);
#endif
