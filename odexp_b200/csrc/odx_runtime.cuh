/* odx_runtime.cuh — host side of the engine: device buffers, launch sequencing, and the extern "C"
 * entry points declared in include/odx.h.  Included by odx_engine.cuh (one instantiation per model). */
#pragma once
#include <dlfcn.h>
#ifdef ODX_MODEL_HAS_LOWRANK
#include "odx_lowrank.cuh"
#endif
#if __has_include(<nccl.h>)
#include <nccl.h>
#define ODX_HAVE_NCCL_H 1
#endif

namespace odx {

enum { SLOT_Y0 = 0, SLOT_Y1 = 1, SLOT_STAGE = 2, SLOT_RATES = 10, NSLOT = 11 };

#ifdef ODX_HAVE_NCCL_H
/* NCCL is bound at run time (dlopen) so that the model library has no link-time dependency and shares the
 * libnccl.so.2 already loaded by the host process (e.g. the one bundled with torch). */
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    Send = (decltype(Send))dlsym(lib, "ncclSend"); Recv = (decltype(Recv))dlsym(lib, "ncclRecv");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart"); GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString && Send && Recv && GroupStart && GroupEnd;
  }
};
static NcclApi g_nccl;
#endif

}  // namespace odx

using namespace odx;

struct odx_event_state;
static void odx_events_free(odx_handle *h);
static void odx_output_free(odx_handle *h);

struct odx_handle {
  odx_event_state *ev_state = nullptr;       /* birth / death scratch (odx_events.cuh), allocated on first use */
  void *out = nullptr;                       /* output path state (odx_output.cuh), allocated on first use */
  odx_handle *replica = nullptr;             /* multi-GPU birth/death: whole ordered population, rebuilt per sweep */
  void *gcol = nullptr; int64_t gcol_cap = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t cap = 0, n = 0;
  int32_t max_id = 0;
  double *Y = nullptr, *YS = nullptr, *ACC = nullptr, *E = nullptr, *AUX = nullptr, *PSI = nullptr, *DEATH = nullptr,
         *REPLI = nullptr, *KS = nullptr, *YALT = nullptr, *K1 = nullptr, *DOUT = nullptr, *YONE = nullptr, *YH = nullptr, *K1B = nullptr;
  int32_t *ID = nullptr;
  double *d_mu = nullptr;
  double h_mu[A1(NPAR)];
  double *partial = nullptr; unsigned int *counter = nullptr; double *sums = nullptr; unsigned long long *rmax = nullptr;
  double *h_pin = nullptr;                 /* pinned scratch: NSLOT*NRED doubles + 16 */
  uint32_t reinit_index = 0;               /* resets at stopping times so far (Philox stream of the re-drawn ICs) */
  double *stage_dev = nullptr; size_t stage_bytes = 0;   /* AoS staging for upload / download */
  double *gather = nullptr; int32_t *gather_id = nullptr; int64_t gather_cap = 0;   /* all-gathered j-particles (%C, multi-GPU) */
  int leap_fused = 1;                                                                /* tau-leaping sweep as one cooperative launch (single GPU) */
  double *pair_sym_part = nullptr; size_t pair_sym_len = 0; int *pair_poison = nullptr; int pair_sym = -1;     /* k_pair_sym: [tile][other tile][npsi][PSYM_TILE], flag, option */
  double *pair_part = nullptr; size_t pair_part_len = 0;                             /* [segment][npsi][cap] raw partial sums of k_pair */
  int64_t all_n[MAX_RANKS] = {0, 0, 0, 0, 0, 0, 0, 0};                               /* particles per rank */
  int sm_count = 148, grid_cap = 148 * 8; int ks_slots = 0;
  uint32_t k0 = 3141592u, k1 = 0u;
  uint32_t eval_index = 0;
  int y_slot = SLOT_Y0; bool ymoments_valid = false;
  int report_slot = SLOT_Y0, rates_slot = SLOT_RATES;
  double eps_abs = 1e-6, eps_rel = 1e-6;
  uint64_t ev_count = 0, ev_failed = 0, rhs_evals = 0; bool k1_valid = false;
  int rank = 0, nranks = 1; double n_global = 0.0; int64_t n_before = 0;   /* particles on lower ranks */
#ifdef ODX_HAVE_NCCL_H
  ncclComm_t comm = nullptr;
#endif
  /* persistent step kernel + peer-memory exchange */
  unsigned long long *ll_words = nullptr;     /* LL exchange words of the persistent kernel (zero-initialised) */
  int use_persistent = 1, coop_ok = 0, coop_blocks = 0, sweep_dir = 0, use_recon = 1;
  int coop_blocks_tma = 0, tma_mode = -1;     /* tma_mode: -1 choose by size, 0 never, 1 always (option "tma") */
  int shard_events = 1;                       /* multi-GPU tau-leaping: 1 sharded sweep + redistribution, 0 replicate the population (option "shard_events") */
  int ssa_mode = -1;                          /* SSA event pick: -1 device for keyed runs / host for replayed ones, 0 host, 1 device (option "ssa_device") */
  int lean_mode = -1;                         /* lean RK4 scheme: -1 with the TMA ring (HBM-resident populations), 0 never, 1 whenever eligible (option "lean") */
  double l2_bytes = 0.0;
  uint32_t epoch = 0;
  void *mail = nullptr;                       /* [2][MAX_RANKS][NRED] doubles | [2][MAX_RANKS] flags | abort int (IPC-exported) */
  void *peer_mail[MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool peers_attached = false;
  /* low-rank coupling scratch + options (opt lrmax / lrabstol / lrreltol) */
  void *lr_plan = nullptr, *lr_rank = nullptr; double *lr_stats = nullptr, *lr_spart = nullptr, *lr_apart = nullptr, *lr_A = nullptr, *lr_x = nullptr, *lr_ones = nullptr; bool lr_attr_done = false;
  int lr_pmax = 31; double lr_abstol = 1e-6, lr_reltol = 1e-6;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t launches = 0;
  volatile int abort_req = 0;                 /* odx_request_abort (SIGINT) */
  /* delay(): history stacks per call site, EVENT_TYPE and *SIM->h as the model code sees them (odx_set_event_type) */
  double *dl_T[A1(M::NDELAY)] = {}, *dl_X[A1(M::NDELAY)] = {}; int *dl_pos[A1(M::NDELAY)] = {};
  int event_type = 0; double event_h = 0.0;
  cudaEvent_t chunk_ev[2] = {nullptr, nullptr};
  HostRand hr; bool hr_seeded = false; uint32_t host_draws = 0;   /* serial stream of the host loop (replay mode) */
  std::string err;
};

namespace odx {

#define ODX_CUDA(h, call)                                                                            \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      char b_[512];                                                                                  \
      snprintf(b_, sizeof b_, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
      (h)->err = b_;                                                                                 \
      return ODX_ECUDA;                                                                              \
    }                                                                                                \
  } while (0)
#define ODX_TRY(expr) do { int s_ = (expr); if (s_ != ODX_SUCCESS) return s_; } while (0)

static inline Pop pop_view(const odx_handle *h) {
  Pop p;
  p.Y = h->Y; p.YS = h->YS; p.YS_OUT = h->YS; p.Y_OUT = h->Y; p.ACC = h->ACC; p.E = h->E; p.AUX = h->AUX; p.PSI = h->PSI; p.DEATH = h->DEATH; p.REPLI = h->REPLI;
  p.KS = h->KS; p.ID = h->ID; p.stride = h->cap; p.n = h->n; p.n_global = h->n_global; p.mu = h->d_mu; p.k0 = h->k0; p.k1 = h->k1; philox_round_keys(h->k0, h->k1, p.rk);
  p.stop = reinterpret_cast<int *>(h->counter + 8);
  for (int k = 0; k < MUV_N; k++) p.muv[k] = (k < NPAR) ? h->h_mu[k] : 0.0;
  for (int k = 0; k < A1(M::NDELAY); k++) { p.dl[k].T = h->dl_T[k]; p.dl[k].X = h->dl_X[k]; p.dl[k].pos = h->dl_pos[k]; p.dl[k].K = M::delay_count(k); }
  return p;
}
/* delay(): what the call sites do besides looking up, before the RHS evaluation at (t, ysrc): refill while an event flag is set,
 * push at the first evaluation of a dde step (see k_delay_fill / k_delay_push) */
static int launch_check(odx_handle *h);
static int delay_before_eval(odx_handle *h, const double *ysrc, double t, double hstep, bool first_eval) {
  if (M::NDELAY == 0 || h->n <= 0) return ODX_SUCCESS;
  if (h->nranks > 1) { h->err = "delay() history stacks are not supported on a sharded population"; return ODX_EUNSUP; }
  if (h->event_type != 0) {
    int64_t g = (h->n * DELAY_SSTACK + BLOCK - 1) / BLOCK; if (g > 8 * h->grid_cap) g = 8 * h->grid_cap;
    k_delay_fill<<<(unsigned)g, BLOCK, 0, h->stream>>>(pop_view(h), ysrc, t, hstep);
    return launch_check(h);
  }
  if (first_eval) {
    int64_t g = (h->n + BLOCK - 1) / BLOCK; if (g > h->grid_cap) g = h->grid_cap;
    k_delay_push<<<(unsigned)g, BLOCK, 0, h->stream>>>(pop_view(h), ysrc, t);
    return launch_check(h);
  }
  return ODX_SUCCESS;
}
static inline Red red_view(const odx_handle *h, int slot) {
  Red r; r.partial = h->partial; r.counter = h->counter; r.sums = h->sums + (size_t)slot * NRED; return r;
}
static inline int grid_for(const odx_handle *h, int64_t n) {
  int64_t g = (n + BLOCK - 1) / BLOCK;
  if (g < 1) g = 1;
  if (g > h->grid_cap) g = h->grid_cap;
  return (int)g;
}
static inline int grid_chain(const odx_handle *h, int64_t n) {      /* two particles per thread */
  int64_t g = (n + 2 * BLOCK - 1) / (2 * BLOCK);
  if (g < 1) g = 1;
  if (g > h->grid_cap) g = h->grid_cap;
  return (int)g;
}
constexpr bool NEEDS_MOMENTS = (NMFD > 0) || M::HAS_BIRTH;

/* cross-GPU sum of the raw %M sums (and birth / rate sums): allreduce(SUM) of NRED doubles per RHS evaluation */
static void fill_peer_args(odx_handle *h, struct odx::PersistArgs &g);
__global__ void k_peer_allreduce(double *sums, int rank, int nranks, uint32_t epoch, double *mailbox, uint32_t *mailflag, PersistArgs g);
static int post_reduce(odx_handle *h, int slot) {
  if (h->nranks <= 1) return ODX_SUCCESS;
  if (h->peers_attached) {
    PersistArgs g; memset(&g, 0, sizeof g);
    fill_peer_args(h, g);
    k_peer_allreduce<<<1, 64, 0, h->stream>>>(h->sums + (size_t)slot * NRED, h->rank, h->nranks, h->epoch, g.mailbox, g.mailflag, g);
    h->launches++;
    ODX_CUDA(h, cudaGetLastError());
    h->epoch += 1;
    return ODX_SUCCESS;
  }
#ifdef ODX_HAVE_NCCL_H
  double *s = h->sums + (size_t)slot * NRED;
  ncclResult_t r = g_nccl.AllReduce(s, s, NRED, ncclDouble, ncclSum, h->comm, h->stream);
  if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
  return ODX_SUCCESS;
#else
  h->err = "built without nccl.h"; return ODX_ENCCL;
#endif
}

static int launch_check(odx_handle *h) {
  h->launches++;
  ODX_CUDA(h, cudaGetLastError());
  return ODX_SUCCESS;
}

#ifdef ODX_MODEL_HAS_LOWRANK
/* the expansion variable of `%C ... lrwexp {var = <auxiliary variable>}`: the generated gather reads myself_->aux after pass 1 of
 * pop_rhs (mvar(), src/odexp:1578, src/datastruct.c:840-850), i.e. ode_rhs of the stage state with zeroed mean fields and the psi
 * of the previous evaluation */
__global__ void __launch_bounds__(BLOCK) k_lr_x_aux(Pop p, const double *__restrict__ ysrc, double t, double *__restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * BLOCK) {
    double y[A1(NVAR)], e[A1(NEXPR)], psi[A1(NPSI)], f[A1(NVAR)], aux[A1(NAUX)], mz[A1(NMFD)], d = 0.0, r = 0.0;
    load_cols<NVAR>(ysrc, p.stride, i, y);
    load_cols<NEXPR>(p.E, p.stride, i, e);
    load_cols<NPSI>(p.PSI, p.stride, i, psi);
#pragma unroll
    for (int k = 0; k < A1(NMFD); k++) mz[k] = 0.0;
    DevCtx c = make_ctx(p, 0.0, 0, DOM_RHS, M::NEEDS_ID ? p.ID[i] : 0);
    M::template rhs<true>(c, t, y, e, psi, mz, f, aux, d, r);
    x[i] = aux[M::LR0_VAR_KIND == 1 ? M::LR0_VAR : 0];
  }
}

/* lrwpkern on the device (see odx_lowrank.cuh); writes PSI[LR0_PSI] */
static int compute_lowrank(odx_handle *h, const double *ysrc, double t) {
  const int64_t n = h->n;
  if (n <= 0 && h->nranks <= 1) return ODX_SUCCESS;
#ifdef ODX_HAVE_NCCL_H
  if (h->nranks > 1 && !h->comm) { h->err = "multi-GPU low-rank %C needs odx_comm_init"; return ODX_ENCCL; }
#endif
  int nblk = (int)((n + LR_MTILE - 1) / LR_MTILE);
  if (nblk > 8 * h->sm_count) nblk = 8 * h->sm_count;
  if (nblk < 1) nblk = 1;
  if (!h->lr_plan) {
    ODX_CUDA(h, cudaMalloc(&h->lr_plan, sizeof(LrPlan)));
    ODX_CUDA(h, cudaMalloc(&h->lr_rank, sizeof(LrRank)));
    ODX_CUDA(h, cudaMemsetAsync(h->lr_rank, 0, sizeof(LrRank), h->stream));
    ODX_CUDA(h, cudaMalloc((void **)&h->lr_stats, sizeof(double) * 8));
    ODX_CUDA(h, cudaMalloc((void **)&h->lr_spart, sizeof(double) * 3 * (h->grid_cap + 1)));
    ODX_CUDA(h, cudaMalloc((void **)&h->lr_apart, sizeof(double) * (size_t)LR_R * LR_P1 * (24 * h->sm_count + 24 * h->sm_count / 32 + 2)));
    ODX_CUDA(h, cudaMalloc((void **)&h->lr_A, sizeof(double) * LR_R * LR_P1));
  }
  /* weights: the U / V columns of E, or one column of ones for the unweighted `lrexp` (same arithmetic as lrexp(): 1.0 * x) */
  if (M::LR0_UNWEIGHTED && !h->lr_ones) {
    std::vector<double> ones((size_t)h->cap, 1.0);
    ODX_CUDA(h, cudaMalloc((void **)&h->lr_ones, sizeof(double) * h->cap));
    ODX_CUDA(h, cudaMemcpy(h->lr_ones, ones.data(), sizeof(double) * h->cap, cudaMemcpyHostToDevice));
  }
  const double *lrU = M::LR0_UNWEIGHTED ? h->lr_ones : h->E + (size_t)(M::LR0_U < 0 ? 0 : M::LR0_U) * h->cap;
  const double *lrV = M::LR0_UNWEIGHTED ? h->lr_ones : h->E + (size_t)(M::LR0_V < 0 ? 0 : M::LR0_V) * h->cap;
  const double *x = ysrc + (size_t)M::LR0_VAR * h->cap;
  if (M::LR0_VAR_KIND == 2) x = h->E + (size_t)M::LR0_VAR * h->cap;
  if (M::LR0_VAR_KIND == 1) {
    if (!h->lr_x) ODX_CUDA(h, cudaMalloc((void **)&h->lr_x, sizeof(double) * h->cap));
    if (n > 0) { k_lr_x_aux<<<grid_for(h, n), BLOCK, 0, h->stream>>>(pop_view(h), ysrc, t, h->lr_x); ODX_TRY(launch_check(h)); }
    x = h->lr_x;
  }
  k_lr_stats<<<grid_for(h, n), BLOCK, 0, h->stream>>>(x, n, h->lr_spart, h->counter + 1, h->lr_stats);
  ODX_TRY(launch_check(h));
#ifdef ODX_HAVE_NCCL_H
  if (h->nranks > 1) {      /* mean / range over all GPUs: SUM of the sums, MAX of {max, -min} */
    ncclResult_t r = g_nccl.AllReduce(h->lr_stats + 2, h->lr_stats + 2, 1, ncclDouble, ncclSum, h->comm, h->stream);
    if (r == ncclSuccess) { double *mm = h->lr_stats + 1; (void)mm; }
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
    /* stats layout [min, max, sum, -min]: reduce {max} and {-min} with MAX (two one-element calls keep the layout simple) */
    r = g_nccl.AllReduce(h->lr_stats + 1, h->lr_stats + 1, 1, ncclDouble, ncclMax, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
    r = g_nccl.AllReduce(h->lr_stats + 3, h->lr_stats + 3, 1, ncclDouble, ncclMax, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
  }
#endif
  LrRank *rk = (LrRank *)h->lr_rank;
  k_lr_rank_plan<<<2 * h->sm_count, BLOCK, 0, h->stream>>>(h->lr_stats, h->n_global, h->lr_pmax, h->lr_abstol, h->lr_reltol, rk, h->counter + 2,
                                                          (LrPlan *)h->lr_plan);
  ODX_TRY(launch_check(h));
  int npart = nblk;
  if (LR_WSPLIT) {
    const int64_t ntiles = (n + LRT_PT - 1) / LRT_PT;
    nblk = ntiles < h->sm_count ? (int)ntiles : h->sm_count;     /* persistent: one CTA per SM, batches strided over the grid */
    if (nblk < 1) nblk = 1;
    npart = nblk;
    const double *Vc = lrV; const LrPlan *plm = (const LrPlan *)h->lr_plan;
    ODX_CUDA(h, cudaMemsetAsync(h->lr_apart, 0, sizeof(double) * (size_t)LR_R * LR_P1 * npart, h->stream));
    if (!h->lr_attr_done) {                       /* the rings need more than the 48 KB static limit */
      ODX_CUDA(h, cudaFuncSetAttribute(k_lr_moments_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lrt_smem_bytes(LR_P1)));
      ODX_CUDA(h, cudaFuncSetAttribute(k_lr_apply_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lra_smem_bytes(LR_P1)));
      h->lr_attr_done = true;
    }
    k_lr_moments_tma<<<nblk, LRT_THREADS, lrt_smem_bytes(LR_P1), h->stream>>>(x, Vc, h->cap, n, plm, h->lr_apart); ODX_TRY(launch_check(h));
  } else {
    k_lr_moments<<<nblk, BLOCK, 0, h->stream>>>(x, lrV, h->cap, n, (const LrPlan *)h->lr_plan, h->lr_apart);
    ODX_TRY(launch_check(h));
  }
  {
    const int qb = (LR_R * LR_P1 + BLOCK - 1) / BLOCK, ng = (npart + 31) / 32;
    double *stage2 = h->lr_apart + (size_t)LR_R * LR_P1 * (24 * h->sm_count);      /* tail of the partial buffer */
    k_lr_moments_sum<<<dim3(qb, ng), BLOCK, 0, h->stream>>>(h->lr_apart, npart, 32, stage2);      /* one pass over all partials was slower */
    ODX_TRY(launch_check(h));
    k_lr_moments_sum<<<dim3(qb, 1), BLOCK, 0, h->stream>>>(stage2, ng, ng, h->lr_A);
    ODX_TRY(launch_check(h));
  }
#ifdef ODX_HAVE_NCCL_H
  if (h->nranks > 1) {      /* moments are sums over particles: allreduce(SUM) of r x (p+1) doubles */
    ncclResult_t r = g_nccl.AllReduce(h->lr_A, h->lr_A, (size_t)LR_R * LR_P1, ncclDouble, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
  }
#endif
  /* a CTA first stages A (r x (p+1) doubles) and the coefficient table in shared memory: with one chunk of 128 particles per CTA
   * that prologue cost as much as the chunk (ncu: long-scoreboard stalls, FP64 pipe 32 %); six resident CTAs per SM loop instead */
  int ga = (int)((n + 127) / 128); if (ga > 6 * h->sm_count) ga = 6 * h->sm_count; if (ga < 1) ga = 1;
  const double *Uc = lrU; double *psi = h->PSI + (size_t)M::LR0_PSI * h->cap;
  const LrPlan *pl = (const LrPlan *)h->lr_plan;
  if (LR_WSPLIT) {
    const int64_t nbatch = (n + LRA_PT - 1) / LRA_PT;
    const int gt = nbatch < h->sm_count ? (nbatch < 1 ? 1 : (int)nbatch) : h->sm_count;
    k_lr_apply_tma<<<gt, LRA_THREADS, lra_smem_bytes(LR_P1), h->stream>>>(x, Uc, h->cap, n, pl, h->lr_A, psi);
    return launch_check(h);
  }
  k_lr_apply<8><<<ga, 128, 0, h->stream>>>(x, Uc, h->cap, n, pl, h->lr_A, psi); ODX_TRY(launch_check(h));
  k_lr_apply<16><<<ga, 128, 0, h->stream>>>(x, Uc, h->cap, n, pl, h->lr_A, psi); ODX_TRY(launch_check(h));
  k_lr_apply<24><<<ga, 128, 0, h->stream>>>(x, Uc, h->cap, n, pl, h->lr_A, psi); ODX_TRY(launch_check(h));
  k_lr_apply<32><<<ga, 128, 0, h->stream>>>(x, Uc, h->cap, n, pl, h->lr_A, psi);
  return launch_check(h);
}
#endif

/* psi for the state in ysrc (all-pairs %C).  Multi-GPU: allgather of the coupled columns first. */
static int compute_psi(odx_handle *h, const double *ysrc, double t) {
#ifdef ODX_MODEL_HAS_LOWRANK
  ODX_TRY(compute_lowrank(h, ysrc, t));
#endif
  if (!M::HAS_PAIR) return ODX_SUCCESS;
  if (h->n <= 0 && h->nranks <= 1) return ODX_SUCCESS;      /* an empty shard still takes part in the collectives below */
  Pop p = pop_view(h);
  const int itiles = h->n > 0 ? (int)((h->n + PAIR_ITILE - 1) / PAIR_ITILE) : 1;
  /* split the j-range of every source segment until the grid is PAIR_WAVES waves of resident CTAs deep: with ~1.5 waves (the
   * first version at N = 2^17) the half-empty last wave idles a quarter of the GPU */
  static int resident = 0;
  if (!resident) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_pair, PAIR_THREADS, 0) != cudaSuccess || resident < 1) resident = 8;
  }
  int js = (PAIR_WAVES * resident * h->sm_count + itiles - 1) / itiles;
  const int64_t jtiles = (h->n + PAIR_TJ - 1) / PAIR_TJ;
  if (js > jtiles) js = (int)jtiles;
  if (js < 1) js = 1;
  if (js > 64) js = 64;
  const int nsrc = h->nranks > 1 ? h->nranks : 1;
  const size_t need = (size_t)nsrc * js * A1(NPSI) * h->cap;
  if (h->pair_part_len < need) {
    if (h->pair_part) cudaFree(h->pair_part);
    h->pair_part = nullptr; h->pair_part_len = 0;
    ODX_CUDA(h, cudaMalloc((void **)&h->pair_part, sizeof(double) * need));
    h->pair_part_len = need;
  }
  const int *only_if = nullptr;
  int64_t nmax = 0;
  const int ncol = PAIR_NCOL;
  if (h->nranks > 1) {
#ifdef ODX_HAVE_NCCL_H
    /* allgather of the coupled columns (nvar[+nexpr] x N doubles per RHS evaluation) into [rank][col][nmax] + ids */
    if (!h->comm) { h->err = "multi-GPU %C needs odx_comm_init (NCCL allgather)"; return ODX_ENCCL; }
    for (int r = 0; r < h->nranks; r++) nmax = h->all_n[r] > nmax ? h->all_n[r] : nmax;
    if (h->gather_cap < nmax) {
      if (h->gather) cudaFree(h->gather);
      if (h->gather_id) cudaFree(h->gather_id);
      h->gather = nullptr; h->gather_id = nullptr; h->gather_cap = 0;
      ODX_CUDA(h, cudaMalloc((void **)&h->gather, sizeof(double) * (size_t)(h->nranks + 1) * ncol * nmax));
      ODX_CUDA(h, cudaMalloc((void **)&h->gather_id, sizeof(int32_t) * (size_t)(h->nranks + 1) * nmax));
      h->gather_cap = nmax;
    }
    double *sendbuf = h->gather + (size_t)h->nranks * ncol * nmax;
    int32_t *sendid = h->gather_id + (size_t)h->nranks * nmax;
    for (int c = 0; c < NVAR; c++)
      ODX_CUDA(h, cudaMemcpyAsync(sendbuf + (size_t)c * nmax, ysrc + (size_t)c * h->cap, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (M::CPL_USES_EXPR)
      for (int c = 0; c < NEXPR; c++)
        ODX_CUDA(h, cudaMemcpyAsync(sendbuf + (size_t)(NVAR + c) * nmax, h->E + (size_t)c * h->cap, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    ODX_CUDA(h, cudaMemcpyAsync(sendid, h->ID, h->n * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    ncclResult_t r = g_nccl.AllGather(sendbuf, h->gather, (size_t)ncol * nmax, ncclDouble, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
    r = g_nccl.AllGather(sendid, h->gather_id, (size_t)nmax, ncclInt32, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
#else
    h->err = "built without nccl.h"; return ODX_ENCCL;
#endif
  }
#ifdef ODX_MODEL_CPL_SYM
  /* antisymmetric couplings: every unordered pair once (k_pair_sym), then the ordered kernels below as a pass that returns at once
   * unless the symmetric one met a non-finite term.  pair_sym: -1 automatic (populations of at least 16 tiles whose partial
   * blocks fit in 8 GiB), 0 never, 1 always.  Sharded populations: every rank takes the same decision from the rank sizes. */
  {
    PsymArgs g; memset(&g, 0, sizeof g);
    const int nrk = h->nranks > 1 ? h->nranks : 1;
    g.rank = h->nranks > 1 ? h->rank : 0; g.nranks = nrk;
    for (int r = 0; r < nrk; r++) {
      g.all_n[r] = h->nranks > 1 ? h->all_n[r] : h->n;
      g.tile_base[r + 1] = g.tile_base[r] + (int)((g.all_n[r] + PSYM_TILE - 1) / PSYM_TILE);
    }
    g.nslots = g.tile_base[nrk];
    const int nown = g.tile_base[g.rank + 1] - g.tile_base[g.rank];
    const size_t TS = (size_t)A1(NPSI) * PSYM_TILE;
    int maxown = 0;
    for (int r = 0; r < nrk; r++) maxown = std::max(maxown, g.tile_base[r + 1] - g.tile_base[r]);
    /* [own][slot] blocks + (sharded) [slot][own] blocks for the peers, their sums per slot, and what the peers send */
    const size_t sym_need = TS * ((size_t)maxown * g.nslots * (nrk > 1 ? 2 : 1) + (nrk > 1 ? (size_t)g.nslots + (size_t)nrk * maxown : 0));
    const bool use_sym = h->pair_sym != 0 && (h->pair_sym > 0 || (g.nslots >= 16 && sym_need * sizeof(double) <= ((size_t)8 << 30)));
    if (use_sym) {
      if (h->pair_sym_len < sym_need) {
        if (h->pair_sym_part) cudaFree(h->pair_sym_part);
        h->pair_sym_part = nullptr; h->pair_sym_len = 0;
        ODX_CUDA(h, cudaMalloc((void **)&h->pair_sym_part, sizeof(double) * sym_need));
        h->pair_sym_len = sym_need;
      }
      if (!h->pair_poison) ODX_CUDA(h, cudaMalloc((void **)&h->pair_poison, sizeof(int)));
      static bool attr_set = false;
      if (!attr_set) {
        ODX_CUDA(h, cudaFuncSetAttribute(k_pair_sym, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PSYM_SMEM));
        attr_set = true;
      }
      g.part = h->pair_sym_part;
      g.part_remote = g.part + TS * (size_t)maxown * g.nslots;
      double *sendsum = g.part_remote + TS * (size_t)maxown * g.nslots, *recv = sendsum + TS * (size_t)g.nslots;
      if (nrk > 1) { g.jsrc = h->gather; g.jcol_stride = nmax; g.jrank_stride = (int64_t)ncol * nmax; }
      else { g.jsrc = ysrc; g.jcol_stride = h->cap; g.jrank_stride = 0; recv = nullptr; }
      ODX_CUDA(h, cudaMemsetAsync(h->pair_poison, 0, sizeof(int), h->stream));
      if (nown > 0) {
        k_pair_sym<<<dim3((unsigned)g.nslots, (unsigned)nown), PSYM_THREADS, PSYM_SMEM, h->stream>>>(p, ysrc, t, g);
        ODX_TRY(launch_check(h));
      }
#ifdef ODX_HAVE_NCCL_H
      if (nrk > 1) {
        if (nown > 0) { k_pair_sym_remote_sum<<<g.nslots, BLOCK, 0, h->stream>>>(g, nown, sendsum); ODX_TRY(launch_check(h)); }
        /* the sums for the peers' tiles go to their owners, theirs for the own tiles come in: nown x NPSI x PSYM_TILE doubles per peer */
        ncclResult_t r = g_nccl.GroupStart();
        for (int q = 0; q < nrk && r == ncclSuccess; q++) {
          if (q == g.rank) continue;
          const int nq = g.tile_base[q + 1] - g.tile_base[q];
          if (nq > 0 && nown > 0) {
            r = g_nccl.Send(sendsum + TS * (size_t)g.tile_base[q], TS * (size_t)nq, ncclDouble, q, h->comm, h->stream);
            if (r == ncclSuccess) r = g_nccl.Recv(recv + TS * (size_t)q * nown, TS * (size_t)nown, ncclDouble, q, h->comm, h->stream);
          }
        }
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) { h->err = std::string("ncclSend / ncclRecv (%C sums): ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2); return ODX_ENCCL; }
      }
#endif
      if (h->n > 0) {
        k_pair_sym_finish<<<grid_for(h, h->n), BLOCK, 0, h->stream>>>(p, ysrc, t, g, nown, recv, h->PSI, h->pair_poison);
        ODX_TRY(launch_check(h));
      }
      only_if = h->pair_poison;
    }
  }
#endif
  if (h->nranks <= 1) {
    k_pair<<<dim3(itiles, js), PAIR_THREADS, 0, h->stream>>>(p, ysrc, ysrc, h->E, h->ID, h->cap, h->n, t, h->pair_part, 0, only_if);
    ODX_TRY(launch_check(h));
  } else {
    /* one launch per source rank in rank order = global list order */
    for (int src = 0; src < h->nranks && h->n > 0; src++) {
      const double *gy = h->gather + (size_t)src * ncol * nmax;
      k_pair<<<dim3(itiles, js), PAIR_THREADS, 0, h->stream>>>(p, ysrc, gy, gy + (size_t)NVAR * nmax, h->gather_id + (size_t)src * nmax, nmax,
                                                               h->all_n[src], t, h->pair_part, src * js, only_if);
      ODX_TRY(launch_check(h));
    }
  }
  if (h->n <= 0) return ODX_SUCCESS;
  k_pair_finish<<<grid_for(h, h->n), BLOCK, 0, h->stream>>>(p, h->pair_part, nsrc * js, h->PSI, only_if);
  return launch_check(h);
}

static int ensure_moments_Y(odx_handle *h, double t) {
  if (!NEEDS_MOMENTS || h->ymoments_valid) return ODX_SUCCESS;
  k_moments<<<grid_for(h, h->n), BLOCK, 0, h->stream>>>(pop_view(h), h->Y, t, red_view(h, h->y_slot));
  ODX_TRY(launch_check(h));
  ODX_TRY(post_reduce(h, h->y_slot));
  h->ymoments_valid = true;
  return ODX_SUCCESS;
}
static inline const double *sums_ptr(const odx_handle *h, int slot) { return NEEDS_MOMENTS ? h->sums + (size_t)slot * NRED : nullptr; }

/* generated pop_rhs once, with outputs (reference call sites methods_odexp.c:364, 687, 709) */
static int eval_rhs_out(odx_handle *h, const double *ysrc, int sums_slot, double t, double hcur, double *fout) {
  ODX_TRY(delay_before_eval(h, ysrc, t, hcur != 0.0 ? hcur : h->event_h, false));
  RhsArgs a; a.sums_in = sums_ptr(h, sums_slot); a.t = t; a.h = hcur; a.eval_index = h->eval_index; a.fout = fout;
  h->eval_index += 2; h->rhs_evals++;
  k_rhs_out<<<grid_for(h, h->n), BLOCK, 0, h->stream>>>(pop_view(h), ysrc, a, red_view(h, SLOT_RATES));
  ODX_TRY(launch_check(h));
  if (M::HAS_DEATH || M::HAS_REPLI) ODX_TRY(post_reduce(h, SLOT_RATES));
  h->report_slot = sums_slot; h->rates_slot = SLOT_RATES;
  return ODX_SUCCESS;
}

static int do_rhs(odx_handle *h, double t, double hcur) {
  if (h->n <= 0) return ODX_SUCCESS;
  ODX_TRY(ensure_moments_Y(h, t));
  ODX_TRY(compute_psi(h, h->Y, t));
  return eval_rhs_out(h, h->Y, h->y_slot, t, hcur, nullptr);
}

constexpr size_t MAIL_BOX_BYTES = sizeof(double) * 2 * MAX_RANKS * NRED;
constexpr size_t MAIL_FLAG_BYTES = sizeof(uint32_t) * 2 * MAX_RANKS;
constexpr size_t MAIL_LL_OFF = ((MAIL_BOX_BYTES + MAIL_FLAG_BYTES + sizeof(int) + 255) / 256) * 256;
constexpr size_t MAIL_LL_BYTES = sizeof(unsigned long long) * 2 * MAX_RANKS * LLW;
constexpr size_t MAIL_BYTES = ((MAIL_LL_OFF + MAIL_LL_BYTES + 255) / 256) * 256;
static inline unsigned long long *mail_ll(void *m) { return (unsigned long long *)((char *)m + MAIL_LL_OFF); }
static inline double *mail_box(void *m) { return (double *)m; }
static inline uint32_t *mail_flag(void *m) { return (uint32_t *)((char *)m + MAIL_BOX_BYTES); }
static inline int *mail_abort(void *m) { return (int *)((char *)m + MAIL_BOX_BYTES + MAIL_FLAG_BYTES); }
static int ensure_mail(odx_handle *h) {
  if (h->mail) return ODX_SUCCESS;
  ODX_CUDA(h, cudaMalloc(&h->mail, MAIL_BYTES));
  ODX_CUDA(h, cudaMemset(h->mail, 0, MAIL_BYTES));
  h->peer_mail[0] = h->mail;
  return ODX_SUCCESS;
}

/* allreduce(SUM) of one sums slot through the peer mailboxes (same protocol and epoch counter as the persistent kernel) */
__global__ void k_peer_allreduce(double *sums, int rank, int nranks, uint32_t epoch, double *mailbox, uint32_t *mailflag,
                                 PersistArgs g) {
  const uint32_t tag = epoch + 1u; const size_t par = (size_t)(epoch & 1u);
  if (threadIdx.x < NRED) {
    for (int r = 0; r < nranks; r++) g.peer_box[r][(par * MAX_RANKS + rank) * NRED + threadIdx.x] = sums[threadIdx.x];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < nranks) { volatile uint32_t *f = g.peer_flag[threadIdx.x] + par * MAX_RANKS + rank; *f = tag; }
  __threadfence_system();
  if (threadIdx.x < nranks) {
    volatile uint32_t *f = mailflag + par * MAX_RANKS + threadIdx.x;
    unsigned long long spins = 0;
    while (*f != tag) { __nanosleep(64); if (++spins > (1ull << 26)) { *g.abort_flag = 1; break; } }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < NRED) {
    double v = 0.0;
    for (int r = 0; r < nranks; r++) v += __ldcv(&mailbox[(par * MAX_RANKS + r) * NRED + threadIdx.x]);
    sums[threadIdx.x] = v;
  }
}

/* Before a persistent launch on a sharded population every rank tells its peers (stage count, step count, epoch) through the same
 * mailbox and compares: a rank that asks for another number of steps would leave the others spinning at a mean-field barrier
 * until the poll limit (tens of seconds).  On a mismatch the abort code 2 is raised, the step kernel that follows returns at
 * once, and odx_sync reports it.  Device-side only: no host synchronisation is added to the launch path. */
__global__ void k_peer_agree(int rank, int nranks, uint32_t epoch, double want0, double want1, double *mailbox, uint32_t *mailflag, PersistArgs g) {
  const uint32_t tag = epoch + 1u; const size_t par = (size_t)(epoch & 1u);
  if (threadIdx.x < 2) {
    for (int r = 0; r < nranks; r++) g.peer_box[r][(par * MAX_RANKS + rank) * NRED + threadIdx.x] = threadIdx.x ? want1 : want0;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < nranks) { volatile uint32_t *f = g.peer_flag[threadIdx.x] + par * MAX_RANKS + rank; *f = tag; }
  __threadfence_system();
  if (threadIdx.x < nranks) {
    volatile uint32_t *f = mailflag + par * MAX_RANKS + threadIdx.x;
    unsigned long long spins = 0;
    while (*f != tag) { __nanosleep(64); if (++spins > (1ull << 26)) { *g.abort_flag = 1; break; } }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < nranks && *g.abort_flag == 0) {
    const double a = __ldcv(&mailbox[(par * MAX_RANKS + threadIdx.x) * NRED + 0]), b = __ldcv(&mailbox[(par * MAX_RANKS + threadIdx.x) * NRED + 1]);
    if (a != want0 || b != want1) *g.abort_flag = 2;
  }
}

static void fill_peer_args(odx_handle *h, PersistArgs &g) {
  g.rank = h->rank; g.nranks = h->nranks; g.epoch0 = h->epoch;
  g.mailbox = h->mail ? mail_box(h->mail) : nullptr; g.mailflag = h->mail ? mail_flag(h->mail) : nullptr;
  g.abort_flag = h->mail ? mail_abort(h->mail) : nullptr;
  for (int r = 0; r < MAX_RANKS; r++) {
    void *m = (r < h->nranks) ? (r == h->rank ? h->mail : h->peer_mail[r]) : nullptr;
    g.peer_box[r] = m ? mail_box(m) : nullptr; g.peer_flag[r] = m ? mail_flag(m) : nullptr;
    g.peer_ll[r] = m ? mail_ll(m) : nullptr;
  }
  g.mail_ll = h->mail ? mail_ll(h->mail) : nullptr;
}

/* nsteps whole steps in one cooperative launch (see k_chain_persist) */
static int chain_run_persist(odx_handle *h, int nst, double t0, double hh, int64_t nsteps, int want_out_last, int iterate = 0) {
  ODX_TRY(ensure_moments_Y(h, t0));
  const int new_slot = h->y_slot ^ 1;
  Pop p = pop_view(h);
  PersistArgs g; memset(&g, 0, sizeof g);
  g.nst = nst; g.nsteps = nsteps; g.t0 = t0; g.h = hh; g.want_out_last = want_out_last; g.eval_index0 = h->eval_index; g.recon = h->use_recon; g.iterate = iterate;
  g.sums_y = sums_ptr(h, h->y_slot); g.partial = h->partial; g.bar = h->counter + 32; g.totals = h->sums + (size_t)(SLOT_STAGE + 6) * NRED;
  g.ll_part = h->ll_words; g.ll_gtot = h->ll_words + (size_t)2 * PERSIST_MAX_CTAS * LLW;
  g.sums_new = h->sums + (size_t)new_slot * NRED; g.sums_report = h->sums + (size_t)SLOT_STAGE * NRED;
  fill_peer_args(h, g);
  /* TMA-ring variant for populations that stream from HBM, plain loads while the working set is L2-resident; the lean RK4
   * scheme once the working set is well beyond L2 (measured on Kuramoto, ms per step scaled to 2^24 particles: N = 2^21 plain
   * 0.446 / ring 0.461 / lean 0.481, 2^22: 0.437 / 0.403 / 0.422, 2^23: 0.446 / 0.413 / 0.386, 2^24: 0.443 / 0.400 / 0.337) */
  /* columns a step touches: Euler reads and writes Y in place (+ E, PSI); the chain methods also keep a stage state and an accumulator.
   * vdP SDE at N = 2^22: 67 MB, L2-resident -- the plain variant runs 0.0465 ms per step there, the ring 0.0513 (profiles/r2x_sde_fast_normal.txt) */
  const double wset = (double)h->n * (nst == 1 ? (NVAR + NEXPR + NPSI) : TMA_NCOL_MAX) * sizeof(double);
  const bool ring = TMA_ON && h->coop_blocks_tma > 0 && (h->tma_mode == 1 || (h->tma_mode < 0 && wset > 0.9 * h->l2_bytes));
  g.lean = (LEAN_OK && ring && nst == 4 && !iterate && (h->lean_mode == 1 || (h->lean_mode < 0 && wset > 1.5 * h->l2_bytes))) ? 1 : 0;   /* ring variant only */
  const int pt = ring ? PT_TMA : PT_LD, per_cta = ring ? TMA_TILE : 2 * PT_LD;
  int64_t grid = (h->n + per_cta - 1) / per_cta;
  if (grid > (ring ? h->coop_blocks_tma : h->coop_blocks)) grid = ring ? h->coop_blocks_tma : h->coop_blocks;
  if (grid > PERSIST_MAX_CTAS) grid = PERSIST_MAX_CTAS;
  if (grid < 1) grid = 1;
  if (h->nranks > 1 && h->peers_attached) {
    k_peer_agree<<<1, 32, 0, h->stream>>>(h->rank, h->nranks, h->epoch, (double)nsteps, (double)nst * 4294967296.0 + (double)h->epoch, g.mailbox, g.mailflag, g);
    ODX_TRY(launch_check(h));
    h->epoch++; g.epoch0 = h->epoch;
  }
  void *args[] = {(void *)&p, (void *)&g};
  ODX_CUDA(h, cudaLaunchCooperativeKernel(ring ? (const void *)k_chain_persist<TMA_ON> : (const void *)k_chain_persist<false>, dim3((unsigned)grid),
                                          dim3(pt), args, ring ? PF_BYTES : 0, h->stream));
  h->launches++;
  h->epoch += (uint32_t)(nst * nsteps);
  h->eval_index += (uint32_t)(2 * nst * nsteps); h->rhs_evals += (uint64_t)nst * nsteps;
  h->report_slot = SLOT_STAGE;
  h->rates_slot = new_slot; h->y_slot = new_slot; h->ymoments_valid = NEEDS_MOMENTS; h->k1_valid = false;
  return ODX_SUCCESS;
}
static inline bool can_persist(const odx_handle *h) {
  /* delay(): the history is edited between evaluations, by kernels of their own */
  return h->use_persistent && h->coop_ok && !M::HAS_PAIR && !M::HAS_LOWRANK && !M::USES_KY && M::NDELAY == 0 && (h->nranks == 1 || h->peers_attached);
}
static int alloc_cols(odx_handle *h, double **ptr, int ncol);
/* models with KY(): neighbours read the stage state while the stage forms the next one, so nothing is updated in place */
static int ky_guard(odx_handle *h) {
  if (!M::USES_KY) return ODX_SUCCESS;
  if (h->nranks > 1) { h->err = "KY() ring neighbours are not supported on a sharded population yet (needs a halo exchange)"; return ODX_EUNSUP; }
  if (!h->YALT) ODX_TRY(alloc_cols(h, &h->YALT, NVAR));
  return ODX_SUCCESS;
}

/* one step of a chain method.  nst = 1: Euler (fe_apply); nst = 4: classic RK4 (dde_apply) */
static int chain_step(odx_handle *h, int nst, double t, double hh, int want_out, int iterate = 0) {
  if (h->n <= 0) return ODX_SUCCESS;
  if (can_persist(h)) return chain_run_persist(h, nst, t, hh, 1, want_out, iterate);
  ODX_TRY(ky_guard(h));
  ODX_TRY(ensure_moments_Y(h, t));
  const int g = grid_chain(h, h->n);
  Pop p = pop_view(h);
  const int new_slot = h->y_slot ^ 1;
  ChainArgs a; a.h = hh; a.want_out = want_out; a.reverse = 0; a.recon = 0.0; a.skip_acc_store = 0; a.iterate = iterate;
  if (nst == 1) {
    ODX_TRY(compute_psi(h, h->Y, t));
    a.sums_in = sums_ptr(h, h->y_slot); a.t_eval = t; a.t_next = t + hh; a.a = hh; a.w = 1.0; a.eval_index = h->eval_index; h->eval_index += 2;
    a.reverse = (h->sweep_dir ^= 1);
    if (M::USES_KY) p.Y_OUT = h->YS;              /* the new state goes to the other buffer, then the buffers change roles */
    ODX_TRY(delay_before_eval(h, h->Y, t, hh, false));       /* fe_apply never sets DELAY_FIRST_EVAL: look-ups only */
    k_chain<true, true><<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, new_slot));
    ODX_TRY(launch_check(h));
    if (M::USES_KY) { double *tmp = h->Y; h->Y = h->YS; h->YS = tmp; }
    h->report_slot = h->y_slot; h->rhs_evals += 1;
  } else {
    /* k[i] = y[i] + (*h)/2.0*k1[i]; ...; y[i] += (*h)/6.0*(k1 + 2.0*(k2+k3) + k4)   (methods_odexp.c:1846-1863) */
    const double half = hh / 2.0;
    ODX_TRY(compute_psi(h, h->Y, t));
    a.sums_in = sums_ptr(h, h->y_slot); a.t_eval = t; a.t_next = t + half; a.a = half; a.w = 1.0; a.eval_index = h->eval_index; h->eval_index += 2;
    a.reverse = (h->sweep_dir ^= 1); a.skip_acc_store = h->use_recon;
    ODX_TRY(delay_before_eval(h, h->Y, t, hh, true));        /* DELAY_FIRST_EVAL = 1 around k1 (src/methods_odexp.c:1842-1844) */
    k_chain<true, false><<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, SLOT_STAGE + 0));
    a.skip_acc_store = 0;
    ODX_TRY(launch_check(h)); if (NEEDS_MOMENTS) ODX_TRY(post_reduce(h, SLOT_STAGE + 0));
    ODX_TRY(compute_psi(h, h->YS, t + half));
    a.sums_in = sums_ptr(h, SLOT_STAGE + 0); a.t_eval = t + half; a.t_next = t + half; a.a = half; a.w = 2.0; a.eval_index = h->eval_index; h->eval_index += 2;
    a.reverse = (h->sweep_dir ^= 1); a.recon = h->use_recon ? 1.0 / half : 0.0;
    if (M::USES_KY) p.YS_OUT = h->YALT;           /* y2 in YS is still read by the neighbours: y3 goes to the other buffer */
    ODX_TRY(delay_before_eval(h, h->YS, t + half, hh, false));
    k_chain<false, false><<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, SLOT_STAGE + 1));
    a.recon = 0.0;
    ODX_TRY(launch_check(h)); if (NEEDS_MOMENTS) ODX_TRY(post_reduce(h, SLOT_STAGE + 1));
    ODX_TRY(compute_psi(h, M::USES_KY ? h->YALT : h->YS, t + half));
    a.sums_in = sums_ptr(h, SLOT_STAGE + 1); a.t_eval = t + half; a.t_next = t + hh; a.a = hh; a.w = 2.0; a.eval_index = h->eval_index; h->eval_index += 2;
    a.reverse = (h->sweep_dir ^= 1);
    if (M::USES_KY) { p.YS = h->YALT; p.YS_OUT = h->YS; }
    ODX_TRY(delay_before_eval(h, p.YS, t + half, hh, false));
    k_chain<false, false><<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, SLOT_STAGE + 2));
    if (M::USES_KY) { p.YS = h->YS; p.YS_OUT = h->YS; }
    ODX_TRY(launch_check(h)); if (NEEDS_MOMENTS) ODX_TRY(post_reduce(h, SLOT_STAGE + 2));
    ODX_TRY(compute_psi(h, h->YS, t + hh));
    a.sums_in = sums_ptr(h, SLOT_STAGE + 2); a.t_eval = t + hh; a.t_next = t + hh; a.a = hh / 6.0; a.w = 1.0; a.eval_index = h->eval_index; h->eval_index += 2;
    a.reverse = (h->sweep_dir ^= 1);
    ODX_TRY(delay_before_eval(h, h->YS, t + hh, hh, false));
    k_chain<false, true><<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, new_slot));
    ODX_TRY(launch_check(h));
    h->report_slot = SLOT_STAGE + 2; h->rhs_evals += 4;
  }
  if (NEEDS_MOMENTS || M::HAS_DEATH || M::HAS_REPLI) ODX_TRY(post_reduce(h, new_slot));
  h->rates_slot = new_slot;
  h->y_slot = new_slot; h->ymoments_valid = NEEDS_MOMENTS;
  h->k1_valid = false;
  return ODX_SUCCESS;
}

/* embedded pairs: coefficient tables as published for GSL's rkf45 / rkck steppers */
struct Tableau { double ah[12]; double b[12][12]; double c[13]; double ec[13]; unsigned order; int nst; int err_diff; };
static const Tableau TAB_RKF45 = {
    {1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0},
    {{1.0 / 4.0},
     {3.0 / 32.0, 9.0 / 32.0},
     {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0},
     {8341.0 / 4104.0, -32832.0 / 4104.0, 29440.0 / 4104.0, -845.0 / 4104.0},
     {-6080.0 / 20520.0, 41040.0 / 20520.0, -28352.0 / 20520.0, 9295.0 / 20520.0, -5643.0 / 20520.0}},
    {902880.0 / 7618050.0, 0.0, 3953664.0 / 7618050.0, 3855735.0 / 7618050.0, -1371249.0 / 7618050.0, 277020.0 / 7618050.0},
    {1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0}, 5, 6, 0};
static const Tableau TAB_RKCK = {
    {1.0 / 5.0, 3.0 / 10.0, 3.0 / 5.0, 1.0, 7.0 / 8.0},
    {{1.0 / 5.0},
     {3.0 / 40.0, 9.0 / 40.0},
     {0.3, -0.9, 1.2},
     {-11.0 / 54.0, 2.5, -70.0 / 27.0, 35.0 / 27.0},
     {1631.0 / 55296.0, 175.0 / 512.0, 575.0 / 13824.0, 44275.0 / 110592.0, 253.0 / 4096.0}},
    {37.0 / 378.0, 0.0, 250.0 / 621.0, 125.0 / 594.0, 0.0, 512.0 / 1771.0},
    {37.0 / 378.0 - 2825.0 / 27648.0, 0.0, 250.0 / 621.0 - 18575.0 / 48384.0, 125.0 / 594.0 - 13525.0 / 55296.0,
     -277.0 / 14336.0, 512.0 / 1771.0 - 0.25}, 5, 6, 0};
/* GSL rk2: k2 at the midpoint, k3 at y + h (-k1 + 2 k2), y += h (k1 + 4 k2 + k3) / 6, yerr = h (k2 - (k1 + 4 k2 + k3) / 6) */
static const Tableau TAB_RK2 = {
    {1.0 / 2.0, 1.0},
    {{1.0 / 2.0},
     {-1.0, 2.0}},
    {1.0 / 6.0, 4.0 / 6.0, 1.0 / 6.0},
    {-1.0 / 6.0, 1.0 / 3.0, -1.0 / 6.0}, 2, 3, 0};
/* GSL rk8pd = Prince-Dormand RK8(7)13M: c = the 8th-order weights (propagated), ec = the 7th-order weights, yerr = h (ksum7 - ksum8) */
static const Tableau TAB_RK8PD = {
    {1.0 / 18.0, 1.0 / 12.0, 1.0 / 8.0, 5.0 / 16.0, 3.0 / 8.0, 59.0 / 400.0, 93.0 / 200.0, 5490023248.0 / 9719169821.0, 13.0 / 20.0,
   1201146811.0 / 1299019798.0, 1.0, 1.0},
    {{1.0 / 18.0},
   {1.0 / 48.0, 1.0 / 16.0},
   {1.0 / 32.0, 0.0, 3.0 / 32.0},
   {5.0 / 16.0, 0.0, -75.0 / 64.0, 75.0 / 64.0},
   {3.0 / 80.0, 0.0, 0.0, 3.0 / 16.0, 3.0 / 20.0},
   {29443841.0 / 614563906.0, 0.0, 0.0, 77736538.0 / 692538347.0, -28693883.0 / 1125000000.0, 23124283.0 / 1800000000.0},
   {16016141.0 / 946692911.0, 0.0, 0.0, 61564180.0 / 158732637.0, 22789713.0 / 633445777.0, 545815736.0 / 2771057229.0,
    -180193667.0 / 1043307555.0},
   {39632708.0 / 573591083.0, 0.0, 0.0, -433636366.0 / 683701615.0, -421739975.0 / 2616292301.0, 100302831.0 / 723423059.0,
    790204164.0 / 839813087.0, 800635310.0 / 3783071287.0},
   {246121993.0 / 1340847787.0, 0.0, 0.0, -37695042795.0 / 15268766246.0, -309121744.0 / 1061227803.0, -12992083.0 / 490766935.0,
    6005943493.0 / 2108947869.0, 393006217.0 / 1396673457.0, 123872331.0 / 1001029789.0},
   {-1028468189.0 / 846180014.0, 0.0, 0.0, 8478235783.0 / 508512852.0, 1311729495.0 / 1432422823.0, -10304129995.0 / 1701304382.0,
    -48777925059.0 / 3047939560.0, 15336726248.0 / 1032824649.0, -45442868181.0 / 3398467696.0, 3065993473.0 / 597172653.0},
   {185892177.0 / 718116043.0, 0.0, 0.0, -3185094517.0 / 667107341.0, -477755414.0 / 1098053517.0, -703635378.0 / 230739211.0,
    5731566787.0 / 1027545527.0, 5232866602.0 / 850066563.0, -4093664535.0 / 808688257.0, 3962137247.0 / 1805957418.0,
    65686358.0 / 487910083.0},
   {403863854.0 / 491063109.0, 0.0, 0.0, -5068492393.0 / 434740067.0, -411421997.0 / 543043805.0, 652783627.0 / 914296604.0,
    11173962825.0 / 925320556.0, -13158990841.0 / 6184727034.0, 3936647629.0 / 1978049680.0, -160528059.0 / 685178525.0,
    248638103.0 / 1413531060.0, 0.0}},
    {14005451.0 / 335480064.0, 0.0, 0.0, 0.0, 0.0, -59238493.0 / 1068277825.0, 181606767.0 / 758867731.0, 561292985.0 / 797845732.0,
   -1041891430.0 / 1371343529.0, 760417239.0 / 1151165299.0, 118820643.0 / 751138087.0, -528747749.0 / 2220607170.0, 1.0 / 4.0},
    {13451932.0 / 455176623.0, 0.0, 0.0, 0.0, 0.0, -808719846.0 / 976000145.0, 1757004468.0 / 5645159321.0, 656045339.0 / 265891186.0,
   -3867574721.0 / 1518517206.0, 465885868.0 / 322736535.0, 53011238.0 / 667516719.0, 2.0 / 45.0, 0.0}, 8, 13, 1};

static int read_rmax(odx_handle *h, double *rmax_out) {
  ODX_CUDA(h, cudaMemcpyAsync(h->h_pin, h->rmax, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  double rmax; memcpy(&rmax, h->h_pin, sizeof rmax);
#ifdef ODX_HAVE_NCCL_H
  if (h->nranks > 1) {
    ODX_CUDA(h, cudaMemcpyAsync(h->rmax, &rmax, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = g_nccl.AllReduce(h->rmax, h->rmax, 1, ncclDouble, ncclMax, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce(MAX): ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
    ODX_CUDA(h, cudaMemcpyAsync(h->h_pin, h->rmax, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(&rmax, h->h_pin, sizeof rmax);
  }
#endif
  *rmax_out = rmax;
  return ODX_SUCCESS;
}

/* one attempted embedded step of size hh from (t, Y): leaves y_new in YS, dydt_out in DOUT, returns rmax */
static int embedded_attempt(odx_handle *h, const Tableau &T, double t, double hh, double *rmax_out) {
  const int g = grid_for(h, h->n), nst = T.nst;
  Pop p = pop_view(h);
  ODX_CUDA(h, cudaMemsetAsync(h->rmax, 0, sizeof(unsigned long long), h->stream));
  if (!h->k1_valid) {
    ODX_TRY(ensure_moments_Y(h, t));
    ODX_TRY(compute_psi(h, h->Y, t));
  }
  for (int s = 1; s <= nst; s++) {
    TabArgs a;
    memset(&a, 0, sizeof a);
    a.s = s; a.h = hh; a.K1 = h->K1; a.eps_abs = h->eps_abs; a.eps_rel = h->eps_rel; a.rmax_bits = h->rmax;
    a.final_ = (s == nst); a.err_diff = T.err_diff;
    a.skip_rhs = (s == 1 && h->k1_valid) ? 1 : 0;
    a.ycur = (s == 1) ? h->Y : h->YS; a.y0src = h->Y;
    a.sums_in = sums_ptr(h, s == 1 ? h->y_slot : SLOT_STAGE + (s - 2) % 5);       /* only the previous stage's sums are read */
    a.t_eval = (s == 1) ? t : t + T.ah[s - 2] * hh;
    a.t_next = (s == nst) ? t + hh : t + T.ah[s - 1] * hh;
    if (s < nst) { for (int j = 0; j < s; j++) a.b[j] = T.b[s - 1][j]; }
    else { for (int j = 0; j < nst; j++) { a.b[j] = T.c[j]; a.ec[j] = T.ec[j]; } }
    a.eval_index = h->eval_index; h->eval_index += 2;
    if (!a.skip_rhs) h->rhs_evals++;
    const int out_slot = (s == nst) ? SLOT_STAGE + 5 : SLOT_STAGE + (s - 1) % 5;  /* the sums of y_new always land in slot 5 */
    /* KY(): neighbours read the stage state this kernel evaluates the RHS at, so the state it forms goes to the other buffer
     * (the stages alternate so that the last one leaves y_new in YS) */
    if (M::USES_KY) { const bool alt = ((nst - s) & 1) != 0; p.YS_OUT = alt ? h->YALT : h->YS; if (s > 1) a.ycur = alt ? h->YS : h->YALT; }
    if (!a.skip_rhs) ODX_TRY(delay_before_eval(h, a.ycur, a.t_eval, hh, false));      /* the GSL steppers never push: look-ups (or refills) only */
    k_tab<<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, out_slot));
    ODX_TRY(launch_check(h));
    if (NEEDS_MOMENTS) ODX_TRY(post_reduce(h, out_slot));
    ODX_TRY(compute_psi(h, p.YS_OUT, a.t_next));
  }
  h->k1_valid = true;   /* K1 now holds f(t, Y) whatever happens to this attempt */
  /* dydt_out = f(t + h, y_new): also refreshes aux / rates / reported mean fields like every reference RHS call */
  ODX_TRY(eval_rhs_out(h, h->YS, SLOT_STAGE + 5, t + hh, hh, h->DOUT));
  return read_rmax(h, rmax_out);
}

/* one classic RK4 (sub)step of size hh from y0src through the tableau kernel: result in YS, its %M sums in out_slot.
 * k1buf holds (or receives) the slope at y0src. */
static int rk4_tab_step(odx_handle *h, const double *y0src, double *k1buf, bool k1_valid, int y0_slot, double t, double hh, int out_slot) {
  static const double cnode[5] = {0.0, 0.5, 0.5, 1.0, 1.0};
  const int g = grid_for(h, h->n);
  Pop p = pop_view(h);
  if (!k1_valid) ODX_TRY(compute_psi(h, y0src, t));
  for (int s = 1; s <= 4; s++) {
    TabArgs a; memset(&a, 0, sizeof a);
    a.s = s; a.h = hh; a.K1 = k1buf; a.y0src = y0src; a.eps_abs = h->eps_abs; a.eps_rel = h->eps_rel; a.rmax_bits = h->rmax;
    a.final_ = 0; a.skip_rhs = (s == 1 && k1_valid) ? 1 : 0;
    a.ycur = (s == 1) ? y0src : h->YS;
    a.sums_in = sums_ptr(h, s == 1 ? y0_slot : SLOT_STAGE + (s - 2));
    a.t_eval = t + cnode[s - 1] * hh; a.t_next = t + cnode[s] * hh;
    if (s == 1) a.b[0] = 0.5;
    else if (s == 2) a.b[1] = 0.5;
    else if (s == 3) a.b[2] = 1.0;
    else { a.b[0] = 1.0 / 6.0; a.b[1] = 1.0 / 3.0; a.b[2] = 1.0 / 3.0; a.b[3] = 1.0 / 6.0; }
    a.eval_index = h->eval_index; h->eval_index += 2;
    if (!a.skip_rhs) h->rhs_evals++;
    const int slot = (s == 4) ? out_slot : SLOT_STAGE + (s - 1);
    if (M::USES_KY) { p.YS_OUT = (s & 1) ? h->YALT : h->YS; if (s > 1) a.ycur = (s & 1) ? h->YS : h->YALT; }
    k_tab<<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, slot));
    ODX_TRY(launch_check(h));
    if (NEEDS_MOMENTS) ODX_TRY(post_reduce(h, slot));
    ODX_TRY(compute_psi(h, p.YS_OUT, a.t_next));
  }
  return ODX_SUCCESS;
}

/* one attempted step of GSL's rk4 (step doubling, 11 RHS evaluations): y_one = RK4(h), y = RK4(h/2) o RK4(h/2),
 * yerr = 4 (y - y_one) / 15.  Leaves y in YS and dydt_out in DOUT like embedded_attempt. */
static int rk4_doubling_attempt(odx_handle *h, double t, double hh, double *rmax_out) {
  const size_t colbytes = sizeof(double) * (size_t)A1(NVAR) * h->cap;
  ODX_CUDA(h, cudaMemsetAsync(h->rmax, 0, sizeof(unsigned long long), h->stream));
  if (!h->k1_valid) ODX_TRY(ensure_moments_Y(h, t));
  ODX_TRY(rk4_tab_step(h, h->Y, h->K1, h->k1_valid, h->y_slot, t, hh, SLOT_STAGE + 3));
  h->k1_valid = true;
  ODX_CUDA(h, cudaMemcpyAsync(h->YONE, h->YS, colbytes, cudaMemcpyDeviceToDevice, h->stream));
  ODX_TRY(rk4_tab_step(h, h->Y, h->K1, true, h->y_slot, t, hh / 2.0, SLOT_STAGE + 4));
  ODX_CUDA(h, cudaMemcpyAsync(h->YH, h->YS, colbytes, cudaMemcpyDeviceToDevice, h->stream));
  ODX_TRY(eval_rhs_out(h, h->YH, SLOT_STAGE + 4, t + hh / 2.0, hh, h->K1B));
  ODX_TRY(rk4_tab_step(h, h->YH, h->K1B, true, SLOT_STAGE + 4, t + hh / 2.0, hh / 2.0, SLOT_STAGE + 5));
  ODX_TRY(eval_rhs_out(h, h->YS, SLOT_STAGE + 5, t + hh, hh, h->DOUT));
  k_err_doubling<<<grid_for(h, h->n), BLOCK, 0, h->stream>>>(pop_view(h), h->YS, h->YONE, h->eps_abs, h->eps_rel, h->rmax);
  ODX_TRY(launch_check(h));
  return read_rmax(h, rmax_out);
}

static int alloc_cols(odx_handle *h, double **ptr, int ncol) {
  ODX_CUDA(h, cudaMalloc((void **)ptr, sizeof(double) * (size_t)A1(ncol) * h->cap));
  ODX_CUDA(h, cudaMemsetAsync(*ptr, 0, sizeof(double) * (size_t)A1(ncol) * h->cap, h->stream));
  return ODX_SUCCESS;
}
static int ensure_stage(odx_handle *h, size_t bytes) {
  if (h->stage_bytes >= bytes) return ODX_SUCCESS;
  if (h->stage_dev) cudaFree(h->stage_dev);
  h->stage_dev = nullptr; h->stage_bytes = 0;
  ODX_CUDA(h, cudaMalloc((void **)&h->stage_dev, bytes));
  h->stage_bytes = bytes;
  return ODX_SUCCESS;
}
static void invalidate_state(odx_handle *h) { h->ymoments_valid = false; h->k1_valid = false; h->ev_count = 0; }

static int upload_cols(odx_handle *h, const double *aos, double *soa, int ncol, int64_t n, int64_t first) {
  if (ncol == 0 || n == 0 || aos == nullptr) return ODX_SUCCESS;
  if (ncol == 1) {          /* a single column is already in device layout: straight into place, no staging pass */
    ODX_CUDA(h, cudaMemcpyAsync(soa + first, aos, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    return ODX_SUCCESS;
  }
  const size_t bytes = sizeof(double) * (size_t)ncol * n;
  ODX_TRY(ensure_stage(h, bytes));
  ODX_CUDA(h, cudaMemcpyAsync(h->stage_dev, aos, bytes, cudaMemcpyHostToDevice, h->stream));
  k_aos_to_soa<<<h->sm_count * 4, 256, 0, h->stream>>>(h->stage_dev, soa, ncol, n, h->cap, first);
  ODX_TRY(launch_check(h));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  return ODX_SUCCESS;
}
static int d2h_pipelined(odx_handle *h, void *dst, const void *src_dev, size_t bytes);      /* odx_output.cuh */
static int download_cols(odx_handle *h, const double *soa, double *aos, int ncol, int64_t n, int64_t first) {
  if (ncol == 0 || n == 0 || aos == nullptr) return ODX_SUCCESS;
  if (ncol == 1) return d2h_pipelined(h, aos, soa + first, sizeof(double) * n);      /* already in host layout */
  const size_t bytes = sizeof(double) * (size_t)ncol * n;
  ODX_TRY(ensure_stage(h, bytes));
  k_soa_to_aos<<<h->sm_count * 4, 256, 0, h->stream>>>(soa, h->stage_dev, ncol, n, h->cap, first);
  ODX_TRY(launch_check(h));
  return d2h_pipelined(h, aos, h->stage_dev, bytes);
}

}  // namespace odx

/* =================================================================================================
 * C ABI
 * =============================================================================================== */
extern "C" {

const odx_model_info *odx_model(void) { return &OdxModel::info(); }

const char *odx_last_error(const odx_handle *h) { return h ? h->err.c_str() : "null handle"; }

int odx_create(odx_handle **out, int device, int64_t capacity) {
  if (!out || capacity <= 0) return ODX_EINVAL;
  odx_handle *h = new (std::nothrow) odx_handle();
  if (!h) return ODX_ENOMEM;
  *out = h;
  h->device = device;
  h->cap = (capacity + 31) / 32 * 32;      /* columns start 256-byte aligned */
  ODX_CUDA(h, cudaSetDevice(device));
  cudaDeviceProp prop;
  ODX_CUDA(h, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { h->err = "this library is built for sm_100a (B200) only"; return ODX_EUNSUP; }
  h->sm_count = prop.multiProcessorCount;
  ODX_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  int occ = 1;
  ODX_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_chain<false, false>, BLOCK, 0));
  if (occ < 1) occ = 1;
  h->grid_cap = h->sm_count * occ;
  {
    int coop = 0, occp = 0;
    ODX_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    ODX_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, k_chain_persist<false>, PT_LD, 0));
    h->coop_ok = coop && occp > 0; h->coop_blocks = h->sm_count * occp;
    if (TMA_ON) {
      int occt = 0;
      ODX_CUDA(h, cudaFuncSetAttribute(k_chain_persist<TMA_ON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_BYTES));
      ODX_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occt, k_chain_persist<TMA_ON>, PT_TMA, PF_BYTES));
      h->coop_blocks_tma = h->sm_count * occt;
    }
    h->l2_bytes = (double)prop.l2CacheSize;
  }
  ODX_TRY(alloc_cols(h, &h->Y, NVAR)); ODX_TRY(alloc_cols(h, &h->YS, NVAR)); ODX_TRY(alloc_cols(h, &h->ACC, NVAR));
  ODX_TRY(alloc_cols(h, &h->E, NEXPR)); ODX_TRY(alloc_cols(h, &h->AUX, NAUX)); ODX_TRY(alloc_cols(h, &h->PSI, NPSI));
  ODX_TRY(alloc_cols(h, &h->DEATH, 1)); ODX_TRY(alloc_cols(h, &h->REPLI, 1));
  ODX_CUDA(h, cudaMalloc((void **)&h->ID, sizeof(int32_t) * h->cap));
  for (int k = 0; k < M::NDELAY; k++) {       /* delay(): DELAY_SSTACK samples of (t, x) + the write position per declared stack */
    const size_t K = (size_t)M::delay_count(k);
    ODX_CUDA(h, cudaMalloc((void **)&h->dl_T[k], sizeof(double) * K * DELAY_SSTACK));
    ODX_CUDA(h, cudaMalloc((void **)&h->dl_X[k], sizeof(double) * K * DELAY_SSTACK));
    ODX_CUDA(h, cudaMalloc((void **)&h->dl_pos[k], sizeof(int) * K));
    ODX_CUDA(h, cudaMemsetAsync(h->dl_T[k], 0, sizeof(double) * K * DELAY_SSTACK, h->stream));      /* a global `stack` is zero-initialised */
    ODX_CUDA(h, cudaMemsetAsync(h->dl_X[k], 0, sizeof(double) * K * DELAY_SSTACK, h->stream));
    ODX_CUDA(h, cudaMemsetAsync(h->dl_pos[k], 0, sizeof(int) * K, h->stream));
  }
  ODX_CUDA(h, cudaMalloc((void **)&h->d_mu, sizeof(double) * A1(NPAR)));
  const odx_model_info *mi = odx_model();
  for (int k = 0; k < NPAR; k++) h->h_mu[k] = mi->par_defaults[k];
  ODX_CUDA(h, cudaMemcpyAsync(h->d_mu, h->h_mu, sizeof(double) * A1(NPAR), cudaMemcpyHostToDevice, h->stream));
  ODX_CUDA(h, cudaMalloc((void **)&h->partial, sizeof(double) * (size_t)NRED * 2 * (h->grid_cap + h->coop_blocks + 8)));
  ODX_CUDA(h, cudaMalloc((void **)&h->counter, sizeof(unsigned int) * (32 + 2 * PERSIST_MAX_CTAS + 8)));
  ODX_CUDA(h, cudaMemsetAsync(h->counter, 0, sizeof(unsigned int) * (32 + 2 * PERSIST_MAX_CTAS + 8), h->stream));
  ODX_CUDA(h, cudaMalloc((void **)&h->ll_words, sizeof(unsigned long long) * ((size_t)2 * PERSIST_MAX_CTAS * LLW + 2 * LLW + 8)));
  ODX_CUDA(h, cudaMemsetAsync(h->ll_words, 0, sizeof(unsigned long long) * ((size_t)2 * PERSIST_MAX_CTAS * LLW + 2 * LLW + 8), h->stream));
  ODX_CUDA(h, cudaMalloc((void **)&h->sums, sizeof(double) * NSLOT * NRED));
  ODX_CUDA(h, cudaMemsetAsync(h->sums, 0, sizeof(double) * NSLOT * NRED, h->stream));
  ODX_CUDA(h, cudaMalloc((void **)&h->rmax, sizeof(unsigned long long) * 2));
  ODX_CUDA(h, cudaMallocHost((void **)&h->h_pin, sizeof(double) * (NSLOT * NRED + 16)));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  return ODX_SUCCESS;
}

void odx_destroy(odx_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
#ifdef ODX_HAVE_NCCL_H
  if (h->comm) g_nccl.CommDestroy(h->comm);
#endif
  odx_events_free(h);
  odx_output_free(h);
  if (h->replica) { odx_destroy(h->replica); h->replica = nullptr; cudaSetDevice(h->device); }
  if (h->gcol) cudaFree(h->gcol);
  if (h->lr_plan) cudaFree(h->lr_plan);
  if (h->lr_rank) cudaFree(h->lr_rank);
  { double *lb[] = {h->lr_stats, h->lr_spart, h->lr_apart, h->lr_A, h->lr_x, h->lr_ones}; for (double *q : lb) if (q) cudaFree(q); }
  for (int r = 0; r < MAX_RANKS; r++) if (h->peer_mail[r] && h->peer_mail[r] != h->mail) cudaIpcCloseMemHandle(h->peer_mail[r]);
  if (h->mail) cudaFree(h->mail);
  if (h->ll_words) cudaFree(h->ll_words);
  double *bufs[] = {h->Y, h->YS, h->ACC, h->E, h->AUX, h->PSI, h->DEATH, h->REPLI, h->KS, h->YALT, h->K1, h->DOUT, h->YONE, h->YH, h->K1B, h->d_mu, h->partial,
                    h->sums, h->stage_dev, h->gather, h->pair_part, h->pair_sym_part};
  if (h->pair_poison) cudaFree(h->pair_poison);
  for (double *b : bufs) if (b) cudaFree(b);
  if (h->ID) cudaFree(h->ID);
  for (int k = 0; k < A1(M::NDELAY); k++) { if (h->dl_T[k]) cudaFree(h->dl_T[k]); if (h->dl_X[k]) cudaFree(h->dl_X[k]); if (h->dl_pos[k]) cudaFree(h->dl_pos[k]); }
  if (h->gather_id) cudaFree(h->gather_id);
  if (h->counter) cudaFree(h->counter);
  if (h->rmax) cudaFree(h->rmax);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  for (int i = 0; i < 8; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 2; i++) if (h->chunk_ev[i]) cudaEventDestroy(h->chunk_ev[i]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int odx_sync(odx_handle *h) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->mail) {
    int ab = 0;
    ODX_CUDA(h, cudaMemcpy(&ab, mail_abort(h->mail), sizeof ab, cudaMemcpyDeviceToHost));
    if (ab == 2) { h->err = "ranks disagree on the steps of a persistent launch: every rank must request the same stepper, step count and history of exchanges (INTEGRATION.md)"; return ODX_EFAILED; }
    if (ab) { h->err = "peer exchange timed out: a rank did not reach the mean-field barrier"; return ODX_EFAILED; }
  }
  return ODX_SUCCESS;
}

int odx_set_option(odx_handle *h, const char *name, double value) {
  if (strcmp(name, "persistent") == 0) { h->use_persistent = value != 0.0; return ODX_SUCCESS; }
  if (strcmp(name, "tma") == 0) { h->tma_mode = (int)value; return ODX_SUCCESS; }
  if (strcmp(name, "lean") == 0) { h->lean_mode = (int)value; return ODX_SUCCESS; }
  if (strcmp(name, "pair_sym") == 0) { h->pair_sym = (int)value; return ODX_SUCCESS; }
  if (strcmp(name, "leap_fused") == 0) { h->leap_fused = (int)value; return ODX_SUCCESS; }
  if (strcmp(name, "shard_events") == 0) { h->shard_events = (int)value; return ODX_SUCCESS; }
  if (strcmp(name, "ssa_device") == 0) { h->ssa_mode = (int)value; return ODX_SUCCESS; }
  if (strcmp(name, "reconstruct") == 0) { h->use_recon = value != 0.0; return ODX_SUCCESS; }
  if (strcmp(name, "lrmax") == 0) { h->lr_pmax = (int)value < 1 ? 1 : ((int)value > 31 ? 31 : (int)value); return ODX_SUCCESS; }
  if (strcmp(name, "lrabstol") == 0) { h->lr_abstol = value; return ODX_SUCCESS; }
  if (strcmp(name, "lrreltol") == 0) { h->lr_reltol = value; return ODX_SUCCESS; }
  h->err = std::string("odx_set_option: unknown option ") + name;
  return ODX_EINVAL;
}

/* peer-memory exchange: every rank exports its mailbox, the caller distributes the 64-byte handles (all ranks, rank
 * order) and every rank attaches to the others (cudaIpcOpenMemHandle; NVLink peer access is enabled lazily) */
int odx_peer_export(odx_handle *h, void *handle64) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_TRY(ensure_mail(h));
  cudaIpcMemHandle_t mh;
  ODX_CUDA(h, cudaIpcGetMemHandle(&mh, h->mail));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  memcpy(handle64, &mh, 64);
  return ODX_SUCCESS;
}
int odx_peer_attach(odx_handle *h, int rank, int nranks, const void *handles64) {
  if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) { h->err = "odx_peer_attach: bad rank / nranks"; return ODX_EINVAL; }
  if (handles64 == nullptr) { h->peers_attached = false; return ODX_SUCCESS; }      /* detach: back to the NCCL exchange */
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_TRY(ensure_mail(h));
  for (int r = 0; r < nranks; r++) {
    if (r == rank) { h->peer_mail[r] = h->mail; continue; }
    cudaIpcMemHandle_t mh; memcpy(&mh, (const char *)handles64 + (size_t)r * 64, 64);
    void *ptr = nullptr;
    ODX_CUDA(h, cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    h->peer_mail[r] = ptr;
  }
  h->rank = rank; h->nranks = nranks; h->peers_attached = nranks > 1;
  return ODX_SUCCESS;
}

int odx_set_params(odx_handle *h, const double *mu, int npar) {
  if (npar != NPAR) { h->err = "odx_set_params: wrong number of parameters"; return ODX_EINVAL; }
  ODX_CUDA(h, cudaSetDevice(h->device));
  for (int k = 0; k < NPAR; k++) h->h_mu[k] = mu[k];
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  ODX_CUDA(h, cudaMemcpyAsync(h->d_mu, h->h_mu, sizeof(double) * A1(NPAR), cudaMemcpyHostToDevice, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  invalidate_state(h);
  return ODX_SUCCESS;
}
/* EVENT_TYPE (SIM->stop_flag, src/odexp.h:387-399) and *SIM->h as the model code sees them during the evaluations that follow: only
 * delay() reads them (`if (EVENT_TYPE && EVENT_T0)` refills the history, src/utils_odexp.c:212-226) */
int odx_set_event_type(odx_handle *h, int event_type, double hstep) { h->event_type = event_type; h->event_h = hstep; return ODX_SUCCESS; }
int odx_set_seed(odx_handle *h, uint64_t seed) { h->k0 = (uint32_t)seed; h->k1 = (uint32_t)(seed >> 32); h->eval_index = 0; return ODX_SUCCESS; }
int odx_set_tolerances(odx_handle *h, double eps_abs, double eps_rel) { h->eps_abs = eps_abs; h->eps_rel = eps_rel; return ODX_SUCCESS; }

int odx_comm_unique_id(void *uid128) {
#ifdef ODX_HAVE_NCCL_H
  if (!g_nccl.load()) return ODX_ENCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ODX_ENCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(uid128, &id, 128);
  return ODX_SUCCESS;
#else
  (void)uid128; return ODX_ENCCL;
#endif
}
int odx_comm_init(odx_handle *h, int rank, int nranks, const void *uid128) {
  if (nranks <= 1) { h->rank = 0; h->nranks = 1; return ODX_SUCCESS; }
#ifdef ODX_HAVE_NCCL_H
  if (!g_nccl.load()) { h->err = "libnccl.so.2 not found"; return ODX_ENCCL; }
  ODX_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId id; memcpy(&id, uid128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (r != ncclSuccess) { h->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
  h->rank = rank; h->nranks = nranks;
  return ODX_SUCCESS;
#else
  (void)rank; (void)uid128; h->err = "built without nccl.h"; return ODX_ENCCL;
#endif
}

int odx_comm_rank(const odx_handle *h, int *rank, int *nranks) {
  if (rank) *rank = h->rank;
  if (nranks) *nranks = h->nranks;
  return ODX_SUCCESS;
}

/* global population size / offsets after any change of the local size (allreduce over ranks) */
static int refresh_global_n(odx_handle *h) {
  h->n_global = (double)h->n; h->n_before = 0; h->all_n[0] = h->n;
#ifdef ODX_HAVE_NCCL_H
  if (h->nranks > 1) {
    /* allgather of the per-rank sizes through a tiny device buffer */
    double *d = h->sums + (size_t)SLOT_RATES * NRED;   /* scratch, rewritten by the next evaluation */
    (void)d;
    unsigned long long *buf = nullptr;
    ODX_CUDA(h, cudaMalloc((void **)&buf, sizeof(unsigned long long) * (h->nranks + 1)));
    unsigned long long mine = (unsigned long long)h->n;
    ODX_CUDA(h, cudaMemcpyAsync(buf + h->nranks, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = g_nccl.AllGather(buf + h->nranks, buf, 1, ncclUint64, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); cudaFree(buf); return ODX_ENCCL; }
    std::vector<unsigned long long> all(h->nranks);
    ODX_CUDA(h, cudaMemcpyAsync(all.data(), buf, sizeof(unsigned long long) * h->nranks, cudaMemcpyDeviceToHost, h->stream));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(buf);
    double tot = 0; int64_t before = 0;
    for (int r2 = 0; r2 < h->nranks; r2++) { tot += (double)all[r2]; if (r2 < h->rank) before += (int64_t)all[r2]; if (r2 < MAX_RANKS) h->all_n[r2] = (int64_t)all[r2]; }
    h->n_global = tot; h->n_before = before;
  }
#endif
  return ODX_SUCCESS;
}

int odx_pop_upload(odx_handle *h, int64_t n, const double *y, const double *expr, const int32_t *ids, int32_t max_id) {
  if (n < 0 || n > h->cap) { h->err = "odx_pop_upload: n exceeds capacity"; return ODX_EINVAL; }
  ODX_CUDA(h, cudaSetDevice(h->device));
  h->n = n;
  ODX_TRY(upload_cols(h, y, h->Y, NVAR, n, 0));
  ODX_TRY(upload_cols(h, expr, h->E, NEXPR, n, 0));
  if (ids) {
    ODX_CUDA(h, cudaMemcpyAsync(h->ID, ids, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  } else if (n > 0) {
    k_iota_ids<<<h->sm_count * 4, 256, 0, h->stream>>>(h->ID, n, max_id - (int32_t)n);
    ODX_TRY(launch_check(h));
  }
  h->max_id = max_id;
  ODX_CUDA(h, cudaMemsetAsync(h->PSI, 0, sizeof(double) * A1(NPSI) * h->cap, h->stream));
  ODX_CUDA(h, cudaMemsetAsync(h->AUX, 0, sizeof(double) * A1(NAUX) * h->cap, h->stream));
  ODX_CUDA(h, cudaMemsetAsync(h->DEATH, 0, sizeof(double) * h->cap, h->stream));
  ODX_CUDA(h, cudaMemsetAsync(h->REPLI, 0, sizeof(double) * h->cap, h->stream));
  invalidate_state(h);
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));      /* the caller's host buffers may be reused as soon as we return */
  return refresh_global_n(h);
}

int odx_pop_download(odx_handle *h, int64_t first, int64_t count, double *y, double *expr, double *aux, double *psi,
                     int32_t *ids, double *death_rate, double *repli_rate) {
  if (first < 0 || count < 0 || first + count > h->n) { h->err = "odx_pop_download: range outside the population"; return ODX_EINVAL; }
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_TRY(download_cols(h, h->Y, y, NVAR, count, first));
  ODX_TRY(download_cols(h, h->E, expr, NEXPR, count, first));
  ODX_TRY(download_cols(h, h->AUX, aux, NAUX, count, first));
  ODX_TRY(download_cols(h, h->PSI, psi, NPSI, count, first));
  if (ids && count) ODX_TRY(d2h_pipelined(h, ids, h->ID + first, sizeof(int32_t) * count));
  if (death_rate && count) ODX_TRY(d2h_pipelined(h, death_rate, h->DEATH + first, sizeof(double) * count));
  if (repli_rate && count) ODX_TRY(d2h_pipelined(h, repli_rate, h->REPLI + first, sizeof(double) * count));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  return ODX_SUCCESS;
}

/* thinned download for trajectory output at scale: particles first, first+step, ... (m of them); only the selected rows cross
 * PCIe (at N = 2^24 a full traj.dat frame is 1.2 GB, src/datastruct.c:655-674) */
static int download_cols_strided(odx_handle *h, const double *soa, double *aos, int ncol, int64_t m, int64_t first, int64_t step) {
  if (ncol == 0 || m == 0 || aos == nullptr) return ODX_SUCCESS;
  const size_t bytes = sizeof(double) * (size_t)ncol * m;
  ODX_TRY(ensure_stage(h, bytes));
  k_soa_to_aos_strided<<<h->sm_count * 4, 256, 0, h->stream>>>(soa, h->stage_dev, ncol, m, h->cap, first, step);
  ODX_TRY(launch_check(h));
  ODX_CUDA(h, cudaMemcpyAsync(aos, h->stage_dev, bytes, cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  return ODX_SUCCESS;
}
int odx_pop_download_strided(odx_handle *h, int64_t first, int64_t step, int64_t m, double *y, double *expr, double *aux, double *psi,
                             int32_t *ids) {
  if (first < 0 || step < 1 || m < 0 || (m > 0 && first + (m - 1) * step >= h->n)) {
    h->err = "odx_pop_download_strided: selection outside the population"; return ODX_EINVAL;
  }
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_TRY(download_cols_strided(h, h->Y, y, NVAR, m, first, step));
  ODX_TRY(download_cols_strided(h, h->E, expr, NEXPR, m, first, step));
  ODX_TRY(download_cols_strided(h, h->AUX, aux, NAUX, m, first, step));
  ODX_TRY(download_cols_strided(h, h->PSI, psi, NPSI, m, first, step));
  if (ids && m > 0) {
    ODX_TRY(ensure_stage(h, sizeof(int32_t) * (size_t)m));
    k_gather_ids<<<h->sm_count * 4, 256, 0, h->stream>>>(h->ID, (int32_t *)h->stage_dev, m, first, step);
    ODX_TRY(launch_check(h));
    ODX_CUDA(h, cudaMemcpyAsync(ids, h->stage_dev, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, h->stream));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return ODX_SUCCESS;
}

int odx_pop_size(odx_handle *h, int64_t *n_local, int64_t *n_global, int32_t *max_id) {
  if (n_local) *n_local = h->n;
  if (n_global) *n_global = (int64_t)h->n_global;
  if (max_id) *max_id = h->max_id;
  return ODX_SUCCESS;
}

/* getpar(id) (src/datastruct.c): local position of the particle with this id, -1 when it is not on this GPU (dead, or owned by
 * another rank) */
int odx_pop_find_id(odx_handle *h, int32_t id, int64_t *pos) {
  if (!pos) return ODX_EINVAL;
  *pos = -1;
  if (h->n <= 0) return ODX_SUCCESS;
  ODX_CUDA(h, cudaSetDevice(h->device));
  long long *d = (long long *)h->rmax;                       /* 8-byte device scratch word, idle between steps */
  long long init = -1;
  ODX_CUDA(h, cudaMemcpyAsync(d, &init, sizeof init, cudaMemcpyHostToDevice, h->stream));
  k_find_id<<<grid_for(h, h->n), BLOCK, 0, h->stream>>>(h->ID, h->n, id, d);
  ODX_TRY(launch_check(h));
  ODX_CUDA(h, cudaMemcpyAsync(h->h_pin, d, sizeof init, cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  long long got; memcpy(&got, h->h_pin, sizeof got);
  *pos = (int64_t)got;
  return ODX_SUCCESS;
}

int odx_pop_init(odx_handle *h, int64_t n, double t, uint32_t seed, int replay) {
  if (n < 0 || n > h->cap) { h->err = "odx_pop_init: n exceeds capacity"; return ODX_EINVAL; }
  ODX_CUDA(h, cudaSetDevice(h->device));
  const odx_model_info *mi = odx_model(); (void)mi;
  if (replay) {
    /* reference order (src/odexp:1045-1069, methods_odexp.c:315-337): srand(seed) (main.c:403), then per particle in
     * list order the EXPR expressions followed by the INIT expressions, all drawing from the one serial stream */
    if (h->nranks > 1) { h->err = "replayed initial conditions are single-GPU (serial stream)"; return ODX_EUNSUP; }
    std::vector<double> y((size_t)A1(NVAR) * n), e((size_t)A1(NEXPR) * n);
    HostRand &hr = h->hr; hr.seed(seed); h->hr_seeded = true;
    for (int64_t i = 0; i < n; i++) {
      HostCtx c; memset(&c, 0, sizeof c);
      c.mu = h->h_mu; c.pop_size = (double)n; c.h = 0.0; c.hr = &hr; c.id = (int32_t)i; c.oid = (int32_t)i; c.atbirth = 1; c.ist0 = 1;
      double yy[A1(NVAR)], ee[A1(NEXPR)];
      for (int v = 0; v < A1(NVAR); v++) yy[v] = 0.0;
      for (int v = 0; v < A1(NEXPR); v++) ee[v] = 0.0;
      OdxModel::ic(c, t, yy, ee);
      for (int v = 0; v < NVAR; v++) y[(size_t)i * NVAR + v] = yy[v];
      for (int v = 0; v < NEXPR; v++) e[(size_t)i * NEXPR + v] = ee[v];
    }
    return odx_pop_upload(h, n, y.data(), e.data(), nullptr, (int32_t)n);
  }
  h->n = n;
  ODX_TRY(refresh_global_n(h));
  h->max_id = (int32_t)h->n_global;
  h->k0 = seed; h->k1 = 0;
  if (n > 0) {
    k_iota_ids<<<h->sm_count * 4, 256, 0, h->stream>>>(h->ID, n, (int32_t)h->n_before);
    ODX_TRY(launch_check(h));
    k_init<<<grid_for(h, n), BLOCK, 0, h->stream>>>(pop_view(h), 0, n, t, 0u, 1);
    ODX_TRY(launch_check(h));
  }
  invalidate_state(h);
  return ODX_SUCCESS;
}

int odx_pop_reinit(odx_handle *h, double t, int replay) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  const int64_t n = h->n;
  if (replay) {
    /* pop_init_conditions over the list in order, drawing from the running serial stream; ids, aux, psi and rates stay */
    if (h->nranks > 1) { h->err = "replayed initial conditions are single-GPU (serial stream)"; return ODX_EUNSUP; }
    if (!h->hr_seeded) { h->hr.seed(3141592u); h->hr_seeded = true; }
    std::vector<double> y((size_t)A1(NVAR) * n), e((size_t)A1(NEXPR) * n);
    std::vector<int32_t> ids((size_t)n + 1);
    ODX_TRY(download_cols(h, h->Y, y.data(), NVAR, n, 0));
    ODX_TRY(download_cols(h, h->E, e.data(), NEXPR, n, 0));
    if (n) ODX_CUDA(h, cudaMemcpyAsync(ids.data(), h->ID, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < n; i++) {
      HostCtx c; memset(&c, 0, sizeof c);
      c.mu = h->h_mu; c.pop_size = h->n_global; c.h = 0.0; c.hr = &h->hr; c.id = ids[i]; c.oid = ids[i];
      double yy[A1(NVAR)], ee[A1(NEXPR)];
      for (int v = 0; v < A1(NVAR); v++) yy[v] = v < NVAR ? y[(size_t)i * NVAR + v] : 0.0;
      for (int v = 0; v < A1(NEXPR); v++) ee[v] = v < NEXPR ? e[(size_t)i * NEXPR + v] : 0.0;
      OdxModel::ic(c, t, yy, ee);
      for (int v = 0; v < NVAR; v++) y[(size_t)i * NVAR + v] = yy[v];
      for (int v = 0; v < NEXPR; v++) e[(size_t)i * NEXPR + v] = ee[v];
    }
    ODX_TRY(upload_cols(h, y.data(), h->Y, NVAR, n, 0));
    ODX_TRY(upload_cols(h, e.data(), h->E, NEXPR, n, 0));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  } else if (n > 0) {
    k_init<<<grid_for(h, n), BLOCK, 0, h->stream>>>(pop_view(h), 0, n, t, h->reinit_index, 0);
    ODX_TRY(launch_check(h));
  }
  h->reinit_index++;
  /* the reference does not re-allocate its gsl evolve object here, so the next embedded step still starts from the
   * derivative left by the previous one (dydt_in <- dydt_out); keep that state, drop only the cached moments */
  const bool k1 = h->k1_valid; const uint64_t count = h->ev_count;
  invalidate_state(h);
  h->k1_valid = k1; h->ev_count = count;
  return ODX_SUCCESS;
}

int odx_stop_flag(odx_handle *h, int *raised) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  int *flag = reinterpret_cast<int *>(h->counter + 8);
  int *pin = reinterpret_cast<int *>(h->h_pin + NSLOT * NRED + 8);
  if (h->nranks > 1) {
#ifdef ODX_HAVE_NCCL_H
    ncclResult_t r = g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMax, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce(MAX): ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
#else
    h->err = "built without nccl.h"; return ODX_ENCCL;
#endif
  }
  ODX_CUDA(h, cudaMemcpyAsync(pin, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (raised) *raised = *pin;
  return ODX_SUCCESS;
}

int odx_event_record(odx_handle *h, int slot) {
  if (slot < 0 || slot >= 8) return ODX_EINVAL;
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (!h->ev[slot]) ODX_CUDA(h, cudaEventCreate(&h->ev[slot]));
  ODX_CUDA(h, cudaEventRecord(h->ev[slot], h->stream));
  return ODX_SUCCESS;
}
int odx_event_elapsed_ms(odx_handle *h, int a, int b, double *ms) {
  if (a < 0 || a >= 8 || b < 0 || b >= 8 || !h->ev[a] || !h->ev[b]) return ODX_EINVAL;
  ODX_CUDA(h, cudaEventSynchronize(h->ev[b]));
  float f = 0.f;
  ODX_CUDA(h, cudaEventElapsedTime(&f, h->ev[a], h->ev[b]));
  *ms = (double)f;
  return ODX_SUCCESS;
}
uint64_t odx_launch_count(const odx_handle *h) { return h->launches; }

double odx_draw_uniform(odx_handle *h, int replay) {
  if (replay) { if (!h->hr_seeded) { h->hr.seed(3141592u); h->hr_seeded = true; } return h->hr.rand01(); }
  PhiloxStream ps; ps.k0 = h->k0; ps.k1 = h->k1; ps.id = 0xFFFFFFFFu; ps.stream = h->host_draws++; ps.dom = 0x484f5354u; ps.slot = 0;
  return ps.rand01();
}
long odx_draw_poisson(odx_handle *h, double lambda, int replay) {
  if (replay) { if (!h->hr_seeded) { h->hr.seed(3141592u); h->hr_seeded = true; } return h->hr.poisson(lambda); }
  PhiloxStream ps; ps.k0 = h->k0; ps.k1 = h->k1; ps.id = 0xFFFFFFFEu; ps.stream = h->host_draws++; ps.dom = 0x484f5354u; ps.slot = 0;
  return ps.poisson(lambda);
}

int odx_rhs(odx_handle *h, double t) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  h->ymoments_valid = false;
  return do_rhs(h, t, 0.0);
}

/* what the adaptive low-rank expansion chose in the last RHS evaluation (lrexprank, src/lrexp.c:82-173): order p, centre, range */
int odx_lowrank_order(odx_handle *h, int *p, double *meanx, double *range) {
#ifdef ODX_MODEL_HAS_LOWRANK
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (!h->lr_plan) { h->err = "odx_lowrank_order: no RHS evaluation yet"; return ODX_EINVAL; }
  struct { double meanx, range; int p; } head;
  static_assert(offsetof(LrPlan, p) == 16, "LrPlan head layout");
  ODX_CUDA(h, cudaMemcpyAsync(&head, h->lr_plan, sizeof head, cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (p) *p = head.p;
  if (meanx) *meanx = head.meanx;
  if (range) *range = head.range;
  return ODX_SUCCESS;
#else
  (void)p; (void)meanx; (void)range; h->err = "odx_lowrank_order: the model has no low-rank %C coupling"; return ODX_EUNSUP;
#endif
}

int odx_meanfields(odx_handle *h, double *mfd) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_CUDA(h, cudaMemcpyAsync(h->h_pin, h->sums, sizeof(double) * NSLOT * NRED, cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int k = 0; k < NMFD; k++) mfd[k] = h->h_pin[(size_t)h->report_slot * NRED + k] / h->n_global;
  if (M::HAS_VECT_MFD) M::mfd_vect(h->h_mu, h->n_global, mfd);
  return ODX_SUCCESS;
}

int odx_rates(odx_handle *h, double *sum_rates, double *pop_birth_rate) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_CUDA(h, cudaMemcpyAsync(h->h_pin, h->sums, sizeof(double) * NSLOT * NRED, cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  /* SIM->pop_birth_rate = (sum of %BIRTH over particles) / POP_SIZE  (src/odexp:829-846) */
  const double birth = M::HAS_BIRTH && h->n_global > 0 ? h->h_pin[(size_t)h->report_slot * NRED + RED_BIRTH] / h->n_global : 0.0;
  const double rates = (M::HAS_DEATH || M::HAS_REPLI) ? h->h_pin[(size_t)h->rates_slot * NRED + RED_RATES] : 0.0;
  if (pop_birth_rate) *pop_birth_rate = birth;
  if (sum_rates) *sum_rates = birth + rates;       /* SSA_timestep: r = pop_birth_rate + sum(death + repli) */
  return ODX_SUCCESS;
}

/* fe_apply (src/methods_odexp.c:1790-1814): h is raised / clipped to land on tnext, then reset to h0 */
int odx_fe_apply(odx_handle *h, double *t, double tnext, double *hstep, double hmin, double h0, int want_outputs) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  const double dt = tnext - *t;
  if (*hstep > (dt - hmin)) *hstep = dt;
  ODX_TRY(chain_step(h, 1, *t, *hstep, want_outputs));
  *t += *hstep;
  *hstep = h0;
  return ODX_SUCCESS;
}
/* dde_apply (src/methods_odexp.c:1825-1874): classic RK4 with the same clipping / reset of h */
int odx_dde_apply(odx_handle *h, double *t, double tnext, double *hstep, double hmin, double h0, int want_outputs) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  const double dt = tnext - *t;
  if (*hstep > (dt - hmin)) *hstep = dt;
  ODX_TRY(chain_step(h, 4, *t, *hstep, want_outputs));
  *t += *hstep;
  *hstep = h0;
  return ODX_SUCCESS;
}

/* iteration_apply (src/methods_odexp.c:1816-1823) calls pop_rhs(t, y, y): input and output alias.  In the generated pop_rhs
 * that makes pass 1 overwrite y with f(y) evaluated at ZERO mean fields before the state is copied into the particles, the
 * %M walks then see those values, and pass 2 applies the map again with these mean fields.  Reproduced as two sweeps:
 *   y1 = f(y; mfd = 0),  mfd = mean(terms(y1)),  y <- f(y1; mfd). */
int odx_iteration_apply(odx_handle *h, double *t, int want_outputs) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (h->n > 0) {
    if (M::HAS_PAIR || M::HAS_LOWRANK) { h->err = "meth iteration with %C coupling is not supported"; return ODX_EUNSUP; }
    if (M::USES_KY) { h->err = "meth iteration with KY() is not supported (pop_rhs(t, y, y) aliases the list the neighbours are read from)"; return ODX_EUNSUP; }
    const int g = grid_chain(h, h->n);
    Pop p = pop_view(h);
    ChainArgs a; memset(&a, 0, sizeof a);
    a.h = 1.0; a.a = 1.0; a.w = 1.0; a.iterate = 1; a.t_eval = *t; a.t_next = *t; a.eval_index = h->eval_index;
    a.sums_in = nullptr; a.want_out = 0;
    k_chain<true, true><<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, SLOT_STAGE + 0));
    ODX_TRY(launch_check(h));
    if (NEEDS_MOMENTS) ODX_TRY(post_reduce(h, SLOT_STAGE + 0));
    a.sums_in = sums_ptr(h, SLOT_STAGE + 0); a.want_out = want_outputs;
    const int new_slot = h->y_slot ^ 1;
    k_chain<true, true><<<g, BLOCK, 0, h->stream>>>(p, a, red_view(h, new_slot));
    ODX_TRY(launch_check(h));
    if (NEEDS_MOMENTS || M::HAS_DEATH || M::HAS_REPLI) ODX_TRY(post_reduce(h, new_slot));
    h->eval_index += 2; h->rhs_evals += 1;
    h->report_slot = SLOT_STAGE + 0; h->rates_slot = new_slot; h->y_slot = new_slot; h->ymoments_valid = NEEDS_MOMENTS; h->k1_valid = false;
  }
  *t += 1;
  return ODX_SUCCESS;
}

int odx_run_fixed(odx_handle *h, int method, double *t, double hstep, int64_t nsteps) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (method != ODX_METH_FE && method != ODX_METH_DDE) { h->err = "odx_run_fixed: fixed-step methods are fe and dde"; return ODX_EINVAL; }
  if (can_persist(h) && h->n > 0 && nsteps > 0) {
    /* one cooperative launch cannot be interrupted: cut the run into launches of ~2e10 particle-stages (~0.1 s) and keep at
     * most two in flight, so that odx_request_abort takes effect within a few tenths of a second */
    const int nst = method == ODX_METH_FE ? 1 : 4;
    int64_t chunk = (int64_t)(2.0e10 / ((double)h->n * nst));
    if (chunk < 1) chunk = 1;
    const double t0 = *t;
    int64_t done = 0;
    for (int k = 0; done < nsteps; k++) {
      if (h->abort_req) break;
      const int64_t m = (nsteps - done < chunk) ? nsteps - done : chunk;
      if (k >= 2) ODX_CUDA(h, cudaEventSynchronize(h->chunk_ev[k & 1]));
      ODX_TRY(chain_run_persist(h, nst, t0 + hstep * (double)done, hstep, m, done + m == nsteps ? 1 : 0));
      if (nsteps > chunk) {
        if (!h->chunk_ev[k & 1]) ODX_CUDA(h, cudaEventCreateWithFlags(&h->chunk_ev[k & 1], cudaEventDisableTiming));
        ODX_CUDA(h, cudaEventRecord(h->chunk_ev[k & 1], h->stream));
      }
      done += m;
    }
    *t = t0 + hstep * (double)done;
    return ODX_SUCCESS;
  }
  for (int64_t s = 0; s < nsteps; s++) {
    if (h->abort_req) break;
    ODX_TRY(chain_step(h, method == ODX_METH_FE ? 1 : 4, *t, hstep, s == nsteps - 1));
    *t += hstep;
  }
  return ODX_SUCCESS;
}

void odx_request_abort(odx_handle *h) { if (h) h->abort_req = 1; }
int odx_abort_requested(const odx_handle *h) { return h ? h->abort_req : 0; }
void odx_clear_abort(odx_handle *h) { if (h) h->abort_req = 0; }

int odx_evolve_reset(odx_handle *h) { h->ev_count = 0; h->k1_valid = false; return ODX_SUCCESS; }
int odx_evolve_stats(odx_handle *h, uint64_t *count, uint64_t *failed_steps, uint64_t *rhs_evals) {
  if (count) *count = h->ev_count;
  if (failed_steps) *failed_steps = h->ev_failed;
  if (rhs_evals) *rhs_evals = h->rhs_evals;
  return ODX_SUCCESS;
}

/* gsl_odeiv2_evolve_apply + standard control (a_y = 1, a_dydt = 0), restated for the resident population.
 * Behaviour follows GSL's documented driver: reuse dydt_out as the next dydt_in, clip the step to reach t1,
 * on rmax > 1.1 shrink by max(0.2, 0.9 rmax^(-1/ord)) and retry, on rmax < 0.5 grow by min(5, max(1, 0.9 rmax^(-1/(ord+1)))). */
int odx_evolve_apply(odx_handle *h, int method, double *t, double t1, double *hstep) {
  ODX_TRY(ky_guard(h));
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (method != ODX_METH_RKF45 && method != ODX_METH_RKCK && method != ODX_METH_RK4 && method != ODX_METH_RK2 && method != ODX_METH_RK8PD) {
    h->err = "odx_evolve_apply: method must be rkf45, rkck, rk8pd, rk4 or rk2"; return ODX_EUNSUP;
  }
  if (h->n <= 0) { *t = t1; return ODX_SUCCESS; }
  const Tableau &T = (method == ODX_METH_RKF45) ? TAB_RKF45 : (method == ODX_METH_RK2) ? TAB_RK2 : (method == ODX_METH_RK8PD) ? TAB_RK8PD : TAB_RKCK;
  const unsigned order = (method == ODX_METH_RK4) ? 4u : T.order;
  const int ks_need = (method == ODX_METH_RK8PD) ? 12 : 5;        /* slope buffers k_2 .. k_nst */
  if (h->KS && h->ks_slots < ks_need) { ODX_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFree(h->KS); h->KS = nullptr; h->k1_valid = false; }
  if (!h->KS) { ODX_TRY(alloc_cols(h, &h->KS, ks_need * A1(NVAR))); h->ks_slots = ks_need; }
  if (!h->K1) { ODX_TRY(alloc_cols(h, &h->K1, NVAR)); ODX_TRY(alloc_cols(h, &h->DOUT, NVAR)); }
  if (method == ODX_METH_RK4 && !h->YONE) {
    ODX_TRY(alloc_cols(h, &h->YONE, NVAR)); ODX_TRY(alloc_cols(h, &h->YH, NVAR)); ODX_TRY(alloc_cols(h, &h->K1B, NVAR));
  }
  const double t0 = *t; double h0 = *hstep; const double dt = t1 - t0; int final_step = 0;
  if ((dt < 0.0 && h0 > 0.0) || (dt > 0.0 && h0 < 0.0)) { h->err = "step direction must match interval direction"; return ODX_EINVAL; }
  if (h->ev_count == 0) h->k1_valid = false;     /* first call: evaluate dydt_in = f(t0, y) */
  for (;;) {
    if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) { h0 = dt; final_step = 1; } else final_step = 0;
    double rmax = 0.0;
    if (method == ODX_METH_RK4) ODX_TRY(rk4_doubling_attempt(h, t0, h0, &rmax));
    else ODX_TRY(embedded_attempt(h, T, t0, h0, &rmax));
    if (rmax < DBL_MIN) rmax = DBL_MIN;
    const double h_old = h0; const double S = 0.9; int dec = 0;
    if (rmax > 1.1) {
      double r = S / pow(rmax, 1.0 / order);
      if (r < 0.2) r = 0.2;
      h0 = r * h_old; dec = 1;
    } else if (rmax < 0.5) {
      double r = S / pow(rmax, 1.0 / (order + 1.0));
      if (r > 5.0) r = 5.0;
      if (r < 1.0) r = 1.0;
      h0 = r * h_old;
    }
    const double tnew = final_step ? t1 : t0 + h_old;
    if (dec) {
      volatile double t_curr = tnew, t_next = tnew + h0;
      if (fabs(h0) < fabs(h_old) && t_next != t_curr) { h->ev_failed++; continue; }   /* y0 = Y is untouched: retry */
      *hstep = h0; h->err = "evolve: cannot reach the requested tolerance"; return ODX_FAILURE;
    }
    /* accept: y_new (YS) becomes the state, dydt_out becomes the next dydt_in */
    std::swap(h->Y, h->YS); std::swap(h->K1, h->DOUT);
    h->ev_count++;
    *t = tnew;
    if (!final_step) *hstep = h0;
    /* moments of the new Y were produced by the final stage kernel (slot STAGE+5) */
    if (NEEDS_MOMENTS) {
      ODX_CUDA(h, cudaMemcpyAsync(h->sums + (size_t)h->y_slot * NRED, h->sums + (size_t)(SLOT_STAGE + 5) * NRED, sizeof(double) * NRED,
                                  cudaMemcpyDeviceToDevice, h->stream));
      h->ymoments_valid = true; h->report_slot = h->y_slot;
    }
    return ODX_SUCCESS;
  }
}

int odx_ssa_apply(odx_handle *h, double t, double u, odx_event_row *row);
int odx_tauleap_apply(odx_handle *h, double t, double dt, uint64_t leap_index, odx_event_row *rows, int64_t max_rows, int64_t *nrows);
int odx_tauleap_apply_view(odx_handle *h, double t, double dt, uint64_t leap_index, const odx_event_row **rows, int64_t *nrows);

}  // extern "C"

#include "odx_events.cuh"
#include "odx_output.cuh"
