/* odx_events.cuh — stochastic birth / death on the resident SoA population.
 *
 * Reference behaviour restated for a parallel machine (src/methods_odexp.c:1556-1762, src/datastruct.c:416-563):
 *
 *  tau-leaping  one sweep over the list in order.  Per particle: u1 < death*dt flags DEATH, u2 < repli*dt flags REPLI;
 *               as coded in the reference ANY flag draws u3, u3 < death/(death+repli) kills the particle, otherwise the
 *               particle is cloned (daughter appended at the end of the list, copying state and rates, new id = max_id++,
 *               initial conditions run on the mother then on the daughter with the mother as `sister`) and THEN the mother
 *               is deleted.  Daughters appended during the sweep are visited later in the same sweep.  Finally
 *               poisson(birth_rate*dt) de-novo births are appended.  Every event writes one stats row.
 *               Here: one flag kernel + exclusive scan + spawn kernel per generation of daughters, then a stable
 *               prefix-scan compaction that preserves list order, ids by exclusive scan on max_id, and the event rows
 *               reconstructed in sweep order from two more scans.  Draws are keyed Philox: (particle id, slot 0..2,
 *               stream = event index, domain LEAP), so the outcome does not depend on the traversal.
 *  SSA          one event per call.  The normalised cumulative sum over [death_0..N-1, repli_0..N-1, birth] and the search
 *               `while (u > r[k]) k++` run on the device for keyed runs (k_ssa_tile_sums / k_ssa_pick: 16 bytes cross PCIe per
 *               event); runs on the replayed serial stream do them on the host from the downloaded rate columns in the reference's
 *               order of additions, so the chosen event is identical to the reference for the same u.  The list edit
 *               (delete / clone / append, initial conditions) runs on the device.
 */
#pragma once
#include <cooperative_groups.h>

namespace odx {

constexpr int SCAN_ITEMS = 4;                      /* items per thread */
constexpr int SCAN_TILE = BLOCK * SCAN_ITEMS;      /* items per block */
enum { OUT_NONE = 0, OUT_DIE = 1, OUT_REPL_DIE = 2, OUT_REPL_KEEP = 3 };

/* ---- three-phase exclusive scan of NCH interleaved int32 channels ----------------------------------------------- */
template <int NCH>
__global__ void __launch_bounds__(BLOCK) k_scan_tiles(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                       int32_t *__restrict__ tile_tot, int64_t n) {
  __shared__ int32_t s_warp[BLOCK / 32][NCH];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS][NCH], tsum[NCH];
#pragma unroll
  for (int c = 0; c < NCH; c++) tsum[c] = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) { v[k][c] = (base + k < n) ? in[(base + k) * NCH + c] : 0; tsum[c] += v[k][c]; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t incl[NCH];
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    int32_t x = tsum[c];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    incl[c] = x;
    if (lane == 31) s_warp[warp][c] = x;
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    int32_t woff = 0;
    for (int w2 = 0; w2 < warp; w2++) woff += s_warp[w2][c];
    int32_t run = woff + incl[c] - tsum[c];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[(base + k) * NCH + c] = run; run += v[k][c]; }
    if (threadIdx.x == BLOCK - 1) tile_tot[(int64_t)blockIdx.x * NCH + c] = woff + incl[c];
  }
}
template <int NCH>
__global__ void __launch_bounds__(BLOCK) k_scan_totals(int32_t *tile_tot, int ntiles, int32_t *grand) {
  /* one block: exclusive scan of the tile totals in place, grand totals to grand[NCH] */
  __shared__ int32_t s_warp[BLOCK / 32][NCH];
  __shared__ int32_t s_carry[NCH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NCH) s_carry[threadIdx.x] = 0;
  __syncthreads();
  for (int b0 = 0; b0 < ntiles; b0 += BLOCK) {
    const int i = b0 + threadIdx.x;
    int32_t mine[NCH], incl[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      mine[c] = (i < ntiles) ? tile_tot[(int64_t)i * NCH + c] : 0;
      int32_t x = mine[c];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      incl[c] = x;
      if (lane == 31) s_warp[warp][c] = x;
    }
    __syncthreads();
    int32_t tot[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      int32_t woff = 0, all = 0;
      for (int w2 = 0; w2 < BLOCK / 32; w2++) { if (w2 < warp) woff += s_warp[w2][c]; all += s_warp[w2][c]; }
      if (i < ntiles) tile_tot[(int64_t)i * NCH + c] = s_carry[c] + woff + incl[c] - mine[c];
      tot[c] = all;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int c = 0; c < NCH; c++) s_carry[c] += tot[c];
    }
    __syncthreads();
  }
  if (threadIdx.x < NCH) grand[threadIdx.x] = s_carry[threadIdx.x];
}
template <int NCH>
__global__ void __launch_bounds__(BLOCK) k_scan_add(int32_t *__restrict__ out, const int32_t *__restrict__ tile_off, int64_t n) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) {
#pragma unroll
      for (int c = 0; c < NCH; c++) out[(base + k) * NCH + c] += tile_off[(int64_t)blockIdx.x * NCH + c];
    }
  }
}

/* ---- tau-leap kernels ----------------------------------------------------------------------------------------- */
/* per-particle outcome for particles [first, first+count); flags3[i] = {replicates, rows, dies} for the scans */
__device__ __forceinline__ int leap_outcome(const Pop &p, int64_t i, double dt, uint32_t stream) {
  const double death = p.DEATH[i], repli = p.REPLI[i];
  PhiloxStream rng;
#pragma unroll
  for (int r = 0; r < 20; r++) rng.rk[r] = p.rk[r];
  rng.has_rk = 1;
  rng.k0 = p.k0; rng.k1 = p.k1; rng.id = (uint32_t)p.ID[i]; rng.stream = stream; rng.dom = DOM_LEAP; rng.slot = 0;
  const double u1 = rng.rand01(), u2 = rng.rand01();
  int o = OUT_NONE;
  if ((u1 < death * dt) || (u2 < repli * dt)) {
    const double u3 = rng.rand01();
    o = (u3 < death / (death + repli)) ? OUT_DIE : OUT_REPL_DIE;
  }
  return o;
}
__global__ void __launch_bounds__(BLOCK) k_leap_flags(Pop p, int64_t first, int64_t count, double dt, uint32_t stream,
                                                       int32_t *__restrict__ outcome, int32_t *__restrict__ repl_flag) {
  for (int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x; q < count; q += (int64_t)gridDim.x * BLOCK) {
    const int64_t i = first + q;
    const int o = leap_outcome(p, i, dt, stream);
    outcome[i] = o;
    repl_flag[q] = (o == OUT_REPL_DIE) ? 1 : 0;
  }
}

/* clone + initial conditions for every replicating particle of [first, first+count): daughter slot = dst0 + off,
 * daughter id = id0 + off.  keep_mother: SSA replication (the mother stays, with her initial conditions re-run). */
__device__ __forceinline__ void spawn_one(const Pop &p, int64_t i, int o, int64_t j, int32_t cid, double t, uint32_t stream,
                                          int32_t *__restrict__ child_id) {
  {
    const int32_t mid = p.ID[i];
    double my[A1(NVAR)], me[A1(NEXPR)], maux[A1(NAUX)], dy[A1(NVAR)], de[A1(NEXPR)];
    load_cols<NVAR>(p.Y, p.stride, i, my);
    load_cols<NEXPR>(p.E, p.stride, i, me);
    load_cols<NAUX>(p.AUX, p.stride, i, maux);
#pragma unroll
    for (int v = 0; v < A1(NVAR); v++) dy[v] = my[v];
#pragma unroll
    for (int v = 0; v < A1(NEXPR); v++) de[v] = me[v];
    /* mother first (sister == NULL), then the daughter with the mother as sister */
    DevCtx cm = make_ctx(p, 0.0, stream, DOM_IC + 1, mid);
    cm.atbirth = 0; cm.atrepli = 1; cm.has_sister = 0;
    M::ic(cm, t, my, me);
    DevCtx cd = make_ctx(p, 0.0, stream, DOM_IC + 2, cid);
    cd.atbirth = 0; cd.atrepli = 1; cd.has_sister = 1; cd.sis_y = my; cd.sis_e = me; cd.sis_aux = maux;
    M::ic(cd, t, dy, de);
    store_cols<NVAR>(p.Y, p.stride, j, dy);
    store_cols<NEXPR>(p.E, p.stride, j, de);
    store_cols<NAUX>(p.AUX, p.stride, j, maux);
    if (NPSI > 0) { double ps[A1(NPSI)]; load_cols<NPSI>(p.PSI, p.stride, i, ps); store_cols<NPSI>(p.PSI, p.stride, j, ps); }
    p.DEATH[j] = p.DEATH[i]; p.REPLI[j] = p.REPLI[i];
    p.ID[j] = cid;
    child_id[i] = cid;
    if (o == OUT_REPL_KEEP) { store_cols<NVAR>(p.Y, p.stride, i, my); store_cols<NEXPR>(p.E, p.stride, i, me); }
  }
}
__global__ void __launch_bounds__(BLOCK) k_spawn(Pop p, int64_t first, int64_t count, const int32_t *__restrict__ outcome,
                                                  const int32_t *__restrict__ off, int64_t dst0, int32_t id0, double t, uint32_t stream,
                                                  int32_t *__restrict__ child_id) {
  for (int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x; q < count; q += (int64_t)gridDim.x * BLOCK) {
    const int64_t i = first + q;
    const int o = outcome[i];
    if (o != OUT_REPL_DIE && o != OUT_REPL_KEEP) continue;
    spawn_one(p, i, o, dst0 + off[q], id0 + off[q], t, stream, child_id);
  }
}

/* de-novo births: slots [dst0, dst0+count), ids id0.. ; par_birth defaults then ode_init_conditions at birth */
__device__ __forceinline__ void birth_one(const Pop &p, int64_t j, int32_t id, double t, uint32_t stream) {
  {
    double y[A1(NVAR)], e[A1(NEXPR)], z[A1(NAUX)], zp[A1(NPSI)];
#pragma unroll
    for (int v = 0; v < A1(NVAR); v++) y[v] = 0.0;
#pragma unroll
    for (int v = 0; v < A1(NEXPR); v++) e[v] = 0.0;
#pragma unroll
    for (int v = 0; v < A1(NAUX); v++) z[v] = 0.0;
#pragma unroll
    for (int v = 0; v < A1(NPSI); v++) zp[v] = 0.0;
    DevCtx c = make_ctx(p, 0.0, stream, DOM_IC + 3, id);
    c.atbirth = 1;
    M::ic(c, t, y, e);
    store_cols<NVAR>(p.Y, p.stride, j, y);
    store_cols<NEXPR>(p.E, p.stride, j, e);
    store_cols<NAUX>(p.AUX, p.stride, j, z);
    store_cols<NPSI>(p.PSI, p.stride, j, zp);
    p.DEATH[j] = 0.0; p.REPLI[j] = 0.0; p.ID[j] = id;
  }
}
__global__ void __launch_bounds__(BLOCK) k_births(Pop p, int64_t dst0, int64_t count, int32_t id0, double t, uint32_t stream) {
  for (int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x; q < count; q += (int64_t)gridDim.x * BLOCK)
    birth_one(p, dst0 + q, id0 + (int32_t)q, t, stream);
}

/* scan inputs over the whole sweep: ch0 = keep, ch1 = rows written, ch2 = net deaths before (outcome DIE) */
__global__ void __launch_bounds__(BLOCK) k_sweep_flags(const int32_t *__restrict__ outcome, int32_t *__restrict__ f3, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * BLOCK) {
    const int o = outcome[i];
    f3[i * 3 + 0] = (o == OUT_NONE || o == OUT_REPL_KEEP) ? 1 : 0;
    f3[i * 3 + 1] = (o == OUT_DIE) ? 1 : ((o == OUT_REPL_DIE) ? 2 : (o == OUT_REPL_KEEP ? 1 : 0));
    f3[i * 3 + 2] = (o == OUT_DIE) ? 1 : 0;
  }
}

/* event rows in sweep order (fwrite_SIM rows of tau_leaping_apply_birthdeath) */
__global__ void __launch_bounds__(BLOCK) k_event_rows(const int32_t *__restrict__ outcome, const int32_t *__restrict__ s3,
                                                       const int32_t *__restrict__ ids, const int32_t *__restrict__ child_id,
                                                       int64_t n, int32_t n0, odx_event_row *__restrict__ rows, int64_t max_rows) {
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * BLOCK) {
    const int o = outcome[i];
    if (o == OUT_NONE) continue;
    const int64_t r = s3[i * 3 + 1];
    const int32_t nbefore = n0 - s3[i * 3 + 2];        /* replicate-then-die is net zero */
    if (o == OUT_DIE) {
      if (r < max_rows) { odx_event_row e = {nbefore - 1, ODX_STOP_DEATH, {ids[i], -1, -1}}; rows[r] = e; }
    } else {
      if (r < max_rows) { odx_event_row e = {nbefore + 1, ODX_STOP_REPLI, {ids[i], 1, child_id[i]}}; rows[r] = e; }
      if (o == OUT_REPL_DIE && r + 1 < max_rows) { odx_event_row e = {nbefore, ODX_STOP_DEATH, {ids[i], -1, -1}}; rows[r + 1] = e; }
    }
  }
}

/* stable compaction: survivor i moves to position s3[i].keep in the alternate buffers */
struct AltBufs { double *Y, *E, *AUX, *PSI, *DEATH, *REPLI; int32_t *ID; };
__global__ void __launch_bounds__(BLOCK) k_compact(Pop p, AltBufs a, const int32_t *__restrict__ outcome, const int32_t *__restrict__ s3,
                                                    int64_t nsweep, int64_t ntotal, int32_t kept_in_sweep) {
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < ntotal; i += (int64_t)gridDim.x * BLOCK) {
    int64_t j;
    if (i < nsweep) {
      const int o = outcome[i];
      if (!(o == OUT_NONE || o == OUT_REPL_KEEP)) continue;
      j = s3[i * 3 + 0];
    } else {
      j = (int64_t)kept_in_sweep + (i - nsweep);       /* de-novo births stay behind the sweep survivors */
    }
#pragma unroll
    for (int c = 0; c < NVAR; c++) a.Y[(int64_t)c * p.stride + j] = p.Y[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NEXPR; c++) a.E[(int64_t)c * p.stride + j] = p.E[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NAUX; c++) a.AUX[(int64_t)c * p.stride + j] = p.AUX[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NPSI; c++) a.PSI[(int64_t)c * p.stride + j] = p.PSI[(int64_t)c * p.stride + i];
    a.DEATH[j] = p.DEATH[i]; a.REPLI[j] = p.REPLI[i]; a.ID[j] = p.ID[i];
  }
}
/* ---- the whole tau-leaping sweep in ONE cooperative launch (single GPU) -----------------------------------------------------
 * The per-generation kernels above cost ~30 launches and one host round trip per generation of daughters (0.86 ms at N = 2^22
 * for ~0.1 ms of memory traffic).  Here every CTA owns a contiguous chunk of the range a phase walks, so that positions in list
 * order are (sum over the CTAs in front) + (running offset inside the chunk) + (block scan inside a tile of LEAP_T items): a
 * chunk total per CTA and one grid barrier replace the three-kernel scans, and the generation loop runs on the device.
 *   generation g over [first, end):  outcomes (keyed draws) + replicators per chunk | barrier | daughters spawned at
 *                                    end + position, ids max_id + position | barrier; next generation = the daughters
 *   de-novo births behind the sweep; survivors / rows / deaths per chunk | barrier |
 *   event rows in sweep order.  result[] = {survivors, rows, deaths, particles swept, max_id after the daughters, error}.
 * The stable compaction into the alternate buffers is a plain launch behind it (k_leap_compact, same chunks), so that the rows can
 * be copied to the host while it runs (written from the kernel straight into pinned host memory they cost 0.4 ms at 2 x 10^5 rows).
 * Same draws, ids, rows and list order as the per-generation path (tests/test_gpu_events.py compares the two). */
#ifndef LEAP_T_N
#define LEAP_T_N 512
#endif
constexpr int LEAP_T = LEAP_T_N;
constexpr int LEAP_MAX_CTAS = 1024;
struct LeapArgs {
  AltBufs a; double dt, t; uint32_t stream; int32_t n0, max_id0; int64_t nbirth;
  int32_t *outcome, *child, *cta_cnt;            /* cta_cnt: [2][LEAP_MAX_CTAS] replicators per chunk (generation parity), [LEAP_MAX_CTAS][3] sweep totals */
  odx_event_row *rows; int64_t rows_cap; long long *result;
};
__device__ __forceinline__ int leap_block_sum(int v, int *s_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < LEAP_T / 32; w++) t += s_w[w];
  return t;
}
/* exclusive scan of v over the block (thread order), total in `total` */
__device__ __forceinline__ unsigned leap_block_scan(unsigned v, unsigned &total, unsigned *s_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  __syncthreads();
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  unsigned woff = 0, all = 0;
#pragma unroll
  for (int w = 0; w < LEAP_T / 32; w++) { const unsigned q = s_w[w]; if (w < warp) woff += q; all += q; }
  total = all;
  return woff + x - v;
}
/* sum of cnt[0 .. nb) and of the entries in front of blockIdx.x (stride between entries) */
__device__ __forceinline__ void leap_cta_prefix(const int32_t *cnt, int stride, int nb, int &before, int &all, int *s_w) {
  int mine_b = 0, mine_a = 0;
  for (int b = threadIdx.x; b < nb; b += LEAP_T) { const int c = cnt[(size_t)b * stride]; mine_a += c; if (b < (int)blockIdx.x) mine_b += c; }
  before = leap_block_sum(mine_b, s_w);
  all = leap_block_sum(mine_a, s_w);
}
__global__ void __launch_bounds__(LEAP_T) k_leap_sweep(Pop p, LeapArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ int s_w[LEAP_T / 32];
  const int ncta = (int)gridDim.x;
  int64_t first = 0, end = p.n;
  int32_t max_id = a.max_id0;
  int gen = 0;
#ifdef LEAP_TIMING
  int mark = 8;
#define LEAP_MARK() do { if (blockIdx.x == 0 && threadIdx.x == 0 && mark < 40) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); a.result[mark++] = (long long)tt; } } while (0)
#else
#define LEAP_MARK() do { } while (0)
#endif
  LEAP_MARK();
  while (end > first) {
    const int64_t count = end - first;
    const int64_t chunk = ((count + ncta - 1) / ncta + LEAP_T - 1) / LEAP_T * LEAP_T;
    int64_t c0 = first + (int64_t)blockIdx.x * chunk, c1 = c0 + chunk;
    if (c0 > end) c0 = end;
    if (c1 > end) c1 = end;
    int32_t *cnt = a.cta_cnt + (size_t)(gen & 1) * LEAP_MAX_CTAS;
    int mine = 0;
    for (int64_t i = c0 + threadIdx.x; i < c1; i += LEAP_T) {
      const int o = leap_outcome(p, i, a.dt, a.stream);
      a.outcome[i] = o;
      mine += (o == OUT_REPL_DIE);
    }
    mine = leap_block_sum(mine, s_w);
    if (threadIdx.x == 0) cnt[blockIdx.x] = mine;
    LEAP_MARK();
    grid.sync();
    LEAP_MARK();
    int before, nd;
    leap_cta_prefix(cnt, 1, ncta, before, nd, s_w);
    if (end + nd > p.stride) { if (blockIdx.x == 0 && threadIdx.x == 0) a.result[5] = 1; return; }      /* the same decision in every CTA */
    if (mine > 0) {
      unsigned run = (unsigned)before;
      for (int64_t t0 = c0; t0 < c1; t0 += LEAP_T) {
        const int64_t i = t0 + threadIdx.x;
        const int o = (i < c1) ? a.outcome[i] : OUT_NONE;
        unsigned tot;
        const unsigned ex = leap_block_scan(o == OUT_REPL_DIE ? 1u : 0u, tot, (unsigned *)s_w);
        if (o == OUT_REPL_DIE) spawn_one(p, i, o, end + run + ex, max_id + (int32_t)(run + ex), a.t, a.stream, a.child);
        run += tot;
      }
    }
    LEAP_MARK();
    first = end; end += nd; max_id += nd; gen++;
    if (nd > 0) grid.sync();                       /* the daughters are the next generation */
  }
  const int64_t nsweep = end;
  if (nsweep + a.nbirth > p.stride) { if (blockIdx.x == 0 && threadIdx.x == 0) a.result[5] = 1; return; }
  for (int64_t q = (int64_t)blockIdx.x * LEAP_T + threadIdx.x; q < a.nbirth; q += (int64_t)ncta * LEAP_T)
    birth_one(p, nsweep + q, max_id + (int32_t)q, a.t, a.stream);
  /* survivors, rows and deaths per chunk of the whole sweep */
  const int64_t chunk = ((nsweep + ncta - 1) / ncta + LEAP_T - 1) / LEAP_T * LEAP_T;
  int64_t c0 = (int64_t)blockIdx.x * chunk, c1 = c0 + chunk;
  if (c0 > nsweep) c0 = nsweep;
  if (c1 > nsweep) c1 = nsweep;
  int32_t *cnt3 = a.cta_cnt + 2 * LEAP_MAX_CTAS;
  {
    int k = 0, r = 0, d = 0;
    for (int64_t i = c0 + threadIdx.x; i < c1; i += LEAP_T) {
      const int o = a.outcome[i];
      k += (o == OUT_NONE); r += (o == OUT_DIE) ? 1 : (o == OUT_REPL_DIE ? 2 : 0); d += (o == OUT_DIE);
    }
    k = leap_block_sum(k, s_w); r = leap_block_sum(r, s_w); d = leap_block_sum(d, s_w);
    if (threadIdx.x == 0) { cnt3[blockIdx.x * 3 + 0] = k; cnt3[blockIdx.x * 3 + 1] = r; cnt3[blockIdx.x * 3 + 2] = d; }
  }
  LEAP_MARK();
  grid.sync();
  LEAP_MARK();
  int kb, ka, rb, ra, db, da;
  leap_cta_prefix(cnt3 + 0, 3, ncta, kb, ka, s_w);
  leap_cta_prefix(cnt3 + 1, 3, ncta, rb, ra, s_w);
  leap_cta_prefix(cnt3 + 2, 3, ncta, db, da, s_w);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.result[0] = ka; a.result[1] = ra; a.result[2] = da; a.result[3] = nsweep; a.result[4] = max_id; a.result[5] = 0;
  }
  if (a.rows == nullptr) return;
  /* event rows in sweep order (fwrite_SIM rows of tau_leaping_apply_birthdeath) */
  int64_t rows_run = rb; int32_t dead_run = db;
  for (int64_t t0 = c0; t0 < c1; t0 += LEAP_T) {
    const int64_t i = t0 + threadIdx.x;
    const int o = (i < c1) ? a.outcome[i] : -1;
    const unsigned rf = (o == OUT_DIE) ? 1u : (o == OUT_REPL_DIE ? 2u : 0u), df = (o == OUT_DIE) ? 1u : 0u;
    unsigned tot;
    const unsigned ex = leap_block_scan(rf | (df << 16), tot, (unsigned *)s_w);
    if (rf) {
      const int64_t r = rows_run + (ex & 0xffffu);
      const int32_t nbefore = a.n0 - (dead_run + (int32_t)(ex >> 16));
      const int32_t id = p.ID[i];
      if (o == OUT_DIE) {
        if (r < a.rows_cap) { odx_event_row e = {nbefore - 1, ODX_STOP_DEATH, {id, -1, -1}}; a.rows[r] = e; }
      } else {
        if (r < a.rows_cap) { odx_event_row e = {nbefore + 1, ODX_STOP_REPLI, {id, 1, a.child[i]}}; a.rows[r] = e; }
        if (r + 1 < a.rows_cap) { odx_event_row e = {nbefore, ODX_STOP_DEATH, {id, -1, -1}}; a.rows[r + 1] = e; }
      }
    }
    rows_run += tot & 0xffffu; dead_run += (int32_t)(tot >> 16);
  }
  LEAP_MARK();
}
/* stable compaction of the swept list into the alternate buffers: same chunks as k_leap_sweep (same grid), survivors in front of
 * a chunk from its per-chunk totals; a plain launch behind the cooperative one, so that the event rows can travel to the host
 * while it runs */
__global__ void __launch_bounds__(LEAP_T) k_leap_compact(Pop p, LeapArgs a) {
  __shared__ int s_w[LEAP_T / 32];
  if (a.result[5] != 0) return;
  const int ncta = (int)gridDim.x;
  const int64_t nsweep = a.result[3];
  const int64_t chunk = ((nsweep + ncta - 1) / ncta + LEAP_T - 1) / LEAP_T * LEAP_T;
  int64_t c0 = (int64_t)blockIdx.x * chunk, c1 = c0 + chunk;
  if (c0 > nsweep) c0 = nsweep;
  if (c1 > nsweep) c1 = nsweep;
  int kb, ka;
  leap_cta_prefix(a.cta_cnt + 2 * LEAP_MAX_CTAS, 3, ncta, kb, ka, s_w);
  int64_t keep_run = kb;
  for (int64_t t0 = c0; t0 < c1; t0 += LEAP_T) {
    const int64_t i = t0 + threadIdx.x;
    const int o = (i < c1) ? a.outcome[i] : -1;
    unsigned tot;
    const unsigned ex = leap_block_scan(o == OUT_NONE ? 1u : 0u, tot, (unsigned *)s_w);
    if (o == OUT_NONE) {
      const int64_t j = keep_run + ex;
#pragma unroll
      for (int c = 0; c < NVAR; c++) a.a.Y[(int64_t)c * p.stride + j] = p.Y[(int64_t)c * p.stride + i];
#pragma unroll
      for (int c = 0; c < NEXPR; c++) a.a.E[(int64_t)c * p.stride + j] = p.E[(int64_t)c * p.stride + i];
#pragma unroll
      for (int c = 0; c < NAUX; c++) a.a.AUX[(int64_t)c * p.stride + j] = p.AUX[(int64_t)c * p.stride + i];
#pragma unroll
      for (int c = 0; c < NPSI; c++) a.a.PSI[(int64_t)c * p.stride + j] = p.PSI[(int64_t)c * p.stride + i];
      a.a.DEATH[j] = p.DEATH[i]; a.a.REPLI[j] = p.REPLI[i]; a.a.ID[j] = p.ID[i];
    }
    keep_run += tot;
  }
  /* de-novo births stay behind the survivors of the sweep */
  for (int64_t q = (int64_t)blockIdx.x * LEAP_T + threadIdx.x; q < a.nbirth; q += (int64_t)ncta * LEAP_T) {
    const int64_t i = nsweep + q, j = (int64_t)ka + q;
#pragma unroll
    for (int c = 0; c < NVAR; c++) a.a.Y[(int64_t)c * p.stride + j] = p.Y[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NEXPR; c++) a.a.E[(int64_t)c * p.stride + j] = p.E[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NAUX; c++) a.a.AUX[(int64_t)c * p.stride + j] = p.AUX[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NPSI; c++) a.a.PSI[(int64_t)c * p.stride + j] = p.PSI[(int64_t)c * p.stride + i];
    a.a.DEATH[j] = p.DEATH[i]; a.a.REPLI[j] = p.REPLI[i]; a.a.ID[j] = p.ID[i];
  }
}
__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

/* ---- SSA event pick on the device (SSA_apply_birthdeath + ncumsum, src/methods_odexp.c:1556-1660, 1764-1777) ---------------
 * The reference normalises the running sum of [death_0..N-1, repli_0..N-1, birth] by its total and takes the first index with
 * u <= r[idx] (`while (u > r[idx]) idx++`).  Here: per-tile sums of the logical array (fixed-order tree inside a tile), one block
 * that scans the tile sums, finds the tile and then the element with the same comparison u > cum / total.  The running sum is
 * formed hierarchically, so it can differ from the serial one in the last bits: a different event only if u falls within
 * round-off of a boundary.  Parity runs (replayed serial stream) therefore keep the host path, which reproduces the serial
 * arithmetic exactly; keyed runs use this one: two 8-byte words cross PCIe instead of 2N doubles. */
constexpr int SSA_TILE = 4096;
__device__ __forceinline__ double ssa_elem(const double *__restrict__ death, const double *__restrict__ repli, int64_t N, double birth, int64_t i) {
  return i < N ? death[i] : (i < 2 * N ? repli[i - N] : (i == 2 * N ? birth : 0.0));
}
__global__ void __launch_bounds__(BLOCK) k_ssa_tile_sums(const double *__restrict__ death, const double *__restrict__ repli, int64_t N, double birth,
                                                          double *__restrict__ tile_sum) {
  __shared__ double s_w[BLOCK / 32];
  const int64_t base = (int64_t)blockIdx.x * SSA_TILE + (int64_t)threadIdx.x * (SSA_TILE / BLOCK);
  double v = 0.0;
#pragma unroll
  for (int k = 0; k < SSA_TILE / BLOCK; k++) v += ssa_elem(death, repli, N, birth, base + k);
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < BLOCK / 32; w++) t += s_w[w]; tile_sum[blockIdx.x] = t; }
}
/* out[0] = chosen index (as a double-sized integer word), out[1] = total rate */
__global__ void __launch_bounds__(BLOCK) k_ssa_pick(const double *__restrict__ death, const double *__restrict__ repli, int64_t N, double birth,
                                                     const double *__restrict__ tile_sum, int ntiles, double u, long long *__restrict__ out_idx,
                                                     double *__restrict__ out_total) {
  __shared__ double s_total, s_prefix, s_thr[BLOCK];
  __shared__ int s_tile;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int t = 0; t < ntiles; t++) tot += tile_sum[t];
    int t = 0; double pre = 0.0;
    if (tot > 0.0) {
      while (t < ntiles - 1 && u > (pre + tile_sum[t]) / tot) { pre += tile_sum[t]; t++; }
    }
    s_total = tot; s_prefix = pre; s_tile = t;
  }
  __syncthreads();
  const double tot = s_total;
  if (!(tot > 0.0)) { if (threadIdx.x == 0) { *out_idx = -1; *out_total = tot; } return; }
  constexpr int PER = SSA_TILE / BLOCK;
  const int64_t base = (int64_t)s_tile * SSA_TILE + (int64_t)threadIdx.x * PER;
  double x[PER], loc = 0.0;
#pragma unroll
  for (int k = 0; k < PER; k++) { x[k] = ssa_elem(death, repli, N, birth, base + k); loc += x[k]; }
  s_thr[threadIdx.x] = loc;
  __syncthreads();
  if (threadIdx.x == 0) {
    /* first thread whose running end reaches u, then the element inside it; the last logical element (2N) ends the search */
    double cum = s_prefix; int th = 0;
    const int64_t last = 2 * N;
    while (th < BLOCK - 1 && u > (cum + s_thr[th]) / tot && (int64_t)s_tile * SSA_TILE + (int64_t)(th + 1) * PER <= last) { cum += s_thr[th]; th++; }
    s_prefix = cum; s_tile = th;
  }
  __syncthreads();
  if (threadIdx.x == s_tile) {
    double cum = s_prefix; int k = 0;
    while (k < PER - 1 && base + k < 2 * N && u > (cum + x[k]) / tot) { cum += x[k]; k++; }
    long long idx = base + k;
    if (idx > 2 * N) idx = 2 * N;
    *out_idx = idx; *out_total = tot;
  }
}

/* ---- host side -------------------------------------------------------------------------------------------------- */
struct EventBufs {
  double *ssa_tiles = nullptr; int64_t ssa_tiles_cap = 0; long long *ssa_out = nullptr;
  int32_t *xch = nullptr; void *xch_pin = nullptr;                  /* count exchanges of the sharded sweep */
  int32_t *cta_cnt = nullptr; long long *leap_result = nullptr; int leap_grid = 0;      /* k_leap_sweep */
  cudaStream_t copy_stream = nullptr; cudaEvent_t ev_swept = nullptr;
  int32_t *outcome = nullptr, *flag = nullptr, *scan = nullptr, *tile = nullptr, *grand = nullptr, *child = nullptr, *f3 = nullptr, *s3 = nullptr;
  odx_event_row *rows = nullptr; int64_t rows_cap = 0;
  odx_event_row *rows_pin = nullptr; int64_t rows_pin_cap = 0;      /* pinned host copy of the event rows of the last sweep */
  AltBufs alt = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ready = false;
};

template <int NCH>
static int exclusive_scan(odx_handle *h, const int32_t *in, int32_t *out, int64_t n, int32_t *tile, int32_t *grand, int32_t *grand_host,
                          bool readback = true) {
  const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  if (n > 0) {
    k_scan_tiles<NCH><<<ntiles, BLOCK, 0, h->stream>>>(in, out, tile, n);
    ODX_TRY(launch_check(h));
  }
  k_scan_totals<NCH><<<1, BLOCK, 0, h->stream>>>(tile, n > 0 ? ntiles : 0, grand);
  ODX_TRY(launch_check(h));
  if (n > 0) {
    k_scan_add<NCH><<<ntiles, BLOCK, 0, h->stream>>>(out, tile, n);
    ODX_TRY(launch_check(h));
  }
  if (!readback) return ODX_SUCCESS;      /* the totals stay in grand[] on the device (the sharded sweep allgathers them from there) */
  int32_t *pin = reinterpret_cast<int32_t *>(h->h_pin + NSLOT * NRED + 12);       /* pinned: a pageable target makes the copy a staged, blocking one */
  ODX_CUDA(h, cudaMemcpyAsync(pin, grand, sizeof(int32_t) * NCH, cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int c = 0; c < NCH; c++) grand_host[c] = pin[c];
  return ODX_SUCCESS;
}

}  // namespace odx

struct odx_event_state { odx::EventBufs b; uint32_t event_index = 0; };
static odx_event_state *event_state(odx_handle *h) {
  if (!h->ev_state) h->ev_state = new odx_event_state();
  return h->ev_state;
}
static void odx_events_free(odx_handle *h) {
  if (!h->ev_state) return;
  odx::EventBufs &b = h->ev_state->b;
  void *ptrs[] = {b.outcome, b.flag, b.scan, b.tile, b.grand, b.child, b.f3, b.s3, b.rows, b.alt.Y, b.alt.E, b.alt.AUX, b.alt.PSI,
                  b.alt.DEATH, b.alt.REPLI, b.alt.ID, b.ssa_tiles, b.ssa_out, b.xch, b.cta_cnt, b.leap_result};
  for (void *q : ptrs) if (q) cudaFree(q);
  if (b.rows_pin) cudaFreeHost(b.rows_pin);
  if (b.xch_pin) cudaFreeHost(b.xch_pin);
  if (b.copy_stream) cudaStreamDestroy(b.copy_stream);
  if (b.ev_swept) cudaEventDestroy(b.ev_swept);
  delete h->ev_state;
  h->ev_state = nullptr;
}

namespace odx {

static int ensure_event_bufs(odx_handle *h, EventBufs &b) {
  if (b.ready) return ODX_SUCCESS;
  const int64_t cap = h->cap;
  const int64_t ntile = (cap + SCAN_TILE - 1) / SCAN_TILE + 1;
  ODX_CUDA(h, cudaMalloc((void **)&b.outcome, sizeof(int32_t) * cap));
  ODX_CUDA(h, cudaMalloc((void **)&b.flag, sizeof(int32_t) * cap));
  ODX_CUDA(h, cudaMalloc((void **)&b.scan, sizeof(int32_t) * cap));
  ODX_CUDA(h, cudaMalloc((void **)&b.child, sizeof(int32_t) * cap));
  ODX_CUDA(h, cudaMalloc((void **)&b.f3, sizeof(int32_t) * 3 * cap));
  ODX_CUDA(h, cudaMalloc((void **)&b.s3, sizeof(int32_t) * 3 * cap));
  ODX_CUDA(h, cudaMalloc((void **)&b.tile, sizeof(int32_t) * 3 * ntile));
  ODX_CUDA(h, cudaMalloc((void **)&b.grand, sizeof(int32_t) * 4));
  ODX_TRY(alloc_cols(h, &b.alt.Y, NVAR)); ODX_TRY(alloc_cols(h, &b.alt.E, NEXPR)); ODX_TRY(alloc_cols(h, &b.alt.AUX, NAUX));
  ODX_TRY(alloc_cols(h, &b.alt.PSI, NPSI)); ODX_TRY(alloc_cols(h, &b.alt.DEATH, 1)); ODX_TRY(alloc_cols(h, &b.alt.REPLI, 1));
  ODX_CUDA(h, cudaMalloc((void **)&b.alt.ID, sizeof(int32_t) * cap));
  b.ready = true;
  return ODX_SUCCESS;
}

/* compaction + bookkeeping shared by tau-leap and SSA.  nsweep = particles covered by `outcome`, nbirth appended after.
 * The event rows land in a pinned host buffer owned by the library (a pageable target turns the 20-byte-per-event copy into a staged
 * one at a few GB/s); `rows` (caller's memory, may be NULL) receives a copy, `view_out` (may be NULL) the pinned pointer itself. */
static int finish_events(odx_handle *h, EventBufs &b, int64_t nsweep, int64_t nbirth, int32_t n0, odx_event_row *rows, int64_t max_rows,
                         int64_t *nrows_out, int32_t *g3, const odx_event_row **view_out = nullptr) {
  Pop p = pop_view(h);
  const int g = grid_for(h, nsweep > 0 ? nsweep : 1);
  if (nsweep > 0) { k_sweep_flags<<<g, BLOCK, 0, h->stream>>>(b.outcome, b.f3, nsweep); ODX_TRY(launch_check(h)); }
  ODX_TRY(exclusive_scan<3>(h, b.f3, b.s3, nsweep, b.tile, b.grand, g3));
  const int32_t kept = g3[0], nrows = g3[1];
  int64_t written = 0;
  const bool want = (rows && max_rows > 0) || view_out;
  if (view_out) max_rows = (int64_t)nrows + nbirth;
  if (want && (int64_t)nrows + nbirth > b.rows_pin_cap) {
    if (b.rows_pin) cudaFreeHost(b.rows_pin);
    b.rows_pin = nullptr; b.rows_pin_cap = 0;
    const int64_t cap2 = 2 * ((int64_t)nrows + nbirth) + 1024;
    ODX_CUDA(h, cudaMallocHost((void **)&b.rows_pin, sizeof(odx_event_row) * (size_t)cap2));
    b.rows_pin_cap = cap2;
  }
  if (want && nrows > 0) {
    if (b.rows_cap < nrows) {
      if (b.rows) cudaFree(b.rows);
      b.rows = nullptr; b.rows_cap = 0;
      ODX_CUDA(h, cudaMalloc((void **)&b.rows, sizeof(odx_event_row) * (size_t)(nrows + 1024)));
      b.rows_cap = nrows + 1024;
    }
    k_event_rows<<<g, BLOCK, 0, h->stream>>>(b.outcome, b.s3, h->ID, b.child, nsweep, n0, b.rows, b.rows_cap);
    ODX_TRY(launch_check(h));
    written = nrows < max_rows ? nrows : max_rows;
    ODX_CUDA(h, cudaMemcpyAsync(b.rows_pin, b.rows, sizeof(odx_event_row) * (size_t)written, cudaMemcpyDeviceToHost, h->stream));
    /* no synchronisation here: the copy overlaps the compaction below, the stream is synchronised before the rows are read */
  }
  const int64_t sweep_rows = written;
  /* births: rows {n+1.., STOP_BIRTH, {-1, +1, id}} follow the sweep rows */
  const int32_t n_after_sweep = n0 - g3[2];
  const int32_t first_birth_id = h->max_id - (int32_t)nbirth;
  for (int64_t k = 0; k < nbirth; k++) {
    if (want && written < max_rows) {
      odx_event_row e = {n_after_sweep + (int32_t)k + 1, ODX_STOP_BIRTH, {-1, 1, first_birth_id + (int32_t)k}};
      b.rows_pin[written++] = e;
    }
  }
  (void)sweep_rows;
  const int64_t ntotal = nsweep + nbirth;
  if (ntotal > 0) {
    k_compact<<<grid_for(h, ntotal), BLOCK, 0, h->stream>>>(p, b.alt, b.outcome, b.s3, nsweep, ntotal, kept);
    ODX_TRY(launch_check(h));
  }
  std::swap(h->Y, b.alt.Y); std::swap(h->E, b.alt.E); std::swap(h->AUX, b.alt.AUX); std::swap(h->PSI, b.alt.PSI);
  std::swap(h->DEATH, b.alt.DEATH); std::swap(h->REPLI, b.alt.REPLI); std::swap(h->ID, b.alt.ID);
  h->n = kept + nbirth;
  invalidate_state(h);
  ODX_TRY(refresh_global_n(h));
  if (want) {
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));                 /* the rows have landed in the pinned buffer */
    if (rows && written > 0) memcpy(rows, b.rows_pin, sizeof(odx_event_row) * (size_t)written);
    if (view_out) *view_out = b.rows_pin;
  }
  if (nrows_out) *nrows_out = written;
  return ODX_SUCCESS;
}

}  // namespace odx

/* ---- multi-GPU birth / death: gather -> replicated sweep -> re-shard -----------------------------------------------------
 * The sweep is O(N) light work next to the steps between events, and its semantics are those of ONE ordered list (ids by
 * sweep order, daughters visited later in the same sweep).  Every rank therefore assembles the whole ordered population
 * (allgather of all columns, rank order = list order), runs the single-GPU sweep on it — keyed draws make the result
 * identical on every rank — and keeps the contiguous shard [lo, hi) of the new list that the balanced partition assigns to
 * it.  Counts, ids and event rows are those of the single-GPU run by construction, and the population is rebalanced across
 * the GPUs after every sweep. */
namespace odx {
#ifdef ODX_HAVE_NCCL_H
static int gather_column(odx_handle *h, const void *src, void *dst, size_t elem, int64_t nmax, ncclDataType_t dt) {
  char *sendbuf = (char *)h->gcol + (size_t)h->nranks * nmax * 8;
  ODX_CUDA(h, cudaMemcpyAsync(sendbuf, src, (size_t)h->n * elem, cudaMemcpyDeviceToDevice, h->stream));
  ncclResult_t r = g_nccl.AllGather(sendbuf, h->gcol, (size_t)nmax, dt, h->comm, h->stream);
  if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return ODX_ENCCL; }
  int64_t off = 0;
  for (int q = 0; q < h->nranks; q++) {
    if (h->all_n[q] > 0)
      ODX_CUDA(h, cudaMemcpyAsync((char *)dst + (size_t)off * elem, (char *)h->gcol + (size_t)q * nmax * elem, (size_t)h->all_n[q] * elem,
                                  cudaMemcpyDeviceToDevice, h->stream));
    off += h->all_n[q];
  }
  return ODX_SUCCESS;
}
static int replica_gather(odx_handle *h) {
  if (!h->comm) { h->err = "multi-GPU birth/death needs odx_comm_init (NCCL allgather)"; return ODX_ENCCL; }
  const int64_t N = (int64_t)h->n_global;
  int64_t nmax = 1;
  for (int q = 0; q < h->nranks; q++) nmax = h->all_n[q] > nmax ? h->all_n[q] : nmax;
  const int64_t want = 2 * N + 4096;
  if (!h->replica || h->replica->cap < N + N / 4 + 1024) {
    if (h->replica) odx_destroy(h->replica);
    h->replica = nullptr;
    int st = odx_create(&h->replica, h->device, want);
    if (st != ODX_SUCCESS) { h->err = std::string("replica: ") + (h->replica ? h->replica->err : "allocation failed"); return st; }
  }
  if (h->gcol_cap < nmax) {
    if (h->gcol) cudaFree(h->gcol);
    h->gcol = nullptr; h->gcol_cap = 0;
    ODX_CUDA(h, cudaMalloc(&h->gcol, (size_t)(h->nranks + 1) * nmax * 8));
    h->gcol_cap = nmax;
  }
  odx_handle *g = h->replica;
  ODX_CUDA(h, cudaStreamSynchronize(g->stream));
  for (int c = 0; c < NVAR; c++) ODX_TRY(gather_column(h, h->Y + (size_t)c * h->cap, g->Y + (size_t)c * g->cap, 8, nmax, ncclDouble));
  for (int c = 0; c < NEXPR; c++) ODX_TRY(gather_column(h, h->E + (size_t)c * h->cap, g->E + (size_t)c * g->cap, 8, nmax, ncclDouble));
  for (int c = 0; c < NAUX; c++) ODX_TRY(gather_column(h, h->AUX + (size_t)c * h->cap, g->AUX + (size_t)c * g->cap, 8, nmax, ncclDouble));
  for (int c = 0; c < NPSI; c++) ODX_TRY(gather_column(h, h->PSI + (size_t)c * h->cap, g->PSI + (size_t)c * g->cap, 8, nmax, ncclDouble));
  ODX_TRY(gather_column(h, h->DEATH, g->DEATH, 8, nmax, ncclDouble));
  ODX_TRY(gather_column(h, h->REPLI, g->REPLI, 8, nmax, ncclDouble));
  ODX_TRY(gather_column(h, h->ID, g->ID, 4, nmax, ncclInt32));
  ODX_CUDA(h, cudaMemcpyAsync(g->sums, h->sums, sizeof(double) * NSLOT * NRED, cudaMemcpyDeviceToDevice, h->stream));
  ODX_CUDA(h, cudaMemcpyAsync(g->d_mu, h->d_mu, sizeof(double) * A1(NPAR), cudaMemcpyDeviceToDevice, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  g->n = N; g->n_global = (double)N; g->all_n[0] = N; g->max_id = h->max_id; g->k0 = h->k0; g->k1 = h->k1;
  g->report_slot = h->report_slot; g->rates_slot = h->rates_slot;
  g->hr = h->hr; g->hr_seeded = h->hr_seeded; g->host_draws = h->host_draws;
  for (int k = 0; k < NPAR; k++) g->h_mu[k] = h->h_mu[k];
  return ODX_SUCCESS;
}
static int replica_scatter(odx_handle *h) {
  odx_handle *g = h->replica;
  ODX_CUDA(h, cudaStreamSynchronize(g->stream));
  const int64_t Nn = g->n, base = Nn / h->nranks, extra = Nn % h->nranks;
  const int64_t lo = h->rank * base + (h->rank < extra ? h->rank : extra), cnt = base + (h->rank < extra ? 1 : 0);
  if (cnt > h->cap) { h->err = "birth/death: this rank's shard exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
  if (cnt > 0) {
    for (int c = 0; c < NVAR; c++) ODX_CUDA(h, cudaMemcpyAsync(h->Y + (size_t)c * h->cap, g->Y + (size_t)c * g->cap + lo, cnt * 8, cudaMemcpyDeviceToDevice, h->stream));
    for (int c = 0; c < NEXPR; c++) ODX_CUDA(h, cudaMemcpyAsync(h->E + (size_t)c * h->cap, g->E + (size_t)c * g->cap + lo, cnt * 8, cudaMemcpyDeviceToDevice, h->stream));
    for (int c = 0; c < NAUX; c++) ODX_CUDA(h, cudaMemcpyAsync(h->AUX + (size_t)c * h->cap, g->AUX + (size_t)c * g->cap + lo, cnt * 8, cudaMemcpyDeviceToDevice, h->stream));
    for (int c = 0; c < NPSI; c++) ODX_CUDA(h, cudaMemcpyAsync(h->PSI + (size_t)c * h->cap, g->PSI + (size_t)c * g->cap + lo, cnt * 8, cudaMemcpyDeviceToDevice, h->stream));
    ODX_CUDA(h, cudaMemcpyAsync(h->DEATH, g->DEATH + lo, cnt * 8, cudaMemcpyDeviceToDevice, h->stream));
    ODX_CUDA(h, cudaMemcpyAsync(h->REPLI, g->REPLI + lo, cnt * 8, cudaMemcpyDeviceToDevice, h->stream));
    ODX_CUDA(h, cudaMemcpyAsync(h->ID, g->ID + lo, cnt * 4, cudaMemcpyDeviceToDevice, h->stream));
  }
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n = cnt; h->max_id = g->max_id; h->hr = g->hr; h->hr_seeded = g->hr_seeded; h->host_draws = g->host_draws;
  invalidate_state(h);
  return refresh_global_n(h);
}

/* ---- sharded tau-leaping sweep ---------------------------------------------------------------------------------------------
 * The sweep has the semantics of ONE ordered list: generation 0 = the particles in list order, generation g + 1 = the daughters of
 * generation g in their mothers' order, appended at the end and visited later in the same sweep; ids and the population size in
 * every event row follow that order.  With the list sharded by contiguous ranges, the global order of generation g is
 * (rank 0's part, rank 1's part, ...), so every position-dependent quantity is a local prefix plus the totals of the
 * (generation, rank) segments in front: each rank flags and spawns its own particles (keyed draws: by particle id, the same on any
 * GPU), and the ranks exchange only counts -- one allgather of one int per generation (daughter ids), one allgather of 3 ints per
 * generation at the end (survivors, rows, deaths: the event rows' positions and population sizes).  The rows of all ranks are
 * combined by an allreduce over the (disjoint) row positions.  The new list -- survivors of generation 0, 1, ... in global order,
 * then the de-novo births -- is cut into balanced contiguous shards again and moved with grouped ncclSend / ncclRecv of the
 * contiguous pieces of every column: the rebalancing of SURVEY 8(e).  Result identical to the single-GPU sweep (tests/mgpu). */
constexpr int SHARD_MAXGEN = 24;
struct SegInfo { int ngen; int64_t first[SHARD_MAXGEN + 1]; int32_t rows_off[SHARD_MAXGEN], deaths_off[SHARD_MAXGEN]; };
__global__ void k_seg_totals(const int32_t *__restrict__ s3, const int32_t *__restrict__ grand, SegInfo sg, int64_t nsweep, int32_t *__restrict__ out) {
  /* out[g][c] = total of channel c over segment g = prefix(first[g+1]) - prefix(first[g]) (exclusive prefixes; the end is the grand total) */
  const int q = threadIdx.x;
  if (q >= sg.ngen * 3) return;
  const int g = q / 3, c = q % 3;
  const int32_t a = sg.first[g] < nsweep ? s3[sg.first[g] * 3 + c] : grand[c];
  const int32_t b = sg.first[g + 1] < nsweep ? s3[sg.first[g + 1] * 3 + c] : grand[c];
  out[q] = b - a;
}
__global__ void __launch_bounds__(BLOCK) k_event_rows_sharded(const int32_t *__restrict__ outcome, const int32_t *__restrict__ s3,
                                                               const int32_t *__restrict__ ids, const int32_t *__restrict__ child_id, int64_t n,
                                                               int32_t n0, SegInfo sg, odx_event_row *__restrict__ rows, int64_t max_rows) {
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * BLOCK) {
    const int o = outcome[i];
    if (o == OUT_NONE) continue;
    int g = 0;
    while (g + 1 < sg.ngen && i >= sg.first[g + 1]) g++;
    const int64_t f = sg.first[g];
    const int64_t r = (int64_t)sg.rows_off[g] + (s3[i * 3 + 1] - s3[f * 3 + 1]);
    const int32_t nbefore = n0 - (sg.deaths_off[g] + (s3[i * 3 + 2] - s3[f * 3 + 2]));
    if (o == OUT_DIE) {
      if (r < max_rows) { odx_event_row e = {nbefore - 1, ODX_STOP_DEATH, {ids[i], -1, -1}}; rows[r] = e; }
    } else {
      if (r < max_rows) { odx_event_row e = {nbefore + 1, ODX_STOP_REPLI, {ids[i], 1, child_id[i]}}; rows[r] = e; }
      if (o == OUT_REPL_DIE && r + 1 < max_rows) { odx_event_row e = {nbefore, ODX_STOP_DEATH, {ids[i], -1, -1}}; rows[r + 1] = e; }
    }
  }
}

#define ODX_NCCL(h, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { (h)->err = std::string("NCCL: ") + g_nccl.GetErrorString(r_); return ODX_ENCCL; } } while (0)

/* all ranks: ints[count] -> all[nranks][count] on the host */
static int allgather_ints(odx_handle *h, EventBufs &b, const int32_t *dev_src, int count, std::vector<int32_t> &all) {
  ODX_NCCL(h, g_nccl.AllGather(dev_src, b.xch, (size_t)count, ncclInt32, h->comm, h->stream));
  all.resize((size_t)count * h->nranks);
  int32_t *pin = reinterpret_cast<int32_t *>(b.xch_pin);
  ODX_CUDA(h, cudaMemcpyAsync(pin, b.xch, sizeof(int32_t) * (size_t)count * h->nranks, cudaMemcpyDeviceToHost, h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  for (size_t k = 0; k < all.size(); k++) all[k] = pin[k];
  return ODX_SUCCESS;
}

/* move the pieces [gstart, gstart + cnt) of the new global list this rank holds at `src` (local offset lstart) to their new owners,
 * and receive this rank's new shard [lo, hi) into `dst`; one column of `elem`-byte entries */
struct Piece { int64_t gstart, cnt, lstart; };
static int exchange_column(odx_handle *h, const char *src, char *dst, size_t elem, const std::vector<std::vector<Piece>> &pieces,
                           const std::vector<int64_t> &lo, const std::vector<int64_t> &hi) {
  const int me = h->rank;      /* the caller brackets all columns of a redistribution with one ncclGroupStart / ncclGroupEnd */
  for (int q = 0; q < h->nranks; q++) {
    /* what I send to q: my pieces cut to [lo[q], hi[q]) */
    for (const Piece &pc : pieces[me]) {
      const int64_t a = std::max(pc.gstart, lo[q]), e = std::min(pc.gstart + pc.cnt, hi[q]);
      if (e <= a) continue;
      const char *sp = src + (size_t)(pc.lstart + (a - pc.gstart)) * elem;
      if (q == me) {
        ODX_CUDA(h, cudaMemcpyAsync(dst + (size_t)(a - lo[me]) * elem, sp, (size_t)(e - a) * elem, cudaMemcpyDeviceToDevice, h->stream));
      } else {
        ODX_NCCL(h, g_nccl.Send(sp, (size_t)(e - a) * elem, ncclChar, q, h->comm, h->stream));
      }
    }
    if (q == me) continue;
    /* what I receive from q: q's pieces cut to my range, in the same order q sends them */
    for (const Piece &pc : pieces[q]) {
      const int64_t a = std::max(pc.gstart, lo[me]), e = std::min(pc.gstart + pc.cnt, hi[me]);
      if (e <= a) continue;
      ODX_NCCL(h, g_nccl.Recv(dst + (size_t)(a - lo[me]) * elem, (size_t)(e - a) * elem, ncclChar, q, h->comm, h->stream));
    }
  }
  return ODX_SUCCESS;
}

static int tauleap_sharded(odx_handle *h, double t, double dt, uint64_t leap_index, odx_event_row *rows, int64_t max_rows, int64_t *nrows_out,
                           const odx_event_row **view_out) {
  if (!h->comm) { h->err = "multi-GPU birth/death needs odx_comm_init"; return ODX_ENCCL; }
  odx_event_state *es = event_state(h);
  EventBufs &b = es->b;
  ODX_TRY(ensure_event_bufs(h, b));
  const int nrk_ = h->nranks, me = h->rank;
  if (!b.xch) {
    ODX_CUDA(h, cudaMalloc((void **)&b.xch, sizeof(int32_t) * (size_t)(3 * SHARD_MAXGEN + 8) * (MAX_RANKS + 1)));
    ODX_CUDA(h, cudaMallocHost((void **)&b.xch_pin, sizeof(int32_t) * (size_t)(3 * SHARD_MAXGEN + 8) * (MAX_RANKS + 1)));
  }
  int32_t *xsend = b.xch + (size_t)(3 * SHARD_MAXGEN + 8) * MAX_RANKS;
  const uint32_t stream = (leap_index == UINT64_MAX) ? es->event_index : (uint32_t)leap_index;
  const int32_t N0 = (int32_t)h->n_global;
  double birth_rate = 0.0;
  ODX_TRY(odx_rates(h, nullptr, &birth_rate));
  /* ---- generations */
  SegInfo sg; memset(&sg, 0, sizeof sg);
  std::vector<std::vector<int32_t>> cgen;          /* cgen[g][r]: daughters born to rank r's generation-g particles */
  int64_t first = 0, end = h->n;
  int32_t max_id = h->max_id;
  std::vector<int32_t> all;
  int32_t g1[4];
  for (int g = 0;; g++) {
    if (g >= SHARD_MAXGEN) { h->err = "sharded tau-leaping: more generations of daughters in one sweep than supported"; return ODX_EUNSUP; }
    sg.first[g] = first;
    const int64_t count = end - first;
    Pop p = pop_view(h);
    if (count > 0) {
      k_leap_flags<<<grid_for(h, count), BLOCK, 0, h->stream>>>(p, first, count, dt, stream, b.outcome, b.flag);
      ODX_TRY(launch_check(h));
    }
    /* the number of daughters stays on the device: scan total -> allgather -> one host round trip per generation for all ranks' counts */
    ODX_TRY(exclusive_scan<1>(h, b.flag, b.scan, count, b.tile, b.grand, g1, false));
    ODX_TRY(allgather_ints(h, b, b.grand, 1, all));
    const int32_t nd = all[me];
    cgen.push_back(all);
    int64_t ctot_ = 0, before = 0;
    for (int r = 0; r < nrk_; r++) { ctot_ += all[r]; if (r < me) before += all[r]; }
    sg.ngen = g + 1;
    if (ctot_ == 0) { sg.first[g + 1] = end; break; }
    if (nd > 0) {
      if (end + nd > h->cap) { h->err = "odx_tauleap_apply: population exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
      k_spawn<<<grid_for(h, count), BLOCK, 0, h->stream>>>(p, first, count, b.outcome, b.scan, end, max_id + (int32_t)before, t, stream, b.child);
      ODX_TRY(launch_check(h));
    }
    max_id += (int32_t)ctot_;
    first = end; end += nd;
  }
  const int ngen_ = sg.ngen;
  const int64_t nsweep = end;
  /* de-novo births: the host loop's draw, the same number on every rank */
  long nnew = M::HAS_BIRTH ? odx_draw_poisson(h, birth_rate * dt, h->hr_seeded ? 1 : 0) : 0;
  if (nnew < 0) nnew = 0;
  /* ---- per-segment totals (survivors, rows, deaths) of every rank */
  int32_t g3[4] = {0, 0, 0, 0};
  if (nsweep > 0) { k_sweep_flags<<<grid_for(h, nsweep), BLOCK, 0, h->stream>>>(b.outcome, b.f3, nsweep); ODX_TRY(launch_check(h)); }
  ODX_TRY(exclusive_scan<3>(h, b.f3, b.s3, nsweep, b.tile, b.grand, g3, false));      /* totals read on the device by k_seg_totals */
  k_seg_totals<<<1, 3 * SHARD_MAXGEN, 0, h->stream>>>(b.s3, b.grand, sg, nsweep, xsend);
  ODX_TRY(launch_check(h));
  ODX_TRY(allgather_ints(h, b, xsend, 3 * ngen_, all));
  auto tot = [&](int r, int g, int c) { return (int64_t)all[(size_t)r * 3 * ngen_ + 3 * g + c]; };
  /* offsets of my segments in the global sweep order (generation-major, then rank), and the pieces of the new list */
  int64_t keep_run = 0, rows_run = 0, deaths_run = 0, my_keep_local = 0;
  std::vector<std::vector<Piece>> pieces(nrk_);
  std::vector<int64_t> keep_local_run(nrk_, 0);
  for (int g = 0; g < ngen_; g++) {
    for (int r = 0; r < nrk_; r++) {
      if (r == me) { sg.rows_off[g] = (int32_t)rows_run; sg.deaths_off[g] = (int32_t)deaths_run; }
      const int64_t k = tot(r, g, 0);
      if (k > 0) pieces[r].push_back(Piece{keep_run, k, keep_local_run[r]});
      keep_local_run[r] += k;
      keep_run += k; rows_run += tot(r, g, 1); deaths_run += tot(r, g, 2);
    }
  }
  my_keep_local = keep_local_run[me];
  const int64_t nrows_sweep = rows_run, Nnew = keep_run + nnew;
  /* ---- event rows: every rank writes its rows at their global positions, an allreduce over the disjoint positions completes them */
  int64_t written = 0;
  const bool want = (rows && max_rows > 0) || view_out;
  if (view_out) max_rows = nrows_sweep + nnew;
  if (want && nrows_sweep + nnew > b.rows_pin_cap) {
    if (b.rows_pin) cudaFreeHost(b.rows_pin);
    b.rows_pin = nullptr; b.rows_pin_cap = 0;
    const int64_t cap2 = 2 * (nrows_sweep + nnew) + 1024;
    ODX_CUDA(h, cudaMallocHost((void **)&b.rows_pin, sizeof(odx_event_row) * (size_t)cap2));
    b.rows_pin_cap = cap2;
  }
  if (want && nrows_sweep > 0) {
    if (b.rows_cap < nrows_sweep) {
      if (b.rows) cudaFree(b.rows);
      b.rows = nullptr; b.rows_cap = 0;
      ODX_CUDA(h, cudaMalloc((void **)&b.rows, sizeof(odx_event_row) * (size_t)(nrows_sweep + 1024)));
      b.rows_cap = nrows_sweep + 1024;
    }
    ODX_CUDA(h, cudaMemsetAsync(b.rows, 0, sizeof(odx_event_row) * (size_t)nrows_sweep, h->stream));
    if (nsweep > 0) {
      k_event_rows_sharded<<<grid_for(h, nsweep), BLOCK, 0, h->stream>>>(b.outcome, b.s3, h->ID, b.child, nsweep, N0, sg, b.rows, nrows_sweep);
      ODX_TRY(launch_check(h));
    }
    static_assert(sizeof(odx_event_row) == 5 * sizeof(int32_t), "event row layout");
    ODX_NCCL(h, g_nccl.AllReduce(b.rows, b.rows, (size_t)nrows_sweep * 5, ncclInt32, ncclSum, h->comm, h->stream));
    written = nrows_sweep < max_rows ? nrows_sweep : max_rows;
    ODX_CUDA(h, cudaMemcpyAsync(b.rows_pin, b.rows, sizeof(odx_event_row) * (size_t)written, cudaMemcpyDeviceToHost, h->stream));
  }
  const int32_t n_after_sweep = N0 - (int32_t)deaths_run;
  /* ---- local stable compaction (alt buffers), then the redistribution into balanced contiguous shards */
  Pop p = pop_view(h);
  if (nsweep > 0) {
    k_compact<<<grid_for(h, nsweep), BLOCK, 0, h->stream>>>(p, b.alt, b.outcome, b.s3, nsweep, nsweep, (int32_t)my_keep_local);
    ODX_TRY(launch_check(h));
  }
  std::vector<int64_t> lo(nrk_), hi(nrk_);
  { const int64_t base = Nnew / nrk_, extra = Nnew % nrk_;
    for (int r = 0; r < nrk_; r++) { lo[r] = r * base + std::min<int64_t>(r, extra); hi[r] = lo[r] + base + (r < extra ? 1 : 0); } }
  const int64_t mine = hi[me] - lo[me];
  if (mine > h->cap) { h->err = "birth/death: this rank's shard exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
  /* survivors occupy global positions [0, keep_run); de-novo births [keep_run, Nnew) are created in place by their new owners */
  std::vector<int64_t> hi_s(nrk_), lo_s(nrk_);
  for (int r = 0; r < nrk_; r++) { lo_s[r] = std::min(lo[r], keep_run); hi_s[r] = std::min(hi[r], keep_run); }
  ODX_NCCL(h, g_nccl.GroupStart());
  for (int c = 0; c < NVAR; c++) ODX_TRY(exchange_column(h, (const char *)(b.alt.Y + (size_t)c * h->cap), (char *)(h->Y + (size_t)c * h->cap), 8, pieces, lo_s, hi_s));
  for (int c = 0; c < NEXPR; c++) ODX_TRY(exchange_column(h, (const char *)(b.alt.E + (size_t)c * h->cap), (char *)(h->E + (size_t)c * h->cap), 8, pieces, lo_s, hi_s));
  for (int c = 0; c < NAUX; c++) ODX_TRY(exchange_column(h, (const char *)(b.alt.AUX + (size_t)c * h->cap), (char *)(h->AUX + (size_t)c * h->cap), 8, pieces, lo_s, hi_s));
  for (int c = 0; c < NPSI; c++) ODX_TRY(exchange_column(h, (const char *)(b.alt.PSI + (size_t)c * h->cap), (char *)(h->PSI + (size_t)c * h->cap), 8, pieces, lo_s, hi_s));
  ODX_TRY(exchange_column(h, (const char *)b.alt.DEATH, (char *)h->DEATH, 8, pieces, lo_s, hi_s));
  ODX_TRY(exchange_column(h, (const char *)b.alt.REPLI, (char *)h->REPLI, 8, pieces, lo_s, hi_s));
  ODX_TRY(exchange_column(h, (const char *)b.alt.ID, (char *)h->ID, 4, pieces, lo_s, hi_s));
  ODX_NCCL(h, g_nccl.GroupEnd());
  /* de-novo births that fall into my new shard */
  {
    const int64_t a = std::max(lo[me], keep_run), e = std::min(hi[me], Nnew);
    if (e > a) {
      h->n = mine;
      k_births<<<grid_for(h, e - a), BLOCK, 0, h->stream>>>(pop_view(h), a - lo[me], e - a, max_id + (int32_t)(a - keep_run), t, stream);
      ODX_TRY(launch_check(h));
    }
  }
  for (long k = 0; k < nnew; k++) {
    if (want && written < max_rows) {
      odx_event_row ev = {n_after_sweep + (int32_t)k + 1, ODX_STOP_BIRTH, {-1, 1, max_id + (int32_t)k}};
      b.rows_pin[written++] = ev;       /* behind the sweep rows (the copy in flight fills only the entries below them) */
    }
  }
  max_id += (int32_t)nnew;
  h->n = mine; h->max_id = max_id;
  h->n_global = (double)Nnew; h->n_before = lo[me];
  for (int r = 0; r < nrk_ && r < MAX_RANKS; r++) h->all_n[r] = hi[r] - lo[r];
  invalidate_state(h);
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (want) {
    if (rows && written > 0) memcpy(rows, b.rows_pin, sizeof(odx_event_row) * (size_t)written);
    if (view_out) *view_out = b.rows_pin;
  }
  if (nrows_out) *nrows_out = written;
  es->event_index++;
  return ODX_SUCCESS;
}
#endif
}  // namespace odx

extern "C" {

/* the sweep as one cooperative launch (k_leap_sweep) + the compaction behind it (k_leap_compact); the host reads the counts as soon
 * as the sweep kernel has ended and copies the event rows on a second stream while the compaction runs.  When the device row buffer
 * turns out too small for this sweep (it holds a quarter of the capacity + 64 Ki rows at first), the rows are rebuilt afterwards from
 * the outcomes by the scan kernels of the per-generation path. */
static int tauleap_fused(odx_handle *h, odx_event_state *es, double t, double dt, uint32_t stream, int32_t n0, double birth_rate,
                         odx_event_row *rows, int64_t max_rows, int64_t *nrows, const odx_event_row **view_out) {
  EventBufs &b = es->b;
  if (!b.cta_cnt) {
    ODX_CUDA(h, cudaMalloc((void **)&b.cta_cnt, sizeof(int32_t) * 5 * LEAP_MAX_CTAS));
    ODX_CUDA(h, cudaMalloc((void **)&b.leap_result, sizeof(long long) * 48));
    ODX_CUDA(h, cudaMemset(b.leap_result, 0, sizeof(long long) * 48));
    ODX_CUDA(h, cudaStreamCreateWithFlags(&b.copy_stream, cudaStreamNonBlocking));
    ODX_CUDA(h, cudaEventCreateWithFlags(&b.ev_swept, cudaEventDisableTiming));
    int occ = 0;
    ODX_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_leap_sweep, LEAP_T, 0));
    if (occ < 1) { h->err = "k_leap_sweep does not fit on an SM"; return ODX_EFAILED; }
    if (occ > 2) occ = 2;
    b.leap_grid = occ * h->sm_count > LEAP_MAX_CTAS ? LEAP_MAX_CTAS : occ * h->sm_count;
  }
  long nnew = M::HAS_BIRTH ? odx_draw_poisson(h, birth_rate * dt, h->hr_seeded ? 1 : 0) : 0;     /* the only host draw of a sweep */
  if (nnew < 0) nnew = 0;
  const bool want = (rows && max_rows > 0) || view_out;
  if (want && b.rows_cap < h->cap / 4 + 65536) {
    if (b.rows) cudaFree(b.rows);
    b.rows = nullptr; b.rows_cap = 0;
    ODX_CUDA(h, cudaMalloc((void **)&b.rows, sizeof(odx_event_row) * (size_t)(h->cap / 4 + 65536)));
    b.rows_cap = h->cap / 4 + 65536;
  }
  Pop p = pop_view(h);
  LeapArgs a;
  a.a = b.alt; a.dt = dt; a.t = t; a.stream = stream; a.n0 = n0; a.max_id0 = h->max_id; a.nbirth = nnew;
  a.outcome = b.outcome; a.child = b.child; a.cta_cnt = b.cta_cnt; a.result = b.leap_result;
  a.rows = want ? b.rows : nullptr; a.rows_cap = want ? b.rows_cap : 0;
  void *args[] = {(void *)&p, (void *)&a};
  ODX_CUDA(h, cudaLaunchCooperativeKernel((const void *)k_leap_sweep, dim3((unsigned)b.leap_grid), dim3(LEAP_T), args, 0, h->stream));
  ODX_CUDA(h, cudaEventRecord(b.ev_swept, h->stream));
  k_leap_compact<<<b.leap_grid, LEAP_T, 0, h->stream>>>(p, a);
  ODX_TRY(launch_check(h));
  h->launches++;
  /* counts first (behind the sweep, not behind the compaction), then the rows while the compaction runs */
  long long *pin = reinterpret_cast<long long *>(h->h_pin + NSLOT * NRED);
  ODX_CUDA(h, cudaStreamWaitEvent(b.copy_stream, b.ev_swept, 0));
  ODX_CUDA(h, cudaMemcpyAsync(pin, b.leap_result, sizeof(long long) * 6, cudaMemcpyDeviceToHost, b.copy_stream));
  ODX_CUDA(h, cudaStreamSynchronize(b.copy_stream));
#ifdef LEAP_TIMING
  { long long tm[40]; cudaMemcpy(tm, b.leap_result, sizeof tm, cudaMemcpyDeviceToHost);
    fprintf(stderr, "k_leap_sweep marks (us):"); for (int k = 9; k < 40 && tm[k] > tm[8]; k++) fprintf(stderr, " %.1f", (tm[k] - tm[k - 1]) * 1e-3); fprintf(stderr, "\n");
    cudaMemset(b.leap_result, 0, sizeof tm); }
#endif
  if (pin[5] != 0) { cudaStreamSynchronize(h->stream); h->err = "odx_tauleap_apply: population exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
  const int64_t kept = pin[0], nr = pin[1], nsweep = pin[3];
  const int32_t ndead = (int32_t)pin[2];
  int64_t written = 0;
  if (want) {
    if (view_out) max_rows = nr + nnew;
    written = nr < max_rows ? nr : max_rows;
    if (b.rows_pin_cap < nr + nnew) {
      if (b.rows_pin) cudaFreeHost(b.rows_pin);
      b.rows_pin = nullptr; b.rows_pin_cap = 0;
      const int64_t cap2 = 2 * (nr + nnew) + 1024;
      ODX_CUDA(h, cudaMallocHost((void **)&b.rows_pin, sizeof(odx_event_row) * (size_t)cap2));
      b.rows_pin_cap = cap2;
    }
    if (nr <= b.rows_cap) {
      if (written > 0) ODX_CUDA(h, cudaMemcpyAsync(b.rows_pin, b.rows, sizeof(odx_event_row) * (size_t)written, cudaMemcpyDeviceToHost, b.copy_stream));
    } else {
      /* more events than the row buffer held: rebuild the rows from the outcomes (the old columns are untouched until the swap) */
      int32_t g3[4];
      const int g = grid_for(h, nsweep);
      k_sweep_flags<<<g, BLOCK, 0, h->stream>>>(b.outcome, b.f3, nsweep); ODX_TRY(launch_check(h));
      ODX_TRY(exclusive_scan<3>(h, b.f3, b.s3, nsweep, b.tile, b.grand, g3));
      cudaFree(b.rows); b.rows = nullptr; b.rows_cap = 0;
      ODX_CUDA(h, cudaMalloc((void **)&b.rows, sizeof(odx_event_row) * (size_t)(2 * nr + 1024)));
      b.rows_cap = 2 * nr + 1024;
      k_event_rows<<<g, BLOCK, 0, h->stream>>>(b.outcome, b.s3, h->ID, b.child, nsweep, n0, b.rows, b.rows_cap); ODX_TRY(launch_check(h));
      ODX_CUDA(h, cudaMemcpyAsync(b.rows_pin, b.rows, sizeof(odx_event_row) * (size_t)written, cudaMemcpyDeviceToHost, h->stream));
    }
  }
  ODX_CUDA(h, cudaStreamSynchronize(h->stream));
  ODX_CUDA(h, cudaStreamSynchronize(b.copy_stream));
  h->max_id = (int32_t)pin[4];
  const int32_t n_after_sweep = n0 - ndead;
  for (int64_t k = 0; k < nnew; k++) {
    if (want && written < max_rows) {
      odx_event_row e = {n_after_sweep + (int32_t)k + 1, ODX_STOP_BIRTH, {-1, 1, h->max_id + (int32_t)k}};
      b.rows_pin[written++] = e;
    }
  }
  h->max_id += (int32_t)nnew;
  std::swap(h->Y, b.alt.Y); std::swap(h->E, b.alt.E); std::swap(h->AUX, b.alt.AUX); std::swap(h->PSI, b.alt.PSI);
  std::swap(h->DEATH, b.alt.DEATH); std::swap(h->REPLI, b.alt.REPLI); std::swap(h->ID, b.alt.ID);
  h->n = kept + nnew;
  invalidate_state(h);
  ODX_TRY(refresh_global_n(h));
  if (rows && written > 0) memcpy(rows, b.rows_pin, sizeof(odx_event_row) * (size_t)written);
  if (view_out) *view_out = b.rows_pin;
  if (nrows) *nrows = written;
  es->event_index++;
  return ODX_SUCCESS;
}
static int tauleap_core(odx_handle *h, double t, double dt, uint64_t leap_index, odx_event_row *rows, int64_t max_rows, int64_t *nrows,
                        const odx_event_row **view_out) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (h->nranks > 1) {
#ifdef ODX_HAVE_NCCL_H
    if (h->shard_events) return tauleap_sharded(h, t, dt, leap_index, rows, max_rows, nrows, view_out);
    ODX_TRY(replica_gather(h));
    event_state(h->replica)->event_index = event_state(h)->event_index;       /* one event counter, whichever path an event takes */
    int st = tauleap_core(h->replica, t, dt, leap_index, rows, max_rows, nrows, view_out);
    if (st != ODX_SUCCESS) { h->err = std::string("replicated sweep: ") + h->replica->err; return st; }
    event_state(h)->event_index = event_state(h->replica)->event_index;
    return replica_scatter(h);
#else
    h->err = "built without nccl.h"; return ODX_ENCCL;
#endif
  }
  odx_event_state *es = event_state(h);
  EventBufs &b = es->b;
  ODX_TRY(ensure_event_bufs(h, b));
  const uint32_t stream = (leap_index == UINT64_MAX) ? es->event_index : (uint32_t)leap_index;
  const int32_t n0 = (int32_t)h->n;
  double birth_rate = 0.0;
  ODX_TRY(odx_rates(h, nullptr, &birth_rate));
  if (h->coop_ok && h->leap_fused != 0) return tauleap_fused(h, es, t, dt, stream, n0, birth_rate, rows, max_rows, nrows, view_out);
  /* generations of the sweep: [first, end) -> daughters appended at [end, end + nd) */
  int64_t first = 0, end = h->n;
  int32_t g1[4];
  while (end > first) {
    const int64_t count = end - first;
    Pop p = pop_view(h);
    k_leap_flags<<<grid_for(h, count), BLOCK, 0, h->stream>>>(p, first, count, dt, stream, b.outcome, b.flag);
    ODX_TRY(launch_check(h));
    ODX_TRY(exclusive_scan<1>(h, b.flag, b.scan, count, b.tile, b.grand, g1));
    const int64_t nd = g1[0];
    if (nd > 0) {
      if (end + nd > h->cap) { h->err = "odx_tauleap_apply: population exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
      k_spawn<<<grid_for(h, count), BLOCK, 0, h->stream>>>(p, first, count, b.outcome, b.scan, end, h->max_id, t, stream, b.child);
      ODX_TRY(launch_check(h));
      h->max_id += (int32_t)nd;
    }
    first = end; end += nd;
  }
  const int64_t nsweep = end;
  /* de-novo births: poisson(pop_birth_rate * dt), drawn by the host loop's stream */
  long nnew = M::HAS_BIRTH ? odx_draw_poisson(h, birth_rate * dt, h->hr_seeded ? 1 : 0) : 0;
  if (nnew < 0) nnew = 0;
  if (nsweep + nnew > h->cap) { h->err = "odx_tauleap_apply: population exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
  if (nnew > 0) {
    k_births<<<grid_for(h, nnew), BLOCK, 0, h->stream>>>(pop_view(h), nsweep, nnew, h->max_id, t, stream);
    ODX_TRY(launch_check(h));
    h->max_id += (int32_t)nnew;
  }
  int32_t g3[4];
  ODX_TRY(finish_events(h, b, nsweep, nnew, n0, rows, max_rows, nrows, g3, view_out));
  es->event_index++;
  return ODX_SUCCESS;
}
int odx_tauleap_apply(odx_handle *h, double t, double dt, uint64_t leap_index, odx_event_row *rows, int64_t max_rows, int64_t *nrows) {
  return tauleap_core(h, t, dt, leap_index, rows, max_rows, nrows, nullptr);
}
/* the same sweep, all event rows left in pinned host memory owned by the library (valid until the next birth / death call) */
int odx_tauleap_apply_view(odx_handle *h, double t, double dt, uint64_t leap_index, const odx_event_row **rows, int64_t *nrows) {
  if (!rows || !nrows) { h->err = "odx_tauleap_apply_view: rows and nrows must not be NULL"; return ODX_EINVAL; }
  *rows = nullptr; *nrows = 0;
  return tauleap_core(h, t, dt, leap_index, nullptr, 0, nrows, rows);
}

int odx_ssa_apply(odx_handle *h, double t, double u, odx_event_row *row) {
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (h->nranks > 1) {
#ifdef ODX_HAVE_NCCL_H
    ODX_TRY(replica_gather(h));
    event_state(h->replica)->event_index = event_state(h)->event_index;
    int st = odx_ssa_apply(h->replica, t, u, row);
    if (st != ODX_SUCCESS) { h->err = std::string("replicated SSA event: ") + h->replica->err; return st; }
    event_state(h)->event_index = event_state(h->replica)->event_index;
    return replica_scatter(h);
#else
    h->err = "built without nccl.h"; return ODX_ENCCL;
#endif
  }
  odx_event_state *es = event_state(h);
  EventBufs &b = es->b;
  ODX_TRY(ensure_event_bufs(h, b));
  const int64_t N = h->n;
  if (row) { row->n = (int32_t)N; row->stop_flag = 0; row->event[0] = row->event[1] = row->event[2] = 0; }
  double birth_rate = 0.0;
  ODX_TRY(odx_rates(h, nullptr, &birth_rate));
  int64_t idx = 0;
  const bool on_device = (h->ssa_mode == 1) || (h->ssa_mode < 0 && !h->hr_seeded);
  if (on_device) {
    /* keyed runs: scan + search on the device (see k_ssa_pick), 16 bytes back */
    const int ntiles = (int)((2 * N + 1 + SSA_TILE - 1) / SSA_TILE);
    if (b.ssa_tiles_cap < ntiles) {
      if (b.ssa_tiles) cudaFree(b.ssa_tiles);
      b.ssa_tiles = nullptr; b.ssa_tiles_cap = 0;
      ODX_CUDA(h, cudaMalloc((void **)&b.ssa_tiles, sizeof(double) * (size_t)(2 * ntiles + 16)));
      b.ssa_tiles_cap = 2 * ntiles + 16;
    }
    if (!b.ssa_out) ODX_CUDA(h, cudaMalloc((void **)&b.ssa_out, 16));
    k_ssa_tile_sums<<<ntiles, BLOCK, 0, h->stream>>>(h->DEATH, h->REPLI, N, birth_rate, b.ssa_tiles);
    ODX_TRY(launch_check(h));
    k_ssa_pick<<<1, BLOCK, 0, h->stream>>>(h->DEATH, h->REPLI, N, birth_rate, b.ssa_tiles, ntiles, u, b.ssa_out, (double *)(b.ssa_out + 1));
    ODX_TRY(launch_check(h));
    long long *pin = reinterpret_cast<long long *>(h->h_pin + NSLOT * NRED);
    ODX_CUDA(h, cudaMemcpyAsync(pin, b.ssa_out, 16, cudaMemcpyDeviceToHost, h->stream));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
    if (pin[0] < 0) return ODX_SUCCESS;            /* total rate 0: nothing happens */
    idx = (int64_t)pin[0];
  } else {
    /* replayed serial stream (parity runs): normalised cumulative sum and linear search exactly as SSA_apply_birthdeath / ncumsum
     * (host, reference order of additions) */
    std::vector<double> r((size_t)(2 * N + 1));
    if (N > 0) {
      ODX_CUDA(h, cudaMemcpyAsync(r.data(), h->DEATH, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream));
      ODX_CUDA(h, cudaMemcpyAsync(r.data() + N, h->REPLI, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream));
      ODX_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    r[2 * N] = birth_rate;
    for (int64_t i = 1; i < 2 * N + 1; i++) r[i] += r[i - 1];
    const double sumr = r[2 * N];
    for (int64_t i = 0; i < 2 * N + 1; i++) r[i] /= r[2 * N];
    if (!(sumr > 0.0)) return ODX_SUCCESS;
    while (u > r[idx]) idx++;
  }
  const uint32_t stream = es->event_index;
  int64_t nbirth = 0;
  if (N > 0) { k_fill_i32<<<grid_for(h, N), BLOCK, 0, h->stream>>>(b.outcome, N, OUT_NONE); ODX_TRY(launch_check(h)); }
  int32_t one;
  if (idx < N) {
    one = OUT_DIE;
    ODX_CUDA(h, cudaMemcpyAsync(b.outcome + idx, &one, sizeof one, cudaMemcpyHostToDevice, h->stream));
  } else if (idx < 2 * N) {
    const int64_t k = idx - N;
    if (N + 1 > h->cap) { h->err = "odx_ssa_apply: population exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
    one = OUT_REPL_KEEP;
    ODX_CUDA(h, cudaMemcpyAsync(b.outcome + k, &one, sizeof one, cudaMemcpyHostToDevice, h->stream));
    ODX_CUDA(h, cudaMemsetAsync(b.scan, 0, sizeof(int32_t) * N, h->stream));
    k_spawn<<<grid_for(h, N), BLOCK, 0, h->stream>>>(pop_view(h), 0, N, b.outcome, b.scan, N, h->max_id, t, stream, b.child);
    ODX_TRY(launch_check(h));
    h->max_id += 1;
    /* the daughter sits at slot N: treat it as an appended particle that is kept */
    nbirth = 1;
  } else {
    if (N + 1 > h->cap) { h->err = "odx_ssa_apply: population exceeds the capacity given to odx_create"; return ODX_ENOMEM; }
    k_births<<<1, BLOCK, 0, h->stream>>>(pop_view(h), N, 1, h->max_id, t, stream);
    ODX_TRY(launch_check(h));
    h->max_id += 1;
    nbirth = 1;
  }
  odx_event_row tmp[2]; int64_t nr = 0; int32_t g3[4];
  const bool repl = (idx >= N && idx < 2 * N);
  /* for a replication the appended slot is not a de-novo birth: no birth row, the sweep row carries the child id */
  ODX_TRY(finish_events(h, b, N, nbirth, (int32_t)N, repl ? tmp : tmp, repl ? 1 : 2, &nr, g3));
  if (row && nr > 0) *row = tmp[0];
  es->event_index++;
  return ODX_SUCCESS;
}

}  // extern "C"
