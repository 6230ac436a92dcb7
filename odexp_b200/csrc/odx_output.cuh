/* odx_output.cuh — the output path at scale (reference: fwrite_all_particles / fwrite_quick, src/datastruct.c:655-674,
 * src/methods_odexp.c:1353-1436; readers src/datastruct.c:747-820).
 *
 *  1. d2h_pipelined: device -> pageable host memory through a ring of pinned chunks.  A plain cudaMemcpy into pageable memory is
 *     staged by the driver at ~5 GB/s; here the DMA engine fills pinned chunks at link speed while a few host threads copy the
 *     chunks that have landed into the caller's buffer, so the slower of the two (the host copy, several threads) sets the pace.
 *  2. trajectory frames assembled ON THE DEVICE in the byte layout of traj.dat (one record per selected particle: step, id, t,
 *     y, aux, psi, mean fields, expr, followed by the three doubles of the current.plot row), copied to pinned host memory on a
 *     second stream and handed to the C driver as one block to fwrite.  The state is snapshotted by the record kernel on the
 *     compute stream, so the copy of frame k and the host's fwrite of frame k-1 overlap the stepping of the next interval. */
#pragma once
#include <atomic>
#include <thread>

namespace odx {

constexpr size_t D2H_CHUNK = (size_t)16 << 20;
constexpr int D2H_SLOTS = 8, D2H_THREADS = 6;

struct OutState {
  /* pinned ring of d2h_pipelined */
  unsigned char *ring = nullptr; cudaEvent_t ring_ev[D2H_SLOTS] = {};
  /* trajectory frames: two device record buffers + two pinned host buffers, second stream */
  cudaStream_t copy_stream = nullptr;
  unsigned char *rec_dev[2] = {nullptr, nullptr}; size_t rec_dev_cap[2] = {0, 0};
  unsigned char *rec_pin[2] = {nullptr, nullptr}; size_t rec_pin_cap[2] = {0, 0};
  cudaEvent_t rec_ready[2] = {nullptr, nullptr}, rec_done[2] = {nullptr, nullptr};
  size_t rec_bytes[2] = {0, 0}; bool rec_pending[2] = {false, false};
  int next_slot = 0;
};

static OutState *out_state(odx_handle *h);

static int d2h_pipelined(odx_handle *h, void *dst, const void *src_dev, size_t bytes) {
  if (bytes == 0) return ODX_SUCCESS;
  if (bytes < 2 * D2H_CHUNK) {                       /* small: one direct copy */
    ODX_CUDA(h, cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, h->stream));
    ODX_CUDA(h, cudaStreamSynchronize(h->stream));
    return ODX_SUCCESS;
  }
  OutState *os = out_state(h);
  if (!os->ring) {
    ODX_CUDA(h, cudaMallocHost((void **)&os->ring, D2H_CHUNK * D2H_SLOTS));
    for (int k = 0; k < D2H_SLOTS; k++) ODX_CUDA(h, cudaEventCreateWithFlags(&os->ring_ev[k], cudaEventDisableTiming));
  }
  const size_t nchunk = (bytes + D2H_CHUNK - 1) / D2H_CHUNK;
  std::atomic<size_t> issued{0}, copied_upto[D2H_THREADS];
  for (int t = 0; t < D2H_THREADS; t++) copied_upto[t].store(0);
  std::atomic<int> failed{0};
  /* worker t copies chunks t, t + T, t + 2T, ... as they land */
  auto worker = [&](int t) {
    cudaSetDevice(h->device);
    for (size_t c = (size_t)t; c < nchunk; c += D2H_THREADS) {
      while (issued.load(std::memory_order_acquire) <= c) { if (failed.load()) return; std::this_thread::yield(); }
      if (cudaEventSynchronize(os->ring_ev[c % D2H_SLOTS]) != cudaSuccess) { failed.store(1); return; }
      const size_t off = c * D2H_CHUNK, len = (bytes - off < D2H_CHUNK) ? bytes - off : D2H_CHUNK;
      memcpy((unsigned char *)dst + off, os->ring + (c % D2H_SLOTS) * D2H_CHUNK, len);
      copied_upto[t].store(c + 1, std::memory_order_release);
    }
  };
  std::thread th[D2H_THREADS];
  for (int t = 0; t < D2H_THREADS; t++) th[t] = std::thread(worker, t);
  cudaError_t err = cudaSuccess;
  for (size_t c = 0; c < nchunk && err == cudaSuccess; c++) {
    if (c >= (size_t)D2H_SLOTS) {                    /* the slot is free once chunk c - SLOTS has been copied out by its worker */
      const size_t prev = c - D2H_SLOTS;
      while (copied_upto[prev % D2H_THREADS].load(std::memory_order_acquire) < prev + 1) { if (failed.load()) break; std::this_thread::yield(); }
    }
    const size_t off = c * D2H_CHUNK, len = (bytes - off < D2H_CHUNK) ? bytes - off : D2H_CHUNK;
    err = cudaMemcpyAsync(os->ring + (c % D2H_SLOTS) * D2H_CHUNK, (const unsigned char *)src_dev + off, len, cudaMemcpyDeviceToHost, h->stream);
    if (err == cudaSuccess) err = cudaEventRecord(os->ring_ev[c % D2H_SLOTS], h->stream);
    if (err != cudaSuccess) { failed.store(1); break; }
    issued.store(c + 1, std::memory_order_release);
  }
  if (err != cudaSuccess) failed.store(1);
  for (int t = 0; t < D2H_THREADS; t++) th[t].join();
  if (err != cudaSuccess || failed.load()) {
    h->err = std::string("pipelined device-to-host copy failed: ") + cudaGetErrorString(err != cudaSuccess ? err : cudaGetLastError());
    return ODX_ECUDA;
  }
  return ODX_SUCCESS;
}

/* one traj.dat record per selected particle, exactly fwrite_all_particles' bytes (src/datastruct.c:655-674):
 *   unsigned step | int id | double t | y[nvar] | aux[naux] | psi[npsi] | meanfield[nmfd] | expr[nexpr] */
constexpr int REC_DOUBLES = NVAR + NAUX + NPSI + NMFD + NEXPR;
constexpr size_t REC_BYTES = 16 + 8 * (size_t)REC_DOUBLES;
constexpr int MFD_ARG_MAX = 24;
struct MfdArg { double v[MFD_ARG_MAX]; };
__global__ void __launch_bounds__(BLOCK) k_traj_records(Pop p, int64_t m, int64_t step_stride, unsigned int nbrsteps, double t, MfdArg mf,
                                                         unsigned char *__restrict__ out) {
  for (int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x; q < m; q += (int64_t)gridDim.x * BLOCK) {
    const int64_t i = q * step_stride;
    unsigned char *r = out + (size_t)q * REC_BYTES;
    *reinterpret_cast<unsigned int *>(r) = nbrsteps;
    *reinterpret_cast<int *>(r + 4) = p.ID[i];
    double *d = reinterpret_cast<double *>(r + 8);
    int k = 0;
    d[k++] = t;
#pragma unroll
    for (int c = 0; c < NVAR; c++) d[k++] = p.Y[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NAUX; c++) d[k++] = p.AUX[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NPSI; c++) d[k++] = p.PSI[(int64_t)c * p.stride + i];
#pragma unroll
    for (int c = 0; c < NMFD; c++) d[k++] = mf.v[c < MFD_ARG_MAX ? c : 0];
#pragma unroll
    for (int c = 0; c < NEXPR; c++) d[k++] = p.E[(int64_t)c * p.stride + i];
  }
}
/* the current.plot row (fwrite_quick, src/methods_odexp.c:1353-1436) for axis indices ng[3]: -1 = T, then the dynamical
 * variables, auxiliary variables and couplings of the particle with id `cp` (getpar: looked up by id), then the mean fields.
 * As coded in the reference a dynamical variable is read from the particle-major state array at y[cp * tx + ng] with
 * tx = nvar + naux + npsi: reproduced, NaN where the reference would read past the end of y or the particle does not exist. */
__global__ void __launch_bounds__(BLOCK) k_quick_row(Pop p, int ngx, int ngy, int ngz, int cp, double t, MfdArg mf, double *__restrict__ row) {
  __shared__ long long s_pos;
  if (threadIdx.x == 0) s_pos = -1;
  __syncthreads();
  for (int64_t i = threadIdx.x; i < p.n; i += BLOCK) if (p.ID[i] == cp) s_pos = (long long)i;
  __syncthreads();
  if (threadIdx.x >= 3) return;
  const int tx = NVAR + NAUX + NPSI;
  const int g = threadIdx.x == 0 ? ngx : (threadIdx.x == 1 ? ngy : ngz);
  const long long pos = s_pos;
  double v = __longlong_as_double(0x7ff8000000000000LL);
  if (g == -1) v = t;
  else if (g >= tx) { if (g - tx < NMFD && g - tx < MFD_ARG_MAX) v = mf.v[g - tx]; }
  else if (pos >= 0) {
    if (g < NVAR) {
      const int64_t el = (int64_t)cp * tx + g, ip = el / A1(NVAR);
      if (ip < p.n) v = p.Y[(int64_t)(el % A1(NVAR)) * p.stride + ip];
    } else if (g < NVAR + NAUX) v = p.AUX[(int64_t)(g - NVAR) * p.stride + pos];
    else v = p.PSI[(int64_t)(g - NVAR - NAUX) * p.stride + pos];
  }
  row[threadIdx.x] = v;
}

}  // namespace odx

static odx::OutState *odx::out_state(odx_handle *h) {
  if (!h->out) h->out = new odx::OutState();
  return (odx::OutState *)h->out;
}
static void odx_output_free(odx_handle *h) {
  if (!h->out) return;
  odx::OutState *os = (odx::OutState *)h->out;
  if (os->copy_stream) cudaStreamSynchronize(os->copy_stream);
  if (os->ring) cudaFreeHost(os->ring);
  for (int k = 0; k < odx::D2H_SLOTS; k++) if (os->ring_ev[k]) cudaEventDestroy(os->ring_ev[k]);
  for (int k = 0; k < 2; k++) {
    if (os->rec_dev[k]) cudaFree(os->rec_dev[k]);
    if (os->rec_pin[k]) cudaFreeHost(os->rec_pin[k]);
    if (os->rec_ready[k]) cudaEventDestroy(os->rec_ready[k]);
    if (os->rec_done[k]) cudaEventDestroy(os->rec_done[k]);
  }
  if (os->copy_stream) cudaStreamDestroy(os->copy_stream);
  delete os;
  h->out = nullptr;
}

extern "C" {

/* Start the output of one trajectory frame: records of particles 0, stride, 2 stride, ... in traj.dat layout + the current.plot
 * row are assembled on the device from the CURRENT state and copied to pinned host memory on a second stream.  Returns at
 * once with a ticket; the caller may go on launching steps.  stride <= 0: no records, only the current.plot row. */
int odx_frame_begin(odx_handle *h, unsigned int nbrsteps, double t, int64_t stride, const double *mfd, int want_quick, int ngx, int ngy, int ngz,
                    int particle, int *ticket) {
  using namespace odx;
  ODX_CUDA(h, cudaSetDevice(h->device));
  if (NMFD > MFD_ARG_MAX) { h->err = "odx_frame_begin: more mean fields than the record kernel takes by value"; return ODX_EUNSUP; }
  OutState *os = out_state(h);
  if (!os->copy_stream) {
    ODX_CUDA(h, cudaStreamCreateWithFlags(&os->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      ODX_CUDA(h, cudaEventCreateWithFlags(&os->rec_ready[k], cudaEventDisableTiming));
      ODX_CUDA(h, cudaEventCreateWithFlags(&os->rec_done[k], cudaEventDisableTiming));
    }
  }
  const int s = os->next_slot;
  if (os->rec_pending[s]) { h->err = "odx_frame_begin: two frames are already in flight (call odx_frame_end first)"; return ODX_EINVAL; }
  const int64_t m = (stride > 0 && h->n > 0) ? (h->n + stride - 1) / stride : 0;
  const size_t bytes = (size_t)m * REC_BYTES + 3 * sizeof(double);
  if (os->rec_dev_cap[s] < bytes) {
    ODX_CUDA(h, cudaStreamSynchronize(os->copy_stream));
    if (os->rec_dev[s]) cudaFree(os->rec_dev[s]);
    if (os->rec_pin[s]) cudaFreeHost(os->rec_pin[s]);
    os->rec_dev[s] = nullptr; os->rec_pin[s] = nullptr; os->rec_dev_cap[s] = os->rec_pin_cap[s] = 0;
    const size_t cap = bytes + bytes / 4 + 4096;
    ODX_CUDA(h, cudaMalloc((void **)&os->rec_dev[s], cap));
    ODX_CUDA(h, cudaMallocHost((void **)&os->rec_pin[s], cap));
    os->rec_dev_cap[s] = os->rec_pin_cap[s] = cap;
  }
  MfdArg mf; memset(&mf, 0, sizeof mf);
  for (int k = 0; k < NMFD && k < MFD_ARG_MAX; k++) mf.v[k] = mfd ? mfd[k] : 0.0;
  Pop p = pop_view(h);
  if (m > 0) { k_traj_records<<<grid_for(h, m), BLOCK, 0, h->stream>>>(p, m, stride, nbrsteps, t, mf, os->rec_dev[s]); ODX_TRY(launch_check(h)); }
  double *qrow = reinterpret_cast<double *>(os->rec_dev[s] + (size_t)m * REC_BYTES);
  if (want_quick) { k_quick_row<<<1, BLOCK, 0, h->stream>>>(p, ngx, ngy, ngz, particle, t, mf, qrow); ODX_TRY(launch_check(h)); }
  else ODX_CUDA(h, cudaMemsetAsync(qrow, 0, 3 * sizeof(double), h->stream));
  ODX_CUDA(h, cudaEventRecord(os->rec_ready[s], h->stream));
  ODX_CUDA(h, cudaStreamWaitEvent(os->copy_stream, os->rec_ready[s], 0));
  ODX_CUDA(h, cudaMemcpyAsync(os->rec_pin[s], os->rec_dev[s], bytes, cudaMemcpyDeviceToHost, os->copy_stream));
  ODX_CUDA(h, cudaEventRecord(os->rec_done[s], os->copy_stream));
  os->rec_bytes[s] = bytes; os->rec_pending[s] = true;
  os->next_slot = s ^ 1;
  if (ticket) *ticket = s;
  return ODX_SUCCESS;
}
/* Wait for a frame started by odx_frame_begin: *records points at nrecords * record_bytes bytes of traj.dat records in pinned
 * memory owned by the library (valid until the ticket is reused, two frames later), quick_row[3] receives the current.plot row. */
int odx_frame_end(odx_handle *h, int ticket, const void **records, int64_t *nrecords, int64_t *record_bytes, double *quick_row) {
  using namespace odx;
  OutState *os = out_state(h);
  if (ticket < 0 || ticket > 1 || !os->rec_pending[ticket]) { h->err = "odx_frame_end: no such frame in flight"; return ODX_EINVAL; }
  ODX_CUDA(h, cudaSetDevice(h->device));
  ODX_CUDA(h, cudaEventSynchronize(os->rec_done[ticket]));
  const size_t body = os->rec_bytes[ticket] - 3 * sizeof(double);
  if (records) *records = os->rec_pin[ticket];
  if (nrecords) *nrecords = (int64_t)(body / REC_BYTES);
  if (record_bytes) *record_bytes = (int64_t)REC_BYTES;
  if (quick_row) memcpy(quick_row, os->rec_pin[ticket] + body, 3 * sizeof(double));
  os->rec_pending[ticket] = false;
  return ODX_SUCCESS;
}

}  // extern "C"
