// Microbenchmark (development tool): FP64 pipe throughput on sm_100a as a function of operand kind and instruction mix.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[16] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1e-9, 2e-9, 3e-9, 4e-9, 1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8};
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double seed, int iseed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double r0 = seed * 1.0000001, r1 = seed * 0.999999;
  int i0 = iseed + threadIdx.x, i1 = i0 * 3, i2 = i0 * 5, i3 = i0 * 7;
  for (int it = 0; it < iters; it++) {
#pragma unroll 4
    for (int u = 0; u < 4; u++) {
      if (MODE == 0) {        // pure register DFMA
        a0 = fma(a0, r0, r1); a1 = fma(a1, r0, r1); a2 = fma(a2, r0, r1); a3 = fma(a3, r0, r1);
        a4 = fma(a4, r0, r1); a5 = fma(a5, r0, r1); a6 = fma(a6, r0, r1); a7 = fma(a7, r0, r1);
      } else if (MODE == 1) { // constant-bank / uniform operand as addend (Horner style)
        a0 = fma(a0, r0, kc[4]); a1 = fma(a1, r0, kc[5]); a2 = fma(a2, r0, kc[6]); a3 = fma(a3, r0, kc[7]);
        a4 = fma(a4, r0, kc[4]); a5 = fma(a5, r0, kc[5]); a6 = fma(a6, r0, kc[6]); a7 = fma(a7, r0, kc[7]);
      } else if (MODE == 2) { // DFMA + one integer op each
        a0 = fma(a0, r0, r1); i0 = i0 * 3 + i1; a1 = fma(a1, r0, r1); i1 = i1 * 5 + i2; a2 = fma(a2, r0, r1); i2 = i2 * 7 + i3; a3 = fma(a3, r0, r1); i3 = i3 * 9 + i0;
        a4 = fma(a4, r0, r1); i0 ^= i2; a5 = fma(a5, r0, r1); i1 ^= i3; a6 = fma(a6, r0, r1); i2 += i0; a7 = fma(a7, r0, r1); i3 += i1;
      } else if (MODE == 3) { // 2 chains only (latency test)
        a0 = fma(a0, r0, r1); a1 = fma(a1, r0, r1); a0 = fma(a0, r0, r1); a1 = fma(a1, r0, r1);
        a0 = fma(a0, r0, r1); a1 = fma(a1, r0, r1); a0 = fma(a0, r0, r1); a1 = fma(a1, r0, r1);
      } else if (MODE == 4) { // 3 distinct register operands per DFMA (register-file ports)
        a0 = fma(a1, a2, a0); a1 = fma(a2, a3, a1); a2 = fma(a3, a4, a2); a3 = fma(a4, a5, a3);
        a4 = fma(a5, a6, a4); a5 = fma(a6, a7, a5); a6 = fma(a7, a0, a6); a7 = fma(a0, a1, a7);
      } else if (MODE == 5) { // DFMA + two integer ops each
        a0 = fma(a0, r0, r1); i0 = i0 * 3 + i1; i1 ^= i0; a1 = fma(a1, r0, r1); i1 = i1 * 5 + i2; i2 ^= i1; a2 = fma(a2, r0, r1); i2 = i2 * 7 + i3; i3 ^= i2; a3 = fma(a3, r0, r1); i3 = i3 * 9 + i0; i0 ^= i3;
        a4 = fma(a4, r0, r1); i0 += i2; i2 ^= i0; a5 = fma(a5, r0, r1); i1 += i3; i3 ^= i1; a6 = fma(a6, r0, r1); i2 += i0; i0 ^= i2; a7 = fma(a7, r0, r1); i3 += i1; i1 ^= i3;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (double)(i0 + i1 + i2 + i3);
}
template <int MODE> void run(const char *name, int warps_per_sm) {
  int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  const int threads = 256, blocks = p.multiProcessorCount * (warps_per_sm * 32 / threads);
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const int iters = 20000;
  k<MODE><<<blocks, threads>>>(out, 100, 1.0, 1);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 1.0, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = (double)blocks * threads * iters * 4 * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("%-44s warps/SM=%2d  %.3f ms  %.2f T DFMA/s  = %.1f DFMA/clk/SM at %d MHz nominal\n", name, warps_per_sm, ms, dfma / (ms * 1e-3) / 1e12,
         dfma / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  for (int w : {8, 16, 24, 32, 64}) {
    run<0>("DFMA reg,reg,reg (8 chains)", w);
    run<1>("DFMA reg,reg,const (8 chains)", w);
    run<4>("DFMA 3 distinct regs (8 chains)", w);
    run<2>("DFMA + 1 int op (8 chains)", w);
    run<5>("DFMA + 2 int ops (8 chains)", w);
    run<3>("DFMA 2 chains (latency)", w);
  }
  return 0;
}
