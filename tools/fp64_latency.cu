// fp64_latency.cu -- dependent-issue latency of DFMA / DADD / DMUL and of a 64-bit SHFL on this device
// (1 warp, one dependent chain): explains the `wait` stalls of the alignment kernels.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x = threadIdx.x * 1e-3, y = 1.0 + threadIdx.x * 1e-3, z = 0.5;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) { y = y + a; y = y + b; y = y + a; y = y + b; }
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) { z = z * a; z = z * a; z = z * a; z = z * a; }
  long long t3 = clock64();
  double w = x;
  for (int i = 0; i < iters; ++i) { w = __shfl_xor_sync(0xffffffffu, w, 1); w = __shfl_xor_sync(0xffffffffu, w, 2); w = __shfl_xor_sync(0xffffffffu, w, 4); w = __shfl_xor_sync(0xffffffffu, w, 8); }
  long long t4 = clock64();
  // two independent chains: does ILP 2 double the rate?
  double p = x, q = y;
  for (int i = 0; i < iters; ++i) { p = fma(p, a, b); q = fma(q, a, b); p = fma(p, a, b); q = fma(q, a, b); }
  long long t5 = clock64();
  out[threadIdx.x] = x + y + z + w + p + q;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 5 * 8);
  const int iters = 4096;
  k_chain<<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  k_chain<<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  printf("{\"dfma_dependent_cycles\": %.2f, \"dadd_dependent_cycles\": %.2f, \"dmul_dependent_cycles\": %.2f, \"shfl64_dependent_cycles\": %.2f, \"dfma_two_chains_cycles_per_instr\": %.2f}\n",
         cyc[0] / (4.0 * iters), cyc[1] / (4.0 * iters), cyc[2] / (4.0 * iters), cyc[3] / (4.0 * iters), cyc[4] / (4.0 * iters));
  return 0;
}
