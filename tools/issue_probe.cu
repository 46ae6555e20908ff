// issue_probe.cu -- does a non-FP64 instruction issue in the shadow of a DFMA?  A DFMA occupies the 16-lane
// FP64 pipe of an SM sub-partition for 2 cycles.  If the scheduler can issue an FFMA / IMAD / LDS in the second
// of those cycles, a stream of D DFMAs + N other instructions costs max(2 D, D + N) issue cycles; if not, 2 D + N.
// This decides what k_batch_level (71 fp64 + 69 other instructions per pixel-iteration) is bound by.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

// D DFMAs (8 independent chains) and N "other" instructions (8 independent chains) per loop trip
template <int KIND, int D, int N>
__global__ void __launch_bounds__(128) k_mix(double* out, int iters, double a, double b, float fa, float fb, int ia, const float* smem_src) {
  __shared__ float sh[128 * 8];
  for (int k = 0; k < 8; ++k) sh[threadIdx.x * 8 + k] = (float)k;
  __syncthreads();
  double x[8]; float f[8]; int q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 1e-3 + k; f[k] = threadIdx.x * 1e-3f + k; q[k] = threadIdx.x + k; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < (D > N ? D : N); ++r) {
      if (r < D) x[r % 8] = fma(x[r % 8], a, b);
      if (r < N) {
        if (KIND == 0) f[r % 8] = fmaf(f[r % 8], fa, fb);
        if (KIND == 1) q[r % 8] = q[r % 8] * ia + 12345;
        if (KIND == 2) f[r % 8] += sh[(threadIdx.x * 8 + ((q[0] + r) & 7))];
        if (KIND == 3) q[r % 8] = (q[r % 8] ^ ia) + (q[(r + 1) % 8] >> 3);     // LOP3 / SHF / IADD mix
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k] + f[k] + q[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int D, int N>
static void run(const char* name, int sms, int clock_khz, double* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 13, ctas = 4;          // 4 CTAs of 128 threads per SM: 4 warps per sub-partition (what k_batch_level has)
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_mix<KIND, D, N><<<sms * ctas, 128>>>(out, iters, 1.0000001, 1e-9, 1.0000001f, 1e-9f, 3, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  // cycles per loop trip per warp-slot: 4 warps per sub-partition share one issue port
  const double cycles = best * 1e-3 * clock_khz * 1e3 / iters / 4.0;
  printf("{\"other\": \"%s\", \"dfma_per_trip\": %d, \"other_per_trip\": %d, \"issue_cycles_per_trip\": %.2f, \"if_overlapped\": %d, \"if_serialised\": %d}\n",
         name, D, N, cycles, (2 * D > D + N ? 2 * D : D + N), 2 * D + N);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4 * 128);
  const int sms = p.multiProcessorCount;
  run<0, 16, 0>("none", sms, clk, out);
  run<0, 0, 16>("ffma only", sms, clk, out);
  run<0, 16, 8>("ffma", sms, clk, out);
  run<0, 16, 16>("ffma", sms, clk, out);
  run<0, 16, 32>("ffma", sms, clk, out);
  run<1, 16, 16>("imad", sms, clk, out);
  run<1, 16, 32>("imad", sms, clk, out);
  run<2, 16, 16>("lds+fadd", sms, clk, out);
  run<3, 16, 16>("int alu (2-3 instr each)", sms, clk, out);
  return 0;
}
