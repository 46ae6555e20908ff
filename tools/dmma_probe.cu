// dmma_probe.cu -- does the FP64 tensor-core MMA (mma.sync m8n8k4 f64, SASS DMMA) run beside the
// vector FP64 pipe (DFMA) on this device, or do they share it?  Three kernels with the same loop
// shape: DFMA only, DMMA only, and NF DFMAs + NM DMMAs per trip.  If the mixed time is close to
// max(dfma, dmma) the pipes are independent; if it is close to the sum they are one pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
  double x[NF > 0 ? NF : 1];
  double c[NM > 0 ? 2 * NM : 2];
#pragma unroll
  for (int k = 0; k < (NF > 0 ? NF : 1); ++k) x[k] = threadIdx.x * 1e-3 + k;
#pragma unroll
  for (int k = 0; k < (NM > 0 ? 2 * NM : 2); ++k) c[k] = 0.;
  const double fa = 1e-3 * (threadIdx.x & 7), fb = 1e-3 * (threadIdx.x & 3);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < NF; ++k) x[k] = fma(x[k], a, b);
#pragma unroll
    for (int k = 0; k < NM; ++k) dmma(c[2 * k], c[2 * k + 1], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < (NF > 0 ? NF : 1); ++k) s += x[k];
#pragma unroll
  for (int k = 0; k < (NM > 0 ? 2 * NM : 2); ++k) s += c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NM>
static void run(const char* name, int sms, double* out, int clock_khz) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 13;
  for (int cps = 1; cps <= 4; cps *= 2) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      k_mix<NF, NM><<<sms * cps, 256>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // cycles per loop trip per SM sub-partition (4 per SM, cps*8 warps per SM -> cps*2 warps each)
    const double cyc = best * 1e-3 * clock_khz * 1e3 / iters / (cps * 2);
    printf("{\"kernel\": \"%s\", \"dfma_per_trip\": %d, \"dmma_per_trip\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"smsp_cycles_per_warp_trip\": %.2f, "
           "\"dfma_warp_instr_per_clk_per_sm\": %.3f, \"dmma_per_clk_per_sm\": %.4f}\n",
           name, NF, NM, cps * 8, best, cyc, NF * (double)sms * cps * 8 * iters / (best * 1e-3) / sms / (clock_khz * 1e3),
           NM * (double)sms * cps * 8 * iters / (best * 1e-3) / sms / (clock_khz * 1e3));
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  run<16, 0>("dfma_only", sms, out, p.clockRate);
  run<0, 4>("dmma_only", sms, out, p.clockRate);
  run<0, 8>("dmma_only8", sms, out, p.clockRate);
  run<16, 4>("mixed_16_4", sms, out, p.clockRate);
  run<16, 2>("mixed_16_2", sms, out, p.clockRate);
  run<16, 8>("mixed_16_8", sms, out, p.clockRate);
  run<28, 0>("dfma_28", sms, out, p.clockRate);
  run<28, 8>("mixed_28_8", sms, out, p.clockRate);
  return 0;
}
