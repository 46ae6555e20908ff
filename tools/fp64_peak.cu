// fp64_peak.cu -- measures the vector FP64 (DFMA) issue rate of the device: the secondary roofline of
// the alignment kernels (MEASURED_PEAKS.json has no fp64 entry).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 14;
  for (int wpb = 1; wpb <= 8; wpb *= 2) {     // CTAs of 256 threads per SM
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      k_dfma<8><<<sms * wpb, 256>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double fmas = (double)sms * wpb * 256 * 8 * iters;
    printf("{\"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.3f, \"dfma_per_clk_per_sm_at_%dMHz\": %.2f}\n",
           wpb, wpb * 8, fmas / (best * 1e-3), 2 * fmas / (best * 1e-3) / 1e12, p.clockRate / 1000, fmas / (best * 1e-3) / sms / (p.clockRate * 1e3));
  }
  return 0;
}
