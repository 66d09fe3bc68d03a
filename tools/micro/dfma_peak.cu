// dfma_peak.cu -- measured fp64 FMA peak of the device (BASELINE.md section 2 asks for the real bound of
// the push / deposit kernels to have a measured denominator): a register-resident DFMA micro-kernel,
// 8 independent chains per thread, no memory traffic.  Prints one JSON line.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dfma_peak tools/micro/dfma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters)
{
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += x[c];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // never true: keeps the chains alive
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 1024 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int   iters = 4096, CH = 8;
  double      best = 0;
  int         best_bps = 0;
  for (int bps : {2, 4, 8}) { // resident CTAs of 256 threads per SM
    const int blocks = p.multiProcessorCount * bps;
    k_dfma<CH><<<blocks, 256>>>(out, 0.999999, 1e-9, 64); // warm-up
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      k_dfma<CH><<<blocks, 256>>>(out, 0.999999, 1e-9, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)blocks * 256 * CH * 16.0 * iters;
      const double rate = fma / (ms * 1e-3);
      if (rate > best) best = rate, best_bps = bps;
    }
  }
  if (cudaGetLastError() != cudaSuccess) {
    std::printf("{\"error\": \"kernel failed\"}\n");
    return 1;
  }
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  std::printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_gfma_per_s\": %.1f, \"fp64_tflops\": %.2f, "
              "\"fma_per_clk_per_sm_at_max_clock\": %.2f, \"max_clock_mhz\": %.0f, \"ctas_per_sm\": %d}\n",
              p.name, p.multiProcessorCount, best / 1e9, 2 * best / 1e12,
              best / p.multiProcessorCount / (clk * 1e3), clk / 1e3, best_bps);
  return 0;
}
