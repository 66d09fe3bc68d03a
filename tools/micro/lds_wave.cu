// Microbenchmark: shared-memory wavefronts of 64/128-bit loads by address pattern, SHFL cost (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_wave lds_wave.cu
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum ./lds_wave
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;
template <int MODE>
__global__ void k(double* out, const int* pat)
{
  __shared__ __align__(16) double s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int off = pat[MODE * 32 + lane];
  double a = 0, b = 0;
  for (int it = 0; it < ITER; it++) {
    if (MODE < 8) {
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(s + off + (it & 63) * 6)));
      a += v;
    } else if (MODE < 12) {
      double v, w;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"((unsigned)__cvta_generic_to_shared(s + off + (it & 63) * 6)));
      a += v; b += w;
    } else if (MODE == 12) {
      a += __shfl_xor_sync(0xffffffffu, a + it, 16);
    } else if (MODE == 13) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(s + off + (it & 63) * 6)));
      a += v;
    } else if (MODE == 14) { // st.shared.f64, lane-contiguous
      asm volatile("st.shared.f64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(s + off + (it & 63) * 32)), "d"(a + it));
    } else if (MODE == 15) { // st.shared.v2.f64 lane-contiguous
      asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"((unsigned)__cvta_generic_to_shared(s + off + (it & 31) * 64)), "d"(a + it), "d"(b));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
}
int main()
{
  int h[16 * 32];
  for (int l = 0; l < 32; l++) {
    h[0 * 32 + l] = 0;                       // 64-bit uniform
    h[1 * 32 + l] = (l & 1) * 6;             // 2 addresses 48 B apart, interleaved lanes
    h[2 * 32 + l] = (l >> 4) * 6;            // 2 addresses, by half-warp
    h[3 * 32 + l] = (l & 1) * 72;            // 2 addresses 576 B apart
    h[4 * 32 + l] = (l & 1) * 72 + ((l >> 1) & 1) * 504; // 4 addresses (0,576,4032,4608) B: bank clash
    h[5 * 32 + l] = (l & 1) * 6 + ((l >> 1) & 1) * 72;   // 4 addresses conflict-free
    h[6 * 32 + l] = l;                       // 32 consecutive doubles
    h[7 * 32 + l] = (l % 10) * 21;           // 10 addresses stride 21 doubles
    h[8 * 32 + l] = 0;                       // 128-bit uniform
    h[9 * 32 + l] = (l & 1) * 6;             // 128-bit 2 addresses
    h[10 * 32 + l] = l * 2;                  // 128-bit consecutive
    h[11 * 32 + l] = (l % 10) * 22;          // 128-bit 10 addresses stride 22 doubles
    h[12 * 32 + l] = 0;
    h[13 * 32 + l] = 0;
    h[14 * 32 + l] = l;
    h[15 * 32 + l] = l * 2;
  }
  int* d; double* o;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, 8 * 148 * 4 * 256);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(M) { k<M><<<148 * 4, 256>>>(o, d); cudaEventRecord(e0); k<M><<<148 * 4, 256>>>(o, d); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); \
    printf("mode %2d: %.3f ms  -> %.2f clk/warp-instr/SM\n", M, ms, ms * 1e-3 * 1.965e9 / (double(ITER) * 4 * 8)); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
