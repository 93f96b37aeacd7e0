// Micro-benchmark (design evidence, round 2): do FP64 instructions occupy the warp scheduler's issue port for the two
// cycles the half-rate FP64 pipe is busy, or can integer / load instructions of other warps issue in the second cycle?
// 8 warps per scheduler; per loop trip 16 DFMA (4 independent chains) + K integer instructions (4 independent chains)
// [+ L shared-memory loads].  Port blocked: t = 2*16 + K (+L) clk per warp-trip per scheduler; not blocked: max(32, 16 + K + L).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int L>
__global__ void __launch_bounds__(1024, 1) k_mix(double *out, long long *clk, double a, double b, unsigned m) {
  __shared__ double sh[1024];
  sh[threadIdx.x] = a;
  __syncthreads();
  double x[4];
  unsigned u[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { x[i] = a + i + threadIdx.x; u[i] = threadIdx.x * 2654435761u + i; }
  double acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 512; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 4; i++) x[i] = fma(x[i], b, a);
#pragma unroll
    for (int r = 0; r < K / 4; r++)
#pragma unroll
      for (int i = 0; i < 4; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(m), "r"(u[(i + 1) & 3]));
#pragma unroll
    for (int l = 0; l < L; l++) {
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sh) + ((u[l & 3] & 1023u) << 3)));
      acc += v * 0.0 == 0.0 ? 0.0 : 1.0;   // keeps the load alive cheaply? (compiler may turn into FP64 ops: counted below)
    }
  }
  const long long t1 = clock64();
  double s = acc;
#pragma unroll
  for (int i = 0; i < 4; i++) s += x[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int K, int L>
void run(double *out, long long *clk) {
  k_mix<K, L><<<148, 1024>>>(out, clk, 1.0, 0.5, 0x5bd1e995u);
  k_mix<K, L><<<148, 1024>>>(out, clk, 1.0, 0.5, 0x5bd1e995u);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("16 DFMA + %2d LOP3 + %d LDS per trip, 8 warps/scheduler: %.1f clk per warp-trip per scheduler (blocked port: %d, free port: %d)\n", K, L,
         h / 512.0 / 8.0, 32 + K + L, (16 + K + L) > 32 ? 16 + K + L : 32);
}
int main() {
  double *out; long long *clk;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaMalloc(&clk, 8);
  run<0, 0>(out, clk); run<8, 0>(out, clk); run<16, 0>(out, clk); run<24, 0>(out, clk); run<32, 0>(out, clk); run<48, 0>(out, clk);
  return 0;
}
