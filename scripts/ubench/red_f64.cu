// Micro-benchmark (design evidence, round 2): throughput of red.global.add.f64, the instruction a HALF neighbour
// list would need for the reaction term (p_j -= df d, three doubles per listed pair; fcalculator.cc:216-222).
//
// Every thread owns one "local" and issues NP x 3 reductions into double4 records of "partners":
//   pattern 0: partners uniformly random over the whole array (worst case)
//   pattern 1: partners within +-W records of the thread's own record (cell-sorted locality, W = 700 ~ 27 cells)
//   pattern 2: as 1, and the 32 lanes of a warp hit 32 different records of one window in ascending order
//              (what a sorted, warp-cooperative row would give)
// Prints reductions per second, clk per reduction per SM and the time 16,034,412 locals x 51 pairs would take.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o red_f64 red_f64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void red_add(double *a, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned hash(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int PATTERN>
__global__ void k(double4 *p, unsigned n, int np, int W) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned lane = threadIdx.x & 31;
  for (int k2 = 0; k2 < np; k2++) {
    unsigned j;
    const unsigned h = hash(i * 977u + k2);
    if (PATTERN == 0) j = h % n;
    else if (PATTERN == 1) j = (i + n - W + h % (2 * W)) % n;
    else j = ((i & ~31u) + n - W + (hash((i >> 5) * 977u + k2) % (2 * W - 32)) + lane) % n;
    double *r = reinterpret_cast<double *>(p + j);
    const double v = 1e-9 * (double)(k2 + 1);
    red_add(r + 0, v);
    red_add(r + 1, -v);
    red_add(r + 2, v);
  }
}

int main(int argc, char **argv) {
  const unsigned n = argc > 1 ? (unsigned)atoll(argv[1]) : 16034412u;
  const int np = argc > 2 ? atoi(argv[2]) : 51;
  double4 *p;
  CK(cudaMalloc(&p, sizeof(double4) * (size_t)n));
  CK(cudaMemset(p, 0, sizeof(double4) * (size_t)n));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const double ghz = prop.clockRate * 1e-6;
  for (int pat = 0; pat < 3; pat++) {
    auto launch = [&]() {
      const unsigned g = (n + 255) / 256;
      if (pat == 0) k<0><<<g, 256>>>(p, n, np, 700);
      if (pat == 1) k<1><<<g, 256>>>(p, n, np, 700);
      if (pat == 2) k<2><<<g, 256>>>(p, n, np, 700);
    };
    launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double reds = 3.0 * np * (double)n;
    printf("pattern %d: %.2f ms for %.3g red.global.add.f64 -> %.1f G red/s, %.2f clk per red per SM (%.3f GHz), "
           "%.1f clk per local of %d pairs; the force kernel's whole budget at 0.5 of the roofline is ~24 clk per local\n",
           pat, ms, reds, reds / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * prop.multiProcessorCount / reds, ghz,
           ms * 1e-3 * ghz * 1e9 * prop.multiProcessorCount / n, np);
  }
  return 0;
}
