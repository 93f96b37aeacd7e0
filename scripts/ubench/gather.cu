// Micro-benchmark (design evidence): cost of a warp-wide gather of particle records from an L1-resident
// buffer as a function of the address pattern and of the load width.  Prints cycles per warp-load per SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double ld256(const char *p) {
  double x, y, z, w;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
  return x + y + z + w;
}
__device__ __forceinline__ double ld128(const char *p) {
  double x, y;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
  return x + y;
}
__device__ __forceinline__ double ld64(const char *p) {
  double x;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(x) : "l"(p));
  return x;
}
__device__ __forceinline__ double lds64(unsigned a) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a));
  return x;
}
__device__ __forceinline__ double lds128(unsigned a) {
  double x, y;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
  return x + y;
}

// offs: [NSET][32] byte offsets per lane (relative to the block's buffer), NSET address sets cycled through
template <int MODE>
__global__ void k(const char *buf, const int *offs, int nset, int iters, int bytes_per_block, double *out,
                  long long *cyc) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31;
  const char *base = buf + (size_t)blockIdx.x * bytes_per_block;
  if (MODE >= 10) {
    for (int i = threadIdx.x; i < bytes_per_block / 8; i += blockDim.x) sm[i] = (double)i;
    __syncthreads();
  }
  int o[8];
  for (int s = 0; s < 8; s++) o[s] = offs[(s % nset) * 32 + lane] + (threadIdx.x >> 5) * 0;
  double acc = 0.0;
  // warm L1
  for (int s = 0; s < 8; s++) acc += ld256(base + (o[s] & ~31));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < 8; s++) {
      o[s] = (o[s] + 512) & (bytes_per_block - 1);  // next 4-line group: same pattern class, new lines
      if (MODE == 0) acc += ld256(base + o[s]);
      if (MODE == 1) { acc += ld128(base + o[s]); acc += ld128(base + o[s] + 16); }
      if (MODE == 2) { acc += ld64(base + o[s]); acc += ld64(base + o[s] + 8); acc += ld64(base + o[s] + 16); }
      if (MODE == 3) acc += ld128(base + o[s]);
      if (MODE == 4) acc += ld64(base + o[s]);
      if (MODE == 10) { unsigned a = (unsigned)__cvta_generic_to_shared(sm) + o[s]; acc += lds128(a); acc += lds64(a + 16); }
      if (MODE == 11) { unsigned a = (unsigned)__cvta_generic_to_shared(sm) + o[s]; acc += lds64(a); acc += lds64(a + 8); acc += lds64(a + 16); }
      if (MODE == 12) { unsigned a = (unsigned)__cvta_generic_to_shared(sm) + o[s]; acc += lds128(a); }
    }
  }
  // the clock is read only after the last load has returned (acc depends on every load)
  if (acc == 1.2345e300) t0 = 0;
  long long t1 = clock64();
  __shared__ long long smin, smax;
  if (threadIdx.x == 0) { smin = t0; smax = t1; }
  __syncthreads();
  if (lane == 0) { atomicMin(&smin, t0); atomicMax(&smax, t1); }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = smax - smin;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const int nblk = 148, tpb = 512, bytes = 32768, iters = 200, NSET = 8;
  char *buf; int *doffs; double *out; long long *cyc;
  cudaMalloc(&buf, (size_t)nblk * bytes); cudaMemset(buf, 0, (size_t)nblk * bytes);
  cudaMalloc(&doffs, NSET * 32 * 4); cudaMalloc(&out, nblk * tpb * 8); cudaMalloc(&cyc, nblk * 8);
  struct Pat { const char *name; int kind; };
  Pat pats[] = {{"random records (32B), random lines", 0}, {"distinct lines, same sector slot", 1},
                {"distinct lines, slot = lane%4", 2}, {"4 lanes per line (8 lines)", 3},
                {"2 lanes per line (16 lines)", 4}, {"all lanes same record (broadcast)", 5},
                {"8 distinct records in 8 lines, 4 lanes each same record", 6},
                {"13 random lines, random slot", 7}, {"random records, 8B-misaligned (24B-style)", 8}};
  const char *modes[] = {"LDG.256", "2xLDG.128", "3xLDG.64", "1xLDG.128(16B)", "1xLDG.64(8B)"};
  const char *smodes[] = {"LDS.128+LDS.64", "3xLDS.64", "1xLDS.128(16B)"};
  srand(1);
  for (auto &pt : pats) {
    std::vector<int> h(NSET * 32);
    const int nrec = bytes / 32, nline = bytes / 128;
    for (int s = 0; s < NSET; s++) {
      std::vector<int> lines(nline);
      for (int i = 0; i < nline; i++) lines[i] = i;
      for (int i = nline - 1; i > 0; i--) { int j = rand() % (i + 1); std::swap(lines[i], lines[j]); }
      for (int l = 0; l < 32; l++) {
        int v = 0;
        switch (pt.kind) {
          case 0: v = (rand() % nrec) * 32; break;
          case 1: v = lines[l] * 128; break;
          case 2: v = lines[l] * 128 + (l % 4) * 32; break;
          case 3: v = lines[l / 4] * 128 + (l % 4) * 32; break;
          case 4: v = lines[l / 2] * 128 + (l % 2) * 32 + (rand() % 2) * 64; break;
          case 5: v = lines[0] * 128; break;
          case 6: v = lines[l / 4] * 128 + ((l / 4) % 4) * 32; break;
          case 7: v = lines[rand() % 13] * 128 + (rand() % 4) * 32; break;
          case 8: v = (rand() % (bytes / 32)) * 32 + 8 * (rand() % 2); break;
        }
        h[s * 32 + l] = v;
      }
    }
    cudaMemcpy(doffs, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    printf("%-58s", pt.name);
    for (int m = 0; m < 5; m++) {
      if (pt.kind == 8 && m != 2 && m != 4) { printf("  %-14s   -  ", modes[m]); continue; }
      if (m == 0) k<0><<<nblk, tpb>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      if (m == 1) k<1><<<nblk, tpb>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      if (m == 2) k<2><<<nblk, tpb>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      if (m == 3) k<3><<<nblk, tpb>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      if (m == 4) k<4><<<nblk, tpb>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      cudaDeviceSynchronize();
      long long c[148]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < nblk; i++) avg += c[i]; avg /= nblk;
      printf("  %s %6.2f", modes[m], avg / (iters * 8.0 * (tpb / 32)));
    }
    printf("\n%-58s", "   shared memory:");
    for (int m = 0; m < 3; m++) {
      if (m == 0) k<10><<<nblk, tpb, bytes>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      if (m == 1) k<11><<<nblk, tpb, bytes>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      if (m == 2) k<12><<<nblk, tpb, bytes>>>(buf, doffs, NSET, iters, bytes, out, cyc);
      cudaDeviceSynchronize();
      long long c[148]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < nblk; i++) avg += c[i]; avg /= nblk;
      printf("  %s %6.2f", smodes[m], avg / (iters * 8.0 * (tpb / 32)));
    }
    printf("\n");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  return 0;
}
