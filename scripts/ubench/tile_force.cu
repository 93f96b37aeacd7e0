// Micro-benchmark (design evidence, round 2): the force walk over a TILE staged in shared memory.
//
// One block owns a tile of tx x ty x tz cells; the positions of the tile and of its one-cell halo are staged
// in shared memory as 32-byte records (x, y, z, y), partner indices are 16-bit indices into that staged
// array, 4 lanes per local, rows in "rainbow" order (partner index mod 4 = lane in group).  Even lane groups
// read the low half of a record first (x, y) and then z; odd groups read the high half first (z, y) and
// then x, so that the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups.
//
// The benchmark builds a jittered FCC configuration at density 0.7 in a (t+2)^3-cell box, lists all partners
// within `keep` by brute force on the host, and times the walk.  MODE 0: full kernel; 1: gathers only;
// 2: FP64 arithmetic only (positions from registers).  Variants: IL = entries of a lane processed stage by stage
// together (instruction-level parallelism over the ~14-deep dependent FP64 chain of one entry), threads per block /
// blocks per SM (register budget).  Also prints the dependent-issue latency of DFMA.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tile_force tile_force.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct FC { double CL2, c24dt, c48dt, c2x8dt, dt2; int cl2_hi; };

__device__ __forceinline__ void lds128(unsigned a, double &x, double &y) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ double lds64(unsigned a) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a));
  return x;
}
__device__ __forceinline__ double rcp_cubic(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  double t = fma(e, e, e);
  return fma(y, t, y);
}

template <int MODE>
__device__ __forceinline__ void pair(unsigned o0, unsigned o1, double a0, double a1, double a2, const FC &fc,
                                     double &f0, double &f1, double &f2) {
  double j0, j1, j2;
  if (MODE != 2) {
    lds128(o0, j0, j1);
    j2 = lds64(o1);
  } else {
    j0 = __hiloint2double(o0, o1); j1 = __hiloint2double(o1, o0); j2 = __hiloint2double(o0 ^ o1, o0);
  }
  if (MODE == 1) { f0 += j0; f1 += j1; f2 += j2; return; }
  const double d0 = j0 - a0, d1 = j1 - a1, d2 = j2 - a2;
  const double r2 = fma(d2, d2, fma(d1, d1, d0 * d0));
  const double s = rcp_cubic(r2);
  const double s2 = s * s;
  const double s3 = s2 * s;
  const double s4 = s2 * s2;
  const double t = fma(-fc.c48dt, s3, fc.c24dt);
  double df = fma(s4, t, fc.c2x8dt);
  // r2 > CL2 contributes nothing: two integer selects on the words of df (the compiler otherwise selects on the
  // three accumulators: six)
  const bool in = __double2hiint(r2) < fc.cl2_hi;
  df = __hiloint2double(in ? __double2hiint(df) : 0, in ? __double2loint(df) : 0);
  f0 = fma(df, d0, f0);
  f1 = fma(df, d1, f1);
  f2 = fma(df, d2, f2);
}


// N entries stage by stage
template <int N>
__device__ __forceinline__ void pairs(const unsigned (&e)[N], unsigned hx, unsigned hx2, double a0, double a1, double a2, const FC &fc,
                                      double &f0, double &f1, double &f2) {
  double d0[N], d1[N], d2[N], r2[N], s[N], df[N];
#pragma unroll
  for (int i = 0; i < N; i++) { lds128(e[i] * 32u + hx, d0[i], d1[i]); d2[i] = lds64(e[i] * 32u + hx2); }
#pragma unroll
  for (int i = 0; i < N; i++) { d0[i] -= a0; d1[i] -= a1; d2[i] -= a2; }
#pragma unroll
  for (int i = 0; i < N; i++) r2[i] = d0[i] * d0[i];
#pragma unroll
  for (int i = 0; i < N; i++) r2[i] = fma(d1[i], d1[i], r2[i]);
#pragma unroll
  for (int i = 0; i < N; i++) r2[i] = fma(d2[i], d2[i], r2[i]);
#pragma unroll
  for (int i = 0; i < N; i++) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s[i]) : "d"(r2[i]));
  double ee[N];
#pragma unroll
  for (int i = 0; i < N; i++) ee[i] = fma(-r2[i], s[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; i++) ee[i] = fma(ee[i], ee[i], ee[i]);
#pragma unroll
  for (int i = 0; i < N; i++) s[i] = fma(s[i], ee[i], s[i]);
  double s2[N], s3[N];
#pragma unroll
  for (int i = 0; i < N; i++) s2[i] = s[i] * s[i];
#pragma unroll
  for (int i = 0; i < N; i++) { s3[i] = s2[i] * s[i]; s2[i] = s2[i] * s2[i]; }
#pragma unroll
  for (int i = 0; i < N; i++) s3[i] = fma(-fc.c48dt, s3[i], fc.c24dt);
#pragma unroll
  for (int i = 0; i < N; i++) df[i] = fma(s2[i], s3[i], fc.c2x8dt);
#pragma unroll
  for (int i = 0; i < N; i++) {
    const bool in = __double2hiint(r2[i]) < fc.cl2_hi;
    df[i] = __hiloint2double(in ? __double2hiint(df[i]) : 0, in ? __double2loint(df[i]) : 0);
  }
#pragma unroll
  for (int i = 0; i < N; i++) { f0 = fma(df[i], d0[i], f0); f1 = fma(df[i], d1[i], f1); f2 = fma(df[i], d2[i], f2); }
}

// the same walk, IL entries of a lane at a time; TPB threads, MINB blocks per SM (register budget 65536 / (TPB MINB))
template <int IL, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
k_tile_il(const double4 *__restrict__ q, int ns, int n_int, const unsigned *__restrict__ list, const int *__restrict__ fw_off,
          const int *__restrict__ fw_len, long long list_stride, double4 *__restrict__ p, FC fc, unsigned long long *tim) {
  extern __shared__ double4 sq[];
  const long long t0 = clock64();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i <= ns; i += blockDim.x) {
    double4 v = (i < ns) ? q[i] : make_double4(1e8, 1e8, 1e8, 1e8);
    v.w = v.y;
    sq[i] = v;
  }
  __syncthreads();
  const long long t1 = clock64();
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sq);
  const unsigned gp = (lane >> 2) & 1u;
  const unsigned hx = sbase + (gp << 4), hx2 = sbase + ((gp ^ 1u) << 4);
  const int nfw = (n_int + 7) >> 3;
  const unsigned *lb = list + (long long)blockIdx.x * list_stride;
  for (int fw = wid; fw < nfw; fw += nw) {
    const int a = fw * 8 + (lane >> 2);
    const unsigned ao = (unsigned)(a < n_int ? a : ns) << 5;
    double a0, a1, a2;
    lds128(ao + hx, a0, a1);
    a2 = lds64(ao + hx2);
    const int len2 = fw_len[fw] >> 1;   // words of 2 steps
    const unsigned *lp = lb + fw_off[fw] + lane;
    double f0 = 0, f1 = 0, f2 = 0;
    if (IL == 2) {
      unsigned wn = __ldg(lp);
#pragma unroll 2
      for (int k = 0; k < len2; k++) {
        const unsigned w = wn;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + 96));
        wn = __ldg(lp + 32);
        lp += 32;
        const unsigned e[2] = {w & 0xffffu, w >> 16};
        pairs<2>(e, hx, hx2, a0, a1, a2, fc, f0, f1, f2);
      }
    } else {
      // 4 at a time: two words; an odd last word is handled two at a time
      unsigned wn0 = __ldg(lp), wn1 = __ldg(lp + 32);
      int k = 0;
#pragma unroll 1
      for (; k + 2 <= len2; k += 2) {
        const unsigned w0 = wn0, w1 = wn1;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + 160));
        wn0 = __ldg(lp + 64);
        wn1 = __ldg(lp + 96);
        lp += 64;
        const unsigned e[4] = {w0 & 0xffffu, w0 >> 16, w1 & 0xffffu, w1 >> 16};
        pairs<4>(e, hx, hx2, a0, a1, a2, fc, f0, f1, f2);
      }
      if (k < len2) {
        const unsigned e[2] = {wn0 & 0xffffu, wn0 >> 16};
        pairs<2>(e, hx, hx2, a0, a1, a2, fc, f0, f1, f2);
      }
    }
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) {
      f0 += __shfl_xor_sync(0xffffffffu, f0, o);
      f1 += __shfl_xor_sync(0xffffffffu, f1, o);
      f2 += __shfl_xor_sync(0xffffffffu, f2, o);
    }
    if ((lane & 3) == 0 && a < n_int) {
      const double fx = gp ? f2 : f0, fz = gp ? f0 : f2;
      double4 *pp = p + (long long)blockIdx.x * n_int + a;
      double4 pa = *pp;
      pa.x += fx; pa.y += f1; pa.z += fz;
      *pp = pa;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&tim[0], (unsigned long long)(t1 - t0));
    atomicAdd(&tim[1], (unsigned long long)(clock64() - t1));
  }
}

// dependent-issue latency of DFMA: one warp, a chain of 4096 dependent DFMAs; with NCH independent chains per thread
template <int NCH>
__global__ void k_dfma_lat(double *out, long long *clk, double a, double b) {
  double x[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) x[i] = a + i + threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int i = 0; i < NCH; i++) x[i] = fma(x[i], b, a);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// list: per block, per force-warp: words (2 entries of 16 bits = 2 steps) lane-strided: word (k, lane) at
// fw_off + 32 k + lane
template <int MODE>
__global__ void __launch_bounds__(512, 2)
k_tile(const double4 *__restrict__ q, int ns, int n_int, const unsigned *__restrict__ list, const int *__restrict__ fw_off,
       const int *__restrict__ fw_len, long long list_stride, double4 *__restrict__ p, FC fc, unsigned long long *tim) {
  extern __shared__ double4 sq[];
  const long long t0 = clock64();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i <= ns; i += blockDim.x) {
    double4 v = (i < ns) ? q[i] : make_double4(1e8, 1e8, 1e8, 1e8);
    v.w = v.y;
    sq[i] = v;
  }
  __syncthreads();
  const long long t1 = clock64();
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sq);
  const unsigned gp = (lane >> 2) & 1u;
  const unsigned hx = sbase + (gp << 4), hx2 = sbase + ((gp ^ 1u) << 4);
  const int nfw = (n_int + 7) >> 3;
  const unsigned *lb = list + (long long)blockIdx.x * list_stride;
  for (int fw = wid; fw < nfw; fw += nw) {
    const int a = fw * 8 + (lane >> 2);
    const unsigned ao = (unsigned)(a < n_int ? a : ns) << 5;
    double a0, a1, a2;
    lds128(ao + hx, a0, a1);
    a2 = lds64(ao + hx2);
    const int len2 = fw_len[fw] >> 1;
    const unsigned *lp = lb + fw_off[fw] + lane;
    double f0 = 0, f1 = 0, f2 = 0;
    // the list allocation is padded, so loading / prefetching up to three words past a row is harmless
    unsigned wn = __ldg(lp);
#pragma unroll 2
    for (int k = 0; k < len2; k++) {
      const unsigned w = wn;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + 96));
      wn = __ldg(lp + 32);
      lp += 32;
      const unsigned e0 = w & 0xffffu, e1 = w >> 16;
      pair<MODE>(e0 * 32u + hx, e0 * 32u + hx2, a0, a1, a2, fc, f0, f1, f2);
      pair<MODE>(e1 * 32u + hx, e1 * 32u + hx2, a0, a1, a2, fc, f0, f1, f2);
    }
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) {
      f0 += __shfl_xor_sync(0xffffffffu, f0, o);
      f1 += __shfl_xor_sync(0xffffffffu, f1, o);
      f2 += __shfl_xor_sync(0xffffffffu, f2, o);
    }
    if ((lane & 3) == 0 && a < n_int) {
      const double fx = gp ? f2 : f0, fz = gp ? f0 : f2;
      double4 *pp = p + (long long)blockIdx.x * n_int + a;
      double4 pa = *pp;
      pa.x += fx; pa.y += f1; pa.z += fz;
      *pp = pa;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&tim[0], (unsigned long long)(t1 - t0));
    atomicAdd(&tim[1], (unsigned long long)(clock64() - t1));
  }
}

int main(int argc, char **argv) {
  const int t = argc > 1 ? atoi(argv[1]) : 3;          // tile edge in cells
  const double keep = argc > 2 ? atof(argv[2]) : 3.15; // list radius (3.3 = outer list)
  const int nblocks = argc > 3 ? atoi(argv[3]) : 148 * 2 * 8;
  const int tpb = argc > 4 ? atoi(argv[4]) : 512;
  const double SL = 3.3, rho = 0.7, cutoff = 3.0, dt = 0.001;
  const int nc = t + 2;
  const double Lb = nc * SL;
  // jittered FCC at density rho
  std::mt19937 rng(1);
  std::uniform_real_distribution<double> U(-0.25, 0.25);
  const double a = cbrt(4.0 / rho);
  struct P { double x, y, z; int cell; bool interior; };
  std::vector<P> ps;
  const int nl = (int)(Lb / a) + 1;
  const double b[4][3] = {{0, 0, 0}, {.5, .5, 0}, {.5, 0, .5}, {0, .5, .5}};
  for (int i = 0; i < nl; i++) for (int j = 0; j < nl; j++) for (int k = 0; k < nl; k++) for (int s = 0; s < 4; s++) {
    P p{(i + b[s][0]) * a + U(rng) + 0.3, (j + b[s][1]) * a + U(rng) + 0.3, (k + b[s][2]) * a + U(rng) + 0.3, 0, false};
    if (p.x < 0 || p.x >= Lb || p.y < 0 || p.y >= Lb || p.z < 0 || p.z >= Lb) continue;
    int cx = (int)(p.x / SL), cy = (int)(p.y / SL), cz = (int)(p.z / SL);
    p.interior = cx >= 1 && cx <= t && cy >= 1 && cy <= t && cz >= 1 && cz <= t;
    p.cell = cx + nc * (cy + nc * cz);
    ps.push_back(p);
  }
  // staged order: interior first (cell order), then the halo (cell order)
  std::stable_sort(ps.begin(), ps.end(), [](const P &u, const P &v) {
    if (u.interior != v.interior) return u.interior;
    return u.cell < v.cell;
  });
  const int ns = (int)ps.size();
  int n_int = 0;
  for (auto &p : ps) n_int += p.interior;
  const int nfw = (n_int + 7) / 8;
  // rows: rainbow order by staged index mod 4
  std::vector<std::vector<int>> rows(n_int);
  long long nent = 0;
  for (int i = 0; i < n_int; i++) {
    std::vector<int> cls[4];
    for (int j = 0; j < ns; j++) {
      if (j == i) continue;
      double dx = ps[j].x - ps[i].x, dy = ps[j].y - ps[i].y, dz = ps[j].z - ps[i].z;
      if (dx * dx + dy * dy + dz * dz <= keep * keep) cls[j & 3].push_back(j);
    }
    size_t m = std::min(std::min(cls[0].size(), cls[1].size()), std::min(cls[2].size(), cls[3].size()));
    std::vector<int> &r = rows[i];
    for (size_t s = 0; s < m; s++) for (int c = 0; c < 4; c++) r.push_back(cls[c][s]);
    for (size_t s = m;; s++) {
      bool any = false;
      for (int c = 0; c < 4; c++) if (s < cls[c].size()) { r.push_back(cls[c][s]); any = true; }
      if (!any) break;
    }
    nent += r.size();
  }
  std::vector<int> fw_off(nfw), fw_len(nfw);
  std::vector<unsigned> words;
  long long walked = 0;
  for (int fw = 0; fw < nfw; fw++) {
    int steps = 0;
    for (int m = 0; m < 8 && fw * 8 + m < n_int; m++) steps = std::max(steps, (int)(rows[fw * 8 + m].size() + 3) / 4);
    steps = (steps + 1) & ~1;
    fw_off[fw] = (int)words.size();
    fw_len[fw] = steps;
    walked += 32LL * steps;
    for (int k2 = 0; k2 < steps / 2; k2++)
      for (int lane = 0; lane < 32; lane++) {
        const int a_ = fw * 8 + lane / 4, c = lane & 3;
        unsigned e[2];
        for (int h = 0; h < 2; h++) {
          const int idx = (2 * k2 + h) * 4 + c;
          e[h] = (a_ < n_int && idx < (int)rows[a_].size()) ? (unsigned)rows[a_][idx] : (unsigned)ns;
        }
        words.push_back(e[0] | (e[1] << 16));
      }
  }
  printf("tile %d^3 cells: staged %d records (%.1f KB), interior %d, force-warps %d, entries/local %.1f, walked/local %.1f (padding %.1f %%)\n",
         t, ns, (ns + 1) * 32 / 1024.0, n_int, nfw, (double)nent / n_int, (double)walked / n_int, 100.0 * (walked - nent) / nent);
  std::vector<double4> hq(ns);
  for (int i = 0; i < ns; i++) hq[i] = make_double4(ps[i].x, ps[i].y, ps[i].z, 0.0);
  double4 *dq, *dp;
  unsigned *dl;
  int *doff, *dlen;
  const long long stride = (long long)words.size();
  CK(cudaMalloc(&dq, sizeof(double4) * ns));
  CK(cudaMalloc(&dp, sizeof(double4) * (size_t)n_int * nblocks));
  CK(cudaMalloc(&dl, sizeof(unsigned) * (stride * nblocks + 256)));
  CK(cudaMalloc(&doff, sizeof(int) * nfw));
  CK(cudaMalloc(&dlen, sizeof(int) * nfw));
  CK(cudaMemcpy(dq, hq.data(), sizeof(double4) * ns, cudaMemcpyHostToDevice));
  CK(cudaMemset(dp, 0, sizeof(double4) * (size_t)n_int * nblocks));
  for (int b_ = 0; b_ < nblocks; b_++) CK(cudaMemcpy(dl + b_ * stride, words.data(), sizeof(unsigned) * stride, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(doff, fw_off.data(), sizeof(int) * nfw, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dlen, fw_len.data(), sizeof(int) * nfw, cudaMemcpyHostToDevice));
  unsigned long long *dtim;
  CK(cudaMalloc(&dtim, 16));
  FC fc;
  fc.CL2 = cutoff * cutoff;
  const double s2 = 1.0 / (cutoff * cutoff), s6 = s2 * s2 * s2, s8 = s6 * s2, s14 = s6 * s6 * s2;
  const double C2 = 6.0 * s14 - 3.0 * s8;
  fc.c24dt = 24 * dt; fc.c48dt = 48 * dt; fc.c2x8dt = 8 * C2 * dt; fc.dt2 = dt / 2;
  long long cb; memcpy(&cb, &fc.CL2, 8); fc.cl2_hi = (int)(cb >> 32);
  const size_t smem = (size_t)(ns + 1) * 32;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const double ghz = prop.clockRate * 1e-6;
  auto run = [&](int mode) {
    auto launch = [&]() {
      if (mode == 0) k_tile<0><<<nblocks, tpb, smem>>>(dq, ns, n_int, dl, doff, dlen, stride, dp, fc, dtim);
      if (mode == 1) k_tile<1><<<nblocks, tpb, smem>>>(dq, ns, n_int, dl, doff, dlen, stride, dp, fc, dtim);
      if (mode == 2) k_tile<2><<<nblocks, tpb, smem>>>(dq, ns, n_int, dl, doff, dlen, stride, dp, fc, dtim);
    };
    if (mode == 0) CK(cudaFuncSetAttribute(k_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (mode == 1) CK(cudaFuncSetAttribute(k_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (mode == 2) CK(cudaFuncSetAttribute(k_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int i = 0; i < 3; i++) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 10;
    CK(cudaMemset(dtim, 0, 16));
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    unsigned long long ht[2];
    CK(cudaMemcpy(ht, dtim, 16, cudaMemcpyDeviceToHost));
    printf("   per block: staging %.0f clk, walk %.0f clk (block-resident time; %d blocks)\n", (double)ht[0] / (reps * (double)nblocks),
           (double)ht[1] / (reps * (double)nblocks), nblocks);
    const double locals = (double)n_int * nblocks;
    const double warp_steps = (double)walked / 32 * nblocks;
    const double clk_per_ws = ms * 1e-3 * ghz * 1e9 * prop.multiProcessorCount / warp_steps;
    printf("mode %d (%s): %.3f ms for %.2f M locals -> %.2f ns/local/SM-normalised, %.2f clk per warp-step per SM (at %.3f GHz nominal), "
           "16,034,412 locals would take %.3f ms\n",
           mode, mode == 0 ? "full" : (mode == 1 ? "gathers only" : "FP64 only"), ms, locals * 1e-6,
           ms * 1e6 / locals * prop.multiProcessorCount, clk_per_ws, ghz, ms * 16034412.0 / locals);
  };
  run(0); run(1); run(2);
  // correctness against the host (first block)
  auto check = [&](const char *name, auto launch1) {
    CK(cudaMemset(dp, 0, sizeof(double4) * (size_t)n_int * nblocks));
    launch1();
    std::vector<double4> hp(n_int);
    CK(cudaMemcpy(hp.data(), dp, sizeof(double4) * n_int, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int i = 0; i < n_int; i++) {
      double fx = 0, fy = 0, fz = 0;
      for (int j : rows[i]) {
        double dx = ps[j].x - ps[i].x, dy = ps[j].y - ps[i].y, dz = ps[j].z - ps[i].z;
        double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > fc.CL2) continue;
        double r6 = r2 * r2 * r2;
        double df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2) + 8 * C2) * dt;
        fx += df * dx; fy += df * dy; fz += df * dz;
      }
      double n = sqrt(fx * fx + fy * fy + fz * fz) + 1e-3 * dt;
      worst = std::max(worst, std::max(fabs(hp[i].x - fx), std::max(fabs(hp[i].y - fy), fabs(hp[i].z - fz))) / n);
    }
    printf("   %s: worst relative impulse error vs host: %.3e\n", name, worst);
  };
  check("mode 0", [&]() { k_tile<0><<<1, tpb, smem>>>(dq, ns, n_int, dl, doff, dlen, stride, dp, fc, dtim); });
  // interleaved variants
  auto run_il = [&](const char *name, auto kern, int tpbv) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, tpbv, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    auto launch = [&]() { kern<<<nblocks, tpbv, smem>>>(dq, ns, n_int, dl, doff, dlen, stride, dp, fc, dtim); };
    for (int i = 0; i < 3; i++) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 10;
    CK(cudaMemset(dtim, 0, 16));
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    unsigned long long ht[2];
    CK(cudaMemcpy(ht, dtim, 16, cudaMemcpyDeviceToHost));
    const double locals = (double)n_int * nblocks;
    const double warp_steps = (double)walked / 32 * nblocks;
    printf("%s: %d regs, %d blocks/SM, staging %.0f clk, walk %.0f clk per block; %.3f ms -> %.2f clk per warp-step per SM, 16,034,412 locals would take %.3f ms\n",
           name, fa.numRegs, nb, (double)ht[0] / (reps * (double)nblocks), (double)ht[1] / (reps * (double)nblocks), ms,
           ms * 1e-3 * ghz * 1e9 * prop.multiProcessorCount / warp_steps, ms * 16034412.0 / locals);
    check(name, [&]() { kern<<<1, tpbv, smem>>>(dq, ns, n_int, dl, doff, dlen, stride, dp, fc, dtim); });
  };
  run_il("IL2 512x2", k_tile_il<2, 512, 2>, 512);
  run_il("IL4 512x2", k_tile_il<4, 512, 2>, 512);
  run_il("IL2 384x2", k_tile_il<2, 384, 2>, 384);
  run_il("IL4 384x2", k_tile_il<4, 384, 2>, 384);
  run_il("IL2 256x2", k_tile_il<2, 256, 2>, 256);
  run_il("IL4 256x2", k_tile_il<4, 256, 2>, 256);
  run_il("IL4 320x2", k_tile_il<4, 320, 2>, 320);
  // DFMA dependent-issue latency and throughput
  {
    double *dout; long long *dclk;
    CK(cudaMalloc(&dout, sizeof(double) * 148 * 1024));
    CK(cudaMalloc(&dclk, 8));
    long long hc;
    auto one = [&](const char *name, auto kern, int nch, int threads, int blocks) {
      kern<<<blocks, threads>>>(dout, dclk, 1.0, 0.5);
      kern<<<blocks, threads>>>(dout, dclk, 1.0, 0.5);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&hc, dclk, 8, cudaMemcpyDeviceToHost));
      printf("DFMA %s: %d chain(s)/thread, %d threads/block: %.2f clk per dependent DFMA round (%.2f clk per warp-instruction per SMSP)\n", name, nch,
             threads, hc / 4096.0, hc / 4096.0 / nch / ((threads + 127) / 128));
    };
    one("latency", k_dfma_lat<1>, 1, 32, 1);
    one("2 chains", k_dfma_lat<2>, 2, 32, 1);
    one("4 chains", k_dfma_lat<4>, 4, 32, 1);
    one("8 warps/SMSP", k_dfma_lat<1>, 1, 1024, 148);
    one("8 warps/SMSP x2", k_dfma_lat<2>, 2, 1024, 148);
    one("4 warps/SMSP", k_dfma_lat<1>, 1, 512, 148);
    one("6 warps/SMSP x2", k_dfma_lat<2>, 2, 768, 148);
  }
  return 0;
}
