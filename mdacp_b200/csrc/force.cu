// LJ force kick: ForceCalculator::CalculateForce (fcalculator.cc:35-50); semantic spec
// ForceCalculator::CalculateForceSorted (fcalculator.cc:182-223):
//   for each listed pair (i,j): d = q_j - q_i, r2 = |d|^2; if r2 > CL2 the pair contributes 0;
//   r6 = r2^3; df = ((24 r6 - 48)/(r6 r6 r2) + 8 C2) dt; p_i += df d; p_j -= df d.
//
// B200 formulation.
//  * FULL neighbour list, no atomics: every local's momentum is owned by exactly one thread, so the
//    summation order is deterministic and ghost momenta are never written (they are scratch in the
//    reference, fcalculator.cc:273-275 / SURVEY.md Q8).  RED.ADD.F64 for the reaction term of a half list
//    would cost ~3 P atomics per particle at ~1.3 clk/lane/SM -- several times this whole kernel.
//  * TWO locals per thread (slots a, a+1 of one cell, matched nearest neighbours, cells.cu:k_cell_match).
//    ncu shows the thread-per-particle version bound by the L1 data pipe on the scattered gather of
//    partner positions (~20 wavefronts per warp-gather, FP64 pipe ~30 %; profiles/r01_force_v*_ncu.txt);
//    the cost tracks gathered bytes whatever the layout.  Sharing each gathered partner between two
//    nearby locals cuts gathers per evaluation to ~0.65 at ~1.3x FP64 work (the union of two rows).
//    The a-b interaction itself is evaluated once and applied to both (Newton's third law in-thread).
//  * one 256-bit load per partner (LDG.E.256); list stored warp-interleaved (entry k of lane l at
//    base + 32 k + l) so index loads are one coalesced 128-byte line per warp; locals cell-sorted.
//  * cutoff / "is a listed partner of a (b)" tests run on the integer pipe (bit-pattern compare of the
//    non-negative r2, flag bits 31/30 of the entry), keeping the FP64 pipe for arithmetic.
//
// Roofline: algorithmic HBM bytes per particle = 4 B x P (half-list indices) + 96 B
// (q read, p read + write), 30 flop per listed unordered pair.
#include <stdlib.h>
#include "unit.h"

namespace mdb {

// 1 / x: hardware seed (MUFU.RCP64H) + one cubic (Halley-type) refinement y (1 + e + e^2);
// GUARD adds one linear Newton step (2 DFMA) that makes the result independent of the seed quality.
template <bool GUARD>
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  double t = fma(e, e, e);
  y = fma(y, t, y);
  if (GUARD) {
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
  }
  return y;
}

struct ForceConsts {
  double CL2;
  double c24dt;   // 24 dt
  double c48dt;   // 48 dt
  double c2x8dt;  // 8 C2 dt
  double dt;
  double c2x8;
};

// VARIANT 0: reference expression with an IEEE division (validation baseline)
// VARIANT 1: s = 1/r2 (seed + cubic + guard step), df = s^4 (24 - 48 s^3) dt + 8 C2 dt
// VARIANT 2: same without the guard step
template <int VARIANT>
__device__ __forceinline__ double pair_df(double r2, const ForceConsts &fc) {
  if (VARIANT == 0) {
    double r6 = r2 * r2 * r2;
    return ((24.0 * r6 - 48.0) / (r6 * r6 * r2) + fc.c2x8) * fc.dt;
  } else {
    double s = fast_rcp<VARIANT == 1>(r2);
    double s2 = s * s;
    double s3 = s2 * s;
    double s4 = s2 * s2;
    double t = fma(-fc.c48dt, s3, fc.c24dt);
    return fma(s4, t, fc.c2x8dt);
  }
}

__device__ __forceinline__ void ld_q256(const double4 *p, double &x, double &y, double &z) {
  double w;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
  (void)w;
}

#ifndef FORCE_UNROLL
#define FORCE_UNROLL 2
#endif
#ifndef FORCE_MINBLK
#define FORCE_MINBLK 8
#endif
template <int VARIANT, bool KICK>
__global__ void __launch_bounds__(128, FORCE_MINBLK)
k_force(const double4 *__restrict__ q, double4 *__restrict__ p, double *__restrict__ fout,
        const unsigned *__restrict__ list, const long long *__restrict__ warp_off, const int *__restrict__ warp_len,
        const unsigned *__restrict__ tp_info, const int *__restrict__ ref_of_slot, int nt, ForceConsts fc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = t >> 5;
  if ((long long)w * 32 >= nt) return;
  const int lane = threadIdx.x & 31;
  const bool act = t < nt;
  const unsigned info = tp_info[act ? t : nt - 1];
  const int a = (int)(info & 0x3FFFFFFFu);
  const bool two = (info & 0x40000000u) != 0;
  double ax, ay, az, bx, by, bz;
  ld_q256(q + a, ax, ay, az);
  ld_q256(q + (two ? a + 1 : a), bx, by, bz);
  const int len = warp_len[w];
  const unsigned *__restrict__ lp = list + warp_off[w] + lane;
  double fax = 0.0, fay = 0.0, faz = 0.0, fbx = 0.0, fby = 0.0, fbz = 0.0;
  const long long cl2b = __double_as_longlong(fc.CL2);

  constexpr int UNR = FORCE_UNROLL;
  unsigned en = (len > 0) ? __ldg(lp) : 0u;
#pragma unroll UNR
  for (int k = 0; k < len; k++) {
    const unsigned e = en;
    if (k + 1 < len) en = __ldg(lp + (long long)(k + 1) * 32);
    double jx, jy, jz;
    ld_q256(q + (e & 0x3FFFFFFFu), jx, jy, jz);
    {
      const double dx = jx - ax, dy = jy - ay, dz = jz - az;
      const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      double df = pair_df<VARIANT>(r2, fc);
      // r2 > CL2 (fcalculator.cc:209-211) or not a listed partner of a: no contribution
      if ((__double_as_longlong(r2) > cl2b) || !(e & 0x80000000u)) df = 0.0;
      fax = fma(df, dx, fax);
      fay = fma(df, dy, fay);
      faz = fma(df, dz, faz);
    }
    {
      const double dx = jx - bx, dy = jy - by, dz = jz - bz;
      const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      double df = pair_df<VARIANT>(r2, fc);
      if ((__double_as_longlong(r2) > cl2b) || !(e & 0x40000000u)) df = 0.0;
      fbx = fma(df, dx, fbx);
      fby = fma(df, dy, fby);
      fbz = fma(df, dz, fbz);
    }
  }
  if (!act) return;
  if (info & 0x80000000u) {  // the two locals of this thread are listed partners of each other
    const double dx = bx - ax, dy = by - ay, dz = bz - az;
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double df = pair_df<VARIANT>(r2, fc);
    if (__double_as_longlong(r2) > cl2b) df = 0.0;
    fax = fma(df, dx, fax); fay = fma(df, dy, fay); faz = fma(df, dz, faz);
    fbx = fma(-df, dx, fbx); fby = fma(-df, dy, fby); fbz = fma(-df, dz, fbz);
  }
  if (KICK) {
    double4 pa = p[a];
    pa.x += fax; pa.y += fay; pa.z += faz;
    p[a] = pa;
    if (two) {
      double4 pb = p[a + 1];
      pb.x += fbx; pb.y += fby; pb.z += fbz;
      p[a + 1] = pb;
    }
  } else {
    const long long ra = ref_of_slot[a];
    fout[3 * ra + 0] = fax; fout[3 * ra + 1] = fay; fout[3 * ra + 2] = faz;
    if (two) {
      const long long rb = ref_of_slot[a + 1];
      fout[3 * rb + 0] = fbx; fout[3 * rb + 1] = fby; fout[3 * rb + 2] = fbz;
    }
  }
}

static int force_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("MDB200_FORCE_VARIANT");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 2) v = 1;
  }
  return v;
}

static ForceConsts make_fc(const Unit *u) {
  ForceConsts fc;
  fc.CL2 = u->c.CL2;
  fc.dt = u->c.dt;
  fc.c24dt = 24.0 * u->c.dt;
  fc.c48dt = 48.0 * u->c.dt;
  fc.c2x8 = u->c.C2 * 8.0;
  fc.c2x8dt = u->c.C2 * 8.0 * u->c.dt;
  return fc;
}

#define FORCE_ARGS(F) u->q.ptr, u->p.ptr, F, u->list.ptr, u->warp_off.ptr, u->warp_len.ptr, u->tp_info.ptr, \
                      u->ref_of_slot.ptr, (int)u->nthreads, fc

int force_prepare(Unit *) { return MDB200_OK; }

int force_kick(Unit *u) {
  if (u->pn == 0 || u->nthreads == 0) return MDB200_OK;
  const int tpb = 128;
  const unsigned grid = (unsigned)((u->nthreads + tpb - 1) / tpb);
  ForceConsts fc = make_fc(u);
  switch (force_variant()) {
    case 0: k_force<0, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
    case 2: k_force<2, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
    default: k_force<1, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
  }
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int force_impulse(Unit *u, double *f3_host, long long cap_rows) {
  if (cap_rows < u->pn) {
    set_error("impulse buffer too small: %lld rows < pn %lld", cap_rows, u->pn);
    return MDB200_ECAPACITY;
  }
  if (u->pn == 0) return MDB200_OK;
  DevBuf<double> f;
  MDB_TRY(f.ensure((size_t)u->pn * 3));
  const int tpb = 128;
  const unsigned grid = (unsigned)((u->nthreads + tpb - 1) / tpb);
  ForceConsts fc = make_fc(u);
  switch (force_variant()) {
    case 0: k_force<0, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
    case 2: k_force<2, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
    default: k_force<1, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
  }
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  MDB_CUDA(cudaMemcpyAsync(f3_host, f.ptr, sizeof(double) * 3 * (size_t)u->pn, cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  f.release();
  return MDB200_OK;
}

}  // namespace mdb
