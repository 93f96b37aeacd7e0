// LJ force kick: ForceCalculator::CalculateForce (fcalculator.cc:35-50); semantic spec
// ForceCalculator::CalculateForceSorted (fcalculator.cc:182-223):
//   for each listed pair (i,j): d = q_j - q_i, r2 = |d|^2; if r2 > CL2 the pair contributes 0;
//   r6 = r2^3; df = ((24 r6 - 48)/(r6 r6 r2) + 8 C2) dt; p_i += df d; p_j -= df d.
//
// B200 formulation.  One thread per LOCAL particle walks its row of the FULL neighbour list
// (both (i,j) and (j,i) are stored), so each thread owns p_i exclusively: no atomics, no
// reaction term, deterministic summation order, and ghost momenta are never written (they are
// scratch in the reference, fcalculator.cc:273-275 / SURVEY.md Q8).  The list is stored
// warp-interleaved (entry t of lane l at base + 32 t + l) so every index load is one coalesced
// 128-byte line per warp, and locals are cell-sorted so the gathered partner positions of a
// warp sit in a few neighbouring cells and hit in L1/L2 instead of HBM.
//
// Roofline: FP64 pipe.  Per evaluated (directed) pair: 3 sub + 3 (r2) + reciprocal
// + 5 (LJ polynomial) + 3 fma; algorithmic HBM bytes per particle = 4 B x P (half-list indices)
// + 96 B (q read, p read + write) with P listed unordered pairs per particle.
#include <stdlib.h>
#include "unit.h"

namespace mdb {

// 1 / x to full double precision: hardware seed (MUFU.RCP64H, ~20+ bits) + one cubic
// (Halley-type) refinement y (1 + e + e^2) + one linear Newton step as a guard; the latter
// costs 2 DFMA and keeps the result within 1 ulp irrespective of the seed quality.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  double t = fma(e, e, e);
  y = fma(y, t, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
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
// VARIANT 1: reciprocal-polynomial form: s = 1/r2, df = s^4 (24 - 48 s^3) dt + 8 C2 dt
template <int VARIANT>
__device__ __forceinline__ double pair_df(double r2, const ForceConsts &fc) {
  if (VARIANT == 0) {
    double r6 = r2 * r2 * r2;
    return ((24.0 * r6 - 48.0) / (r6 * r6 * r2) + fc.c2x8) * fc.dt;
  } else {
    double s = fast_rcp(r2);
    double s2 = s * s;
    double s3 = s2 * s;
    double s4 = s2 * s2;
    double t = fma(-fc.c48dt, s3, fc.c24dt);
    return fma(s4, t, fc.c2x8dt);
  }
}

// one 256-bit load per gathered partner (LDG.E.256, sm_100+): a single L1 request per pair
// instead of a 128-bit + 64-bit pair of requests that each walk the same 32 sectors
__device__ __forceinline__ void ld_q256(const double4 *p, double &x, double &y, double &z) {
  double w;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
}

__global__ void k_split_positions(const double4 *__restrict__ q, double *__restrict__ x, double *__restrict__ y,
                                  double *__restrict__ z, double2 *__restrict__ xy, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 v = q[i];
  x[i] = v.x; y[i] = v.y; z[i] = v.z;
  xy[i] = make_double2(v.x, v.y);
}

template <int VARIANT, bool KICK>
__global__ void __launch_bounds__(128)
k_force(const double4 *__restrict__ q, const double *__restrict__ sx, const double *__restrict__ sy,
        const double *__restrict__ sz, const double2 *__restrict__ sxy, double4 *__restrict__ p, double *__restrict__ fout,
        const int *__restrict__ list, const long long *__restrict__ warp_off, const int *__restrict__ warp_len,
        const int *__restrict__ ref_of_slot, int pn, ForceConsts fc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = i >> 5;
  if ((long long)w * 32 >= pn) return;
  const int lane = threadIdx.x & 31;
  const bool act = i < pn;
  const int ii = act ? i : pn - 1;
  const double2 qxy = __ldg(reinterpret_cast<const double2 *>(q + ii));
  const double qz = __ldg(reinterpret_cast<const double *>(q + ii) + 2);
  const int len = warp_len[w];
  const int *__restrict__ lp = list + warp_off[w] + lane;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  const long long cl2b = __double_as_longlong(fc.CL2);

  int jn = (len > 0) ? __ldg(lp) : 0;
  for (int t = 0; t < len; t++) {
    const int j = jn;
    if (t + 1 < len) jn = __ldg(lp + (long long)(t + 1) * 32);
    double jx, jy, jz;
    if (VARIANT == 3) {
      jx = __ldg(sx + j); jy = __ldg(sy + j); jz = __ldg(sz + j);
    } else if (VARIANT == 4) {
      const double2 jxy = __ldg(sxy + j);
      jx = jxy.x; jy = jxy.y; jz = __ldg(sz + j);
    } else if (VARIANT == 2) {
      ld_q256(q + j, jx, jy, jz);
    } else {
      const double2 jxy = __ldg(reinterpret_cast<const double2 *>(q + j));
      jx = jxy.x; jy = jxy.y;
      jz = __ldg(reinterpret_cast<const double *>(q + j) + 2);
    }
    const double dx = jx - qxy.x;
    const double dy = jy - qxy.y;
    const double dz = jz - qz;
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double df = pair_df<VARIANT>(r2, fc);
    // r2 > CL2 on non-negative doubles == integer compare of the bit patterns: runs on the
    // integer pipe, keeping the FP64 pipe for arithmetic (fcalculator.cc:209-211: df = 0)
    if (__double_as_longlong(r2) > cl2b) df = 0.0;
    fx = fma(df, dx, fx);
    fy = fma(df, dy, fy);
    fz = fma(df, dz, fz);
  }
  if (!act) return;
  if (KICK) {
    double4 pp = p[i];
    pp.x += fx; pp.y += fy; pp.z += fz;
    p[i] = pp;
  } else {
    const int r = ref_of_slot[i];
    fout[3 * (long long)r + 0] = fx;
    fout[3 * (long long)r + 1] = fy;
    fout[3 * (long long)r + 2] = fz;
  }
}

static int force_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("MDB200_FORCE_VARIANT");
    v = e ? atoi(e) : 2;
    if (v < 0 || v > 4) v = 2;
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

static DevBuf<double> g_sx, g_sy, g_sz;
static DevBuf<double2> g_sxy;
int force_prepare(Unit *u) {
  const int v = force_variant();
  if (v != 3 && v != 4) return MDB200_OK;
  MDB_TRY(g_sx.ensure((size_t)u->cap_t + 2)); MDB_TRY(g_sy.ensure((size_t)u->cap_t + 2));
  MDB_TRY(g_sz.ensure((size_t)u->cap_t + 2)); MDB_TRY(g_sxy.ensure((size_t)u->cap_t + 2));
  const long long n = u->cap_t + 1;
  k_split_positions<<<(unsigned)((n + 255) / 256), 256, 0, u->stream>>>(u->q.ptr, g_sx.ptr, g_sy.ptr, g_sz.ptr, g_sxy.ptr, n);
  u->launches++;
  return MDB200_OK;
}
#define FORCE_ARGS(F) u->q.ptr, g_sx.ptr, g_sy.ptr, g_sz.ptr, g_sxy.ptr, u->p.ptr, F, u->list.ptr, u->warp_off.ptr, \
                      u->warp_len.ptr, u->ref_of_slot.ptr, (int)u->pn, fc
int force_kick(Unit *u) {
  if (u->pn == 0) return MDB200_OK;
  const int tpb = 128;
  const unsigned grid = (unsigned)((u->pn + tpb - 1) / tpb);
  ForceConsts fc = make_fc(u);
  switch (force_variant()) {
    case 0: k_force<0, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
    case 1: k_force<1, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
    case 3: k_force<3, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
    case 4: k_force<4, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
    default: k_force<2, true><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr)); break;
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
  const unsigned grid = (unsigned)((u->pn + tpb - 1) / tpb);
  ForceConsts fc = make_fc(u);
  MDB_TRY(force_prepare(u));
  switch (force_variant()) {
    case 0: k_force<0, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
    case 1: k_force<1, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
    case 3: k_force<3, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
    case 4: k_force<4, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
    default: k_force<2, false><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(f.ptr)); break;
  }
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  MDB_CUDA(cudaMemcpyAsync(f3_host, f.ptr, sizeof(double) * 3 * (size_t)u->pn, cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  f.release();
  return MDB200_OK;
}

}  // namespace mdb
