// Observables: KineticEnergyObserver, PotentialEnergyObserver, VirialObserver (observer.cc:5-79).
//
//   K = sum_{i<pn} sum_{d<4} 0.5 p_id^2                                     observer.cc:5-15
//   U = sum_pairs [r2 <= CL2] 4 (r2^-6 - r2^-3 + C2 r2 + C0) * w            observer.cc:18-48
//   W = sum_pairs [r2 <= CL2] ((24 r6 - 48)/(r6 r6 r2) + 8 C2) r2 * w / 3   observer.cc:51-79
// with w = 0.5 when the pair touches a ghost.  The reference walks its half list (each
// local-local pair once with w = 1, each local-ghost pair once with w = 0.5); the full list
// used here holds each local-local pair twice and each local-ghost pair once, so a uniform
// w = 0.5 over all directed entries gives the same sums.  Positions are used as they are at
// the time of the call: locals after the second half drift, ghosts as of the last refresh
// (half a step older) -- the reference's "stale ghost" definition (SURVEY.md finding 2).
//
// Reduction is deterministic: fixed grid, per-block partials in block order, one final block.
#include "comm.h"
#include "unit.h"

namespace mdb {

static const int TPB = 256;
static const int MAXBLK = 1184;  // 8 x 148 SMs

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += sh[w];
  __syncthreads();
  return r;  // valid on thread 0
}

// 1 / x to full double precision without the division sequence: hardware seed (MUFU.RCP64H), one cubic and one
// linear Newton step (as in force.cu)
__device__ __forceinline__ double rcp_full(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, fma(e, e, e), y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// observer.cc:18-79 with s3 = r^-6:  U term 4 (s3^2 - s3 + C2 r2 + C0);
// W term ((24 r6 - 48) / (r6 r6 r2) + 8 C2) r2 = 24 s3 - 48 s3^2 + 8 C2 r2
__device__ __forceinline__ void pair_uw(double dx, double dy, double dz, double CL2, double C2, double C0,
                                        double &ue, double &we) {
  const double r2 = dx * dx + dy * dy + dz * dz;
  if (r2 > CL2) return;
  const double s = rcp_full(r2);
  const double s3 = s * s * s;
  const double c2r2 = C2 * r2;
  ue += 4.0 * (fma(s3, s3, -s3) + c2r2 + C0);
  we += fma(-48.0 * s3, s3, 24.0 * s3) + 8.0 * c2r2;
}

// K over locals (grid-stride over particles) and U, W over the list (grid-stride over force-warps: kG lanes
// per local, each lane takes every kG-th entry of the row, unit.h).
__global__ void __launch_bounds__(TPB)
k_observe(const double4 *__restrict__ q, const double4 *__restrict__ p, const unsigned *__restrict__ list,
          const long long *__restrict__ warp_off, const int *__restrict__ warp_len, int pn, long long nwarps,
          int have_list, double CL2, double C2, double C0, double *__restrict__ partial) {
  __shared__ double sh[TPB / 32];
  double k = 0.0, uu = 0.0, ww = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pn;
       i += (long long)gridDim.x * blockDim.x) {
    const double4 pp = p[i];
    k += 0.5 * pp.x * pp.x;
    k += 0.5 * pp.y * pp.y;
    k += 0.5 * pp.z * pp.z;
    k += 0.5 * pp.w * pp.w;
  }
  if (have_list) {
    const int lane = threadIdx.x & 31;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nwarps; w += wstride) {
      const long long a = w * kPW + lane / kG;
      const double4 qa = q[a < pn ? a : (long long)pn - 1];
      const uint2 *__restrict__ lp = reinterpret_cast<const uint2 *>(list + warp_off[w]) + lane;
      const int len2 = warp_len[w] >> 1;  // one 8-byte load = the entries of two steps (unit.h)
      double ue = 0.0, we = 0.0;
      for (int k = 0; k < len2; k++) {
        // padding entries point at the far-away sentinel particle: r2 > CL2, no contribution
        const uint2 e = __ldg(lp + (long long)k * 32);
        const double4 q0 = q[e.x], q1 = q[e.y];
        pair_uw(q0.x - qa.x, q0.y - qa.y, q0.z - qa.z, CL2, C2, C0, ue, we);
        pair_uw(q1.x - qa.x, q1.y - qa.y, q1.z - qa.z, CL2, C2, C0, ue, we);
      }
      // every directed entry carries weight 0.5 (see the header)
      uu += 0.5 * ue;
      ww += 0.5 * we;
    }
  }
  double s0 = block_sum(k, sh), s1 = block_sum(uu, sh), s2 = block_sum(ww, sh);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x + 0] = s0;
    partial[3 * blockIdx.x + 1] = s1;
    partial[3 * blockIdx.x + 2] = s2;
  }
}

__global__ void __launch_bounds__(TPB) k_observe_final(const double *__restrict__ partial, int nb, Ctrl *ctrl) {
  __shared__ double sh[TPB / 32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    a += partial[3 * i + 0];
    b += partial[3 * i + 1];
    c += partial[3 * i + 2];
  }
  double s0 = block_sum(a, sh), s1 = block_sum(b, sh), s2 = block_sum(c, sh);
  if (threadIdx.x == 0) {
    ctrl->obs[0] = s0;
    ctrl->obs[1] = s1;
    ctrl->obs[2] = s2 / 3.0;  // observer.cc:78
  }
}

static int observe_impl(Unit *u, double out[4], bool with_list) {
  out[0] = out[1] = out[2] = 0.0;
  out[3] = (double)u->pn;
  if (u->pn == 0) return MDB200_OK;
  long long nb = (u->pn * kG + TPB - 1) / TPB;
  if (nb > MAXBLK) nb = MAXBLK;
  MDB_TRY(u->partial.ensure((size_t)3 * MAXBLK + 8));
  k_observe<<<(unsigned)nb, TPB, 0, u->stream>>>(u->q.ptr, u->p.ptr, u->list.ptr, u->warp_off.ptr, u->warp_len.ptr,
                                                  (int)u->pn, u->nwarps,
                                                  (with_list && u->list_initialized) ? 1 : 0, u->c.CL2, u->c.C2, u->c.C0,
                                                  u->partial.ptr);
  k_observe_final<<<1, TPB, 0, u->stream>>>(u->partial.ptr, (int)nb, u->ctrl);
  u->launches += 2;
  MDB_CUDA(cudaGetLastError());
  MDB_CUDA(cudaMemcpyAsync(u->ctrl_host->obs, u->ctrl->obs, 3 * sizeof(double), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  out[0] = u->ctrl_host->obs[0];
  out[1] = u->ctrl_host->obs[1];
  out[2] = u->ctrl_host->obs[2];
  return MDB200_OK;
}

int observe_unit(Unit *u, double out[4]) { return observe_impl(u, out, true); }
// K only (Temperature() inside the Nose-Hoover step, mdmanager.cc:161,169)
int observe_kinetic(Unit *u, double out[4]) { return observe_impl(u, out, false); }

}  // namespace mdb
