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
// Reduction is deterministic: one block per tile, per-block partials in block order, one final block.
#include "comm.h"
#include "tile.cuh"

namespace mdb {

TileGrid tile_grid(const Unit *u);  // tile_search.cu

static const int TPB = 512;   // two blocks of 512 threads per SM next to their staged tiles
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

// K over locals (grid-stride over particles) and U, W over the OUTER list, one block per tile (tile.cuh): the
// positions of the tile and its halo are staged in shared memory, kG lanes per local, each lane takes every
// kG-th entry of the row (unit.h).
__global__ void __launch_bounds__(TPB, 2)
k_observe(TileGrid g, Consts c, const int *__restrict__ cell_start, const int *__restrict__ gcell_start,
          const int *__restrict__ gitems, const int *__restrict__ tile_order, const int *__restrict__ tile_fw0,
          const double4 *__restrict__ q,
          const double4 *__restrict__ p, const uint2 *__restrict__ list, const long long *__restrict__ fw_off,
          const int *__restrict__ fw_len, int pn, int have_list, double *__restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double sh[TPB / 32];
  TileShared &ts = *reinterpret_cast<TileShared *>(smem);
  double4 *sq = reinterpret_cast<double4 *>(smem + tile_shared_offset());
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
    const int tile = tile_order[blockIdx.x];  // tiles with interior locals only (tile_search.cu)
    tile_setup(ts, tile, g, c, cell_start, gcell_start);
    tile_stage_tma(ts, q, gitems, sq);
    tile_stage_wait(ts);
    const int lane = threadIdx.x & 31;
    const int nfw = (ts.n_int + kPW - 1) / kPW, gfw0 = tile_fw0[tile];
    for (int fw = threadIdx.x >> 5; fw < nfw; fw += blockDim.x >> 5) {
      const int a = fw * kPW + lane / kG;
      const bool act = a < ts.n_int;  // the last force-warp of a tile may own fewer than 8 locals
      const double4 qa = sq[act ? a : ts.n_all];
      const uint2 *__restrict__ lp = list + fw_off[gfw0 + fw] + lane;
      const int steps = act ? fw_len[gfw0 + fw] : 0;
      double ue = 0.0, we = 0.0;
      for (int s0 = 0; s0 < steps; s0 += 4) {
        // padding entries point at the far-away sentinel record: r2 > CL2, no contribution
        const uint2 w = __ldg(lp + 8 * s0);  // chunk s0 / 4: 32 words per chunk
#pragma unroll
        for (int s = 0; s < 4; s++) {
          if (s0 + s < steps) {
            const unsigned j = ((s < 2 ? w.x : w.y) >> (16 * (s & 1))) & 0xffffu;
            const double4 qj = sq[j];
            pair_uw(qj.x - qa.x, qj.y - qa.y, qj.z - qa.z, c.CL2, c.C2, c.C0, ue, we);
          }
        }
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
  const bool walk = with_list && u->list_initialized && u->ntiles > 0;
  long long nb = walk ? std::max(u->n_active_tiles, 1) : (u->pn + TPB - 1) / TPB;
  if (!walk && nb > MAXBLK) nb = MAXBLK;
  MDB_TRY(u->partial.ensure((size_t)3 * nb + 8));
  const size_t sm = walk ? tile_smem_bytes_f64(u->tile_cap) : tile_shared_offset();
  MDB_CUDA(cudaFuncSetAttribute(k_observe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes_f64(walk ? u->tile_cap : 0)));
  k_observe<<<(unsigned)nb, TPB, sm, u->stream>>>(tile_grid(u), u->c, u->cell_start.ptr, u->gcell_start.ptr, u->gcell_items.ptr,
                                                  u->tile_order.ptr, u->tile_fw0.ptr, u->q.ptr, u->p.ptr, u->list.ptr, u->fw_off.ptr, u->fw_len.ptr,
                                                  (int)u->pn, walk ? 1 : 0, u->partial.ptr);
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
