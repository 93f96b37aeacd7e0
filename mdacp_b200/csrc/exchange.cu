// Ghost (border particle) refresh: MDUnit::MakeBufferForBorderParticles (mdunit.cc:137-176) fused
// with MDUnit::ReceiveBorderParticles (mdunit.cc:179-191).  The gather kernel reads the fixed
// border list built at the last rebuild, adds the periodic shift of mdunit.cc:141-163 and stores
// straight into the destination unit's ghost window -- which may be this unit (periodic self
// neighbour), another unit on the same device, a peer device mapped over NVLink, or a pack
// buffer that NCCL ships to another process.  24 B of payload per ghost (32 B stored: double4).
#include <stdlib.h>
#include "unit.h"

namespace mdb {

// dq2 (may be null): the same ghost window in the destination's second position buffer -- positions are
// double-buffered (force.cu writes the drifted locals into the other buffer), ghosts live in both.
__global__ void k_ghost_put(const double4 *__restrict__ sq, const int *__restrict__ border,
                            double4 *__restrict__ dq, double4 *__restrict__ dq2, int n, double sx, double sy,
                            double sz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 v = sq[border[i]];
  // q + diff per component, as mdunit.cc:167 (adding 0.0 is exact)
  v = make_double4(v.x + sx, v.y + sy, v.z + sz, v.y + sy);  // lane 3 repeats y (tile.cuh)
  dq[i] = v;
  if (dq2) dq2[i] = v;
}

int ghost_put_to(Unit *s, int dir, double4 *dst, double4 *dst2, cudaStream_t st) {
  long long n = s->n_border[dir];
  if (n == 0) return MDB200_OK;
  k_ghost_put<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->q.ptr, s->border[dir].ptr, dst, dst2, (int)n,
                                                          s->shift[dir][0], s->shift[dir][1], s->shift[dir][2]);
  s->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int ghost_put(Unit *s, int dir, Unit *d, long long dst_off) {
  return ghost_put_to(s, dir, d->q.ptr + dst_off, d->q_alt.ptr + dst_off, s->stream);
}

// ---------------------------------------------------------------------------------------------------------
// One unit that is its own neighbour in all six directions (a periodic box on one GPU).  The reference refreshes the
// ghosts direction by direction because the y borders include ghosts received in x, and the z borders ghosts of both
// (mdmanager.cc:283-292).  The border lists are fixed between rebuilds, so that chain is resolved ONCE per rebuild:
// ghost g is the image of local ghost_src[g] displaced by ghost_shift[g] -- the shifts of the directions it travelled
// through, at most one per axis, so each component is still ONE addition q + shift exactly as in mdunit.cc:167 --
// and the per-step refresh is a single gather.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_ghost_compose(const int *__restrict__ border, int n, int pn, int dst0, int *__restrict__ gsrc,
                                double4 *__restrict__ gshift, double sx, double sy, double sz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = border[i];
  int src = s;
  double4 sh = make_double4(sx, sy, sz, 0.0);
  if (s >= pn) {  // a ghost of an earlier direction: its source and its shift so far
    src = gsrc[s - pn];
    const double4 t = gshift[s - pn];
    sh.x += t.x; sh.y += t.y; sh.z += t.z;   // disjoint components: x + 0.0 etc., exact
  }
  gsrc[dst0 + i] = src;
  gshift[dst0 + i] = sh;
}

__global__ void k_ghost_all(const double4 *__restrict__ sq, const int *__restrict__ gsrc, const double4 *__restrict__ gshift,
                            double4 *__restrict__ dq, double4 *__restrict__ dq2, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 v = sq[gsrc[i]];
  const double4 sh = gshift[i];
  const double4 w = make_double4(v.x + sh.x, v.y + sh.y, v.z + sh.z, v.y + sh.y);  // lane 3 repeats y (tile.cuh)
  dq[i] = w;
  dq2[i] = w;
}

int ghost_compose(Unit *u) {
  u->ghost_composed = false;
  static const bool off = getenv("MDB200_NO_GHOST_COMPOSE") != nullptr;
  if (off) return MDB200_OK;
  for (int dir = 0; dir < 6; dir++)
    if (u->nbr[dir] != u || u->src[dir] != u) return MDB200_OK;
  const long long ng = u->tn - u->pn;
  if (ng <= 0) return MDB200_OK;
  MDB_TRY(u->ghost_src.ensure((size_t)ng));
  MDB_TRY(u->ghost_shift.ensure((size_t)ng));
  for (int dir = 0; dir < 6; dir++) {
    const long long n = u->n_border[dir];
    if (n == 0) continue;
    if (n != u->recv_cnt[dir]) { set_error("ghost_compose: border / window size mismatch"); return MDB200_EINVAL; }
    k_ghost_compose<<<(unsigned)((n + 255) / 256), 256, 0, u->stream>>>(u->border[dir].ptr, (int)n, (int)u->pn,
                                                                       (int)(u->recv_off[dir] - u->pn), u->ghost_src.ptr,
                                                                       u->ghost_shift.ptr, u->shift[dir][0], u->shift[dir][1],
                                                                       u->shift[dir][2]);
    u->launches++;
  }
  MDB_CUDA(cudaGetLastError());
  u->ghost_composed = true;
  return MDB200_OK;
}

int ghost_refresh_composed(Unit *u, cudaStream_t st) {
  const long long ng = u->tn - u->pn;
  if (ng <= 0) return MDB200_OK;
  k_ghost_all<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(u->q.ptr, u->ghost_src.ptr, u->ghost_shift.ptr, u->q.ptr + u->pn,
                                                           u->q_alt.ptr + u->pn, (int)ng);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int ghost_pack(Unit *s, int dir) {
  MDB_TRY(s->send_buf.ensure((size_t)s->n_border[dir] + 1));
  return ghost_put_to(s, dir, s->send_buf.ptr, nullptr, s->stream);
}

}  // namespace mdb
