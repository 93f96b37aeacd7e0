// Ghost (border particle) refresh: MDUnit::MakeBufferForBorderParticles (mdunit.cc:137-176) fused
// with MDUnit::ReceiveBorderParticles (mdunit.cc:179-191).  The gather kernel reads the fixed
// border list built at the last rebuild, adds the periodic shift of mdunit.cc:141-163 and stores
// straight into the destination unit's ghost window -- which may be this unit (periodic self
// neighbour), another unit on the same device, a peer device mapped over NVLink, or a pack
// buffer that NCCL ships to another process.  24 B of payload per ghost (32 B stored: double4).
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

int ghost_pack(Unit *s, int dir) {
  MDB_TRY(s->send_buf.ensure((size_t)s->n_border[dir] + 1));
  return ghost_put_to(s, dir, s->send_buf.ptr, nullptr, s->stream);
}

}  // namespace mdb
