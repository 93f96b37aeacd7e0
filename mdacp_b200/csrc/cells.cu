// Rebuild path: migration bookkeeping, periodic wrap, GPU cell sort (counting sort of particles
// into cells) and ghost (border) lists.
//
// Reference semantics restated here (file:line under the reference tree):
//   migration predicate   MDRect::IsOverBoundary            mdrect.cc:5-21
//   stable compaction     MDUnit::MakeBufferForSendingParticle mdunit.cc:78-101
//   append of arrivals    MDUnit::ReceiveParticles          mdunit.cc:104-122
//   periodic wrap + check Variables::AdjustPeriodicBoundary variables.cc:59-79
//   cell geometry         MeshList::ChangeScale             meshlist.cc:31-55
//   counting sort         MeshList::MakeMesh                meshlist.cc:153-205
//   ghost predicate       MDRect::IsInsideEdge              mdrect.cc:24-47
//   border lists          MDUnit::FindBorderParticles       mdunit.cc:125-134
//   (pair search and list build: tile_search.cu)
//   snapshot              PairList::Init                    pairlist.cc:11-23
//
// Index spaces.  "reference index" r is the index the reference's Variables arrays would use
// (initial order, stable compaction at migration, arrivals appended).  "slot" is where the
// particle lives in device memory: locals are kept CELL-SORTED in slots [0,pn) so that the
// force kernel's gathers of partner positions hit nearby lines; ghosts keep slot == reference
// index in [pn,tn).  ref_of_slot / slot_of_ref translate; results handed back through the
// C-ABI are always in reference indices.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <utility>
#include "unit.h"

namespace mdb {

static const int TPB = 256;
static inline unsigned nblk(long long n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

// ---------------------------------------------------------------------------------------------
// migration
// ---------------------------------------------------------------------------------------------
// psrc (late momenta, unit.h): upload row of the local in slot r -- straight after an upload slot = reference index = row
__global__ void k_begin_order(const int *__restrict__ slot_of_ref, int *__restrict__ ord, int *__restrict__ psrc, int pn) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < pn) {
    ord[r] = slot_of_ref[r];
    if (psrc) psrc[r] = r;
  }
}

__device__ __forceinline__ bool over_boundary(int dir, const double4 &q, const Consts &c) {
  switch (dir) {  // mdrect.cc:5-21
    case 0: return q.x < c.s[0];
    case 1: return q.x >= c.e[0];
    case 2: return q.y < c.s[1];
    case 3: return q.y >= c.e[1];
    case 4: return q.z < c.s[2];
    default: return q.z >= c.e[2];
  }
}

__global__ void k_mig_flag(const double4 *__restrict__ q, const int *__restrict__ ord, int *__restrict__ flags,
                           int pn, int dir, Consts c, Ctrl *ctrl) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  int f = 0;
  if (k < pn) {
    f = over_boundary(dir, q[ord[k]], c) ? 1 : 0;
    flags[k] = f;
  }
  int s = __syncthreads_count(f);
  if (threadIdx.x == 0 && s) atomicAdd(&ctrl->scratch[0], s);
}

__global__ void k_mig_compact(const double4 *__restrict__ q, const double4 *__restrict__ p,
                              const int *__restrict__ type, const int *__restrict__ ord,
                              const int *__restrict__ flags, const int *__restrict__ scanv,
                              int *__restrict__ ord_new, double4 *__restrict__ sq, double4 *__restrict__ sp,
                              int *__restrict__ st, const int *__restrict__ psrc, int *__restrict__ ss, int pn) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= pn) return;
  int h = ord[k];
  int j = scanv[k];
  if (flags[k]) {
    sq[j] = q[h];
    if (psrc) ss[j] = psrc[h];  // late momenta: the upload row travels instead
    else sp[j] = p[h];
    st[j] = type[h];
  } else {
    ord_new[k - j] = h;
  }
}

__global__ void k_mig_unpack(double4 *__restrict__ q, double4 *__restrict__ p, int *__restrict__ type,
                             int *__restrict__ ord, const double4 *__restrict__ rq,
                             const double4 *__restrict__ rp, const int *__restrict__ rt, int *__restrict__ psrc,
                             const int *__restrict__ rs, int n, int pn, int tail) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  q[tail + i] = rq[i];
  if (psrc) psrc[tail + i] = rs[i];
  else p[tail + i] = rp[i];
  type[tail + i] = rt[i];
  ord[pn + i] = tail + i;
}

int rebuild_begin_order(Unit *u) {
  MDB_TRY(u->ord.ensure((size_t)u->cap_p + 1));
  MDB_TRY(u->ord_alt.ensure((size_t)u->cap_p + 1));
  // migration buffers sized once, well above the few per mille of pn that move per rebuild
  {
    const size_t m = (size_t)(u->pn / 64) + 4096;
    MDB_TRY(u->mig_send_q.ensure(m)); MDB_TRY(u->mig_send_p.ensure(m)); MDB_TRY(u->mig_send_t.ensure(m));
    MDB_TRY(u->mig_recv_q.ensure(m)); MDB_TRY(u->mig_recv_p.ensure(m)); MDB_TRY(u->mig_recv_t.ensure(m));
  }
  if (u->late_p) {
    MDB_TRY(u->psrc.ensure((size_t)u->cap_p + 1));
    MDB_TRY(u->psrc_alt.ensure((size_t)u->cap_p + 1));
    MDB_TRY(u->mig_send_s.ensure(u->mig_send_t.cap));
  }
  if (u->pn > 0) {
    k_begin_order<<<nblk(u->pn), TPB, 0, u->stream>>>(u->slot_of_ref.ptr, u->ord.ptr, u->late_p ? u->psrc.ptr : nullptr, (int)u->pn);
    u->launches++;
  }
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int migrate_flag_and_count(Unit *u, int dir, long long *n_leave) {
  *n_leave = 0;
  if (u->pn == 0) return MDB200_OK;
  MDB_TRY(u->flags.ensure((size_t)u->cap_t + 2));
  MDB_CUDA(cudaMemsetAsync(&u->ctrl->scratch[0], 0, sizeof(int), u->stream));
  k_mig_flag<<<nblk(u->pn), TPB, 0, u->stream>>>(u->q.ptr, u->ord.ptr, u->flags.ptr, (int)u->pn, dir, u->c,
                                                  u->ctrl);
  u->launches++;
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[0], &u->ctrl->scratch[0], sizeof(int),
                           cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  *n_leave = u->ctrl_host->scratch[0];
  return MDB200_OK;
}

int migrate_compact_pack(Unit *u, int dir, long long n_leave) {
  (void)dir;
  if (n_leave == 0) return MDB200_OK;
  MDB_TRY(u->scan.ensure((size_t)u->cap_t + 2));
  MDB_TRY(u->scan_ws.ensure(2 * scan_workspace_ints(u->cap_t + 2)));
  MDB_TRY(u->mig_send_q.ensure((size_t)n_leave));
  MDB_TRY(u->mig_send_p.ensure((size_t)n_leave));
  MDB_TRY(u->mig_send_t.ensure((size_t)n_leave));
  if (u->late_p) MDB_TRY(u->mig_send_s.ensure((size_t)n_leave));
  MDB_TRY(exclusive_scan_i32(u->flags.ptr, u->scan.ptr, u->pn, u->scan_ws.ptr, u->stream, &u->launches));
  k_mig_compact<<<nblk(u->pn), TPB, 0, u->stream>>>(u->q.ptr, u->p.ptr, u->type.ptr, u->ord.ptr, u->flags.ptr,
                                                     u->scan.ptr, u->ord_alt.ptr, u->mig_send_q.ptr,
                                                     u->mig_send_p.ptr, u->mig_send_t.ptr,
                                                     u->late_p ? u->psrc.ptr : nullptr, u->mig_send_s.ptr, (int)u->pn);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  std::swap(u->ord, u->ord_alt);
  u->pn -= n_leave;
  return MDB200_OK;
}

// ---------------------------------------------------------------------------------------------
// capacity growth (rebuild time only)
// ---------------------------------------------------------------------------------------------
__global__ void k_write_sentinel(double4 *q, long long idx) {
  // far-away dummy particle: every padded list entry points here; r^2 is huge so the cutoff
  // select zeroes its contribution without any special case in the force loop
  q[idx] = make_double4(1.0e8, 1.0e8, 1.0e8, 0.0);
}

int ensure_capacity(Unit *u, long long need_p, long long need_t) {
  bool grow = false;
  long long ncp = u->cap_p, nct = u->cap_t;
  // MDB200_DEBUG_SMALL_CAPACITY (test hook, see mdb200_upload): no head-room, so that rebuilds keep re-allocating
  static const bool small_cap = getenv("MDB200_DEBUG_SMALL_CAPACITY") != nullptr;
  if (need_p > u->cap_p) { ncp = small_cap ? need_p + 16 : need_p + need_p / 4 + 1024; grow = true; }
  if (need_t > u->cap_t || ncp > u->cap_t) {
    long long want = need_t > ncp ? need_t : ncp;
    nct = small_cap ? want + 16 : want + want / 4 + 4096;
    grow = true;
  }
  if (!grow) return MDB200_OK;
  // momenta still travelling into p (mdb200_upload_begin): let them arrive before the buffer is replaced
  if (u->p_pending && u->ustream) MDB_CUDA(cudaStreamSynchronize(u->ustream));
  cudaStream_t st = u->stream;
  MDB_TRY(u->q.ensure((size_t)nct + 1, true, st));
  MDB_TRY(u->q_alt.ensure((size_t)nct + 1, true, st));  // ghosts live in both position buffers
  MDB_TRY(u->q0.ensure((size_t)nct + 1, true, st));
  MDB_TRY(u->p.ensure((size_t)ncp + 1, true, st));
  MDB_TRY(u->p_alt.ensure((size_t)ncp + 1, false, st));
  MDB_TRY(u->type.ensure((size_t)ncp + 1, true, st));
  MDB_TRY(u->type_alt.ensure((size_t)ncp + 1, false, st));
  MDB_TRY(u->ref_of_slot.ensure((size_t)ncp + 1, true, st));
  MDB_TRY(u->slot_of_ref.ensure((size_t)ncp + 1, true, st));
  MDB_TRY(u->ord.ensure((size_t)ncp + 1, true, st));
  MDB_TRY(u->ord_alt.ensure((size_t)ncp + 1, false, st));
  MDB_TRY(u->flags.ensure((size_t)nct + 2, false, st));
  MDB_TRY(u->scan.ensure((size_t)nct + 2, false, st));
  MDB_TRY(u->scan_ws.ensure(2 * scan_workspace_ints(nct + 2), false, st));
  MDB_TRY(u->cell_of.ensure((size_t)nct + 1, false, st));
  MDB_TRY(u->cell_items.ensure((size_t)ncp + 1, false, st));
  MDB_TRY(u->gcell_items.ensure((size_t)nct + 1, false, st));
  MDB_TRY(u->row_count.ensure((size_t)ncp + 32, false, st));
  MDB_TRY(u->disp.ensure((size_t)nct + 1, false, st));
  // use the smallest real capacity so that the sentinel index is valid in both q buffers
  u->cap_p = ncp;
  u->cap_t = nct;
  k_write_sentinel<<<1, 1, 0, st>>>(u->q.ptr, u->cap_t);
  k_write_sentinel<<<1, 1, 0, st>>>(u->q_alt.ptr, u->cap_t);
  u->launches += 2;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int migrate_unpack(Unit *d, const double4 *rq, const double4 *rp, const int *rt, long long n, const int *rs) {
  if (n == 0) return MDB200_OK;
  if (d->late_p && !rs) { set_error("late momenta: a migration without upload rows"); return MDB200_EINVAL; }
  // arrivals go to fresh slots past everything in use; the cell sort compacts afterwards
  long long tail = d->scratch_tail;
  MDB_TRY(ensure_capacity(d, tail + n, tail + n));
  if (d->late_p) MDB_TRY(d->psrc.ensure((size_t)(tail + n) + 1, true, d->stream));
  k_mig_unpack<<<nblk(n), TPB, 0, d->stream>>>(d->q.ptr, d->p.ptr, d->type.ptr, d->ord.ptr, rq, rp, rt,
                                               d->late_p ? d->psrc.ptr : nullptr, rs, (int)n, (int)d->pn, (int)tail);
  d->launches++;
  MDB_CUDA(cudaGetLastError());
  d->pn += n;
  d->scratch_tail = tail + n;
  return MDB200_OK;
}

// ---------------------------------------------------------------------------------------------
// periodic wrap (variables.cc:59-79)
// ---------------------------------------------------------------------------------------------
__global__ void k_wrap(double4 *__restrict__ q, const int *__restrict__ ord, int pn, Consts c, Ctrl *ctrl) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= pn) return;
  int h = ord[k];
  double4 v = q[h];
  if (v.x < 0.0) v.x += c.L[0]; else if (v.x > c.L[0]) v.x -= c.L[0];
  if (v.y < 0.0) v.y += c.L[1]; else if (v.y > c.L[1]) v.y -= c.L[1];
  if (v.z < 0.0) v.z += c.L[2]; else if (v.z > c.L[2]) v.z -= c.L[2];
  v.w = v.y;
  q[h] = v;
  bool bad = (v.x < 0.0 || v.x > c.L[0] || v.y < 0.0 || v.y > c.L[1] || v.z < 0.0 || v.z > c.L[2]);
  // NaN positions also count as invalid here (the reference would wander on)
  bad = bad || !(v.x == v.x) || !(v.y == v.y) || !(v.z == v.z);
  if (bad) atomicOr(&ctrl->err, (int)DE_POSITION);
}

int wrap_periodic(Unit *u) {
  if (u->pn == 0) return MDB200_OK;
  k_wrap<<<nblk(u->pn), TPB, 0, u->stream>>>(u->q.ptr, u->ord.ptr, (int)u->pn, u->c, u->ctrl);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

// ---------------------------------------------------------------------------------------------
// counting sort into cells
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double x, double s, double inv, int lo, int hi) {
  double t = floor((x - s) * inv);
  // clamp in floating point first so that wild values cannot overflow the int conversion
  t = fmin(fmax(t, -2.0), (double)(hi + 2));
  int cx = (int)t + 1;
  return min(max(cx, lo), hi);
}

__device__ __forceinline__ int cell_index(const double4 &v, const Consts &c, int halo) {
  // halo = 1: ghosts may use the outer layer [0, m-1]; locals are clamped to [1, ms]
  int lo = halo ? 0 : 1;
  int cx = cell_coord(v.x, c.s[0], c.inv_edge[0], lo, halo ? c.m[0] - 1 : c.ms[0]);
  int cy = cell_coord(v.y, c.s[1], c.inv_edge[1], lo, halo ? c.m[1] - 1 : c.ms[1]);
  int cz = cell_coord(v.z, c.s[2], c.inv_edge[2], lo, halo ? c.m[2] - 1 : c.ms[2]);
  return cx + c.m[0] * (cy + c.m[1] * cz);
}

// locals: item k (reference index) lives at q[ord[k]]
__global__ void k_cell_count_locals(const double4 *__restrict__ q, const int *__restrict__ ord,
                                    int *__restrict__ cell_of, int *__restrict__ cell_count, int pn, Consts c) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= pn) return;
  int ci = cell_index(q[ord[k]], c, 0);
  cell_of[k] = ci;
  atomicAdd(&cell_count[ci], 1);
}
// ghosts: item g in [pn, tn) lives at q[g]
__global__ void k_cell_count_ghosts(const double4 *__restrict__ q, int *__restrict__ cell_of,
                                    int *__restrict__ cell_count, int pn, int tn, Consts c) {
  int g = pn + blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= tn) return;
  int ci = cell_index(q[g], c, 1);
  cell_of[g] = ci;
  atomicAdd(&cell_count[ci], 1);
}
__global__ void k_cell_scatter(const int *__restrict__ cell_of, const int *__restrict__ cell_start,
                               int *__restrict__ cursor, int *__restrict__ items, int first, int last) {
  int k = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= last) return;
  int ci = cell_of[k];
  int pos = cell_start[ci] + atomicAdd(&cursor[ci], 1);
  items[pos] = k;
}
// The scatter order inside a cell depends on atomic arrival order; sort each (small) bucket
// ascending so that the result equals a STABLE counting sort -- the reference's `sortbuf`
// (meshlist.cc:199-204: ascending particle index inside a cell) -- and is run-to-run deterministic.
__global__ void k_cell_sort_buckets(const int *__restrict__ cell_start, int *__restrict__ items, int ncell) {
  int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= ncell) return;
  int b = cell_start[ci], e = cell_start[ci + 1];
  for (int i = b + 1; i < e; i++) {
    int v = items[i];
    int j = i - 1;
    while (j >= b && items[j] > v) {
      items[j + 1] = items[j];
      j--;
    }
    items[j + 1] = v;
  }
}

__global__ void k_permute_locals(const double4 *__restrict__ q, const double4 *__restrict__ p,
                                 const int *__restrict__ type, const int *__restrict__ ord,
                                 const int *__restrict__ items, double4 *__restrict__ q2,
                                 double4 *__restrict__ p2, int *__restrict__ type2,
                                 int *__restrict__ ref_of_slot, int *__restrict__ slot_of_ref,
                                 const int *__restrict__ psrc, int *__restrict__ psrc2, int pn) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= pn) return;
  int k = items[s];
  int h = ord[k];
  double4 qq = q[h];
  qq.w = qq.y;  // lane 3 of a position record repeats y (tile.cuh)
  q2[s] = qq;
  if (psrc) psrc2[s] = psrc[h];  // late momenta: gathered from the upload rows once they have arrived (api.cu)
  else p2[s] = p[h];
  type2[s] = type[h];
  ref_of_slot[s] = k;
  slot_of_ref[k] = s;
}

static int bin_items(Unit *u, DevBuf<int> &count, DevBuf<int> &start, DevBuf<int> &cursor, DevBuf<int> &items,
                     long long first, long long last, bool ghosts) {
  const int nc = u->c.ncell;
  MDB_TRY(count.ensure((size_t)nc + 2));
  MDB_TRY(start.ensure((size_t)nc + 2));
  MDB_TRY(cursor.ensure((size_t)nc + 2));
  MDB_TRY(u->scan_ws.ensure(2 * scan_workspace_ints(nc + 2)));
  MDB_CUDA(cudaMemsetAsync(count.ptr, 0, sizeof(int) * (nc + 1), u->stream));
  MDB_CUDA(cudaMemsetAsync(cursor.ptr, 0, sizeof(int) * (nc + 1), u->stream));
  long long n = last - first;
  if (n > 0) {
    if (ghosts)
      k_cell_count_ghosts<<<nblk(n), TPB, 0, u->stream>>>(u->q.ptr, u->cell_of.ptr, count.ptr, (int)first,
                                                          (int)last, u->c);
    else
      k_cell_count_locals<<<nblk(n), TPB, 0, u->stream>>>(u->q.ptr, u->ord.ptr, u->cell_of.ptr, count.ptr,
                                                          (int)last, u->c);
    u->launches++;
  }
  MDB_TRY(exclusive_scan_i32(count.ptr, start.ptr, nc, u->scan_ws.ptr, u->stream, &u->launches));
  if (n > 0) {
    // ghost items are stored relative to the start of the ghost bucket array
    k_cell_scatter<<<nblk(n), TPB, 0, u->stream>>>(u->cell_of.ptr, start.ptr, cursor.ptr, items.ptr, (int)first,
                                                   (int)last);
    k_cell_sort_buckets<<<nblk(nc), TPB, 0, u->stream>>>(start.ptr, items.ptr, nc);
    u->launches += 2;
  }
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int cell_sort_locals(Unit *u) {
  MDB_TRY(ensure_capacity(u, u->pn, u->pn));
  MDB_TRY(bin_items(u, u->cell_count, u->cell_start, u->cell_cursor, u->cell_items, 0, u->pn, false));
  if (u->pn > 0) {
    k_permute_locals<<<nblk(u->pn), TPB, 0, u->stream>>>(u->q.ptr, u->p.ptr, u->type.ptr, u->ord.ptr,
                                                          u->cell_items.ptr, u->q_alt.ptr, u->p_alt.ptr,
                                                          u->type_alt.ptr, u->ref_of_slot.ptr, u->slot_of_ref.ptr,
                                                          u->late_p ? u->psrc.ptr : nullptr, u->psrc_alt.ptr, (int)u->pn);
    u->launches++;
    MDB_CUDA(cudaGetLastError());
    std::swap(u->q, u->q_alt);
    std::swap(u->p, u->p_alt);
    std::swap(u->type, u->type_alt);
    if (u->late_p) std::swap(u->psrc, u->psrc_alt);
  }
  u->tn = u->pn;
  u->scratch_tail = u->pn;
  return MDB200_OK;
}

int bin_ghosts(Unit *u) {
  return bin_items(u, u->gcell_count, u->gcell_start, u->gcell_cursor, u->gcell_items, u->pn, u->tn, true);
}

// ---------------------------------------------------------------------------------------------
// border (ghost source) lists
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool inside_edge(int dir, const double4 &q, const Consts &c) {
  switch (dir) {  // mdrect.cc:24-47 (strict on both sides)
    case 0: return (q.x < c.s[0] + c.SL && q.x > c.s[0]);
    case 1: return (q.x > c.e[0] - c.SL && q.x < c.e[0]);
    case 2: return (q.y < c.s[1] + c.SL && q.y > c.s[1]);
    case 3: return (q.y > c.e[1] - c.SL && q.y < c.e[1]);
    case 4: return (q.z < c.s[2] + c.SL && q.z > c.s[2]);
    default: return (q.z > c.e[2] - c.SL && q.z < c.e[2]);
  }
}

// scans REFERENCE indices k in [0,tn): the list order must be the reference's (mdunit.cc:129-133)
__global__ void k_border_flag(const double4 *__restrict__ q, const int *__restrict__ slot_of_ref,
                              int *__restrict__ flags, int pn, int tn, int dir, Consts c) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= tn) return;
  int s = (k < pn) ? slot_of_ref[k] : k;
  flags[k] = inside_edge(dir, q[s], c) ? 1 : 0;
}
__global__ void k_border_compact(const int *__restrict__ slot_of_ref, const int *__restrict__ flags,
                                 const int *__restrict__ scanv, int *__restrict__ border, int pn, int tn) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= tn) return;
  if (flags[k]) border[scanv[k]] = (k < pn) ? slot_of_ref[k] : k;
}

int find_border(Unit *u, int dir, long long *n_border) {
  *n_border = 0;
  u->n_border[dir] = 0;
  if (u->tn == 0) return MDB200_OK;
  k_border_flag<<<nblk(u->tn), TPB, 0, u->stream>>>(u->q.ptr, u->slot_of_ref.ptr, u->flags.ptr, (int)u->pn,
                                                     (int)u->tn, dir, u->c);
  u->launches++;
  MDB_TRY(exclusive_scan_i32(u->flags.ptr, u->scan.ptr, u->tn, u->scan_ws.ptr, u->stream, &u->launches));
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[1], u->scan.ptr + u->tn, sizeof(int), cudaMemcpyDeviceToHost,
                           u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  long long nb = u->ctrl_host->scratch[1];
  MDB_TRY(u->border[dir].ensure((size_t)nb + 1));
  if (nb > 0) {
    k_border_compact<<<nblk(u->tn), TPB, 0, u->stream>>>(u->slot_of_ref.ptr, u->flags.ptr, u->scan.ptr,
                                                        u->border[dir].ptr, (int)u->pn, (int)u->tn);
    u->launches++;
  }
  MDB_CUDA(cudaGetLastError());
  u->n_border[dir] = nb;
  *n_border = nb;
  return MDB200_OK;
}

int snapshot_positions(Unit *u) {
  if (u->tn > 0)
    MDB_CUDA(cudaMemcpyAsync(u->q0.ptr, u->q.ptr, sizeof(double4) * (size_t)u->tn, cudaMemcpyDeviceToDevice,
                             u->stream));
  return MDB200_OK;
}

}  // namespace mdb
