// Rebuild path: migration bookkeeping, periodic wrap, GPU cell sort (counting sort of particles
// into cells), ghost (border) lists, and the on-device pair-list build.
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
//   pair search           MeshList::SearchMesh/RegisterPair meshlist.cc:220-271, 288-310
//   CSR bucketing         MeshList::MakeList                meshlist.cc:58-87
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
__global__ void k_begin_order(const int *__restrict__ slot_of_ref, int *__restrict__ ord, int pn) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < pn) ord[r] = slot_of_ref[r];
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
                              int *__restrict__ st, int pn) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= pn) return;
  int h = ord[k];
  int j = scanv[k];
  if (flags[k]) {
    sq[j] = q[h];
    sp[j] = p[h];
    st[j] = type[h];
  } else {
    ord_new[k - j] = h;
  }
}

__global__ void k_mig_unpack(double4 *__restrict__ q, double4 *__restrict__ p, int *__restrict__ type,
                             int *__restrict__ ord, const double4 *__restrict__ rq,
                             const double4 *__restrict__ rp, const int *__restrict__ rt, int n, int pn,
                             int tail) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  q[tail + i] = rq[i];
  p[tail + i] = rp[i];
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
  if (u->pn > 0) {
    k_begin_order<<<nblk(u->pn), TPB, 0, u->stream>>>(u->slot_of_ref.ptr, u->ord.ptr, (int)u->pn);
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
  MDB_TRY(exclusive_scan_i32(u->flags.ptr, u->scan.ptr, u->pn, u->scan_ws.ptr, u->stream, &u->launches));
  k_mig_compact<<<nblk(u->pn), TPB, 0, u->stream>>>(u->q.ptr, u->p.ptr, u->type.ptr, u->ord.ptr, u->flags.ptr,
                                                     u->scan.ptr, u->ord_alt.ptr, u->mig_send_q.ptr,
                                                     u->mig_send_p.ptr, u->mig_send_t.ptr, (int)u->pn);
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
  if (need_p > u->cap_p) { ncp = need_p + need_p / 4 + 1024; grow = true; }
  if (need_t > u->cap_t || ncp > u->cap_t) {
    long long want = need_t > ncp ? need_t : ncp;
    nct = want + want / 4 + 4096;
    grow = true;
  }
  if (!grow) return MDB200_OK;
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

int migrate_unpack(Unit *d, const double4 *rq, const double4 *rp, const int *rt, long long n) {
  if (n == 0) return MDB200_OK;
  // arrivals go to fresh slots past everything in use; the cell sort compacts afterwards
  long long tail = d->scratch_tail;
  MDB_TRY(ensure_capacity(d, tail + n, tail + n));
  k_mig_unpack<<<nblk(n), TPB, 0, d->stream>>>(d->q.ptr, d->p.ptr, d->type.ptr, d->ord.ptr, rq, rp, rt, (int)n,
                                               (int)d->pn, (int)tail);
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
                                 int *__restrict__ ref_of_slot, int *__restrict__ slot_of_ref, int pn) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= pn) return;
  int k = items[s];
  int h = ord[k];
  double4 qq = q[h];
  qq.w = 0.0;
  q2[s] = qq;
  p2[s] = p[h];
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
                                                          (int)u->pn);
    u->launches++;
    MDB_CUDA(cudaGetLastError());
    std::swap(u->q, u->q_alt);
    std::swap(u->p, u->p_alt);
    std::swap(u->type, u->type_alt);
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

// ---------------------------------------------------------------------------------------------
// pair search
// ---------------------------------------------------------------------------------------------
// Listed iff r^2 <= SL^2 with r^2 = (dx*dx + dy*dy) + dz*dz evaluated WITHOUT fused
// multiply-add, exactly as the scalar x86-64 reference does (meshlist.cc:262-266), so the
// pair SET is bit-identical to the reference's.  Ghost-ghost pairs never arise because rows
// exist for locals only (meshlist.cc:261).
//
// Search: one thread per local slot walks the 27 neighbouring cells (9 contiguous (y,z) ranges of
// locals, then the same 9 ranges of binned ghosts).  Two passes (count, then fill) because the list
// layout depends on the per-warp maxima.  Layout of the list: see unit.h (kG lanes per local, kPW
// locals per force-warp).
//
// Order of a row ("rainbow" order, kG == 4).  A position record is 32 bytes, so a partner's index mod 4
// is the sector slot of its 128-byte line = the L1 data-bank group the gather hits.  Measured on B200
// (scripts/ubench/gather.cu): a warp-wide LDG.256 gather costs ~18.6 clk with random slots, ~11 clk when
// every slot is hit by exactly 8 lanes, 32 clk when all lanes hit one slot -- the number of distinct
// lines hardly matters.  So the 4 lanes of a local take, at every step, partners of 4 DIFFERENT classes
// c = j mod 4 for as long as all classes last (steps < m = min class count: lane c takes the t-th
// smallest partner of class c at step t); what is left over is dealt out class-round-robin.  No padding
// is added and within a class the partners stay ascending.
__device__ __forceinline__ bool within_search(const double4 &a, const double4 &b, double SL2) {
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return !(r2 > SL2);
}

// Entry index (step * kG + lane-in-group) of a row's partners; see the header above.
struct RowOrder {
  unsigned cls = 0;       // count pass: partners per class, 4 x 8 bits; fill pass: running ordinals
  unsigned tot = 0;       // fill pass: the class counts of the finished row
  int m = 0, seq = 0;
  bool overflow = false;
  __device__ __forceinline__ void count(int j) {
    if (kG != 4) return;
    const int sh = 8 * (j & 3);
    if (((cls >> sh) & 0xFFu) == 0xFFu) overflow = true;
    else cls += 1u << sh;
  }
  __device__ __forceinline__ void init(unsigned totals) {
    tot = totals;
    cls = 0;
    m = (int)min(min(tot & 0xFFu, (tot >> 8) & 0xFFu), min((tot >> 16) & 0xFFu, (tot >> 24) & 0xFFu));
  }
  __device__ __forceinline__ int place(int j) {
    if (kG != 4) return seq++;
    const int c = j & 3;
    const int t = (int)((cls >> (8 * c)) & 0xFFu);
    cls += 1u << (8 * c);
    if (t < m) return 4 * t + c;
    // leftovers in (ordinal, class) order: how many leftover entries precede (t, c)
    int r = 0;
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
      const int left = (int)((tot >> (8 * cc)) & 0xFFu) - m;
      r += min(left, (t - m) + (cc < c ? 1 : 0));
    }
    return 4 * m + r;
  }
};

template <bool FILL>
__global__ void __launch_bounds__(128)
k_pair_search(const double4 *__restrict__ q, const int *__restrict__ cell_start,
              const int *__restrict__ gcell_start, const int *__restrict__ gcell_items,
              const int *__restrict__ ref_of_slot, int *__restrict__ row_count, unsigned *__restrict__ row_cls,
              const long long *__restrict__ warp_off, int *warp_len, unsigned *__restrict__ list,
              int *__restrict__ warp_flag, int pn, int sentinel, Consts c, Ctrl *ctrl) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;  // local slot
  const int lane = threadIdx.x & 31;
  const int fw = a / kPW;                               // force-warp of this slot
  const int flane0 = (a % kPW) * kG;                    // first lane of its group
  const bool act = a < pn;
  const bool in_tail = (long long)fw * kPW < pn;        // slot belongs to an existing force-warp
  int n = 0, n_local_entries = 0;
  unsigned long long nref = 0;
  long long base = 0;
  int wl = 0;
  RowOrder ro;
  if (FILL && in_tail) {
    base = warp_off[fw];
    wl = warp_len[fw];
  }
  if (act) {
    if (FILL) ro.init(row_cls[a]);
    const double4 qa = q[a];
    const int ra = ref_of_slot[a];
    const int ci = cell_index(qa, c, 0);
    const int cx = ci % c.m[0];
    const int cy = (ci / c.m[0]) % c.m[1];
    const int cz = ci / (c.m[0] * c.m[1]);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, c.m[0] - 1);
    for (int dz = -1; dz <= 1; dz++) {
      const int z = cz + dz;
      if (z < 0 || z >= c.m[2]) continue;
      for (int dy = -1; dy <= 1; dy++) {
        const int y = cy + dy;
        if (y < 0 || y >= c.m[1]) continue;
        const int rowc = c.m[0] * (y + c.m[1] * z);
        // the three x-neighbours are contiguous in cell order: one range per (y,z)
        const int jb = cell_start[rowc + x0], je = cell_start[rowc + x1 + 1];
        for (int j = jb; j < je; j++) {
          if (j == a) continue;
          if (within_search(qa, q[j], c.SL2)) {
            if (FILL) {
              const int e = ro.place(j);
              list[base + list_word(e / kG, flane0 + e % kG)] = (unsigned)j;
            } else {
              nref += (ref_of_slot[j] > ra) ? 1ull : 0ull;
              ro.count(j);
            }
            n++;
          }
        }
      }
    }
    n_local_entries = n;
    // ghosts after all locals, so that a row is ascending in slot
    for (int dz = -1; dz <= 1; dz++) {
      const int z = cz + dz;
      if (z < 0 || z >= c.m[2]) continue;
      for (int dy = -1; dy <= 1; dy++) {
        const int y = cy + dy;
        if (y < 0 || y >= c.m[1]) continue;
        const int rowc = c.m[0] * (y + c.m[1] * z);
        const int gb = gcell_start[rowc + x0], ge = gcell_start[rowc + x1 + 1];
        for (int k = gb; k < ge; k++) {
          const int g = gcell_items[k];
          if (within_search(qa, q[g], c.SL2)) {
            if (FILL) {
              const int e = ro.place(g);
              list[base + list_word(e / kG, flane0 + e % kG)] = (unsigned)g;
            } else {
              nref += 1ull;
              ro.count(g);
            }
            n++;
          }
        }
      }
    }
    if (!FILL) {
      row_count[a] = n;
      row_cls[a] = ro.cls;
      if (ro.overflow) atomicOr(&ctrl->err, (int)DE_CAPACITY);
    }
  }
  if (FILL) {
    if (in_tail)
      for (int e = n; e < wl * kG; e++) list[base + list_word(e / kG, flane0 + e % kG)] = (unsigned)sentinel;
    // boundary warp = one of its rows holds a ghost
    int hg = (n > n_local_entries) ? 1 : 0;
#pragma unroll
    for (int o = kPW / 2; o > 0; o >>= 1) hg |= __shfl_xor_sync(0xffffffffu, hg, o);
    if (in_tail && (a % kPW) == 0) warp_flag[fw] = hg;
  } else {
    // steps of this local, maximum over the kPW locals of its force-warp, rounded up to even
    int mx = (n + kG - 1) / kG;
    int mrow = n;
#pragma unroll
    for (int o = kPW / 2; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mrow = max(mrow, __shfl_xor_sync(0xffffffffu, mrow, o));
      nref += __shfl_xor_sync(0xffffffffu, nref, o);
    }
    if (in_tail && (a % kPW) == 0) warp_len[fw] = (mx + 1) & ~1;
    if (lane == 0) {
      if (nref) atomicAdd(&ctrl->npairs, nref);
      atomicMax(&ctrl->max_row, mrow);
    }
  }
}

// One-pass variant for every rebuild after the first.  Rows get a FIXED capacity of S steps per force-warp
// (S = longest row of the previous build + head-room), so the layout is known up front (warp_off[w] =
// 32 S w) and the distance tests run once: each thread buffers its hits in local memory, then deals them out
// in rainbow order.  The force kernel reads warp_len[w] steps of each row as before, i.e. no extra traffic;
// only the allocation is larger.  A warp that needs more than S steps raises ctrl->scratch[6] and writes no
// entries; its counts are still complete, so the host falls back to scan + fill of the two-pass build (and the
// next rebuild uses a larger S).
#define ONEPASS_HCAP 192
__global__ void __launch_bounds__(128)
k_pair_search_onepass(const double4 *__restrict__ q, const int *__restrict__ cell_start,
                      const int *__restrict__ gcell_start, const int *__restrict__ gcell_items,
                      int *__restrict__ row_count, unsigned *__restrict__ row_cls, long long *__restrict__ warp_off,
                      int *__restrict__ warp_len, unsigned *__restrict__ list, int *__restrict__ warp_flag, int pn,
                      int sentinel, int S, Consts c, Ctrl *ctrl) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;  // local slot
  const int lane = threadIdx.x & 31;
  const int fw = a / kPW;                               // force-warp of this slot
  const int flane0 = (a % kPW) * kG;                    // first lane of its group
  const bool act = a < pn;
  const bool in_tail = (long long)fw * kPW < pn;        // slot belongs to an existing force-warp
  const int cap = min(ONEPASS_HCAP, S * kG);
  unsigned hits[ONEPASS_HCAP];
  int n = 0, n_local_entries = 0;
  unsigned cls = 0;
  if (act) {
    const double4 qa = q[a];
    const int ci = cell_index(qa, c, 0);
    const int cx = ci % c.m[0];
    const int cy = (ci / c.m[0]) % c.m[1];
    const int cz = ci / (c.m[0] * c.m[1]);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, c.m[0] - 1);
#pragma unroll 1
    for (int ghosts = 0; ghosts < 2; ghosts++) {  // ghosts after all locals
      const int *__restrict__ start = ghosts ? gcell_start : cell_start;
      for (int dz = -1; dz <= 1; dz++) {
        const int z = cz + dz;
        if (z < 0 || z >= c.m[2]) continue;
        for (int dy = -1; dy <= 1; dy++) {
          const int y = cy + dy;
          if (y < 0 || y >= c.m[1]) continue;
          const int rowc = c.m[0] * (y + c.m[1] * z);
          // the three x-neighbours are contiguous in cell order: one range per (y,z)
          const int jb = start[rowc + x0], je = start[rowc + x1 + 1];
          // Chunks of 32 candidates: the test loop is branch-free (hits go into a bit mask) and takes four
          // candidates per trip so that their loads are in flight together -- with one load per test followed by
          // a branch, every test waited for its own load (ncu: issue 41 %, the stall samples on the first DADD).
          // Indices past the range repeat its last candidate and are masked off.
#pragma unroll 1
          for (int j0 = jb; j0 < je; j0 += 32) {
            const int c32 = min(32, je - j0);
            unsigned m = 0u;
#pragma unroll 1
            for (int b = 0; b < c32; b += 4) {
              int jj[4];
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const int idx = min(j0 + b + i, je - 1);
                jj[i] = ghosts ? gcell_items[idx] : idx;
              }
              double4 qj[4];
#pragma unroll
              for (int i = 0; i < 4; i++) qj[i] = q[jj[i]];
#pragma unroll
              for (int i = 0; i < 4; i++) m |= (within_search(qa, qj[i], c.SL2) ? 1u : 0u) << (b + i);
            }
            m &= (c32 == 32) ? 0xffffffffu : ((1u << c32) - 1u);
            while (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1u;
              const int j = ghosts ? gcell_items[j0 + b] : j0 + b;
              if (j == a) continue;
              if (n < cap) hits[n] = (unsigned)j;
              cls += 1u << (8 * (j & 3));
              n++;
            }
          }
        }
      }
      if (!ghosts) n_local_entries = n;
    }
  }
  // steps of this local, maximum over the kPW locals of its force-warp, rounded up to even
  int wl = (n + kG - 1) / kG;
  int over = (n > cap) ? 1 : 0;
  int hg = (n > n_local_entries) ? 1 : 0;
#pragma unroll
  for (int o = kPW / 2; o > 0; o >>= 1) {
    wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
    over |= __shfl_xor_sync(0xffffffffu, over, o);
    hg |= __shfl_xor_sync(0xffffffffu, hg, o);
  }
  wl = (wl + 1) & ~1;
  if (wl > S) over = 1;
  int mrow = n;
  // reference-style pair count in half-pairs (see k_pair_search)
  unsigned long long nref = (unsigned long long)(n_local_entries + 2 * (n - n_local_entries));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mrow = max(mrow, __shfl_xor_sync(0xffffffffu, mrow, o));
    nref += __shfl_xor_sync(0xffffffffu, nref, o);
  }
  if (lane == 0) {
    if (nref) atomicAdd(&ctrl->npairs, nref);
    if (mrow > *(volatile int *)&ctrl->max_row) atomicMax(&ctrl->max_row, mrow);
  }
  // counts are complete even when a row overflowed its buffer: the fallback only needs scan + fill
  if (in_tail && (a % kPW) == 0) {
    warp_len[fw] = wl;
    warp_flag[fw] = hg;  // boundary warp = one of its rows holds a ghost
  }
  if (act) {
    row_count[a] = n;
    row_cls[a] = cls;
    if (((cls & 0xFFu) + ((cls >> 8) & 0xFFu) + ((cls >> 16) & 0xFFu) + (cls >> 24)) != (unsigned)n)
      atomicOr(&ctrl->err, (int)DE_CAPACITY);  // a class count wrapped its 8 bits
  }
  if (over) {
    if ((a % kPW) == 0) ctrl->scratch[6] = 1;
    return;
  }
  const long long base = (long long)fw * S * 32;
  if (in_tail && (a % kPW) == 0) warp_off[fw] = base;
  if (act) {
    RowOrder ro;
    ro.init(cls);
    for (int k = 0; k < n; k++) {
      const unsigned j = hits[k];
      const int e = ro.place((int)j);
      list[base + list_word(e / kG, flane0 + e % kG)] = j;
    }
  }
  if (in_tail)
    for (int e = n; e < wl * kG; e++) list[base + list_word(e / kG, flane0 + e % kG)] = (unsigned)sentinel;
}

__global__ void k_warp_order(const int *__restrict__ flag, const int *__restrict__ scanv, int *__restrict__ order,
                             int nw, int n_interior) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const int nb = scanv[w];  // boundary warps before w
  if (flag[w]) order[n_interior + nb] = w;
  else order[w - nb] = w;
}

int build_pair_list(Unit *u) {
  const long long pn = u->pn;
  MDB_CUDA(cudaMemsetAsync(&u->ctrl->npairs, 0, sizeof(unsigned long long), u->stream));
  MDB_CUDA(cudaMemsetAsync(&u->ctrl->max_row, 0, sizeof(int), u->stream));
  u->list_entries = 0;
  u->npairs = 0;
  u->nwarps = 0;
  if (pn == 0) return MDB200_OK;
  const long long nw = (pn + kPW - 1) / kPW;
  u->nwarps = nw;
  MDB_TRY(u->row_count.ensure((size_t)pn + 32));
  MDB_TRY(u->row_cls.ensure((size_t)pn + 32));
  MDB_TRY(u->warp_len.ensure((size_t)nw + 2));
  MDB_TRY(u->warp_off.ensure((size_t)nw + 2));
  MDB_TRY(u->warp_flag.ensure((size_t)nw + 2));
  MDB_TRY(u->warp_order.ensure((size_t)nw + 2));
  MDB_TRY(u->scan.ensure((size_t)nw + 2));
  MDB_TRY(u->scan_ws.ensure(2 * scan_workspace_ints(nw + 2)));
  MDB_TRY(u->scan_ws64.ensure(scan_workspace_ints(nw + 2)));
  const int tpb = 128;
  const long long nthr = nw * kPW;  // threads past pn only pad the tail warp
  bool built = false, counted = false;
  // ---- one pass with fixed-capacity rows, sized from the previous build (every rebuild but the first)
  if (u->max_row > 0 && !getenv("MDB200_NO_ONEPASS")) {
    const int S = (((u->max_row + kG - 1) / kG + 3) + 1) & ~1;  // longest row so far + 3 steps, even
    if (S * kG <= ONEPASS_HCAP) {
      const long long entries = nw * (long long)S * 32;
      if ((size_t)entries + 64 > u->list.cap) {
        if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] list grows: %lld entries > capacity %zu\n", entries, u->list.cap);
        MDB_TRY(u->list.ensure((size_t)entries + (size_t)entries / 8 + 64));
      }
      MDB_CUDA(cudaMemsetAsync(&u->ctrl->scratch[6], 0, sizeof(int), u->stream));
      k_pair_search_onepass<<<nblk(nthr, tpb), tpb, 0, u->stream>>>(
          u->q.ptr, u->cell_start.ptr, u->gcell_start.ptr, u->gcell_items.ptr, u->row_count.ptr, u->row_cls.ptr,
          u->warp_off.ptr, u->warp_len.ptr, u->list.ptr, u->warp_flag.ptr, (int)pn, sentinel_index(u), S, u->c, u->ctrl);
      u->launches++;
      MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[6], &u->ctrl->scratch[6], sizeof(int), cudaMemcpyDeviceToHost, u->stream));
      MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->npairs, &u->ctrl->npairs, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                               u->stream));
      MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->max_row, &u->ctrl->max_row, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
      MDB_CUDA(cudaStreamSynchronize(u->stream));
      u->max_row = std::max(u->max_row, u->ctrl_host->max_row);  // also after an overflow: the next S fits
      if (u->ctrl_host->scratch[6] == 0) {
        built = true;
        u->list_entries = entries;
        u->npairs = (long long)(u->ctrl_host->npairs / 2);
      } else {
        if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] one-pass list build overflowed S = %d steps; scan + fill\n", S);
        counted = true;  // row_count / row_cls / warp_len / npairs / max_row are complete
        u->npairs = (long long)(u->ctrl_host->npairs / 2);
      }
    }
  }
  // ---- two passes (count, scan, fill): packed rows
  if (!built) {
    if (!counted) {
      k_pair_search<false><<<nblk(nthr, tpb), tpb, 0, u->stream>>>(
          u->q.ptr, u->cell_start.ptr, u->gcell_start.ptr, u->gcell_items.ptr, u->ref_of_slot.ptr, u->row_count.ptr, u->row_cls.ptr,
          u->warp_off.ptr, u->warp_len.ptr, u->list.ptr, u->warp_flag.ptr, (int)pn, sentinel_index(u), u->c, u->ctrl);
      u->launches++;
    }
    MDB_TRY(exclusive_scan_i32_to_i64(u->warp_len.ptr, u->warp_off.ptr, nw, u->scan_ws64.ptr, u->stream, &u->launches, 32));
    MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[4], u->warp_off.ptr + nw, sizeof(long long), cudaMemcpyDeviceToHost,
                             u->stream));
    MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->npairs, &u->ctrl->npairs, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             u->stream));
    MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->max_row, &u->ctrl->max_row, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
    MDB_CUDA(cudaStreamSynchronize(u->stream));
    long long entries;
    memcpy(&entries, &u->ctrl_host->scratch[4], sizeof(long long));
    u->list_entries = entries;
    if (!counted) u->npairs = (long long)u->ctrl_host->npairs;
    u->max_row = std::max(u->max_row, u->ctrl_host->max_row);
    if ((size_t)entries + 64 > u->list.cap && getenv("MDB200_DEBUG"))
      fprintf(stderr, "[mdb200] list grows: %lld entries > capacity %zu\n", entries, u->list.cap);
    // Head-room: rows of a freshly generated lattice grow by ~20 % once it thermalises (the next FCC shell
    // sits just outside the search radius), the fixed-capacity rows of the one-pass rebuilds need more still,
    // and re-allocating a multi-GB list inside the step loop stalls for hundreds of ms (cudaMalloc + cudaFree
    // of several GB) -- so allocate 70 % more than needed.
    if ((size_t)entries + 64 > u->list.cap) MDB_TRY(u->list.ensure((size_t)entries + (size_t)entries * 7 / 10 + 64));
    k_pair_search<true><<<nblk(nthr, tpb), tpb, 0, u->stream>>>(
        u->q.ptr, u->cell_start.ptr, u->gcell_start.ptr, u->gcell_items.ptr, u->ref_of_slot.ptr, u->row_count.ptr, u->row_cls.ptr,
        u->warp_off.ptr, u->warp_len.ptr, u->list.ptr, u->warp_flag.ptr, (int)pn, sentinel_index(u), u->c, u->ctrl);
    u->launches++;
  }
  // interior warps first, boundary warps last (stable within each class)
  MDB_TRY(exclusive_scan_i32(u->warp_flag.ptr, u->scan.ptr, nw, u->scan_ws.ptr, u->stream, &u->launches));
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[5], u->scan.ptr + nw, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  u->n_interior = nw - u->ctrl_host->scratch[5];
  k_warp_order<<<nblk(nw), TPB, 0, u->stream>>>(u->warp_flag.ptr, u->scan.ptr, u->warp_order.ptr, (int)nw,
                                                (int)u->n_interior);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int snapshot_positions(Unit *u) {
  if (u->tn > 0)
    MDB_CUDA(cudaMemcpyAsync(u->q0.ptr, u->q.ptr, sizeof(double4) * (size_t)u->tn, cudaMemcpyDeviceToDevice,
                             u->stream));
  return MDB200_OK;
}

// ---------------------------------------------------------------------------------------------
// parity helper: the list as the reference's raw (key, partner) arrays, reference indices
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_pair(int ri, int rj, int *key, int *partner, unsigned long long *cursor,
                                          unsigned long long cap) {
  unsigned long long pos = atomicAdd(cursor, 1ull);
  if (pos < cap) {
    key[pos] = ri < rj ? ri : rj;
    partner[pos] = ri < rj ? rj : ri;
  }
}
__global__ void k_extract_pairs(const unsigned *__restrict__ list, const long long *__restrict__ warp_off,
                                const int *__restrict__ row_count, const int *__restrict__ ref_of_slot,
                                int *__restrict__ key, int *__restrict__ partner, unsigned long long *cursor,
                                unsigned long long cap, int pn) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= pn) return;
  const int ra = ref_of_slot[a];
  const long long base = warp_off[a / kPW];
  const int flane0 = (a % kPW) * kG;
  const int n = row_count[a];
  for (int e = 0; e < n; e++) {
    const int j = (int)list[base + list_word(e / kG, flane0 + e % kG)];
    const int rj = (j < pn) ? ref_of_slot[j] : j;
    if (ra < rj) emit_pair(ra, rj, key, partner, cursor, cap);
  }
}

int extract_pairs(Unit *u, int *key, int *partner, long long cap, long long *npairs) {
  *npairs = u->npairs;
  if (!u->list_initialized) { *npairs = 0; return MDB200_OK; }
  if (cap < u->npairs) return MDB200_OK;  // caller only wanted the count
  if (u->npairs == 0) return MDB200_OK;
  DevBuf<int> dk, dp;
  MDB_TRY(dk.ensure((size_t)u->npairs));
  MDB_TRY(dp.ensure((size_t)u->npairs));
  unsigned long long *cur = nullptr;
  MDB_CUDA(cudaMalloc(&cur, sizeof(unsigned long long)));
  MDB_CUDA(cudaMemsetAsync(cur, 0, sizeof(unsigned long long), u->stream));
  k_extract_pairs<<<nblk(u->pn), TPB, 0, u->stream>>>(u->list.ptr, u->warp_off.ptr, u->row_count.ptr,
                                                       u->ref_of_slot.ptr, dk.ptr, dp.ptr, cur,
                                                       (unsigned long long)u->npairs, (int)u->pn);
  u->launches++;
  unsigned long long got = 0;
  MDB_CUDA(cudaMemcpyAsync(&got, cur, sizeof(got), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaMemcpyAsync(key, dk.ptr, sizeof(int) * (size_t)u->npairs, cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaMemcpyAsync(partner, dp.ptr, sizeof(int) * (size_t)u->npairs, cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  cudaFree(cur);
  dk.release();
  dp.release();
  if ((long long)got != u->npairs) {
    set_error("pair extraction count mismatch: %llu vs %lld", got, u->npairs);
    return MDB200_EINVAL;
  }
  return MDB200_OK;
}

}  // namespace mdb
