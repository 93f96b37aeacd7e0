// Tiles: the unit of work of the search, force and observe kernels (device code shared by tile_search.cu,
// force.cu and observe.cu).
//
// The reference walks its CSR list with one gather of q[j] per listed pair (fcalculator.cc:182-223) and builds
// the list by testing every particle of the 27 surrounding cells (meshlist.cc:220-271).  On the B200 both loops
// are bound by the L1 data pipe when the partner positions are gathered from global memory (round 1: 20
// wavefronts per warp-wide gather, 63 % sector hits, profiles/r01c_force_ncu.txt).  So a thread block here
// owns a TILE of up to 3 x 3 x 3 cells, copies the positions of those cells and of the one-cell halo around them
// into shared memory once (coalesced, ~3,150 records of 32 bytes at density 0.7), and every partner is then
// addressed by a 16-bit index into that STAGED array:
//   * a gather is two bank-conflict-free shared-memory loads (8 wavefronts per warp-gather, no tags, no misses);
//   * the list stream halves (2 bytes per entry);
//   * the search reads its candidates as shared-memory broadcasts.
//
// Staged order of a tile: the locals of its interior cells (cell order z, y, x; inside a cell the slot order),
// then the locals of the halo cells, then the ghosts binned into the halo cells, then one far-away SENTINEL
// record that list padding points at.  It follows from cell_start / gcell_start / gcell_items alone, which only
// change at a rebuild, so a staged index means the same particle in every kernel until then.
#pragma once
#include "unit.h"

namespace mdb {

struct TileGrid {
  int t[3];     // cells per tile edge (1..3)
  int nt[3];    // tiles per axis
  int ntiles;
};

constexpr int kTileMaxEdge = 3;
constexpr int kTileMaxCells = (kTileMaxEdge + 2) * (kTileMaxEdge + 2) * (kTileMaxEdge + 2);
constexpr int kTileMaxInterior = kTileMaxEdge * kTileMaxEdge * kTileMaxEdge;
constexpr int kTileMaxRows = 144;    // force-warps of a tile (8 interior locals each): tile_search.cu checks the census

struct TileCell {                 // one cell of the tile's halo box (16 bytes: the table shares the SM's shared
  int src_l, src_g;               // memory with the staged tile) -- first slot of the locals / first entry of gcell_items
  unsigned short n_l, n_g;        // counts
  unsigned short off_l, off_g;    // first staged index of the locals / ghosts
};

struct TileShared {          // per-block bookkeeping in shared memory
  int lo[3];                 // first cell of the halo box (cell coordinates of the unit's grid)
  int n[3];                  // cells of the halo box per axis (interior cells + 2)
  int ncells;
  int n_int;                 // staged interior locals: [0, n_int)
  int n_loc;                 // staged locals:          [0, n_loc)
  int n_all;                 // staged records:         [0, n_all); the sentinel sits at n_all
  int nicell;                // interior cells
  int next_row;              // work counter (dynamic row hand-out)
  int rows_done;             // persistent force kernel: rows of the tile finished so far (the buffer is free at all of them)
  TileCell cell[kTileMaxCells];
  int icell_off[kTileMaxInterior + 1];  // staged offset of interior cell i (ascending)
  int icell_src[kTileMaxInterior];      // its first slot
  int icell_k[kTileMaxInterior];        // its index in cell[]
  unsigned char interior[kTileMaxCells + 3];  // 1 = one of the tile's own cells
  // row blocks of the tile's force-warps in the list being walked (filled by tile_stage_rows): first uint2 word
  // relative to the tile's first row block, and steps
  long long row_base;
  int gfw0;                  // first force-warp (row) of the tile in the unit's row arrays
  unsigned long long mbar;              // mbarrier of the bulk copies that stage the tile (tile_stage_tma)
  int row_off[kTileMaxRows];
  unsigned short row_len[kTileMaxRows];
  unsigned char row_cell[kTileMaxRows];   // interior cell (index into icell_*) of the row's first local (persistent force kernel)
};

// Shared memory of a kernel that stages FP64 records: bookkeeping + (records + sentinel) x 32 bytes
__host__ __device__ inline size_t tile_shared_offset() { return (sizeof(TileShared) + 127) & ~(size_t)127; }
inline size_t tile_smem_bytes_f64(int cap_records) { return tile_shared_offset() + (size_t)(cap_records + 1) * 32; }

// ---- mbarrier / bulk-copy (TMA) primitives ------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (16-byte aligned, a multiple of 16 bytes); completion is counted on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

// k / n for the small operands of a halo box (k < 128, n <= 25) without the integer-division sequence:
// floor(k M / 65536) = floor(k / n) for every M in [ceil(65536 / n), ceil(65536 / n) + 1] while 2 k n < 65536,
// and floor(65536.f / n) + 2 lies in that range whatever the last bit of the single-precision quotient is
__device__ __forceinline__ int small_magic(int n) { return (int)__fdividef(65536.0f, (float)n) + 2; }
__device__ __forceinline__ int small_div(int k, int magic) { return (k * magic) >> 16; }

// second half of the set-up, one warp: exclusive prefix sums per category (interior locals / halo locals / ghosts) in
// cell order over the cell table the first half filled
__device__ __forceinline__ void tile_setup_scan(TileShared &ts, int tid, int ncells) {
  int base_a = 0, base_b = 0, base_c = 0, base_i = 0;
  for (int k0 = 0; k0 < ncells; k0 += 32) {
    const int k = k0 + tid;
    const bool valid = k < ncells;
    const bool inter = valid && ts.interior[k];
    const int nl = valid ? ts.cell[k].n_l : 0, ng = valid ? ts.cell[k].n_g : 0;
    int a = inter ? nl : 0, b = inter ? 0 : nl, cg = ng, fi = inter ? 1 : 0;
    int sa = a, sb = b, sc = cg, si = fi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, sa, o), tb = __shfl_up_sync(0xffffffffu, sb, o);
      const int tc = __shfl_up_sync(0xffffffffu, sc, o), ti = __shfl_up_sync(0xffffffffu, si, o);
      if (tid >= o) { sa += ta; sb += tb; sc += tc; si += ti; }
    }
    if (valid) {
      ts.cell[k].off_l = (unsigned short)(inter ? base_a + sa - a : base_b + sb - b);  // halo locals: n_int added below
      ts.cell[k].off_g = (unsigned short)(base_c + sc - cg);                            // n_loc added below
      if (inter) {
        const int i = base_i + si - 1;
        ts.icell_off[i] = base_a + sa - a;
        ts.icell_src[i] = ts.cell[k].src_l;
        ts.icell_k[i] = k;
      }
    }
    base_a += __shfl_sync(0xffffffffu, sa, 31);
    base_b += __shfl_sync(0xffffffffu, sb, 31);
    base_c += __shfl_sync(0xffffffffu, sc, 31);
    base_i += __shfl_sync(0xffffffffu, si, 31);
  }
  __syncwarp();
  for (int k = tid; k < ncells; k += 32) {
    if (!ts.interior[k]) ts.cell[k].off_l = (unsigned short)(ts.cell[k].off_l + base_a);
    ts.cell[k].off_g = (unsigned short)(ts.cell[k].off_g + base_a + base_b);
  }
  if (tid == 0) {
    ts.n_int = base_a;
    ts.n_loc = base_a + base_b;
    ts.n_all = base_a + base_b + base_c;
    ts.nicell = base_i;
    ts.icell_off[base_i] = base_a;
  }
}

// what the tile set-up needs of Consts (the persistent force kernel keeps a copy in shared memory)
struct TileDims {
  int m[3];    // cells per axis including the ghost layers
  int ms[3];   // the unit's own cells per axis
};

// halo box of tile `tile`: first cell and cells per axis
template <class C>
__device__ __forceinline__ void tile_box(int tile, const TileGrid &g, const C &c, int (&lo)[3], int (&n)[3]) {
  int T[3];
  T[2] = tile / (g.nt[0] * g.nt[1]);
  const int rem = tile - T[2] * (g.nt[0] * g.nt[1]);
  T[1] = rem / g.nt[0];
  T[0] = rem - T[1] * g.nt[0];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int first = 1 + T[d] * g.t[d];                  // first interior cell (cells 1..ms are the unit's own)
    const int last = min(first + g.t[d], c.ms[d] + 1);    // one past the last
    lo[d] = first - 1;
    n[d] = last - first + 2;
  }
}

// The set-up by ONE warp (persistent force kernel: warp 0 sets the next tile up while the others walk): fills ts like
// tile_setup except for the mbarrier, which that kernel initialises once.  Ends with __syncwarp.
template <class C>
__device__ __forceinline__ void tile_setup_warp(TileShared &ts, int tile, const TileGrid &g, const C &c,
                                                const int *__restrict__ cell_start, const int *__restrict__ gcell_start) {
  const int lane = threadIdx.x & 31;
  int lo[3], n[3];
  tile_box(tile, g, c, lo, n);
  const int n01 = n[0] * n[1];
  const int ncells = n01 * n[2];
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { ts.lo[d] = lo[d]; ts.n[d] = n[d]; }
    ts.ncells = ncells;
  }
  constexpr int kTrips = (kTileMaxCells + 31) / 32;
  const int m01 = small_magic(n01), m0 = small_magic(n[0]);
  int l0[kTrips], l1[kTrips], g0[kTrips], g1[kTrips];
  // all the loads first: four dependent round trips to L2 otherwise
#pragma unroll
  for (int u = 0; u < kTrips; u++) {
    const int k = lane + 32 * u;
    if (k < ncells) {
      const int kz = small_div(k, m01), r = k - kz * n01, ky = small_div(r, m0), kx = r - ky * n[0];
      const int gc = (lo[0] + kx) + c.m[0] * ((lo[1] + ky) + c.m[1] * (lo[2] + kz));
      l0[u] = __ldg(cell_start + gc); l1[u] = __ldg(cell_start + gc + 1);
      g0[u] = __ldg(gcell_start + gc); g1[u] = __ldg(gcell_start + gc + 1);
    }
  }
#pragma unroll
  for (int u = 0; u < kTrips; u++) {
    const int k = lane + 32 * u;
    if (k < ncells) {
      const int kz = small_div(k, m01), r = k - kz * n01, ky = small_div(r, m0), kx = r - ky * n[0];
      TileCell tc;
      tc.src_l = l0[u];
      tc.n_l = (unsigned short)min(l1[u] - l0[u], 65535);
      tc.src_g = g0[u];
      tc.n_g = (unsigned short)min(g1[u] - g0[u], 65535);
      tc.off_l = 0;
      tc.off_g = 0;
      ts.interior[k] = (kx >= 1 && kx <= n[0] - 2 && ky >= 1 && ky <= n[1] - 2 && kz >= 1 && kz <= n[2] - 2) ? 1 : 0;
      ts.cell[k] = tc;
    }
  }
  __syncwarp();
  tile_setup_scan(ts, lane, ncells);
  __syncwarp();
}

// Fills ts for tile `tile`.  Called by every thread of the block (blockDim.x >= 128); ends with a barrier.
__device__ __forceinline__ void tile_setup(TileShared &ts, int tile, const TileGrid &g, const Consts &c,
                                           const int *__restrict__ cell_start, const int *__restrict__ gcell_start) {
  const int tid = threadIdx.x;
  int lo[3], n[3];
  tile_box(tile, g, c, lo, n);
  const int n01 = n[0] * n[1];
  const int ncells = n01 * n[2];
  if (tid == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { ts.lo[d] = lo[d]; ts.n[d] = n[d]; }
    ts.ncells = ncells;
    ts.next_row = 0;
    mbar_init(&ts.mbar, 1);  // made visible to the async proxy by the fence + barriers below
  }
  if (tid < ncells) {
    const int kz = small_div(tid, small_magic(n01)), r = tid - kz * n01, ky = small_div(r, small_magic(n[0])), kx = r - ky * n[0];
    const int gc = (lo[0] + kx) + c.m[0] * ((lo[1] + ky) + c.m[1] * (lo[2] + kz));
    TileCell tc;
    tc.src_l = cell_start[gc];
    tc.n_l = (unsigned short)min(cell_start[gc + 1] - tc.src_l, 65535);   // the census bounds a tile to 6,400 records
    tc.src_g = gcell_start[gc];
    tc.n_g = (unsigned short)min(gcell_start[gc + 1] - tc.src_g, 65535);
    tc.off_l = 0;
    tc.off_g = 0;
    ts.interior[tid] = (kx >= 1 && kx <= n[0] - 2 && ky >= 1 && ky <= n[1] - 2 && kz >= 1 && kz <= n[2] - 2) ? 1 : 0;
    ts.cell[tid] = tc;
  }
  __syncthreads();
  if (tid < 32) tile_setup_scan(ts, tid, ncells);
  __syncthreads();
}

// the same starting from the interior cell i0 of an earlier local (a row's first): a row of 8 locals spans few cells
__device__ __forceinline__ int tile_interior_slot_from(const TileShared &ts, int a, int i0) {
  int i = i0;
  while (i + 1 < ts.nicell && ts.icell_off[i + 1] <= a) i++;
  return ts.icell_src[i] + (a - ts.icell_off[i]);
}
// interior cell of the interior local with staged index a
__device__ __forceinline__ int tile_interior_cell(const TileShared &ts, int a) {
  int i = 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1)
    if (i + s < ts.nicell && ts.icell_off[i + s] <= a) i += s;
  return i;
}
// slot of the interior local with staged index a (binary search over <= 27 interior cells)
__device__ __forceinline__ int tile_interior_slot(const TileShared &ts, int a) {
  int i = 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1)
    if (i + s < ts.nicell && ts.icell_off[i + s] <= a) i += s;
  return ts.icell_src[i] + (a - ts.icell_off[i]);
}

// 256-bit global accesses (sm_100): a double4 record is one instruction and, for the 8 consecutive records a force-warp's
// writer lanes touch, one pass over each 128-byte line instead of two
__device__ __forceinline__ double4 ld_256(const double4 *p) {
  double4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_256(double4 *p, const double4 &v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ double4 ld_q256(const double4 *p) {
  double4 v;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

// Copies the tile's records into sq as (x, y, z, y): the duplicate y lets odd lane groups read (z, y) first and
// x second while even groups read (x, y) and z -- see force.cu.  The sentinel goes to sq[n_all].  sslot (may be
// null): the slot of every staged record.  No barrier at the end.
constexpr int kStageBatch = 2;
__device__ __forceinline__ void tile_stage_f64(const TileShared &ts, const double4 *__restrict__ q,
                                               const int *__restrict__ gitems, double4 *sq, int *sslot) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ncells = ts.ncells;
  // two cells per trip, their loads issued together: a thread copies only ~6 records, so the copy is pure latency
  for (int k0 = wid; k0 < ncells; k0 += kStageBatch * nw) {
    int src[kStageBatch], off[kStageBatch], nl[kStageBatch];
    double4 v[kStageBatch];
#pragma unroll
    for (int u = 0; u < kStageBatch; u++) {
      const int k = k0 + u * nw;
      nl[u] = 0;
      if (k < ncells) { src[u] = ts.cell[k].src_l; off[u] = ts.cell[k].off_l; nl[u] = ts.cell[k].n_l; }
      if (lane < nl[u]) v[u] = ld_q256(q + src[u] + lane);
    }
#pragma unroll
    for (int u = 0; u < kStageBatch; u++) {
      if (lane < nl[u]) {
        v[u].w = v[u].y;
        sq[off[u] + lane] = v[u];
        if (sslot) sslot[off[u] + lane] = src[u] + lane;
      }
    }
    // the rest: cells with more than 32 locals, and the ghosts binned into the (outermost) cells
#pragma unroll
    for (int u = 0; u < kStageBatch; u++) {
      const int k = k0 + u * nw;
      if (k >= ncells) continue;
      for (int l = lane + 32; l < nl[u]; l += 32) {
        double4 w = ld_q256(q + src[u] + l);
        w.w = w.y;
        sq[off[u] + l] = w;
        if (sslot) sslot[off[u] + l] = src[u] + l;
      }
      const int ng = ts.cell[k].n_g;
      if (ng > 0) {
        const int sg = ts.cell[k].src_g, og = ts.cell[k].off_g;
        for (int l = lane; l < ng; l += 32) {
          const int s = gitems[sg + l];
          double4 w = ld_q256(q + s);
          w.w = w.y;
          sq[og + l] = w;
          if (sslot) sslot[og + l] = s;
        }
      }
    }
  }
  if (threadIdx.x == 0) sq[ts.n_all] = make_double4(1.0e8, 1.0e8, 1.0e8, 1.0e8);
}

// The same staging without the load/store units: position records carry (x, y, z, y) in global memory already
// (every kernel that writes a position keeps lane 3 = y), so the locals of a cell are ONE contiguous bulk copy
// (TMA, cp.async.bulk) into their place in the staged array; the lanes of warp 0 issue one copy per cell, the
// mbarrier in ts counts the bytes.  Ghosts (binned by index, boundary tiles only) are gathered by the other warps
// as before.  Every thread must call tile_stage_wait before it reads sq; no barrier here.
__device__ __forceinline__ void tile_stage_tma(TileShared &ts, const double4 *__restrict__ q, const int *__restrict__ gitems,
                                               double4 *sq) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ncells = ts.ncells;
  if (wid == 0) {
    if (lane == 0) mbar_arrive_expect_tx(&ts.mbar, (unsigned)ts.n_loc * 32u);
    __syncwarp();
    for (int k = lane; k < ncells; k += 32) {
      const int nl = ts.cell[k].n_l;
      if (nl > 0) bulk_copy_g2s(sq + ts.cell[k].off_l, q + ts.cell[k].src_l, (unsigned)nl * 32u, &ts.mbar);
    }
    if (lane == 0) sq[ts.n_all] = make_double4(1.0e8, 1.0e8, 1.0e8, 1.0e8);
  } else if (ts.n_all > ts.n_loc) {
    for (int k = wid - 1; k < ncells; k += nw - 1) {
      const int ng = ts.cell[k].n_g;
      if (ng > 0) {
        const int sg = ts.cell[k].src_g, og = ts.cell[k].off_g;
        for (int l = lane; l < ng; l += 32) sq[og + l] = ld_q256(q + gitems[sg + l]);
      }
    }
  }
}
// completes tile_stage_tma: the bulk copies have landed and every thread's own stores are visible
// (one warp polls the mbarrier, the others sleep at the block barrier: sixteen polling warps were a fifth of all
// instructions the force kernel issued)
__device__ __forceinline__ void tile_stage_wait(TileShared &ts) {
  if (threadIdx.x < 32) mbar_wait(&ts.mbar, 0u);
  __syncthreads();
}

// Row table of the tile in shared memory: the walk then starts a row without waiting for two dependent global loads.
// No barrier at the end.
__device__ __forceinline__ void tile_stage_rows(TileShared &ts, int gfw0, const long long *__restrict__ fw_off,
                                                const int *__restrict__ fw_len) {
  const int nfw = (ts.n_int + kPW - 1) / kPW;
  const long long base = fw_off[gfw0];
  if (threadIdx.x == 0) ts.row_base = base;
  for (int r = threadIdx.x; r < nfw; r += blockDim.x) {
    ts.row_off[r] = (int)(fw_off[gfw0 + r] - base);
    ts.row_len[r] = (unsigned short)fw_len[gfw0 + r];
  }
}

// ---------------------------------------------------------------------------------------------------------
// list layout (unit.h): entry e of the local in group m of a force-warp: lane 4 m + (e & 3), step e >> 2; the
// four steps 4 c .. 4 c + 3 of a lane are the four 16-bit fields of uint2 word (c, lane) at fw_off + 32 c + lane.
// ---------------------------------------------------------------------------------------------------------
// byte offset of entry e inside the local's 32-byte sector of chunk e >> 4
__host__ __device__ inline int sector_byte(int e) { return ((e & 3) << 3) + (((e >> 2) & 3) << 1); }
// storage steps of a row block walked for `steps` steps
__host__ __device__ inline int store_steps(int steps) { return (steps + 3) & ~3; }

// "Rainbow" row order of the inner list: tot = partners per class (staged index mod 4), 4 x 8 bits; m = smallest
// class count.  Steps < m hold one partner of every class (lane c of the group takes class c: with 32-byte
// records the class is the record's 16-byte bank group pair, so the four lanes of a group never conflict);
// what is left over is dealt out in (ordinal, class) order.
__device__ __forceinline__ int rainbow_min(unsigned tot) {
  return (int)min(min(tot & 0xFFu, (tot >> 8) & 0xFFu), min((tot >> 16) & 0xFFu, (tot >> 24) & 0xFFu));
}
__device__ __forceinline__ int rainbow_place(unsigned tot, int m, int c, int t) {
  if (t < m) return 4 * t + c;
  int r = 0;
#pragma unroll
  for (int cc = 0; cc < 4; cc++) {
    const int left = (int)((tot >> (8 * cc)) & 0xFFu) - m;
    r += min(left, (t - m) + (cc < c ? 1 : 0));
  }
  return 4 * m + r;
}

}  // namespace mdb
