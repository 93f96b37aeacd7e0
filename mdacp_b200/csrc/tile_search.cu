// Pair search on tiles: MeshList::MakeList / SearchMesh / RegisterPair (meshlist.cc:58-87, 220-271, 288-310).
//
// Reference semantics.  A pair is listed iff r^2 <= SL^2 with r^2 = (dx*dx + dy*dy) + dz*dz evaluated WITHOUT
// fused multiply-add, exactly as the scalar x86-64 reference does (meshlist.cc:262-266), and at least one of the
// two is a local (meshlist.cc:261) -- so the pair SET is bit-identical to the reference's.  Which cells the
// reference happens to visit does not matter: its cells are at least SL wide, so it finds every such pair.
//
// B200 formulation (tile.cuh).  One block per tile of up to 3 x 3 x 3 cells.  The block stages the tile and its
// halo in shared memory as single-precision coordinates relative to the centre of the halo box; a warp takes
// the (<= 32) locals of one interior cell, one per lane, and walks the 27 surrounding cells: every candidate is
// ONE broadcast shared-memory load and 6 FP32 operations per lane.  The FP32 distance decides everything except
// a shell of +-eps around SL^2 (a few candidates per thousand locals); those are re-tested with the exact FP64
// predicate on the global positions.  Hits are collected as bit masks per 32 candidates and processed
// afterwards, so the test loop is branch-free.
//   pass 1 (k_tile_search<false>): row lengths -> steps per force-warp -> row-block offsets (scan on the host
//          side of the stream);
//   pass 2 (k_tile_search<true>): the same tests again, every lane appends its hits to a 32-byte staging
//          sector in shared memory and writes whole sectors of its row (16 entries of 16 bits, layout in unit.h).
// The OUTER list written here is in arrival order; the force kernel walks the INNER list that
// force.cu re-deals (k_redeal) and filters (k_tile_compact) it.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "tile.cuh"

namespace mdb {

static const int TPB = 256;
static inline unsigned nblk(long long n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

// ---------------------------------------------------------------------------------------------
// census: interior locals / staged records / ghosts of every tile
// ---------------------------------------------------------------------------------------------
__global__ void k_tile_census(TileGrid g, Consts c, const int *__restrict__ cell_start, const int *__restrict__ gcell_start,
                              int *__restrict__ tile_nfw, int *__restrict__ tile_flag, Ctrl *ctrl) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.ntiles) return;
  int T[3] = {t % g.nt[0], (t / g.nt[0]) % g.nt[1], t / (g.nt[0] * g.nt[1])};
  int first[3], last[3];
  for (int d = 0; d < 3; d++) {
    first[d] = 1 + T[d] * g.t[d];
    last[d] = min(first[d] + g.t[d], c.ms[d] + 1);
  }
  int n_int = 0, n_all = 0, n_gh = 0;
  for (int z = first[2] - 1; z <= last[2]; z++)
    for (int y = first[1] - 1; y <= last[1]; y++) {
      const int row = c.m[0] * (y + c.m[1] * z);
      const bool irow = z >= first[2] && z < last[2] && y >= first[1] && y < last[1];
      // cells first[0]-1 .. last[0] of this row are contiguous in cell order
      n_all += cell_start[row + last[0] + 1] - cell_start[row + first[0] - 1];
      n_gh += gcell_start[row + last[0] + 1] - gcell_start[row + first[0] - 1];
      if (irow) n_int += cell_start[row + last[0]] - cell_start[row + first[0]];
    }
  tile_nfw[t] = (n_int + kPW - 1) / kPW;
  tile_flag[t] = n_gh > 0 ? 1 : 0;
  atomicMax(&ctrl->scratch[2], n_all + n_gh);
  atomicMax(&ctrl->scratch[3], n_int);
}

// class of a tile: 0 = has rows and stages no ghost (interior), 1 = has rows and stages ghosts (boundary); a tile without
// interior locals (vacuum) is in neither class and appears in no launch
__global__ void k_tile_class(const int *__restrict__ tile_nfw, const int *__restrict__ flag, int *__restrict__ cls_in,
                             int *__restrict__ cls_bd, int nt) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const bool rows = tile_nfw[t] > 0;
  cls_in[t] = (rows && !flag[t]) ? 1 : 0;
  cls_bd[t] = (rows && flag[t]) ? 1 : 0;
}
__global__ void k_tile_order(const int *__restrict__ cls_in, const int *__restrict__ cls_bd, const int *__restrict__ scan_in,
                             const int *__restrict__ scan_bd, int *__restrict__ order, int nt, int n_interior) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  if (cls_in[t]) order[scan_in[t]] = t;
  else if (cls_bd[t]) order[n_interior + scan_bd[t]] = t;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool within_search(const double4 &a, const double4 &b, double SL2) {
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return !(r2 > SL2);
}

struct SearchShared {
  float lo_t, hi_t;     // FP32 thresholds: r2 <= lo_t is a sure hit, r2 > hi_t a sure miss
  unsigned maxrel_bits;
  int pad;
};

// PASS 0 = count (row lengths), 1 = fill (rows at the offsets the host scanned from the counts), 2 = ONE PASS: every
// row block has the same capacity of cap_chunks chunks (sized from the previous build), so the offsets are known up
// front and the tests run once; a row longer than that raises ctrl->scratch[6] and the host falls back to 0 + 1.
// smem: TileShared | SearchShared | float4 sf[cap + 1] | int row_n[maxint + 8] | (fill) ushort stg[threads][16]
template <int PASS>
__global__ void __launch_bounds__(TPB, 3)
k_tile_search(TileGrid g, Consts c, const double4 *__restrict__ q, const int *__restrict__ cell_start,
              const int *__restrict__ gcell_start, const int *__restrict__ gitems, const int *__restrict__ tile_order,
              const int *__restrict__ tile_fw0,
              int *__restrict__ row_count, int *__restrict__ fw_len, int *__restrict__ fw_store, int *__restrict__ fw_sent,
              long long *__restrict__ fw_off, uint2 *__restrict__ list, int cap, int maxint, int cap_chunks, Ctrl *ctrl) {
  constexpr bool FILL = PASS != 0;
  constexpr bool ONEPASS = PASS == 2;
  extern __shared__ __align__(128) unsigned char smem[];
  TileShared &ts = *reinterpret_cast<TileShared *>(smem);
  SearchShared &ss = *reinterpret_cast<SearchShared *>(smem + tile_shared_offset());
  float4 *sf = reinterpret_cast<float4 *>(smem + tile_shared_offset() + 128);
  int *row_n = reinterpret_cast<int *>(sf + cap + 1);
  unsigned short *stg = reinterpret_cast<unsigned short *>(row_n + ((maxint + 8 + 7) & ~7));
  // a count pass may visit only every cap_chunks-th tile (the sample that sizes the first one-pass build)
  const int tile = tile_order[(PASS == 0 && cap_chunks > 1) ? (int)blockIdx.x * cap_chunks : (int)blockIdx.x];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  tile_setup(ts, tile, g, c, cell_start, gcell_start);
  if (threadIdx.x == 0) ss.maxrel_bits = 0u;
  __syncthreads();
  // ---- stage single-precision coordinates relative to the centre of the halo box
  double org[3], edge[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    edge[d] = 1.0 / c.inv_edge[d];
    org[d] = c.s[d] + ((double)(ts.lo[d] - 1) + 0.5 * (double)ts.n[d]) * edge[d];
  }
  {
    float mr = 0.f;
    for (int k = wid; k < ts.ncells; k += nw) {
      const TileCell tc = ts.cell[k];
      for (int l = lane; l < tc.n_l; l += 32) {
        const double4 v = q[tc.src_l + l];
        const float4 f = make_float4((float)(v.x - org[0]), (float)(v.y - org[1]), (float)(v.z - org[2]), 0.f);
        mr = fmaxf(mr, fmaxf(fabsf(f.x), fmaxf(fabsf(f.y), fabsf(f.z))));
        sf[tc.off_l + l] = f;
      }
      for (int l = lane; l < tc.n_g; l += 32) {
        const double4 v = q[gitems[tc.src_g + l]];
        const float4 f = make_float4((float)(v.x - org[0]), (float)(v.y - org[1]), (float)(v.z - org[2]), 0.f);
        mr = fmaxf(mr, fmaxf(fabsf(f.x), fmaxf(fabsf(f.y), fabsf(f.z))));
        sf[tc.off_g + l] = f;
      }
    }
    if (!(mr == mr)) mr = 3.0e38f;  // NaN coordinates: everything goes through the exact predicate
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
    if (lane == 0) atomicMax(&ss.maxrel_bits, __float_as_uint(mr));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // FP32 error of r2 near r = SL: a staged coordinate carries <= 6e-8 * extent, a difference twice that plus its
    // own rounding (<= 1.8e-7 * extent in all), r2 = sum d^2 then <= 2 sqrt(3) SL * 1.8e-7 * extent + 4e-7 SL^2
    // (three products, two sums); eps is that bound times four.  A tile whose records reach further than its halo box should (particles
    // clamped into the outermost cells) takes the exact predicate for every candidate.
    const float extent = __uint_as_float(ss.maxrel_bits);
    const double emax = fmax(edge[0], fmax(edge[1], edge[2]));
    const double nominal = 0.5 * (double)max(ts.n[0], max(ts.n[1], ts.n[2])) * emax + emax;
    if ((double)extent > nominal) {
      ss.lo_t = -1.f;
      ss.hi_t = 3.0e38f;
    } else {
      const double eps = 4.0 * (6.2e-7 * nominal * c.SL + 4.0e-7 * c.SL2);
      ss.lo_t = (float)(c.SL2 - eps);
      ss.hi_t = (float)(c.SL2 + eps);
    }
  }
  __syncthreads();
  const float lo_t = ss.lo_t, hi_t = ss.hi_t;
  const int gfw0 = tile_fw0[tile];
  const int sentinel = ts.n_all;
  unsigned long long nref = 0ull;   // reference-style pair count in half pairs: local-local hits 1, local-ghost 2
  int mrow = 0;
  // ---- one warp per interior cell, one lane per local of it (cells with more than 32 locals: several rounds)
  for (;;) {
    int ic = 0;
    if (lane == 0) ic = atomicAdd(&ts.next_row, 1);
    ic = __shfl_sync(0xffffffffu, ic, 0);
    if (ic >= ts.nicell) break;
    const int k = ts.icell_k[ic];
    const TileCell own = ts.cell[k];
    const int n01 = ts.n[0] * ts.n[1];
    const int kz = small_div(k, small_magic(n01)), kr = k - kz * n01, ky = small_div(kr, small_magic(ts.n[0])), kx = kr - ky * ts.n[0];
    for (int l0 = 0; l0 < own.n_l; l0 += 32) {
      const int l = l0 + lane;
      const bool act = l < own.n_l;
      const int a = own.off_l + (act ? l : 0);  // staged index of this lane's local
      const float4 me = sf[a];
      int n = 0, nloc = 0;
      // FILL: row position
      int e = 0;
      uint4 *rowp = nullptr;
      int store_chunks = 0;
      unsigned short *my = stg + threadIdx.x * 16;
      if (FILL && act) {
        const int gfw = gfw0 + (a >> 3);
        if (ONEPASS) {
          rowp = reinterpret_cast<uint4 *>(list + (long long)gfw * cap_chunks * 32 + 4 * (a & 7));
          store_chunks = cap_chunks;
        } else {
          rowp = reinterpret_cast<uint4 *>(list + fw_off[gfw] + 4 * (a & 7));
          store_chunks = store_steps(fw_len[gfw]) >> 2;
        }
      }
#pragma unroll 1
      for (int nb = 0; nb < 54; nb++) {
        // neighbour cell nb % 27, its locals (nb < 27) or its ghosts
        const int cc = nb < 27 ? nb : nb - 27;
        const int kk = (kx + cc % 3 - 1) + ts.n[0] * ((ky + (cc / 3) % 3 - 1) + ts.n[1] * (kz + cc / 9 - 1));
        const TileCell tc = ts.cell[kk];
        const bool gh = nb >= 27;
        const int jb = gh ? tc.off_g : tc.off_l, cnt_all = gh ? tc.n_g : tc.n_l;
#pragma unroll 1
        for (int j0 = 0; j0 < cnt_all; j0 += 32) {
          const int cnt = min(32, cnt_all - j0);
          const float4 *cand = sf + jb + j0;
          // bit t: candidate t is not a sure miss (r2 <= hi_t, NaN included): a superset of the hits
          unsigned hm = 0u, bit = 1u;
          int t = 0;
          for (; t + 4 <= cnt; t += 4) {
            const float4 s0 = cand[t], s1 = cand[t + 1], s2 = cand[t + 2], s3 = cand[t + 3];
            const float x0 = s0.x - me.x, y0 = s0.y - me.y, z0 = s0.z - me.z;
            const float x1 = s1.x - me.x, y1 = s1.y - me.y, z1 = s1.z - me.z;
            const float x2 = s2.x - me.x, y2 = s2.y - me.y, z2 = s2.z - me.z;
            const float x3 = s3.x - me.x, y3 = s3.y - me.y, z3 = s3.z - me.z;
            const float q0 = fmaf(z0, z0, fmaf(y0, y0, x0 * x0)), q1 = fmaf(z1, z1, fmaf(y1, y1, x1 * x1));
            const float q2 = fmaf(z2, z2, fmaf(y2, y2, x2 * x2)), q3 = fmaf(z3, z3, fmaf(y3, y3, x3 * x3));
            if (!(q0 > hi_t)) hm |= bit;
            if (!(q1 > hi_t)) hm |= bit << 1;
            if (!(q2 > hi_t)) hm |= bit << 2;
            if (!(q3 > hi_t)) hm |= bit << 3;
            bit <<= 4;
          }
          for (; t < cnt; t++) {
            const float4 s0 = cand[t];
            const float x0 = s0.x - me.x, y0 = s0.y - me.y, z0 = s0.z - me.z;
            if (!(fmaf(z0, z0, fmaf(y0, y0, x0 * x0)) > hi_t)) hm |= bit;
            bit <<= 1;
          }
          if (!act) hm = 0u;
          const unsigned self = (unsigned)(a - (jb + j0));
          if (self < 32u) hm &= ~(1u << self);
          if (!FILL) {
            // row lengths from the superset: at most a few entries per thousand rows too many, which can only make
            // a row block one step longer than needed
            n += __popc(hm);
          } else {
            while (hm) {
              const int b = __ffs(hm) - 1;
              hm &= hm - 1u;
              const int j = jb + j0 + b;
              // the shell around SL^2 (r2 between the sure-hit and the sure-miss threshold): exact FP64 predicate
              // on the global records
              const float4 sj = sf[j];
              const float xj = sj.x - me.x, yj = sj.y - me.y, zj = sj.z - me.z;
              if (!(fmaf(zj, zj, fmaf(yj, yj, xj * xj)) <= lo_t)) {
                const int gj = gh ? gitems[tc.src_g + j0 + b] : tc.src_l + j0 + b;
                if (!within_search(q[own.src_l + l], q[gj], c.SL2)) continue;
              }
              n++;
              if (ONEPASS && e >= 16 * store_chunks) continue;  // row longer than its block: counted, reported below
              my[sector_byte(e) >> 1] = (unsigned short)j;
              e++;
              if ((e & 15) == 0) {
                const uint4 *s4 = reinterpret_cast<const uint4 *>(my);
                uint4 *dst = rowp + 16 * ((e >> 4) - 1);  // chunk stride: 32 words x 8 bytes = 16 uint4
                dst[0] = s4[0];
                dst[1] = s4[1];
              }
            }
          }
        }
        if (nb == 26) nloc = n;
      }
      if (FILL) {
        if (act) {
          // pad the row to its storage length with the sentinel
          if (e & 15) {
            for (; e & 15; e++) my[sector_byte(e) >> 1] = (unsigned short)sentinel;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(my);
            uint4 *dst = rowp + 16 * ((e >> 4) - 1);
            dst[0] = s4[0];
            dst[1] = s4[1];
          }
          const unsigned sw = (unsigned)sentinel | ((unsigned)sentinel << 16);
          const uint4 s4 = make_uint4(sw, sw, sw, sw);
          for (int ch = e >> 4; ch < store_chunks; ch++) {
            rowp[16 * ch] = s4;
            rowp[16 * ch + 1] = s4;
          }
        }
        if (act) {
          row_count[own.src_l + l] = n;   // exact
          nref += (unsigned long long)(nloc + 2 * (n - nloc));
          if (ONEPASS) {
            row_n[a] = n;
            mrow = max(mrow, n);
            if (n > 16 * store_chunks) ctrl->scratch[6] = 1;
          }
        }
      } else {
        if (act) {
          row_n[a] = n;                   // upper bound (see above)
          mrow = max(mrow, n);
        }
      }
    }
  }
  if (FILL) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nref += __shfl_xor_sync(0xffffffffu, nref, o);
    if (lane == 0 && nref) atomicAdd(&ctrl->npairs, nref);
  }
  if (!FILL || ONEPASS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mrow = max(mrow, __shfl_xor_sync(0xffffffffu, mrow, o));
    if (lane == 0 && mrow > *(volatile int *)&ctrl->max_row) atomicMax(&ctrl->max_row, mrow);
    __syncthreads();
    // steps of every force-warp: the longest of its (up to) 8 rows in steps of 4 entries, rounded up to even
    const int nfw_t = (ts.n_int + kPW - 1) / kPW;
    for (int fw = threadIdx.x; fw < nfw_t; fw += blockDim.x) {
      int mx = 0;
      for (int m = 0; m < kPW; m++) {
        const int a = fw * kPW + m;
        if (a < ts.n_int) mx = max(mx, (row_n[a] + kG - 1) / kG);
      }
      mx = (mx + 1) & ~1;
      if (ONEPASS) mx = min(mx, 4 * cap_chunks);
      fw_len[gfw0 + fw] = mx;
      if (ONEPASS) fw_off[gfw0 + fw] = (long long)(gfw0 + fw) * cap_chunks * 32;
      else fw_store[gfw0 + fw] = store_steps(mx) * 8;  // uint2 words of the row block
      fw_sent[gfw0 + fw] = sentinel;
    }
  }
  if (FILL) {
    // lanes of the last force-warp that own no local: sentinel rows
    const int nfw_t = (ts.n_int + kPW - 1) / kPW;
    const int a = ts.n_int + threadIdx.x;
    if (a < nfw_t * kPW) {
      const int gfw = gfw0 + (a >> 3);
      uint4 *rowp = ONEPASS ? reinterpret_cast<uint4 *>(list + (long long)gfw * cap_chunks * 32 + 4 * (a & 7))
                            : reinterpret_cast<uint4 *>(list + fw_off[gfw] + 4 * (a & 7));
      const int store_chunks = ONEPASS ? cap_chunks : store_steps(fw_len[gfw]) >> 2;
      const unsigned sw = (unsigned)sentinel | ((unsigned)sentinel << 16);
      const uint4 s4 = make_uint4(sw, sw, sw, sw);
      for (int ch = 0; ch < store_chunks; ch++) {
        rowp[16 * ch] = s4;
        rowp[16 * ch + 1] = s4;
      }
    }
  }
}

static size_t search_smem_bytes(int cap, int maxint, bool fill) {
  size_t b = tile_shared_offset() + 128 + (size_t)(cap + 1) * 16 + (size_t)((maxint + 8 + 7) & ~7) * 4;
  if (fill) b += (size_t)TPB * 32;
  return b;
}

// tile shapes tried in this order (balanced per axis afterwards); a shape is taken if its largest tile fits two
// blocks per SM in the force kernel
static const int kShapes[][3] = {{3, 3, 3}, {3, 3, 2}, {3, 2, 2}, {2, 2, 2}, {2, 2, 1}, {2, 1, 1}, {1, 1, 1}};
// two force blocks per SM: (233,472 B per SM - 1 KB reserved per block) / 2 - bookkeeping (tile_shared_offset), in records
static const int kCapTwoBlocks = (int)(((233472 - 2048) / 2 - (long)tile_shared_offset()) / 32) - 1;
static const int kCapOneBlock = 6400;

static void set_shape(Unit *u, const int want[3]) {
  // want[] sorted descending is matched to the axes sorted by cell count; each axis is then balanced:
  // nt = ceil(ms / t), t = ceil(ms / nt)
  int ax[3] = {0, 1, 2};
  std::sort(ax, ax + 3, [&](int a, int b) { return u->c.ms[a] > u->c.ms[b]; });
  u->ntiles = 1;
  for (int i = 0; i < 3; i++) {
    const int d = ax[i];
    int t = std::min(want[i], u->c.ms[d]);
    const int nt = (u->c.ms[d] + t - 1) / t;
    t = (u->c.ms[d] + nt - 1) / nt;
    u->tdim[d] = t;
    u->tnum[d] = nt;
    u->ntiles *= nt;
  }
}

TileGrid tile_grid(const Unit *u) {
  TileGrid g;
  for (int d = 0; d < 3; d++) { g.t[d] = u->tdim[d]; g.nt[d] = u->tnum[d]; }
  g.ntiles = u->ntiles;
  return g;
}

int build_pair_list(Unit *u) {
  const long long pn = u->pn;
  MDB_CUDA(cudaMemsetAsync(&u->ctrl->npairs, 0, sizeof(unsigned long long), u->stream));
  MDB_CUDA(cudaMemsetAsync(&u->ctrl->max_row, 0, sizeof(int), u->stream));
  u->list_entries = 0;
  u->npairs = 0;
  u->nfw = 0;
  u->ntiles = 0;
  u->n_active_tiles = 0;
  u->n_interior_tiles = 0;
  u->inner_valid = false;
  u->outer_rainbow = false;
  u->inner_is_outer = false;
  if (pn == 0) return MDB200_OK;
  // ---- tile shape: the largest that leaves enough blocks for the device and whose tiles fit shared memory
  int nshapes = (int)(sizeof(kShapes) / sizeof(kShapes[0]));
  int forced[3] = {0, 0, 0};
  if (const char *e = getenv("MDB200_TILE")) {
    if (sscanf(e, "%d,%d,%d", &forced[0], &forced[1], &forced[2]) != 3) forced[0] = 0;
    for (int d = 0; d < 3; d++) forced[d] = std::max(1, std::min(kTileMaxEdge, forced[d]));
  }
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, u->device);
  bool ok = false;
  for (int s = 0; s < nshapes && !ok; s++) {
    const int *want = forced[0] ? forced : kShapes[s];
    set_shape(u, want);
    const bool last = forced[0] || s == nshapes - 1;
    // small systems: prefer more, smaller tiles until every SM has a few blocks
    if (!last && u->ntiles < 4 * sm_count) continue;
    MDB_TRY(u->tile_fw0.ensure((size_t)u->ntiles + 2));
    MDB_TRY(u->tile_flag.ensure((size_t)u->ntiles + 2));
    MDB_TRY(u->tile_order.ensure((size_t)u->ntiles + 2));
    MDB_TRY(u->scan.ensure((size_t)u->ntiles + 2));
    MDB_TRY(u->flags.ensure((size_t)u->ntiles + 2));
    MDB_TRY(u->scan_ws.ensure(2 * scan_workspace_ints(u->ntiles + 2)));
    MDB_CUDA(cudaMemsetAsync(&u->ctrl->scratch[2], 0, 2 * sizeof(int), u->stream));
    k_tile_census<<<nblk(u->ntiles), TPB, 0, u->stream>>>(tile_grid(u), u->c, u->cell_start.ptr, u->gcell_start.ptr,
                                                           u->flags.ptr, u->tile_flag.ptr, u->ctrl);
    u->launches++;
    MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[2], &u->ctrl->scratch[2], 2 * sizeof(int), cudaMemcpyDeviceToHost, u->stream));
    MDB_CUDA(cudaStreamSynchronize(u->stream));
    u->tile_cap = u->ctrl_host->scratch[2];
    u->tile_maxint = u->ctrl_host->scratch[3];
    ok = (u->tile_cap <= kCapTwoBlocks && u->tile_maxint <= kTileMaxRows * kPW) || last;
    if (forced[0]) break;
  }
  if (u->tile_maxint > kTileMaxRows * kPW) {
    set_error("a tile holds %d particles of its own: more than %d force-warp rows", u->tile_maxint, kTileMaxRows);
    return MDB200_ECAPACITY;
  }
  if (u->tile_cap > kCapOneBlock || u->tile_cap >= 65535) {
    set_error("a tile of one cell and its halo holds %d particles: more than shared memory can stage", u->tile_cap);
    return MDB200_ECAPACITY;
  }
  if (getenv("MDB200_DEBUG"))
    fprintf(stderr, "[mdb200] tiles %d x %d x %d of %d x %d x %d cells, largest stages %d records (%d interior)\n", u->tnum[0],
            u->tnum[1], u->tnum[2], u->tdim[0], u->tdim[1], u->tdim[2], u->tile_cap, u->tile_maxint);
  // ---- force-warps per tile -> first force-warp of every tile; the order of the launches: interior tiles with rows
  //      first, then boundary tiles with rows; tiles without interior locals (vacuum) are left out
  MDB_TRY(exclusive_scan_i32(u->flags.ptr, u->tile_fw0.ptr, u->ntiles, u->scan_ws.ptr, u->stream, &u->launches));
  MDB_TRY(u->tile_tmp.ensure(3 * ((size_t)u->ntiles + 2)));
  int *cls_in = u->tile_tmp.ptr, *cls_bd = cls_in + u->ntiles + 2, *scan_bd = cls_bd + u->ntiles + 2;
  k_tile_class<<<nblk(u->ntiles), TPB, 0, u->stream>>>(u->flags.ptr, u->tile_flag.ptr, cls_in, cls_bd, u->ntiles);
  u->launches++;
  MDB_TRY(exclusive_scan_i32(cls_in, u->scan.ptr, u->ntiles, u->scan_ws.ptr, u->stream, &u->launches));
  MDB_TRY(exclusive_scan_i32(cls_bd, scan_bd, u->ntiles, u->scan_ws.ptr, u->stream, &u->launches));
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[4], u->tile_fw0.ptr + u->ntiles, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[5], u->scan.ptr + u->ntiles, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[6], scan_bd + u->ntiles, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  u->nfw = u->ctrl_host->scratch[4];
  u->n_interior_tiles = u->ctrl_host->scratch[5];
  u->n_active_tiles = u->n_interior_tiles + u->ctrl_host->scratch[6];
  k_tile_order<<<nblk(u->ntiles), TPB, 0, u->stream>>>(cls_in, cls_bd, u->scan.ptr, scan_bd, u->tile_order.ptr, u->ntiles, u->n_interior_tiles);
  u->launches++;
  if (getenv("MDB200_DEBUG") && u->n_active_tiles != u->ntiles)
    fprintf(stderr, "[mdb200] %d of %d tiles hold particles of their own\n", u->n_active_tiles, u->ntiles);
  const long long nfw = u->nfw;
  MDB_TRY(u->row_count.ensure((size_t)pn + 32));
  MDB_TRY(u->fw_len.ensure((size_t)nfw + 2));
  MDB_TRY(u->fw_len_in.ensure((size_t)nfw + 2));
  MDB_TRY(u->fw_sent.ensure((size_t)nfw + 2));
  MDB_TRY(u->fw_off.ensure((size_t)nfw + 2));
  MDB_TRY(u->scan.ensure((size_t)nfw + 2));
  MDB_TRY(u->scan_ws64.ensure(scan_workspace_ints(nfw + 2)));
  const size_t sm1 = search_smem_bytes(u->tile_cap, u->tile_maxint, false), sm2 = search_smem_bytes(u->tile_cap, u->tile_maxint, true);
  MDB_CUDA(cudaFuncSetAttribute(k_tile_search<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  MDB_CUDA(cudaFuncSetAttribute(k_tile_search<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  MDB_CUDA(cudaFuncSetAttribute(k_tile_search<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
#define SEARCH_ARGS(CAPCH) tile_grid(u), u->c, u->q.ptr, u->cell_start.ptr, u->gcell_start.ptr, u->gcell_items.ptr, u->tile_order.ptr, \
                           u->tile_fw0.ptr,                                                                                         \
                           u->row_count.ptr, u->fw_len.ptr, u->scan.ptr, u->fw_sent.ptr, u->fw_off.ptr, u->list.ptr, u->tile_cap,   \
                           u->tile_maxint, CAPCH, u->ctrl
  bool built = false;
  static const bool no_onepass = getenv("MDB200_NO_ONEPASS") != nullptr;
  // ---- first build of a store: no previous longest row to size the row blocks of a one-pass build with.  The count
  //      pass over every 8th tile gives one (a few entries short of the true maximum at worst: the capacity below adds
  //      12 % + 8, and a row that outgrows it all the same sends the build through the two passes).  Not worth it for
  //      small units (MDB200_SAMPLE_TILES = smallest tile count it is done for; tests force it with 1).
  {
    const char *e = getenv("MDB200_SAMPLE_TILES");
    const int min_tiles = e ? atoi(e) : 1024;
    const int stride = 8;
    if (u->max_row_exact == 0 && !no_onepass && min_tiles > 0 && u->n_active_tiles >= min_tiles) {
      k_tile_search<0><<<(u->n_active_tiles + stride - 1) / stride, TPB, sm1, u->stream>>>(SEARCH_ARGS(stride));
      u->launches++;
      MDB_CUDA(cudaGetLastError());
      MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->max_row, &u->ctrl->max_row, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
      MDB_CUDA(cudaMemsetAsync(&u->ctrl->npairs, 0, sizeof(unsigned long long), u->stream));
      MDB_CUDA(cudaMemsetAsync(&u->ctrl->max_row, 0, sizeof(int), u->stream));
      MDB_CUDA(cudaStreamSynchronize(u->stream));
      u->max_row_exact = u->ctrl_host->max_row;
      if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] first build: longest row of the sampled tiles %d\n", u->max_row_exact);
    }
  }
  // ---- one pass with row blocks of fixed capacity, sized from the longest row of the previous build (or of the sample
  //      above): the distance tests run once
  if (u->max_row_exact > 0 && !no_onepass) {
    const int cap_steps = ((u->max_row_exact + u->max_row_exact / 8 + 8 + kG - 1) / kG + 1) & ~1;  // + 12 % + 8 entries
    const int cap_chunks = (cap_steps + 3) >> 2;
    if (cap_chunks <= 16) {
      const size_t need = (size_t)nfw * cap_chunks * 32 + 8192;
      if (need > u->list.cap) {
        if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] list grows: %zu words > capacity %zu\n", need, u->list.cap);
        // a first allocation gets the head-room of the two-pass build below (rows of a lattice grow as it thermalises)
        MDB_TRY(u->list.ensure(u->list.cap == 0 ? need + need * 6 / 10 : need + need / 10));
      }
      MDB_TRY(u->list_in.ensure(u->list.cap));
      MDB_CUDA(cudaMemsetAsync(&u->ctrl->scratch[6], 0, sizeof(int), u->stream));
      k_tile_search<2><<<u->n_active_tiles, TPB, sm2, u->stream>>>(SEARCH_ARGS(cap_chunks));
      u->launches++;
      MDB_CUDA(cudaGetLastError());
      MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[6], &u->ctrl->scratch[6], sizeof(int), cudaMemcpyDeviceToHost, u->stream));
      MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->max_row, &u->ctrl->max_row, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
      MDB_CUDA(cudaStreamSynchronize(u->stream));
      if (u->ctrl_host->scratch[6] == 0) {
        built = true;
        u->max_row = u->ctrl_host->max_row;
        u->max_row_exact = u->max_row;
        u->list_entries = (long long)nfw * cap_chunks * 128;
      } else {
        if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] one-pass list build: a row of %d entries outgrew %d chunks; two passes\n", u->ctrl_host->max_row, cap_chunks);
        u->max_row_exact = std::max(u->max_row_exact, u->ctrl_host->max_row);
        MDB_CUDA(cudaMemsetAsync(&u->ctrl->npairs, 0, sizeof(unsigned long long), u->stream));
        MDB_CUDA(cudaMemsetAsync(&u->ctrl->max_row, 0, sizeof(int), u->stream));
      }
    }
  }
  if (!built) {
    // ---- pass 1: row lengths
    k_tile_search<0><<<u->n_active_tiles, TPB, sm1, u->stream>>>(SEARCH_ARGS(0));
    u->launches++;
    MDB_CUDA(cudaGetLastError());
    MDB_TRY(exclusive_scan_i32_to_i64(u->scan.ptr, u->fw_off.ptr, nfw, u->scan_ws64.ptr, u->stream, &u->launches, 1));
    MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->scratch[4], u->fw_off.ptr + nfw, sizeof(long long), cudaMemcpyDeviceToHost, u->stream));
    MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->max_row, &u->ctrl->max_row, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
    MDB_CUDA(cudaStreamSynchronize(u->stream));
    long long words;
    memcpy(&words, &u->ctrl_host->scratch[4], sizeof(long long));
    u->list_entries = words * 4;
    u->max_row = u->ctrl_host->max_row;   // upper bound of the longest row
    u->max_row_exact = u->max_row;        // (the count pass counts a superset: good enough to size the next one-pass build)
    if (u->max_row > 252) {
      set_error("a particle has %d partners within the search length: rows above 252 entries are not supported", u->max_row);
      return MDB200_ECAPACITY;
    }
    // Head-room: rows of a freshly generated lattice grow by ~20 % once it thermalises, and re-allocating a multi-GB
    // list inside the step loop stalls for hundreds of ms; the walk reads chunks past a row block, hence the slack.
    const size_t need = (size_t)words + 8192;
    if (need > u->list.cap) {
      if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] list grows: %zu words > capacity %zu\n", need, u->list.cap);
      MDB_TRY(u->list.ensure(need + need * 6 / 10));
    }
    MDB_TRY(u->list_in.ensure(u->list.cap));
    // ---- pass 2: rows
    k_tile_search<1><<<u->n_active_tiles, TPB, sm2, u->stream>>>(SEARCH_ARGS(0));
    u->launches++;
    MDB_CUDA(cudaGetLastError());
  }
#undef SEARCH_ARGS
  // the reference-style pair count (exact: counted by the fill pass); the caller synchronises anyway
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->npairs, &u->ctrl->npairs, sizeof(unsigned long long), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  u->npairs = (long long)(u->ctrl_host->npairs / 2);
  return MDB200_OK;
}

// ---------------------------------------------------------------------------------------------
// parity helper: the list as the reference's raw (key, partner) arrays, reference indices
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_tile_extract(TileGrid g, Consts c, const int *__restrict__ cell_start, const int *__restrict__ gcell_start,
               const int *__restrict__ gitems, const int *__restrict__ tile_fw0, const int *__restrict__ row_count,
               const long long *__restrict__ fw_off, const uint2 *__restrict__ list, const int *__restrict__ ref_of_slot,
               int *__restrict__ key, int *__restrict__ partner, unsigned long long *cursor, unsigned long long cap, int pn) {
  extern __shared__ __align__(128) unsigned char smem[];
  TileShared &ts = *reinterpret_cast<TileShared *>(smem);
  int *sslot = reinterpret_cast<int *>(smem + tile_shared_offset());
  const int tile = blockIdx.x;
  tile_setup(ts, tile, g, c, cell_start, gcell_start);
  for (int k = threadIdx.x >> 5; k < ts.ncells; k += blockDim.x >> 5) {
    const TileCell tc = ts.cell[k];
    for (int l = threadIdx.x & 31; l < tc.n_l; l += 32) sslot[tc.off_l + l] = tc.src_l + l;
    for (int l = threadIdx.x & 31; l < tc.n_g; l += 32) sslot[tc.off_g + l] = gitems[tc.src_g + l];
  }
  __syncthreads();
  const int gfw0 = tile_fw0[tile];
  for (int a = threadIdx.x; a < ts.n_int; a += blockDim.x) {
    const int slot = sslot[a];
    const int ra = ref_of_slot[slot];
    const int n = row_count[slot];
    const unsigned char *row = reinterpret_cast<const unsigned char *>(list + fw_off[gfw0 + (a >> 3)] + 4 * (a & 7));
    for (int e = 0; e < n; e++) {
      const int j = *reinterpret_cast<const unsigned short *>(row + 256 * (e >> 4) + sector_byte(e));
      const int sj = sslot[j];
      const int rj = (sj < pn) ? ref_of_slot[sj] : sj;
      if (ra < rj) {
        const unsigned long long pos = atomicAdd(cursor, 1ull);
        if (pos < cap) { key[pos] = ra; partner[pos] = rj; }
      }
    }
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
  const size_t sm = tile_shared_offset() + (size_t)(u->tile_cap + 1) * 4;
  MDB_CUDA(cudaFuncSetAttribute(k_tile_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_tile_extract<<<u->ntiles, TPB, sm, u->stream>>>(tile_grid(u), u->c, u->cell_start.ptr, u->gcell_start.ptr, u->gcell_items.ptr,
                                                    u->tile_fw0.ptr, u->row_count.ptr, u->fw_off.ptr, u->list.ptr,
                                                    u->ref_of_slot.ptr, dk.ptr, dp.ptr, cur, (unsigned long long)u->npairs, (int)u->pn);
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
