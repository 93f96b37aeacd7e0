// LJ force kick: ForceCalculator::CalculateForce (fcalculator.cc:35-50); semantic spec
// ForceCalculator::CalculateForceSorted (fcalculator.cc:182-223):
//   for each listed pair (i,j): d = q_j - q_i, r2 = |d|^2; if r2 > CL2 the pair contributes 0;
//   r6 = r2^3; df = ((24 r6 - 48)/(r6 r6 r2) + 8 C2) dt; p_i += df d; p_j -= df d.
//
// B200 formulation (round 2; tile.cuh has the why).
//  * One block per TILE of up to 3 x 3 x 3 cells: the positions of the tile and of its one-cell halo are staged
//    in shared memory once (32-byte records), partners are 16-bit indices into that staged array, and every
//    gather is two shared-memory loads.  Round 1 gathered from global memory and was bound by the L1 data pipe
//    at 20 wavefronts per warp-gather (profiles/r01c_force_ncu.txt); here a gather costs 8.
//  * FULL neighbour list, no atomics: every local's momentum is owned by one lane group, the summation order is
//    deterministic and ghost momenta are never written (they are scratch in the reference, fcalculator.cc:273-275).
//    red.global.add.f64 for the reaction term of a half list was measured at 1.2-1.3 clk per reduction per SM
//    (profiles/r02_ubench_red_f64.txt): ~180 clk per local against ~30 for this whole kernel.
//  * kG = 4 lanes per local; a force-warp owns 8 consecutive staged interior locals.  A staged record is
//    (x, y, z, y): EVEN lane groups read the low half (x, y) with one LDS.128 and z with one LDS.64, ODD groups
//    read the high half (z, y) first and x second, and simply compute with x and z exchanged (r2 and the three
//    accumulators do not care; the writer lane swaps them back).  The 8 lanes of a quarter-warp then touch 8
//    different 16-byte bank groups as long as the 4 lanes of a group read records of 4 different classes
//    (staged index mod 4) -- which is how k_redeal deals the rows out ("rainbow" order).
//  * the cutoff test is one integer compare on the high word of r2 (exact when the low word of CL2 is zero, as
//    for cutoff 3.0 and 2.5; a full 64-bit compare otherwise), and zeroes df with two integer selects.
//  * list words for four steps come from one 8-byte load per lane (a contiguous 256-byte run per warp), loaded
//    one chunk ahead and prefetched three ahead.
//
// Roofline: algorithmic HBM bytes per particle = 4 B x P (half-list indices) + 96 B
// (q read, p read + write), 30 flop per listed unordered pair.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <string.h>
#include <map>
#include <mutex>
#include <utility>
#include "rng.cuh"
#include "tile.cuh"

namespace mdb {

TileGrid tile_grid(const Unit *u);  // tile_search.cu

#ifndef FORCE_TPB
#define FORCE_TPB 512
#endif

struct ForceConsts {
  double dt2;     // dt / 2
  double CL2;
  double c24dt;   // 24 dt
  double c48dt;   // 48 dt
  double c2x8dt;  // 8 C2 dt
  int cl2_hi;     // high word of CL2
  int exact_cut;  // low word of CL2 != 0: compare all 64 bits
  // FORCE_KICK_LANGEVIN: ForceCalculator::Langevin (fcalculator.cc:765-778) in the epilogue
  double dt, gamma, amp;
  unsigned key0, key1, step_lo, step_hi;
  double keep2;   // deriving launch: (cutoff + delta)^2
};

__device__ __forceinline__ void lds128(unsigned a, double &x, double &y) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ double lds64(unsigned a) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a));
  return x;
}

// 1 / x: hardware seed (MUFU.RCP64H, ~2^-20) + one cubic (Halley-type) refinement y (1 + e + e^2): ~2^-60
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// one listed entry: o0 / o1 = shared-memory addresses of the record's first / second half as this lane reads it
template <bool EXACT>
__device__ __forceinline__ double pair_kick(unsigned o0, unsigned o1, double a0, double a1, double a2, const ForceConsts &fc,
                                            double &f0, double &f1, double &f2) {
  double j0, j1;
  lds128(o0, j0, j1);
  const double j2 = lds64(o1);
  const double d0 = j0 - a0, d1 = j1 - a1, d2 = j2 - a2;
  const double r2 = fma(d2, d2, fma(d1, d1, d0 * d0));
  // s = 1 / r2; df = s^4 (24 - 48 s^3) dt + 8 C2 dt  (= ((24 r6 - 48) / (r6 r6 r2) + 8 C2) dt)
  const double s = fast_rcp(r2);
  const double s2 = s * s;
  const double s3 = s2 * s;
  const double s4 = s2 * s2;
  const double t = fma(-fc.c48dt, s3, fc.c24dt);
  double df = fma(s4, t, fc.c2x8dt);
  // r2 > CL2: the pair contributes nothing (fcalculator.cc:209-211); also kills the sentinel padding
  bool in;
  if (EXACT) in = __double_as_longlong(r2) <= __double_as_longlong(fc.CL2);
  else in = __double2hiint(r2) < fc.cl2_hi;
  df = __hiloint2double(in ? __double2hiint(df) : 0, in ? __double2loint(df) : 0);
  f0 = fma(df, d0, f0);
  f1 = fma(df, d1, f1);
  f2 = fma(df, d2, f2);
  return r2;
}

// MODE 0: impulses to fout (parity probe).  MODE 1: kick, p += F dt (ForceCalculator::CalculateForce).
// MODE 2: kick fused with the second half drift of MDManager::CalculateForce (mdmanager.cc:151-157):
//   q_out[a] = q[a] + p_new dt/2 with the same non-fused mul/add as k_drift_half (bit-identical), written to
//   the OTHER position buffer because q is still being staged by other blocks; the caller swaps the
//   buffers.  It also leaves max |p_new|^2 in ctrl->vmax2_bits for the next step's velocity-budget test
//   (pairlist.cc:48-65), so that test needs no pass of its own.
// MODE 3: as MODE 2 plus the FIRST half drift of the NEXT step (mdmanager.cc:128-131), speculatively:
//   q_out[a] = (q[a] + p_new dt/2) + p_new dt/2, two separately rounded additions exactly like two
//   k_drift_half passes.  The end-of-step positions q[a] + p_new dt/2 are not stored; whoever needs them
//   (rebuild, observation, download, displacement test) recomputes them bit-identically from the untouched
//   input buffer and p_new (api.cu:materialize, integrate.cu).
// MODE 4: the Langevin step of MDManager::CalculateLangevin (mdmanager.cc:186-192): kick, then p += (-gamma p + xi) dt
//   (rng.cuh), NO second half drift (SURVEY.md finding 5); the first half drift of the NEXT step is done
//   speculatively, q_out[a] = q[a] + p_new dt/2, so the end-of-step positions are simply the untouched input buffer.
enum { FORCE_IMPULSE = 0, FORCE_KICK = 1, FORCE_KICK_DRIFT = 2, FORCE_KICK_DRIFT2 = 3, FORCE_KICK_LANGEVIN = 4 };

struct TileArgs {
  TileGrid g;
  const int *cell_start, *gcell_start, *gitems, *tile_fw0, *tile_order;
  int tile_base;     // first entry of tile_order this launch covers
  int cap;           // staged-record capacity the shared memory was sized for
};

// The list words of a row two chunks (8 steps) ahead of their use: a chunk takes a warp ~500 clk at 8 warps per
// scheduler, a miss to HBM ~1000.  Two named buffers, reloaded in place (rotating one into the other makes the copy wait
// for the younger load: ncu showed 14 % of all stall samples there).  The allocation is padded: loading three chunks
// past the last row is harmless.
struct RowStream {
  const uint2 *lp;   // this lane's word of the row's first chunk not yet in wa / wb ... - 64
  uint2 wa, wb;
};
__device__ __forceinline__ void row_stream_open(RowStream &rs, const uint2 *lp) {
  rs.lp = lp;
  rs.wa = __ldg(lp);
  rs.wb = __ldg(lp + 32);
}

// Walks `steps` steps of the row rs points at, 4 lanes per local, and leaves rs at the row whose first word (of this
// lane) is nlp: its first two chunks are requested before the last trip of this row, so that they travel while it is
// finished, reduced and written back.  (The rows of the inner list keep the row offsets of the outer list, i.e. they
// are NOT packed back to back: the stream cannot simply run on.)
// What a DERIVING walk keeps per lane (k_tile_force_p<.., DERIVE>): the launch walks the re-dealt OUTER list and, on the
// side, writes the inner list -- every lane packs the entries of its column with r2 <= fc.keep2 into words of the
// inner row block (same row offsets): the compaction of k_tile_compact for the price of the longer walk, since r2 is
// computed here anyway.
struct DeriveLane {
  unsigned long long acc;          // the output word being filled
  int nk;                          // kept entries of this lane
  bool on;                         // this lane's local exists (the last row of a tile has phantom locals)
  unsigned long long *dst;         // this lane's word 0 of the inner row block
};

template <bool EXACT, bool DERIVE>
__device__ __forceinline__ void walk_row(RowStream &rs, const uint2 *__restrict__ nlp, int steps, unsigned hx, unsigned hx2, double a0,
                                         double a1, double a2, const ForceConsts &fc, double &f0, double &f1, double &f2,
                                         DeriveLane &dl) {
#define KEEP_ENTRY(R2, E)                                                         \
  do {                                                                            \
    if (DERIVE && dl.on && (R2) <= fc.keep2) {                                    \
      dl.acc |= (unsigned long long)((E) >> 5) << (16 * (dl.nk & 3));             \
      dl.nk++;                                                                    \
      if ((dl.nk & 3) == 0) {                                                     \
        dl.dst[32 * ((dl.nk >> 2) - 1)] = dl.acc;                                 \
        dl.acc = 0ull;                                                            \
      }                                                                           \
    }                                                                             \
  } while (0)
#define WALK_CHUNK(W, HALF_ONLY)                                                  \
  do {                                                                            \
    const unsigned e0 = ((W).x & 0xffffu) << 5, e1 = ((W).x >> 16) << 5;          \
    const unsigned e2 = ((W).y & 0xffffu) << 5, e3 = ((W).y >> 16) << 5;          \
    const double r0 = pair_kick<EXACT>(e0 + hx, e0 + hx2, a0, a1, a2, fc, f0, f1, f2); \
    KEEP_ENTRY(r0, e0);                                                           \
    const double r1 = pair_kick<EXACT>(e1 + hx, e1 + hx2, a0, a1, a2, fc, f0, f1, f2); \
    KEEP_ENTRY(r1, e1);                                                           \
    if (!(HALF_ONLY)) {                                                           \
      const double r2 = pair_kick<EXACT>(e2 + hx, e2 + hx2, a0, a1, a2, fc, f0, f1, f2); \
      KEEP_ENTRY(r2, e2);                                                         \
      const double r3 = pair_kick<EXACT>(e3 + hx, e3 + hx2, a0, a1, a2, fc, f0, f1, f2); \
      KEEP_ENTRY(r3, e3);                                                         \
    }                                                                             \
  } while (0)
  const uint2 *__restrict__ lp = rs.lp;
  uint2 wa = rs.wa, wb = rs.wb;
  // rem = steps still to walk (even)
#pragma unroll 1
  for (int rem = steps; rem > 0; rem -= 8) {
    // the last trip reloads the buffers with the first two chunks of the next row instead of words past this one
    const uint2 *__restrict__ src = (rem <= 8) ? nlp - 64 : lp;
    {
      const uint2 w = wa;
      wa = __ldg(src + 64);
      WALK_CHUNK(w, rem <= 2);
    }
    {
      const uint2 w = wb;
      wb = __ldg(src + 96);
      if (rem > 4) WALK_CHUNK(w, rem <= 6);
    }
    lp += 64;
  }
#undef WALK_CHUNK
#undef KEEP_ENTRY
  if (steps <= 0) {
    wa = __ldg(nlp);
    wb = __ldg(nlp + 32);
  }
  rs.lp = nlp;
  rs.wa = wa;
  rs.wb = wb;
}

// end of a deriving row: the word in progress topped up with the sentinel, sentinel words up to the row's new length
// (the longest column, rounded up to even), which lane 0 records
__device__ __forceinline__ void derive_finish_row(DeriveLane &dl, int sentinel, int lane, int *len_slot) {
  int steps_in = dl.nk;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) steps_in = max(steps_in, __shfl_xor_sync(0xffffffffu, steps_in, o));
  steps_in = (steps_in + 1) & ~1;
  const int nchunk_in = (steps_in + 3) >> 2;
  const unsigned long long sent = (unsigned long long)sentinel;
  const unsigned long long sent4 = sent | (sent << 16) | (sent << 32) | (sent << 48);
  int wi = dl.nk >> 2;
  if (dl.nk & 3) {
    dl.dst[32 * wi] = dl.acc | (sent4 & (~0ull << (16 * (dl.nk & 3))));
    wi++;
  }
  for (; wi < nchunk_in; wi++) dl.dst[32 * wi] = sent4;
  if (lane == 0) *len_slot = steps_in;
}

// End of a row: the four lanes of a local add up, the writer lane (lane & 3 == 0) applies the MODE's update to the local
// with staged index a / slot `slot` (see the MODE list above).  pa = the writer lane's momentum, loaded by the caller.
template <int MODE>
__device__ __forceinline__ void finish_row(double f0, double f1, double f2, bool writer, unsigned gp, int a, int slot, double4 pa,
                                           const double4 *sq, double4 *__restrict__ p, double *__restrict__ fout,
                                           const int *__restrict__ ref_of_slot, double4 *__restrict__ q_out, const ForceConsts &fc,
                                           double &v2max) {
#pragma unroll
  for (int o = kG / 2; o > 0; o >>= 1) {
    f0 += __shfl_xor_sync(0xffffffffu, f0, o);
    f1 += __shfl_xor_sync(0xffffffffu, f1, o);
    f2 += __shfl_xor_sync(0xffffffffu, f2, o);
  }
  if (!writer) return;
  const double fx = gp ? f2 : f0, fy = f1, fz = gp ? f0 : f2;
  if (MODE == FORCE_IMPULSE) {
    const long long ra = ref_of_slot[slot];
    fout[3 * ra + 0] = fx; fout[3 * ra + 1] = fy; fout[3 * ra + 2] = fz;
    return;
  }
  pa.x += fx; pa.y += fy; pa.z += fz;
  if (MODE == FORCE_KICK_LANGEVIN)
    langevin_update(pa, (unsigned)ref_of_slot[slot], fc.gamma, fc.amp, fc.dt, fc.key0, fc.key1, fc.step_lo, fc.step_hi);
  st_256(p + slot, pa);
  if (MODE == FORCE_KICK_DRIFT || MODE == FORCE_KICK_DRIFT2 || MODE == FORCE_KICK_LANGEVIN) {
    const double4 qa = sq[a];
    // fcalculator.cc:22-32, no FMA contraction: identical to k_drift_half
    const double sx = __dmul_rn(pa.x, fc.dt2), sy = __dmul_rn(pa.y, fc.dt2), sz = __dmul_rn(pa.z, fc.dt2);
    double nx = __dadd_rn(qa.x, sx), ny = __dadd_rn(qa.y, sy), nz = __dadd_rn(qa.z, sz);
    if (MODE == FORCE_KICK_DRIFT2) { nx = __dadd_rn(nx, sx); ny = __dadd_rn(ny, sy); nz = __dadd_rn(nz, sz); }
    st_256(q_out + slot, make_double4(nx, ny, nz, ny));  // lane 3 repeats y (tile.cuh)
    v2max = fmax(v2max, __dadd_rn(__dadd_rn(__dmul_rn(pa.x, pa.x), __dmul_rn(pa.y, pa.y)), __dmul_rn(pa.z, pa.z)));
  }
}

// max |p|^2 of the launch for the next step's velocity-budget test
__device__ __forceinline__ void publish_v2max(double v2max, int lane, Ctrl *ctrl) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v2max = fmax(v2max, __shfl_xor_sync(0xffffffffu, v2max, o));
  if (lane == 0) {
    // non-negative doubles order like their bit patterns; the maximum settles after a few blocks, so the
    // plain read filters out almost every atomic
    const unsigned long long b = (unsigned long long)__double_as_longlong(v2max);
    if (b > *(volatile unsigned long long *)&ctrl->vmax2_bits) atomicMax(&ctrl->vmax2_bits, b);
  }
}

// ---------------------------------------------------------------------------------------------------------
// One block per tile (two blocks of 512 threads per SM): small units, tiles too large for two staged copies.
// ---------------------------------------------------------------------------------------------------------
template <int MODE, bool EXACT>
__global__ void __launch_bounds__(FORCE_TPB, 2)
k_tile_force(TileArgs ta, Consts c, const double4 *__restrict__ q, double4 *__restrict__ p, double *__restrict__ fout,
             const uint2 *__restrict__ list, const long long *__restrict__ fw_off, const int *__restrict__ fw_len,
             const int *__restrict__ ref_of_slot, double4 *__restrict__ q_out, Ctrl *ctrl, ForceConsts fc) {
  extern __shared__ __align__(128) unsigned char smem[];
  TileShared &ts = *reinterpret_cast<TileShared *>(smem);
  double4 *sq = reinterpret_cast<double4 *>(smem + tile_shared_offset());
  const int tile = ta.tile_order ? ta.tile_order[ta.tile_base + blockIdx.x] : (int)blockIdx.x;
  tile_setup(ts, tile, ta.g, c, ta.cell_start, ta.gcell_start);
  const int gfw0 = ta.tile_fw0[tile];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n_int = ts.n_int;
  const int nfw = (n_int + kPW - 1) / kPW;
  // A warp walks a CONTIGUOUS range of the tile's rows (neighbouring row blocks: one L2 prefetch covers what comes
  // next).  The first two chunks are requested here, before the tile is staged.
  const int rbase = nfw / nwarps, rrem = nfw - rbase * nwarps;
  const int r0 = wid * rbase + min(wid, rrem), r1 = r0 + rbase + (wid < rrem ? 1 : 0);
  RowStream rs;
  row_stream_open(rs, list + (r0 < r1 ? fw_off[gfw0 + r0] : fw_off[gfw0]) + lane);
  tile_stage_rows(ts, gfw0, fw_off, fw_len);
  tile_stage_tma(ts, q, ta.gitems, sq);
  tile_stage_wait(ts);
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sq);
  const unsigned gp = (lane >> 2) & 1u;                                // odd groups: x and z exchanged
  const unsigned hx = sbase + (gp << 4), hx2 = sbase + ((gp ^ 1u) << 4);
  double v2max = 0.0;
  for (int fw = r0; fw < r1; fw++) {
    // keep the stream ahead in L2: the row block two rows further on (one instruction, a 128-byte line per lane)
    if (fw + 2 < r1) {
      const int bytes = ((ts.row_len[fw + 2] + 3) >> 2) * 256;
      if (128 * lane < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(list + ts.row_base + ts.row_off[fw + 2]) + 128 * lane));
    }
    const int a = fw * kPW + (lane >> 2);
    const bool act = a < n_int;
    const unsigned ao = (unsigned)(act ? a : ts.n_all) << 5;
    double a0, a1;
    lds128(ao + hx, a0, a1);
    const double a2 = lds64(ao + hx2);
    const int steps = ts.row_len[fw];
    // the writer lane's momentum is needed at the end of the row: pull its line (and the next row's) towards the SM
    const bool writer = act && (lane & 3) == 0;
    int slot = 0;
    if (writer) {
      slot = tile_interior_slot(ts, a);
      if (MODE != FORCE_IMPULSE) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + slot + kPW));
    }
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    const uint2 *__restrict__ nlp = (fw + 1 < r1) ? list + ts.row_base + ts.row_off[fw + 1] + lane : rs.lp;
    DeriveLane no_dl;
    walk_row<EXACT, false>(rs, nlp, steps, hx, hx2, a0, a1, a2, fc, f0, f1, f2, no_dl);
    double4 pa = make_double4(0.0, 0.0, 0.0, 0.0);
    if (MODE != FORCE_IMPULSE && writer) pa = ld_256(p + slot);   // requested before the reduction needs it
    finish_row<MODE>(f0, f1, f2, writer, gp, a, slot, pa, sq, p, fout, ref_of_slot, q_out, fc, v2max);
  }
  if (MODE == FORCE_KICK_DRIFT || MODE == FORCE_KICK_DRIFT2 || MODE == FORCE_KICK_LANGEVIN) publish_v2max(v2max, lane, ctrl);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant: one block per SM walks a strided sequence of tiles through TWO staged copies.
//
// With one block per tile a third of the block's life is not walking: the tile set-up and the staging (every warp
// waits), and the wait for the slowest warp at the end (84 rows over 16 warps: a warp has 5 or 6).  Here the rows of
// consecutive tiles are one stream of TICKETS handed out by a shared-memory counter: a walker warp that has finished a
// row takes the next ticket, whichever tile it belongs to, and only waits if that tile is not published yet.  One extra
// warp is the PRODUCER: it sets tile n + 2 up (cell table, prefix sums, row table: TileShared has three slots) while
// tiles n and n + 1 are being walked; as soon as every row of tile n is done it issues the bulk copies of tile n + 2
// into the buffer tile n occupied, gathers the ghosts, waits for the copies and publishes the tile with a release
// store.  Nobody waits at a block barrier, and the last rows of a tile overlap the first of the next.
// ---------------------------------------------------------------------------------------------------------
// timing of the persistent kernel's phases (MDB200_DEBUG_FORCE prints them): clocks summed over blocks
__device__ unsigned long long g_persist_clk[8];

constexpr int kPersistWalkers = 24;                       // walker warps (72 registers)
constexpr int kPersistWalkersDerive = 24;                 // ... of a deriving launch (same budget: 8-16 bytes spilled)
constexpr int kPersistProducers = 3;                      // producer warps: producer j sets up the tiles seq = j, j + 3, ...
constexpr int kPersistThreads = 32 * (kPersistWalkers + kPersistProducers);
constexpr int kPersistThreadsDerive = 32 * (kPersistWalkersDerive + kPersistProducers);
struct PersistShared {
  int staged;        // tiles of this block published so far (st.release / ld.acquire)
  int ticket;        // next row ticket
  int my_tiles;      // tiles of this block
  int pad0;
  int tile_done[4];  // ring: seq + 1 once every row of tile seq is done (its buffer and its TileShared slot are free)
  int row_begin[4];  // ring over the tile sequence number: first ticket of the tile, one past its last
  int row_end[4];
  unsigned long long v2max_bits;   // max |p|^2 over the block's locals
  unsigned long long mbar[2];      // mbarrier of the bulk copies into buffer 0 / 1
  unsigned long long clk[4];       // producer: set-up, wait for the buffer, copies; walkers: wait for a tile
  // what the producer needs (a copy of the kernel's arguments: persist_producer is not inlined, so that its registers
  // and the walk's are allocated separately)
  TileArgs ta;
  TileDims dims;
  const double4 *q;
  const uint2 *list;
  const long long *fw_off;
  const int *fw_len;
};
constexpr size_t kPersistShared = 256;
static_assert(sizeof(PersistShared) <= kPersistShared, "PersistShared outgrew its slot");
__host__ __device__ inline size_t persist_head_bytes() { return 3 * tile_shared_offset() + kPersistShared; }
static size_t persist_smem_bytes(int cap_records) { return persist_head_bytes() + 2 * (size_t)(cap_records + 1) * 32; }

__device__ __forceinline__ int ld_acquire_shared(const int *p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(int *p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// slot of the block's tile number seq in the ring of three TileShared
__device__ __forceinline__ int persist_slot(int seq) { return seq - 3 * ((seq * 21846) >> 16); }   // seq % 3, seq < 32768

// a producer warp: the tiles seq = first, first + kPersistProducers, ... of the block.  A set-up takes a warp longer than
// the walkers need for a tile, so three of them work on consecutive tiles; they publish in order.
__device__ __noinline__ void persist_producer(unsigned char *smem, int first) {
  PersistShared &ps = *reinterpret_cast<PersistShared *>(smem + 3 * tile_shared_offset());
  const TileArgs &ta = ps.ta;
  const double4 *__restrict__ q = ps.q;
  const long long *__restrict__ fw_off = ps.fw_off;
  const int *__restrict__ fw_len = ps.fw_len;
  const int lane = threadIdx.x & 31;
  const int my_tiles = ps.my_tiles;
  static_assert(kPersistProducers == 3, "producer j owns TileShared slot j");
  TileShared &ts = *reinterpret_cast<TileShared *>(smem + first * tile_shared_offset());
  for (int seq = first; seq < my_tiles; seq += kPersistProducers) {
    double4 *sq = reinterpret_cast<double4 *>(smem + persist_head_bytes() + (seq & 1) * ((size_t)(ta.cap + 1) * 32));
    const long long t0 = clock64();
    // (1) set-up: needs the TileShared slot, which tile seq - 3 occupied
    if (seq >= 3)
      while (ld_acquire_shared(&ps.tile_done[(seq - 3) & 3]) != seq - 2) __nanosleep(32);
    const int ti = (int)blockIdx.x + seq * (int)gridDim.x;
    const int tile = ta.tile_order ? ta.tile_order[ta.tile_base + ti] : ti;
    const int gfw0 = ta.tile_fw0[tile];
    tile_setup_warp(ts, tile, ta.g, ps.dims, ta.cell_start, ta.gcell_start);
    const int ncells = ts.ncells, n_loc = ts.n_loc, n_all = ts.n_all;
    const int nfw = (ts.n_int + kPW - 1) / kPW;
    const long long base = fw_off[gfw0];
    for (int r = lane; r < nfw; r += 32) {
      ts.row_off[r] = (int)(fw_off[gfw0 + r] - base);
      ts.row_len[r] = (unsigned short)fw_len[gfw0 + r];
      ts.row_cell[r] = (unsigned char)tile_interior_cell(ts, r * kPW);
    }
    if (lane == 0) {
      ts.row_base = base;
      ts.gfw0 = gfw0;
      ts.rows_done = 0;
    }
    // (2) the buffer: free once every row of tile seq - 2 is done
    const long long t1 = clock64();
    if (seq >= 2)
      while (ld_acquire_shared(&ps.tile_done[(seq - 2) & 3]) != seq - 1) __nanosleep(32);
    const long long t2 = clock64();
    // the generic-proxy accesses above (and the walkers' reads of the buffer's previous tile) before the bulk copies
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    unsigned long long *bar = &ps.mbar[seq & 1];
    if (lane == 0) mbar_arrive_expect_tx(bar, (unsigned)n_loc * 32u);
    __syncwarp();
    for (int k = lane; k < ncells; k += 32) {
      const int nl = ts.cell[k].n_l;
      if (nl > 0) bulk_copy_g2s(sq + ts.cell[k].off_l, q + ts.cell[k].src_l, (unsigned)nl * 32u, bar);
    }
    // the first rows of the tile towards L2
    if (nfw > 0) {
      const int head = ts.row_off[min(nfw, 2 * kPersistWalkers) - 1];   // words before the last of them
      for (long long b = 128ll * lane; b <= (long long)head * 8; b += 4096)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(ps.list + base) + b));
    }
    // ghosts (boundary tiles): staged index -> cell by binary search over the cells' ghost offsets
    for (int i0 = n_loc; i0 < n_all; i0 += 128) {
      int s[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + 32 * u + lane;
        s[u] = -1;
        if (i < n_all) {
          int k = 0;
#pragma unroll
          for (int st = 64; st > 0; st >>= 1)
            if (k + st < ncells && ts.cell[k + st].off_g <= i) k += st;
          s[u] = ta.gitems[ts.cell[k].src_g + (i - ts.cell[k].off_g)];
        }
      }
      double4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (s[u] >= 0) v[u] = ld_q256(q + s[u]);
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (s[u] >= 0) sq[i0 + 32 * u + lane] = v[u];
    }
    if (lane == 0) sq[n_all] = make_double4(1.0e8, 1.0e8, 1.0e8, 1.0e8);
    mbar_wait(bar, (unsigned)(seq >> 1) & 1u);   // the buffer's mbarrier completes one phase per tile
    const long long t3 = clock64();
    // publish, in order
    while (ld_acquire_shared(&ps.staged) != seq) __nanosleep(32);
    __syncwarp();
    if (lane == 0) {
      const int begin = seq ? ps.row_end[(seq - 1) & 3] : 0;
      ps.row_begin[seq & 3] = begin;
      ps.row_end[seq & 3] = begin + nfw;
      if (nfw == 0) ps.tile_done[seq & 3] = seq + 1;   // nobody walks it
      st_release_shared(&ps.staged, seq + 1);
      atomicAdd(&ps.clk[0], (unsigned long long)(t1 - t0));
      atomicAdd(&ps.clk[1], (unsigned long long)(t2 - t1));
      atomicAdd(&ps.clk[2], (unsigned long long)(t3 - t2));
    }
    __syncwarp();
  }
}

template <int MODE, bool EXACT, bool DERIVE>
__global__ void __launch_bounds__(DERIVE ? kPersistThreadsDerive : kPersistThreads, 1)
k_tile_force_p(TileArgs ta, Consts c, const double4 *__restrict__ q, double4 *__restrict__ p, double *__restrict__ fout,
               const uint2 *__restrict__ list, const long long *__restrict__ fw_off, const int *__restrict__ fw_len,
               const int *__restrict__ ref_of_slot, double4 *__restrict__ q_out, Ctrl *ctrl, ForceConsts fc, int ntiles,
               uint2 *__restrict__ list_out, int *__restrict__ len_out) {
  constexpr int kWalkers = DERIVE ? kPersistWalkersDerive : kPersistWalkers;
  extern __shared__ __align__(128) unsigned char smem[];
  PersistShared &ps = *reinterpret_cast<PersistShared *>(smem + 3 * tile_shared_offset());
  const int lane = threadIdx.x & 31;
#define TS_OF(SEQ) (*reinterpret_cast<TileShared *>(smem + persist_slot(SEQ) * tile_shared_offset()))
#define SQ_OF(SEQ) (reinterpret_cast<double4 *>(smem + persist_head_bytes() + ((SEQ) & 1) * ((size_t)(ta.cap + 1) * 32)))
  if (threadIdx.x == 0) {
    ps.staged = 0;
    ps.ticket = 0;
    ps.my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles blockIdx.x + k gridDim.x
    ps.v2max_bits = 0ull;
    ps.clk[0] = ps.clk[1] = ps.clk[2] = ps.clk[3] = 0ull;
    ps.tile_done[0] = ps.tile_done[1] = ps.tile_done[2] = ps.tile_done[3] = 0;
    ps.ta = ta;
#pragma unroll
    for (int d = 0; d < 3; d++) { ps.dims.m[d] = c.m[d]; ps.dims.ms[d] = c.ms[d]; }
    ps.q = q;
    ps.list = list;
    ps.fw_off = fw_off;
    ps.fw_len = fw_len;
    mbar_init(&ps.mbar[0], 1);
    mbar_init(&ps.mbar[1], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x >= 32 * kWalkers) {
    persist_producer(smem, (int)(threadIdx.x >> 5) - kWalkers);
  } else {
    int cur = 0;
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&ps.ticket, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    RowStream rs;
    rs.lp = nullptr;   // not open at any row
    rs.wa = rs.wb = make_uint2(0u, 0u);
    const int my_tiles = ps.my_tiles;
    int staged = 0;   // tiles known to be published
    while (cur < my_tiles) {
      // the tile the ticket belongs to
      if (staged <= cur) {
        staged = ld_acquire_shared(&ps.staged);
        if (staged <= cur) {
          const long long w0 = clock64();
          while ((staged = ld_acquire_shared(&ps.staged)) <= cur) __nanosleep(32);
          if (lane == 0) atomicAdd(&ps.clk[3], (unsigned long long)(clock64() - w0));
        }
      }
      if (ticket >= ps.row_end[cur & 3]) {
        cur++;
        continue;
      }
      // the next ticket now: its row's first chunks are requested before this row ends
      int tnext = 0;
      if (lane == 0) tnext = atomicAdd(&ps.ticket, 1);
      tnext = __shfl_sync(0xffffffffu, tnext, 0);
      double f0 = 0.0, f1 = 0.0, f2 = 0.0;
      int slot = 0;
      DeriveLane dl;
      {
        const TileShared &ts = TS_OF(cur);
        const int begin = ps.row_begin[cur & 3], nrows = ps.row_end[cur & 3] - begin;
        const int fw = ticket - begin;
        const uint2 *tl = list + ts.row_base;
        // (the buffers hold the words at rs.lp and rs.lp + 32: the stream is open at this row if the walk before left it here)
        if (rs.lp != tl + ts.row_off[fw] + lane) row_stream_open(rs, tl + ts.row_off[fw] + lane);
        const unsigned gp = (lane >> 2) & 1u;                                // odd groups: x and z exchanged
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(SQ_OF(cur));
        const unsigned hx = sbase + (gp << 4), hx2 = sbase + ((gp ^ 1u) << 4);
        const int a = fw * kPW + (lane >> 2);
        const unsigned ao = (unsigned)(a < ts.n_int ? a : ts.n_all) << 5;
        double a0, a1;
        lds128(ao + hx, a0, a1);
        const double a2 = lds64(ao + hx2);
        // the writer lane's momentum is needed at the end of the row: pull its line towards the SM
        if (a < ts.n_int && (lane & 3) == 0) {
          slot = tile_interior_slot_from(ts, a, ts.row_cell[fw]);
          if (MODE != FORCE_IMPULSE) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + slot));
        }
        const bool more = tnext - begin < nrows;
        const uint2 *__restrict__ nlp = more ? tl + ts.row_off[tnext - begin] + lane : rs.lp;
        if (DERIVE) {
          dl.acc = 0ull;
          dl.nk = 0;
          dl.on = a < ts.n_int;
          dl.dst = reinterpret_cast<unsigned long long *>(list_out + ts.row_base + ts.row_off[fw]) + lane;
        }
        walk_row<EXACT, DERIVE>(rs, nlp, (int)ts.row_len[fw], hx, hx2, a0, a1, a2, fc, f0, f1, f2, dl);
      }
      // (nothing read from shared memory above is to be kept in a register across the walk: it is read again below)
      asm volatile("" ::: "memory");
      {
        const TileShared &ts = TS_OF(cur);
        if (DERIVE) derive_finish_row(dl, ts.n_all, lane, len_out + ts.gfw0 + (ticket - ps.row_begin[cur & 3]));
        const int a = (ticket - ps.row_begin[cur & 3]) * kPW + (lane >> 2);
        const bool writer = a < ts.n_int && (lane & 3) == 0;
        double4 pa = make_double4(0.0, 0.0, 0.0, 0.0);
        if (MODE != FORCE_IMPULSE && writer) pa = ld_256(p + slot);
        double v2 = 0.0;
        finish_row<MODE>(f0, f1, f2, writer, (lane >> 2) & 1u, a, slot, pa, SQ_OF(cur), p, fout, ref_of_slot, q_out, fc, v2);
        if (MODE == FORCE_KICK_DRIFT || MODE == FORCE_KICK_DRIFT2 || MODE == FORCE_KICK_LANGEVIN) {
          // non-negative doubles order like their bit patterns
          const unsigned long long b = (unsigned long long)__double_as_longlong(v2);
          if (writer && b > *(volatile unsigned long long *)&ps.v2max_bits) atomicMax(&ps.v2max_bits, b);
        }
        // This warp has read everything it needs from the tile's buffer (its loads have returned: their values were
        // used above), so no fence: a fence would also wait for the row's global stores.
        __syncwarp();
        if (lane == 0) {
          const int nrows = ps.row_end[cur & 3] - ps.row_begin[cur & 3];
          if (atomicAdd(const_cast<int *>(&ts.rows_done), 1) == nrows - 1) st_release_shared(&ps.tile_done[cur & 3], cur + 1);
        }
      }
      ticket = tnext;
    }
  }
#undef TS_OF
#undef SQ_OF
  __syncthreads();
  if (threadIdx.x == 0) {
    if (MODE == FORCE_KICK_DRIFT || MODE == FORCE_KICK_DRIFT2 || MODE == FORCE_KICK_LANGEVIN) {
      // the maximum settles after a few blocks, so the plain read filters out almost every atomic
      const unsigned long long b = ps.v2max_bits;
      if (b > *(volatile unsigned long long *)&ctrl->vmax2_bits) atomicMax(&ctrl->vmax2_bits, b);
    }
    for (int i = 0; i < 4; i++) atomicAdd(&g_persist_clk[i], ps.clk[i]);
    atomicAdd(&g_persist_clk[4], (unsigned long long)ps.my_tiles);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Inner list.  ~25 % of the listed pairs lie between the cutoff and the search length and contribute exactly
// zero.  The force kernel walks the INNER list: the entries of the outer list with r2 <= (cutoff + delta)^2 when it
// was derived, at the same row offsets, with their own step counts in fw_len_in.  Until some pair can have closed
// the distance delta (velocity budget kept by the host, api.cu) the force launches give identical contributions:
// the dropped entries are exact zeros.
//
// Two steps.  (1) Once per rebuild k_redeal puts every row of the OUTER list (written by the search in arrival
// order) into "rainbow" order in place (tile.cuh): lane c of a local's group then holds the partners of class c
// (staged index mod 4), which is what makes the walk's shared-memory reads conflict-free.  Pure index shuffling:
// no positions, no tile staging, one warp per row through a row block in shared memory.  (2) Every ~16 steps
// k_tile_compact filters the re-dealt outer list lane by lane (below).
// ---------------------------------------------------------------------------------------------------------
#define REDEAL_TPB 256
template <int DUMMY>
__global__ void __launch_bounds__(REDEAL_TPB)
k_redeal(const long long *__restrict__ fw_off, int *__restrict__ fw_len, const int *__restrict__ fw_sent, uint2 *__restrict__ list,
         long long nfw, int max_chunks) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint2 *rb = reinterpret_cast<uint2 *>(smem) + (size_t)wid * max_chunks * 32;  // row block of this warp
  unsigned short *rb16 = reinterpret_cast<unsigned short *>(rb);
  const int c4 = lane & 3, glane0 = lane & ~3;
  constexpr int NB = 8;  // chunks held in registers (rows of up to 32 steps; longer rows take another trip)
  const long long nw = (long long)gridDim.x * (REDEAL_TPB / 32);
  for (long long gfw = (long long)blockIdx.x * (REDEAL_TPB / 32) + wid; gfw < nfw; gfw += nw) {
    const int steps = fw_len[gfw];
    const int nchunk = (steps + 3) >> 2;
    const unsigned sent = (unsigned)fw_sent[gfw];
    const uint2 sent2 = make_uint2(sent | (sent << 16), sent | (sent << 16));
    uint2 *__restrict__ lp = list + fw_off[gfw] + lane;
    // ---- pass A: the real entries of this lane (bit 4 c + s of the mask: step 4 c + s), how many per class
    unsigned long long keep = 0ull;   // rows of at most 16 chunks = 64 steps = 256 entries (derive_inner checks)
    unsigned kcnt = 0u;
    uint2 w[NB];
#pragma unroll 1
    for (int cb = 0; cb < nchunk; cb += NB) {
#pragma unroll
      for (int i = 0; i < NB; i++) w[i] = (cb + i < nchunk) ? lp[32 * (cb + i)] : sent2;  // all in flight together
#pragma unroll
      for (int i = 0; i < NB; i++) {
        const int ch = cb + i;
        if (ch < nchunk) {
          unsigned m = 0u;
#pragma unroll
          for (int s2 = 0; s2 < 4; s2++) {
            const unsigned j = ((s2 < 2 ? w[i].x : w[i].y) >> (16 * (s2 & 1))) & 0xffffu;
            if (j != sent && 4 * ch + s2 < steps) { m |= 1u << s2; kcnt += 1u << (8 * (j & 3u)); }
          }
          keep |= (unsigned long long)m << (4 * ch);
        }
      }
    }
    // ---- ordinals: inclusive scan of the packed class counts over the 4 lanes of the group
    unsigned incl = kcnt;
    unsigned t = __shfl_up_sync(0xffffffffu, incl, 1, kG);
    if (c4 >= 1) incl += t;
    t = __shfl_up_sync(0xffffffffu, incl, 2, kG);
    if (c4 >= 2) incl += t;
    const unsigned tot = __shfl_sync(0xffffffffu, incl, kG - 1, kG);  // class totals of the local
    unsigned run = incl - kcnt;                                        // first ordinals of this lane's entries
    const int nk = (int)((tot & 0xFFu) + ((tot >> 8) & 0xFFu) + ((tot >> 16) & 0xFFu) + (tot >> 24));
    int steps_out = (nk + kG - 1) / kG;
#pragma unroll
    for (int o = kG; o < 32; o <<= 1) steps_out = max(steps_out, __shfl_xor_sync(0xffffffffu, steps_out, o));
    steps_out = (steps_out + 1) & ~1;
    const int nchunk_out = (steps_out + 3) >> 2;
    // ---- row block <- sentinel, then the entries at their rainbow positions
    for (int i = lane; i < nchunk_out * 32; i += 32) rb[i] = sent2;
    __syncwarp();
    const int m = rainbow_min(tot);
#pragma unroll 1
    for (int cb = 0; cb < nchunk; cb += NB) {
      if (nchunk > NB) {  // long row: the words of this trip again (the short ones are still in registers)
#pragma unroll
        for (int i = 0; i < NB; i++) w[i] = (cb + i < nchunk) ? lp[32 * (cb + i)] : sent2;
      }
#pragma unroll
      for (int i = 0; i < NB; i++) {
        const int ch = cb + i;
        const unsigned km = (ch < nchunk) ? (unsigned)((keep >> (4 * ch)) & 0xFull) : 0u;
#pragma unroll
        for (int s2 = 0; s2 < 4; s2++) {
          if (km & (1u << s2)) {
            const unsigned j = ((s2 < 2 ? w[i].x : w[i].y) >> (16 * (s2 & 1))) & 0xffffu;
            const int cc = (int)(j & 3u);
            const int tt = (int)((run >> (8 * cc)) & 0xFFu);
            run += 1u << (8 * cc);
            const int ee = rainbow_place(tot, m, cc, tt);  // entry index of the local: lane ee & 3, step ee >> 2
            const int st = ee >> 2;
            rb16[((st >> 2) * 32 + glane0 + (ee & 3)) * 4 + (st & 3)] = (unsigned short)j;
          }
        }
      }
    }
    __syncwarp();
    // written back in place: every word of the row was read above
    for (int i = lane; i < nchunk_out * 32; i += 32) lp[i - lane] = rb[i];
    if (lane == 0) fw_len[gfw] = steps_out;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// The periodic derivation: per-lane compaction of the RE-DEALT outer list.  Once per rebuild k_redeal re-deals
// the outer list in place into rainbow order (keep = everything); after that a lane's column already holds the
// partners of "its" class, so deriving the inner list is a filter that needs no exchange between lanes: every lane
// walks its own column, keeps the entries with r2 <= keep2 and packs them into output words held in registers; the
// row block is written out at the end, whole chunks per warp.  Columns end up with slightly different lengths (the
// row is as long as its longest column: ~1.5 steps in 22 more than a perfect re-deal) -- the price of a derivation
// that costs a third of a force walk instead of three.
// ---------------------------------------------------------------------------------------------------------
template <int DUMMY>
__global__ void __launch_bounds__(FORCE_TPB, 2)
k_tile_compact(TileArgs ta, Consts c, const double4 *__restrict__ q, const uint2 *__restrict__ list,
               const long long *__restrict__ fw_off, const int *__restrict__ fw_len, uint2 *__restrict__ list_in,
               int *__restrict__ fw_len_in, double keep2) {
  extern __shared__ __align__(128) unsigned char smem[];
  TileShared &ts = *reinterpret_cast<TileShared *>(smem);
  double4 *sq = reinterpret_cast<double4 *>(smem + tile_shared_offset());
  const int tile = ta.tile_order ? ta.tile_order[ta.tile_base + blockIdx.x] : (int)blockIdx.x;
  tile_setup(ts, tile, ta.g, c, ta.cell_start, ta.gcell_start);
  const int gfw0 = ta.tile_fw0[tile];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n_int = ts.n_int;
  const int nfw = (n_int + kPW - 1) / kPW;
  const int rbase = nfw / nwarps, rrem = nfw - rbase * nwarps;
  const int r0 = wid * rbase + min(wid, rrem), r1 = r0 + rbase + (wid < rrem ? 1 : 0);
  tile_stage_rows(ts, gfw0, fw_off, fw_len);
  tile_stage_tma(ts, q, ta.gitems, sq);
  tile_stage_wait(ts);
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sq);
  const unsigned gp = (lane >> 2) & 1u;  // same lane-group parity as the force walk: conflict-free on rainbow rows
  const unsigned hx = sbase + (gp << 4), hx2 = sbase + ((gp ^ 1u) << 4);
  const unsigned long long sent = (unsigned long long)ts.n_all;
  const unsigned long long sent4 = sent | (sent << 16) | (sent << 32) | (sent << 48);
  for (int fw = r0; fw < r1; fw++) {
    const int a = fw * kPW + (lane >> 2);
    const bool act = a < n_int;
    const unsigned ao = (unsigned)(act ? a : ts.n_all) << 5;
    double a0, a1;
    lds128(ao + hx, a0, a1);
    const double a2 = lds64(ao + hx2);
    const int steps = ts.row_len[fw];
    const int nchunk = (steps + 3) >> 2;
    const uint2 *__restrict__ lp = list + ts.row_base + ts.row_off[fw] + lane;
    unsigned long long *__restrict__ dst = reinterpret_cast<unsigned long long *>(list_in + ts.row_base + ts.row_off[fw]) + lane;
    if (fw + 1 < r1) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(list + ts.row_base + ts.row_off[fw + 1]) + 128 * lane));
    unsigned long long acc = 0ull;   // the output word being filled
    int nk = 0;                      // kept entries of this lane
    // A finished word goes straight to its place in the row block: the lanes of a warp fill their words at nearly the
    // same pace, so these 8-byte stores mostly fall into the same one or two chunks, and L2 merges what is left.
#define COMPACT_CHUNK(W, CH)                                                                     \
  do {                                                                                           \
    _Pragma("unroll") for (int s2 = 0; s2 < 4; s2++) {                                           \
      const unsigned j = ((s2 < 2 ? (W).x : (W).y) >> (16 * (s2 & 1))) & 0xffffu;                \
      double j0, j1;                                                                             \
      lds128((j << 5) + hx, j0, j1);                                                             \
      const double j2 = lds64((j << 5) + hx2);                                                   \
      const double d0 = j0 - a0, d1 = j1 - a1, d2 = j2 - a2;                                     \
      const double r2 = fma(d2, d2, fma(d1, d1, d0 * d0));                                       \
      if (act && r2 <= keep2 && 4 * (CH) + s2 < steps) {                                         \
        acc |= (unsigned long long)j << (16 * (nk & 3));                                         \
        nk++;                                                                                    \
        if ((nk & 3) == 0) {                                                                     \
          dst[32 * ((nk >> 2) - 1)] = acc;                                                       \
          acc = 0ull;                                                                            \
        }                                                                                        \
      }                                                                                          \
    }                                                                                            \
  } while (0)
    // two named buffers reloaded in place, two chunks ahead (as in the force walk)
    uint2 wa = __ldg(lp), wb = __ldg(lp + 32);
#pragma unroll 1
    for (int ch = 0; ch < nchunk; ch += 2) {
      {
        const uint2 w = wa;
        wa = __ldg(lp + 32 * (ch + 2));
        COMPACT_CHUNK(w, ch);
      }
      if (ch + 1 < nchunk) {
        const uint2 w = wb;
        wb = __ldg(lp + 32 * (ch + 3));
        COMPACT_CHUNK(w, ch + 1);
      }
    }
#undef COMPACT_CHUNK
    int steps_in = nk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) steps_in = max(steps_in, __shfl_xor_sync(0xffffffffu, steps_in, o));
    steps_in = (steps_in + 1) & ~1;
    const int nchunk_in = (steps_in + 3) >> 2;
    // the word in progress, topped up with the sentinel, and sentinel words up to the row's length
    int wi = nk >> 2;
    if (nk & 3) {
      dst[32 * wi] = acc | (sent4 & (~0ull << (16 * (nk & 3))));
      wi++;
    }
    for (; wi < nchunk_in; wi++) dst[32 * wi] = sent4;
    if (lane == 0) fw_len_in[gfw0 + fw] = steps_in;
  }
}

static ForceConsts make_fc(const Unit *u) {
  ForceConsts fc;
  fc.CL2 = u->c.CL2;
  fc.dt2 = u->c.dt2;
  fc.c24dt = 24.0 * u->c.dt;
  fc.c48dt = 48.0 * u->c.dt;
  fc.c2x8dt = u->c.C2 * 8.0 * u->c.dt;
  long long b;
  memcpy(&b, &fc.CL2, sizeof(b));
  fc.cl2_hi = (int)(b >> 32);
  fc.exact_cut = (unsigned)(b & 0xffffffffll) != 0u;
  // Langevin constants as in integrate.cu:langevin_kick
  fc.dt = u->c.dt;
  fc.gamma = u->prm.heatbath_gamma;
  fc.amp = sqrt(2.0 * u->prm.heatbath_gamma * u->prm.aimed_temperature / u->c.dt);  // hb_D, fcalculator.cc:771
  const unsigned long long sc = (unsigned long long)u->step_count;
  fc.key0 = (unsigned)(u->seed + u->prm.unit_id);  // MDUnit::SetSeed: seed + id (mdunit.h:58)
  fc.key1 = 0x6d646163u;                            // "mdac"
  fc.step_lo = (unsigned)(sc & 0xffffffffu);
  fc.step_hi = (unsigned)(sc >> 32);
  return fc;
}

static TileArgs make_ta(const Unit *u, int part) {
  TileArgs ta;
  ta.g = tile_grid(u);
  ta.cell_start = u->cell_start.ptr;
  ta.gcell_start = u->gcell_start.ptr;
  ta.gitems = u->gcell_items.ptr;
  ta.tile_fw0 = u->tile_fw0.ptr;
  ta.tile_order = u->tile_order.ptr;  // active tiles only (tile_search.cu)
  ta.tile_base = (part == 2) ? u->n_interior_tiles : 0;
  ta.cap = u->tile_cap;
  return ta;
}

// opt-in dynamic shared memory of a kernel, raised when a launch needs more than any before it (per device)
template <typename K>
static int set_smem(K kernel, size_t bytes) {
  static std::map<std::pair<const void *, int>, size_t> granted;
  static std::mutex mtx;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mtx);
  size_t &g = granted[std::make_pair(reinterpret_cast<const void *>(kernel), dev)];
  if (bytes <= g) return MDB200_OK;
  MDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  g = bytes;
  return MDB200_OK;
}

// k_redeal launch: the outer list into rainbow order, in place
static int launch_redeal(Unit *u) {
  const int max_steps = (((u->max_row + kG - 1) / kG) + 1) & ~1;
  const int max_chunks = (max_steps + 3) >> 2;
  if (max_chunks > 16) { set_error("rows of %d entries are too long for the inner-list derivation", u->max_row); return MDB200_ECAPACITY; }
  const size_t sm = (size_t)(REDEAL_TPB / 32) * max_chunks * 256;
  MDB_TRY(set_smem(k_redeal<0>, sm));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, u->device);
  const long long want = (u->nfw + REDEAL_TPB / 32 - 1) / (REDEAL_TPB / 32);
  const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms * 64);
  k_redeal<0><<<grid, REDEAL_TPB, sm, u->stream>>>(u->fw_off.ptr, u->fw_len.ptr, u->fw_sent.ptr, u->list.ptr, u->nfw, max_chunks);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

// the outer list into rainbow order, once per rebuild (no-op afterwards)
int ensure_rainbow(Unit *u) {
  if (u->outer_rainbow || u->pn == 0 || u->ntiles == 0) return MDB200_OK;
  MDB_TRY(launch_redeal(u));
  u->outer_rainbow = true;
  return MDB200_OK;
}

// Inner list = the outer entries within keep_len of the CURRENT positions.  The first call after a rebuild re-deals the
// outer list in place into rainbow order (k_redeal); every call
// then filters it lane by lane (k_tile_compact).  keep_len >= search length: the inner list IS the outer list.
int derive_inner(Unit *u, double keep_len) {
  u->inner_valid = false;
  u->inner_is_outer = false;
  if (u->pn == 0 || u->ntiles == 0) { u->inner_valid = true; u->inner_keep = keep_len; return MDB200_OK; }
  static const bool dbg = getenv("MDB200_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "[mdb200] derive: keep %.4f, longest row <= %d entries, %s%s\n", keep_len, u->max_row,
                   u->outer_rainbow ? "" : "re-deal + ", keep_len >= u->c.SL ? "walk the outer list" : "compaction");
  MDB_TRY(ensure_rainbow(u));
  if (keep_len >= u->c.SL) {
    u->inner_is_outer = true;
  } else {
    MDB_TRY(u->list_in.ensure(u->list.cap));
    const size_t sm = tile_smem_bytes_f64(u->tile_cap);
    MDB_TRY(set_smem(k_tile_compact<0>, sm));
    k_tile_compact<0><<<u->n_active_tiles, FORCE_TPB, sm, u->stream>>>(make_ta(u, 0), u->c, u->q.ptr, u->list.ptr, u->fw_off.ptr, u->fw_len.ptr,
                                                               u->list_in.ptr, u->fw_len_in.ptr, keep_len * keep_len);
    u->launches++;
    MDB_CUDA(cudaGetLastError());
  }
  u->inner_valid = true;
  u->inner_keep = keep_len;
  return MDB200_OK;
}

// whether a launch over ntl tiles of this unit takes the persistent kernel (one block per SM, two staged tiles): where
// two copies of the largest tile fit the SM's shared memory and a tile has rows enough to keep the walker warps busy;
// one block per tile otherwise (MDB200_FORCE_PERSIST=0/1 overrides the second condition; read at every launch so that
// tests can switch it)
static bool force_persistent(const Unit *u, int ntl, int *sms_out, size_t *psm_out) {
  const char *persist_env = getenv("MDB200_FORCE_PERSIST");
  int sms = 148, smem_max = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, u->device);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, u->device);
  const size_t psm = persist_smem_bytes(u->tile_cap);
  bool persist = psm <= (size_t)smem_max && u->tile_maxint >= kPersistWalkers * kPW * 2 && ntl >= 2 * sms;
  if (persist_env) persist = psm <= (size_t)smem_max && atoi(persist_env) != 0;
  if (sms_out) *sms_out = sms;
  if (psm_out) *psm_out = psm;
  return persist;
}

// A force launch that also derives the inner list for keep_len (k_tile_force_p<.., DERIVE>): possible where every launch
// of the step -- the whole unit, or its interior and its boundary tiles (split) -- takes the persistent kernel and the
// outer list is already in rainbow order (MDB200_FUSED_DERIVE=0: never).
bool force_can_derive(const Unit *u, double keep_len, bool split) {
  const char *env = getenv("MDB200_FUSED_DERIVE");  // read at every call so that tests can switch it
  const bool off = env && atoi(env) == 0;
  if (off || u->pn == 0 || u->ntiles == 0 || !u->outer_rainbow || keep_len >= u->c.SL) return false;
  if (!split) return force_persistent(u, u->n_active_tiles, nullptr, nullptr);
  const int n_in = u->n_interior_tiles, n_bd = u->n_active_tiles - u->n_interior_tiles;
  // (an empty part launches nothing)
  return (n_in == 0 || force_persistent(u, n_in, nullptr, nullptr)) && (n_bd == 0 || force_persistent(u, n_bd, nullptr, nullptr));
}

// part: 0 = every tile, 1 = interior tiles only, 2 = boundary tiles only
// derive_keep > 0 (drifting modes, persistent kernel only: force_can_derive): walk the outer list and write the inner
// list for that radius on the side; with split launches BOTH parts must be given the same derive_keep
static int force_launch(Unit *u, int mode, double *fout, int part = 0, double derive_keep = 0.0) {
  const int ntl = (part == 0) ? u->n_active_tiles : (part == 1 ? u->n_interior_tiles : u->n_active_tiles - u->n_interior_tiles);
  const bool deriving = derive_keep > 0.0;
  if (ntl <= 0) {
    if (deriving && part != 1) { u->inner_valid = true; u->inner_is_outer = false; u->inner_keep = derive_keep; }
    return MDB200_OK;
  }
  if (!deriving && !u->inner_valid) MDB_TRY(derive_inner(u, u->c.SL));  // no budgeted inner list: a full one
  ForceConsts fc = make_fc(u);
  fc.keep2 = derive_keep * derive_keep;
  const TileArgs ta = make_ta(u, part);
  const bool outer = deriving || u->inner_is_outer;
  const uint2 *walk_list = outer ? u->list.ptr : u->list_in.ptr;
  const int *walk_len = outer ? u->fw_len.ptr : u->fw_len_in.ptr;
  int sms = 148;
  size_t psm = 0;
  const bool persist = force_persistent(u, ntl, &sms, &psm);
  const size_t sm = tile_smem_bytes_f64(u->tile_cap);
  if (deriving) {
    if (!persist || mode < FORCE_KICK_DRIFT) { set_error("deriving force launch not available here"); return MDB200_EINVAL; }
    MDB_TRY(u->list_in.ensure(u->list.cap));
  }
#define PLAUNCH(M, E, D)                                                                                                    \
  do {                                                                                                                      \
    MDB_TRY(set_smem(k_tile_force_p<M, E, D>, psm));                                                                        \
    k_tile_force_p<M, E, D><<<std::min(ntl, sms), D ? kPersistThreadsDerive : kPersistThreads, psm, u->stream>>>(           \
        ta, u->c, u->q.ptr, u->p.ptr, fout, walk_list, u->fw_off.ptr, walk_len, u->ref_of_slot.ptr, u->q_alt.ptr, u->ctrl, fc, \
        ntl, u->list_in.ptr, u->fw_len_in.ptr);                                                                             \
  } while (0)
#define KLAUNCH(M, E)                                                                                                       \
  do {                                                                                                                      \
    if (persist) {                                                                                                          \
      PLAUNCH(M, E, false);                                                                                                 \
    } else {                                                                                                                \
      MDB_TRY(set_smem(k_tile_force<M, E>, sm));                                                                            \
      k_tile_force<M, E><<<ntl, FORCE_TPB, sm, u->stream>>>(ta, u->c, u->q.ptr, u->p.ptr, fout, walk_list, u->fw_off.ptr,   \
                                                            walk_len, u->ref_of_slot.ptr, u->q_alt.ptr, u->ctrl, fc);       \
    }                                                                                                                       \
  } while (0)
#define LAUNCH(E)                                                           \
  do {                                                                      \
    if (deriving) {                                                         \
      if (mode == FORCE_KICK_DRIFT) PLAUNCH(FORCE_KICK_DRIFT, E, true);     \
      else if (mode == FORCE_KICK_LANGEVIN) PLAUNCH(FORCE_KICK_LANGEVIN, E, true); \
      else PLAUNCH(FORCE_KICK_DRIFT2, E, true);                             \
    }                                                                       \
    else if (mode == FORCE_IMPULSE) KLAUNCH(FORCE_IMPULSE, E);              \
    else if (mode == FORCE_KICK) KLAUNCH(FORCE_KICK, E);                    \
    else if (mode == FORCE_KICK_DRIFT) KLAUNCH(FORCE_KICK_DRIFT, E);        \
    else if (mode == FORCE_KICK_LANGEVIN) KLAUNCH(FORCE_KICK_LANGEVIN, E);  \
    else KLAUNCH(FORCE_KICK_DRIFT2, E);                                     \
  } while (0)
  if (fc.exact_cut) LAUNCH(true);
  else LAUNCH(false);
#undef LAUNCH
#undef KLAUNCH
#undef PLAUNCH
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  if (deriving && part != 1) {  // (after the interior part the boundary rows are still to come)
    u->inner_valid = true;
    u->inner_is_outer = false;
    u->inner_keep = derive_keep;
  }
  return MDB200_OK;
}

// ForceCalculator::CalculateForce (fcalculator.cc:35-50)
int force_kick(Unit *u) {
  u->vmax_ready = false;
  if (u->pn == 0 || u->ntiles == 0) return MDB200_OK;
  return force_launch(u, FORCE_KICK, nullptr);
}

// CalculateForce + UpdatePositionHalf of MDManager::CalculateForce (mdmanager.cc:151-157) in one kernel;
// predrift = also the first half drift of the next step (MODE 3; the caller sets u->predrifted).
// part 0 = every tile; 1 = interior tiles only (no ghost staged: may run before the ghost exchange has
// finished); 2 = boundary tiles.  The buffers swap after part 0 or 2.
int force_kick_drift(Unit *u, bool predrift, int part, double derive_keep) {
  if (u->pn == 0 || u->ntiles == 0) return MDB200_OK;
  static const bool dbg = getenv("MDB200_DEBUG_FORCE") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, u->stream); }
  MDB_TRY(force_launch(u, predrift ? FORCE_KICK_DRIFT2 : FORCE_KICK_DRIFT, nullptr, part, derive_keep));
  if (dbg) {
    cudaEventRecord(e1, u->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[mdb200] force launch: part %d keep %.4f%s  %.3f ms\n", part, u->inner_keep, derive_keep > 0.0 ? " (deriving)" : "", ms);
    unsigned long long h[8] = {0};
    if (cudaMemcpyFromSymbol(h, g_persist_clk, sizeof(h)) == cudaSuccess && h[4]) {
      fprintf(stderr, "[mdb200]   persistent: per tile: set-up %.0f clk, wait for the buffer %.0f, copies %.0f; walkers waited %.0f warp-clk per tile (%llu tiles)\n",
              (double)h[0] / h[4], (double)h[1] / h[4], (double)h[2] / h[4], (double)h[3] / h[4], h[4]);
      memset(h, 0, sizeof(h));
      cudaMemcpyToSymbol(g_persist_clk, h, sizeof(h));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  if (part != 1) {
    std::swap(u->q, u->q_alt);  // ghosts live in both buffers (exchange.cu)
    u->vmax_ready = true;
  }
  return MDB200_OK;
}

// MDManager::CalculateLangevin (mdmanager.cc:186-192) in one kernel: kick, Langevin update of the momenta and --
// speculatively -- the next step's first half drift (the caller sets u->predrifted = 2: the end-of-step positions
// are the untouched input buffer).  Swaps q / q_alt.
int force_kick_langevin(Unit *u, double derive_keep) {
  if (u->pn == 0 || u->ntiles == 0) return MDB200_OK;
  MDB_TRY(force_launch(u, FORCE_KICK_LANGEVIN, nullptr, 0, derive_keep));
  std::swap(u->q, u->q_alt);  // ghosts live in both buffers (exchange.cu)
  u->vmax_ready = true;
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
  // parity probe, called at arbitrary points between steps: walk a fresh copy of the WHOLE outer list rather than an
  // inner list whose velocity budget the caller's momenta changes may not be covered by
  MDB_TRY(derive_inner(u, u->c.SL));
  MDB_TRY(force_launch(u, FORCE_IMPULSE, f.ptr));
  u->inner_valid = false;
  MDB_CUDA(cudaMemcpyAsync(f3_host, f.ptr, sizeof(double) * 3 * (size_t)u->pn, cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  f.release();
  return MDB200_OK;
}

}  // namespace mdb
