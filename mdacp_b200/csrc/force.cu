// LJ force kick: ForceCalculator::CalculateForce (fcalculator.cc:35-50); semantic spec
// ForceCalculator::CalculateForceSorted (fcalculator.cc:182-223):
//   for each listed pair (i,j): d = q_j - q_i, r2 = |d|^2; if r2 > CL2 the pair contributes 0;
//   r6 = r2^3; df = ((24 r6 - 48)/(r6 r6 r2) + 8 C2) dt; p_i += df d; p_j -= df d.
//
// B200 formulation.
//  * FULL neighbour list, no atomics: every local's momentum is owned by one lane group, so the summation
//    order is deterministic and ghost momenta are never written (they are scratch in the reference,
//    fcalculator.cc:273-275 / SURVEY.md Q8).  RED.ADD.F64 for the reaction term of a half list would cost
//    ~3 P atomics per particle at ~1.3 clk/lane/SM -- several times this whole kernel.
//  * kG = 4 LANES PER LOCAL (unit.h): a warp owns 8 consecutive cell-sorted locals; the 4 lanes of a group
//    take 4 consecutive entries of that local's ascending row per step and a 2-stage shuffle reduction
//    combines them at the end.  Why: ncu shows the thread-per-local version bound by the L1 data pipe on the
//    scattered gather of partner positions (~20 distinct 128-byte lines = ~20 wavefronts per warp-gather,
//    FP64 pipe ~30 %; profiles/r01_force_v*_ncu.txt).  Consecutive entries of one row sit in the same or
//    adjacent lines and 8 neighbouring locals share most of their partner lines, which cuts the distinct
//    lines per warp-gather to ~13 (scripts/gather_model.py) at identical FP64 work.
//  * one 256-bit load per partner (LDG.E.256); list words for two steps come from one 8-byte load per lane
//    (a contiguous 256-byte run per warp); locals cell-sorted.
//  * the cutoff test runs on the integer pipe (bit-pattern compare of the non-negative r2), keeping the
//    FP64 pipe for arithmetic.
//
// Roofline: algorithmic HBM bytes per particle = 4 B x P (half-list indices) + 96 B
// (q read, p read + write), 30 flop per listed unordered pair.
#include <stdlib.h>
#include <utility>
#include "unit.h"

namespace mdb {

// tuning knobs (scripts/build_variants.sh)
#ifndef FORCE_UNROLL
#define FORCE_UNROLL 1
#endif
#ifndef FORCE_MINBLK
#define FORCE_MINBLK 10
#endif
#ifndef FORCE_PF_L1
#define FORCE_PF_L1 2
#endif
#ifndef FORCE_PF_CHUNK
#define FORCE_PF_CHUNK 0
#endif
#ifndef FORCE_PERSIST
#define FORCE_PERSIST 0
#endif
#ifndef FORCE_POS_EVICT_LAST
#define FORCE_POS_EVICT_LAST 0
#endif
#ifndef FORCE_LIST_EVICT_FIRST
#define FORCE_LIST_EVICT_FIRST 0
#endif
#ifndef FORCE_TPB
#define FORCE_TPB 128
#endif

// 1 / x: hardware seed (MUFU.RCP64H) + one cubic (Halley-type) refinement y (1 + e + e^2);
// GUARD adds one linear Newton step (2 DFMA) that makes the result independent of the seed quality.
template <bool GUARD>
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  double t = fma(e, e, e);
  y = fma(y, t, y);
  if (GUARD) {
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
  }
  return y;
}

struct ForceConsts {
  double dt2;     // dt / 2
  double CL2;
  double c24dt;   // 24 dt
  double c48dt;   // 48 dt
  double c2x8dt;  // 8 C2 dt
  double dt;
  double c2x8;
};

// VARIANT 0: reference expression with an IEEE division (validation baseline)
// VARIANT 1: s = 1/r2 (seed + cubic + guard step), df = s^4 (24 - 48 s^3) dt + 8 C2 dt
// VARIANT 2: same without the guard step
template <int VARIANT>
__device__ __forceinline__ double pair_df(double r2, const ForceConsts &fc) {
  if (VARIANT == 0) {
    double r6 = r2 * r2 * r2;
    return ((24.0 * r6 - 48.0) / (r6 * r6 * r2) + fc.c2x8) * fc.dt;
  } else {
    double s = fast_rcp<VARIANT == 1>(r2);
    double s2 = s * s;
    double s3 = s2 * s;
    double s4 = s2 * s2;
    double t = fma(-fc.c48dt, s3, fc.c24dt);
    return fma(s4, t, fc.c2x8dt);
  }
}

__device__ __forceinline__ void ld_q256(const double4 *p, double &x, double &y, double &z) {
  double w;
#if FORCE_POS_EVICT_LAST
  asm("ld.global.nc.L1::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
#else
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
#endif
  (void)w;
}


// list words are streamed once: FORCE_LIST_EVICT_FIRST marks their L1 lines as first to go, so that they
// displace fewer position lines
__device__ __forceinline__ uint2 ld_list(const uint2 *p) {
#if FORCE_LIST_EVICT_FIRST
  uint2 v;
  asm("ld.global.nc.L1::evict_first.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

template <int VARIANT>
__device__ __forceinline__ void pair_accumulate(double ax, double ay, double az, double jx, double jy, double jz,
                                                long long cl2b, const ForceConsts &fc, double &fx, double &fy,
                                                double &fz) {
  const double dx = jx - ax, dy = jy - ay, dz = jz - az;
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
  double df = pair_df<VARIANT>(r2, fc);
  // r2 > CL2: the pair contributes nothing (fcalculator.cc:209-211); also kills the sentinel padding
  if (__double_as_longlong(r2) > cl2b) df = 0.0;
  fx = fma(df, dx, fx);
  fy = fma(df, dy, fy);
  fz = fma(df, dz, fz);
}

// MODE 0: impulses to fout (parity probe).  MODE 1: kick, p += F dt (ForceCalculator::CalculateForce).
// MODE 2: kick fused with the second half drift of MDManager::CalculateForce (mdmanager.cc:151-157):
//   q_out[a] = q[a] + p_new dt/2 with the same non-fused mul/add as k_drift_half (bit-identical), written to
//   the OTHER position buffer because q is still being gathered by other warps; the caller swaps the
//   buffers.  It also leaves max |p_new|^2 in ctrl->vmax2_bits for the next step's velocity-budget test
//   (pairlist.cc:48-65), so that test needs no pass of its own.
// MODE 3: as MODE 2 plus the FIRST half drift of the NEXT step (mdmanager.cc:128-131), speculatively:
//   q_out[a] = (q[a] + p_new dt/2) + p_new dt/2, two separately rounded additions exactly like two
//   k_drift_half passes.  The end-of-step positions q[a] + p_new dt/2 are not stored; whoever needs them
//   (rebuild, observation, download, displacement test) recomputes them bit-identically from the untouched
//   input buffer and p_new (api.cu:materialize, integrate.cu).
enum { FORCE_IMPULSE = 0, FORCE_KICK = 1, FORCE_KICK_DRIFT = 2, FORCE_KICK_DRIFT2 = 3 };
#define FORCE_PARAMS                                                                                              \
  const double4 *__restrict__ q, double4 *__restrict__ p, double *__restrict__ fout,                             \
      const unsigned *__restrict__ list, const long long *__restrict__ warp_off, const int *__restrict__ warp_len, \
      const int *__restrict__ ref_of_slot, int pn, long long nwarps, double4 *__restrict__ q_out, Ctrl *ctrl
#define FORCE_PASS q, p, fout, list, warp_off, warp_len, ref_of_slot, pn, nwarps, q_out, ctrl

// One force-warp: the kPW locals in slots [w kPW, (w+1) kPW), kG lanes each.
template <int VARIANT, int MODE>
__device__ __forceinline__ void force_warp(long long w, int lane, FORCE_PARAMS, const ForceConsts &fc) {
  const long long a = w * kPW + lane / kG;
  const bool act = a < pn;
  double ax, ay, az;
  ld_q256(q + (act ? a : (long long)pn - 1), ax, ay, az);
  const int len2 = warp_len[w] >> 1;
  const uint2 *__restrict__ lp = reinterpret_cast<const uint2 *>(list + warp_off[w]) + lane;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  const long long cl2b = __double_as_longlong(fc.CL2);

  constexpr int UNR = FORCE_UNROLL;
#if FORCE_PF_CHUNK
  // one instruction: lane l pulls 128-byte line l of this warp's list chunk (len2 x 256 B) towards the SM
  if (lane * 128 < len2 * 256)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(lp - lane) + lane * 128));
#endif
  uint2 en = (len2 > 0) ? ld_list(lp) : make_uint2(0u, 0u);
#pragma unroll UNR
  for (int k = 0; k < len2; k++) {
    const uint2 e = en;
#if FORCE_PF_L1
    // fire-and-forget: the list words of step-pair k + FORCE_PF_L1 into L1, so that the (compiler-scheduled)
    // load of the next words never waits on L2 / HBM -- ptxas otherwise sinks that load below the arithmetic,
    // which serialises its latency behind the gather's (1.15 -> 1.42 ms at 5.5 M particles)
    if (k + FORCE_PF_L1 < len2) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + (long long)(k + FORCE_PF_L1) * 32));
#endif
    if (k + 1 < len2) en = ld_list(lp + (long long)(k + 1) * 32);
    double j0x, j0y, j0z, j1x, j1y, j1z;
    ld_q256(q + e.x, j0x, j0y, j0z);
    ld_q256(q + e.y, j1x, j1y, j1z);
    pair_accumulate<VARIANT>(ax, ay, az, j0x, j0y, j0z, cl2b, fc, fx, fy, fz);
    pair_accumulate<VARIANT>(ax, ay, az, j1x, j1y, j1z, cl2b, fc, fx, fy, fz);
  }
#pragma unroll
  for (int o = kG / 2; o > 0; o >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
    fz += __shfl_xor_sync(0xffffffffu, fz, o);
  }
  const bool writer = act && (lane % kG) == 0;
  if (MODE == FORCE_IMPULSE) {
    if (writer) {
      const long long ra = ref_of_slot[a];
      fout[3 * ra + 0] = fx; fout[3 * ra + 1] = fy; fout[3 * ra + 2] = fz;
    }
    return;
  }
  double v2 = 0.0;
  if (writer) {
    double4 pa = p[a];
    pa.x += fx; pa.y += fy; pa.z += fz;
    p[a] = pa;
    if (MODE == FORCE_KICK_DRIFT || MODE == FORCE_KICK_DRIFT2) {
      // fcalculator.cc:22-32, no FMA contraction: identical to k_drift_half
      const double sx = __dmul_rn(pa.x, fc.dt2), sy = __dmul_rn(pa.y, fc.dt2), sz = __dmul_rn(pa.z, fc.dt2);
      double nx = __dadd_rn(ax, sx), ny = __dadd_rn(ay, sy), nz = __dadd_rn(az, sz);
      if (MODE == FORCE_KICK_DRIFT2) { nx = __dadd_rn(nx, sx); ny = __dadd_rn(ny, sy); nz = __dadd_rn(nz, sz); }
      q_out[a] = make_double4(nx, ny, nz, 0.0);
      v2 = __dadd_rn(__dadd_rn(__dmul_rn(pa.x, pa.x), __dmul_rn(pa.y, pa.y)), __dmul_rn(pa.z, pa.z));
    }
  }
  if (MODE == FORCE_KICK_DRIFT || MODE == FORCE_KICK_DRIFT2) {
#pragma unroll
    for (int o = 16; o >= kG; o >>= 1) v2 = fmax(v2, __shfl_xor_sync(0xffffffffu, v2, o));
    if (lane == 0) {
      // non-negative doubles order like their bit patterns; the maximum settles after a few warps, so the
      // plain read filters out almost every atomic
      const unsigned long long b = (unsigned long long)__double_as_longlong(v2);
      if (b > *(volatile unsigned long long *)&ctrl->vmax2_bits) atomicMax(&ctrl->vmax2_bits, b);
    }
  }
}

// Plain grid: one force-warp per warp.
// worder (may be null): the launch covers force-warps worder[wbase .. wbase + wcount) -- the interior or
// the boundary part of the unit (cells.cu:k_warp_order); null = force-warps [0, wcount).
template <int VARIANT, int MODE>
__global__ void __launch_bounds__(FORCE_TPB, FORCE_MINBLK)
k_force(FORCE_PARAMS, const int *__restrict__ worder, long long wbase, long long wcount, ForceConsts fc) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= wcount) return;
  const long long w = worder ? (long long)worder[wbase + i] : i;
  force_warp<VARIANT, MODE>(w, threadIdx.x & 31, FORCE_PASS, fc);
}

// PERSISTENT grid with one work queue per SM (FORCE_PERSIST=1).  The force-warps are split into one
// contiguous range per SM; every resident warp of SM s takes the next force-warp of range s (atomic ticket),
// so at any moment the ~40 warps of an SM work on ~320 ADJACENT locals whose partners are the same few cell
// rows and share L1 lines (a round-robin grid puts neighbouring blocks on different SMs).  A warp whose range
// is exhausted steals from the following ranges, so the result does not depend on block placement.
// queue[0..nsm) is this launch's tickets; queue_next[0..nsm) is zeroed for the next launch.
template <int VARIANT, int MODE>
__global__ void __launch_bounds__(FORCE_TPB, FORCE_MINBLK)
k_force_persist(FORCE_PARAMS, int nsm, const int *__restrict__ sm_rank, int *queue, int *queue_next,
                ForceConsts fc) {
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < nsm; i += blockDim.x) queue_next[i] = 0;
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  const long long per = (nwarps + nsm - 1) / nsm;
  int cur = sm_rank[smid & 1023];  // dense rank of this SM (hardware SM ids may have gaps)
  int ticket = 0;
  if (lane == 0) ticket = atomicAdd(&queue[cur], 1);
  for (;;) {
    const int t = __shfl_sync(0xffffffffu, ticket, 0);
    const long long w = (long long)cur * per + t;
    if (t >= per || w >= nwarps) {
      // own range exhausted: look for a range that still has tickets, 32 queues per probe
      int found = -1;
      for (int off = 1; off < nsm && found < 0; off += 32) {
        const int cand = (cur + off + lane) % nsm;
        const bool has = (off + lane < nsm) && (*(volatile int *)&queue[cand] < per) &&
                         ((long long)cand * per + *(volatile int *)&queue[cand] < nwarps);
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (m) found = (cur + off + (__ffs(m) - 1)) % nsm;
      }
      if (found < 0) break;
      cur = found;
      if (lane == 0) ticket = atomicAdd(&queue[cur], 1);
      continue;
    }
    if (lane == 0) ticket = atomicAdd(&queue[cur], 1);  // next ticket, in flight while this one is processed
    force_warp<VARIANT, MODE>(w, lane, FORCE_PASS, fc);
  }
}

// Hardware SM ids (%smid) need not be dense: record which ids exist on this device and rank them.
__global__ void k_probe_smid(int *seen) {
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) seen[smid & 1023] = 1;
  // stay resident for a moment so that the blocks of this probe spread over every SM
  const long long t0 = clock64();
  while (clock64() - t0 < 20000) {}
}

static int force_queue_init(Unit *u) {
  int nsm = 0;
  MDB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, u->device));
  MDB_TRY(u->force_queue.ensure(2 * 1024));
  MDB_TRY(u->sm_rank.ensure(1024));
  MDB_CUDA(cudaMemsetAsync(u->force_queue.ptr, 0, sizeof(int) * 2 * 1024, u->stream));
  MDB_CUDA(cudaMemsetAsync(u->sm_rank.ptr, 0, sizeof(int) * 1024, u->stream));
  k_probe_smid<<<nsm * 16, 128, 0, u->stream>>>(u->sm_rank.ptr);
  int seen[1024];
  MDB_CUDA(cudaMemcpyAsync(seen, u->sm_rank.ptr, sizeof(seen), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  int rank[1024], n = 0;
  for (int i = 0; i < 1024; i++) rank[i] = seen[i] ? n++ : 0;
  if (n == 0) { set_error("SM probe saw no SM"); return MDB200_ECUDA; }
  for (int i = 0, k = 0; i < 1024; i++)
    if (!seen[i]) rank[i] = (k++) % n;
  MDB_CUDA(cudaMemcpyAsync(u->sm_rank.ptr, rank, sizeof(rank), cudaMemcpyHostToDevice, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  u->nsm = n;
  u->force_queue_parity = 0;
  return MDB200_OK;
}

static int force_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("MDB200_FORCE_VARIANT");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 2) v = 1;
  }
  return v;
}

static ForceConsts make_fc(const Unit *u) {
  ForceConsts fc;
  fc.CL2 = u->c.CL2;
  fc.dt = u->c.dt;
  fc.dt2 = u->c.dt2;
  fc.c24dt = 24.0 * u->c.dt;
  fc.c48dt = 48.0 * u->c.dt;
  fc.c2x8 = u->c.C2 * 8.0;
  fc.c2x8dt = u->c.C2 * 8.0 * u->c.dt;
  return fc;
}

#define FORCE_ARGS(F, QOUT) u->q.ptr, u->p.ptr, F, u->list.ptr, u->warp_off.ptr, u->warp_len.ptr, u->ref_of_slot.ptr, \
                            (int)u->pn, u->nwarps, QOUT, u->ctrl

// part: 0 = every force-warp, 1 = interior warps only, 2 = boundary warps only
static int force_launch(Unit *u, int mode, double *fout, int part = 0) {
  const int tpb = FORCE_TPB;
  const int *worder = part ? u->warp_order.ptr : nullptr;
  const long long wbase = (part == 2) ? u->n_interior : 0;
  const long long wcount = (part == 0) ? u->nwarps : (part == 1 ? u->n_interior : u->nwarps - u->n_interior);
  if (wcount <= 0) return MDB200_OK;
  const long long want = (wcount * 32 + tpb - 1) / tpb;
  ForceConsts fc = make_fc(u);
#if FORCE_PERSIST
  if (part != 0) { set_error("FORCE_PERSIST build has no split launch"); return MDB200_EINVAL; }
  if (!u->force_queue.ptr) MDB_TRY(force_queue_init(u));
  const int nsm = u->nsm;
  int *qcur = u->force_queue.ptr + 1024 * u->force_queue_parity;
  int *qnext = u->force_queue.ptr + 1024 * (1 - u->force_queue_parity);
  u->force_queue_parity ^= 1;
  const long long resident = (long long)nsm * FORCE_MINBLK;
  const unsigned grid = (unsigned)(want < resident ? want : resident);
#define KLAUNCH(V, M, F, QOUT) \
  k_force_persist<V, M><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(F, QOUT), nsm, u->sm_rank.ptr, qcur, qnext, fc)
#else
  (void)force_queue_init;
  const unsigned grid = (unsigned)want;
#define KLAUNCH(V, M, F, QOUT) k_force<V, M><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(F, QOUT), worder, wbase, wcount, fc)
#endif
#define LAUNCH(V)                                                                  \
  do {                                                                             \
    if (mode == FORCE_IMPULSE) KLAUNCH(V, FORCE_IMPULSE, fout, nullptr);           \
    else if (mode == FORCE_KICK) KLAUNCH(V, FORCE_KICK, nullptr, nullptr);         \
    else if (mode == FORCE_KICK_DRIFT) KLAUNCH(V, FORCE_KICK_DRIFT, nullptr, u->q_alt.ptr); \
    else KLAUNCH(V, FORCE_KICK_DRIFT2, nullptr, u->q_alt.ptr);                     \
  } while (0)
  switch (force_variant()) {
    case 0: LAUNCH(0); break;
    case 2: LAUNCH(2); break;
    default: LAUNCH(1); break;
  }
#undef LAUNCH
#undef KLAUNCH
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

// ForceCalculator::CalculateForce (fcalculator.cc:35-50)
int force_kick(Unit *u) {
  u->vmax_ready = false;
  if (u->pn == 0 || u->nwarps == 0) return MDB200_OK;
  return force_launch(u, FORCE_KICK, nullptr);
}

// CalculateForce + UpdatePositionHalf of MDManager::CalculateForce (mdmanager.cc:151-157) in one kernel;
// predrift = also the first half drift of the next step (MODE 3; the caller sets u->predrifted).
// part 0 = every force-warp; 1 = interior warps only (no ghost in any row: may run before the ghost
// exchange has finished); 2 = boundary warps.  The buffers swap after part 0 or 2.
int force_kick_drift(Unit *u, bool predrift, int part) {
  if (u->pn == 0 || u->nwarps == 0) return MDB200_OK;
  MDB_TRY(force_launch(u, predrift ? FORCE_KICK_DRIFT2 : FORCE_KICK_DRIFT, nullptr, part));
  if (part != 1) {
    std::swap(u->q, u->q_alt);  // ghosts live in both buffers (exchange.cu)
    u->vmax_ready = true;
  }
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
  MDB_TRY(force_launch(u, FORCE_IMPULSE, f.ptr));
  MDB_CUDA(cudaMemcpyAsync(f3_host, f.ptr, sizeof(double) * 3 * (size_t)u->pn, cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  f.release();
  return MDB200_OK;
}

}  // namespace mdb
