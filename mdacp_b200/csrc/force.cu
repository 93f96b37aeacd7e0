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
//    take one entry each of that local's row per step and a 2-stage shuffle reduction combines them at the
//    end.  Why: ncu shows the kernel bound by the L1 data pipe on the scattered gather of partner positions
//    (FP64 pipe below 50 %; profiles/r01_force_*_ncu.txt, r01c_force_ncu.txt), and the cost of a warp-wide
//    gather of 32-byte records is set by conflicts between the four 32-byte sector slots of a line
//    (profiles/r01_ubench_gather.txt).  Rows are therefore stored in "rainbow" order (cells.cu: RowOrder):
//    at every step the 4 lanes of a local take partners of 4 different slots (index mod 4).
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
#ifndef FORCE_PRUNE_MINBLK
#define FORCE_PRUNE_MINBLK 8
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
  double keep2;   // PRUNE: (cutoff + delta)^2
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
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
  (void)w;
}


// list words are streamed once (L1 eviction hints on them or on the positions made no difference, see
// profiles/r01_force_experiments.md)
__device__ __forceinline__ uint2 ld_list(const uint2 *p) { return __ldg(p); }

// returns the bit pattern of r2 (non-negative doubles order like their bit patterns)
template <int VARIANT>
__device__ __forceinline__ long long pair_accumulate(double ax, double ay, double az, double jx, double jy,
                                                     double jz, long long cl2b, const ForceConsts &fc, double &fx,
                                                     double &fy, double &fz) {
  const double dx = jx - ax, dy = jy - ay, dz = jz - az;
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
  double df = pair_df<VARIANT>(r2, fc);
  // r2 > CL2: the pair contributes nothing (fcalculator.cc:209-211); also kills the sentinel padding
  const long long r2b = __double_as_longlong(r2);
  if (r2b > cl2b) df = 0.0;
  fx = fma(df, dx, fx);
  fy = fma(df, dy, fy);
  fz = fma(df, dz, fz);
  return r2b;
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
// PRUNE (with MODE 2 / 3): the launch walks the full ("outer") list and, on the side, derives the INNER list:
//   the entries with r2 <= keep2 = (cutoff + delta)^2, re-dealt in rainbow order into list_in at the same row
//   offsets, with their own step counts in warp_len_in.  25 % of the listed pairs lie between the cutoff and
//   the search length and contribute exactly zero; until some pair can have closed the distance delta
//   (velocity budget kept by the host, api.cu) the following launches walk the inner list and give
//   identical contributions (the dropped ones are exact zeros) at ~0.8x the gathers.
enum { FORCE_IMPULSE = 0, FORCE_KICK = 1, FORCE_KICK_DRIFT = 2, FORCE_KICK_DRIFT2 = 3 };
#define FORCE_PARAMS                                                                                              \
  const double4 *__restrict__ q, double4 *__restrict__ p, double *__restrict__ fout,                             \
      const unsigned *__restrict__ list, const long long *__restrict__ warp_off, const int *__restrict__ warp_len, \
      const int *__restrict__ ref_of_slot, int pn, long long nwarps, double4 *__restrict__ q_out, Ctrl *ctrl,    \
      unsigned *__restrict__ list_in, int *__restrict__ warp_len_in, int sentinel
#define FORCE_PASS q, p, fout, list, warp_off, warp_len, ref_of_slot, pn, nwarps, q_out, ctrl, list_in, warp_len_in, sentinel

// One force-warp: the kPW locals in slots [w kPW, (w+1) kPW), kG lanes each.
template <int VARIANT, int MODE, bool PRUNE>
__device__ __forceinline__ void force_warp(long long w, int lane, FORCE_PARAMS, const ForceConsts &fc) {
  const long long a = w * kPW + lane / kG;
  const bool act = a < pn;
  double ax, ay, az;
  ld_q256(q + (act ? a : (long long)pn - 1), ax, ay, az);
  const int len2 = warp_len[w] >> 1;
  const long long base = warp_off[w];
  const uint2 *__restrict__ lp = reinterpret_cast<const uint2 *>(list + base) + lane;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  const long long cl2b = __double_as_longlong(fc.CL2);
  // PRUNE: which of this lane's entries stay (bit 2k: first entry of step-pair k, bit 2k+1: second) and how
  // many per class (index mod 4), 4 x 8 bits
  const long long keep2b = __double_as_longlong(fc.keep2);
  unsigned long long keepmask = 0ull;
  unsigned kcnt = 0u;

  constexpr int UNR = FORCE_UNROLL;
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
    const long long r0 = pair_accumulate<VARIANT>(ax, ay, az, j0x, j0y, j0z, cl2b, fc, fx, fy, fz);
    const long long r1 = pair_accumulate<VARIANT>(ax, ay, az, j1x, j1y, j1z, cl2b, fc, fx, fy, fz);
    if (PRUNE) {
      if (r0 <= keep2b) { keepmask |= 1ull << (2 * k); kcnt += 1u << (8 * (e.x & 3u)); }
      if (r1 <= keep2b) { keepmask |= 2ull << (2 * k); kcnt += 1u << (8 * (e.y & 3u)); }
    }
  }
#pragma unroll
  for (int o = kG / 2; o > 0; o >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
    fz += __shfl_xor_sync(0xffffffffu, fz, o);
  }
  if (PRUNE) {
    // ---- deal the kept entries of each local out again in rainbow order (cells.cu) --------------------------
    const int c = lane % kG;
    unsigned incl = kcnt;  // inclusive scan of the packed class counts over the 4 lanes of the group
    unsigned t = __shfl_up_sync(0xffffffffu, incl, 1, kG);
    if (c >= 1) incl += t;
    t = __shfl_up_sync(0xffffffffu, incl, 2, kG);
    if (c >= 2) incl += t;
    const unsigned tot = __shfl_sync(0xffffffffu, incl, kG - 1, kG);  // class totals of the local
    unsigned run = incl - kcnt;                                        // first ordinals of this lane's entries
    const int nk = (int)((tot & 0xFFu) + ((tot >> 8) & 0xFFu) + ((tot >> 16) & 0xFFu) + (tot >> 24));
    int steps = (nk + kG - 1) / kG;
#pragma unroll
    for (int o = kG; o < 32; o <<= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
    steps = (steps + 1) & ~1;
    if (lane == 0) warp_len_in[w] = steps;
    const int m = rainbow_min(tot);
    const int glane0 = lane - c;
    for (int k = 0; k < len2; k++) {
      if (((keepmask >> (2 * k)) & 3ull) == 0ull) continue;
      const uint2 e = __ldg(lp + (long long)k * 32);  // the row was just read: L1 / L2 hit
      if ((keepmask >> (2 * k)) & 1ull) {
        const int cc = (int)(e.x & 3u);
        const int tt = (int)((run >> (8 * cc)) & 0xFFu);
        run += 1u << (8 * cc);
        const int ee = rainbow_place(tot, m, cc, tt);
        list_in[base + list_word(ee / kG, glane0 + ee % kG)] = e.x;
      }
      if ((keepmask >> (2 * k)) & 2ull) {
        const int cc = (int)(e.y & 3u);
        const int tt = (int)((run >> (8 * cc)) & 0xFFu);
        run += 1u << (8 * cc);
        const int ee = rainbow_place(tot, m, cc, tt);
        list_in[base + list_word(ee / kG, glane0 + ee % kG)] = e.y;
      }
    }
    for (int ee = nk + c; ee < steps * kG; ee += kG) list_in[base + list_word(ee / kG, glane0 + ee % kG)] = (unsigned)sentinel;
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

// worder (may be null): the launch covers force-warps worder[wbase .. wbase + wcount) -- the interior or
// the boundary part of the unit (cells.cu:k_warp_order); null = force-warps [0, wcount).
template <int VARIANT, int MODE>
__global__ void __launch_bounds__(FORCE_TPB, FORCE_MINBLK)
k_force(FORCE_PARAMS, const int *__restrict__ worder, long long wbase, long long wcount, ForceConsts fc) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= wcount) return;
  const long long w = worder ? (long long)worder[wbase + i] : i;
  force_warp<VARIANT, MODE, false>(w, threadIdx.x & 31, FORCE_PASS, fc);
}
// the same, deriving the inner list on the side (a few launches per rebuild interval)
template <int VARIANT, int MODE>
__global__ void __launch_bounds__(FORCE_TPB, FORCE_PRUNE_MINBLK)
k_force_prune(FORCE_PARAMS, const int *__restrict__ worder, long long wbase, long long wcount, ForceConsts fc) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= wcount) return;
  const long long w = worder ? (long long)worder[wbase + i] : i;
  force_warp<VARIANT, MODE, true>(w, threadIdx.x & 31, FORCE_PASS, fc);
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

#define FORCE_ARGS(F, QOUT, LIST, WLEN) u->q.ptr, u->p.ptr, F, LIST, u->warp_off.ptr, WLEN, u->ref_of_slot.ptr, \
                                        (int)u->pn, u->nwarps, QOUT, u->ctrl, u->list_in.ptr, u->warp_len_in.ptr, sentinel_index(u)

// part: 0 = every force-warp, 1 = interior warps only, 2 = boundary warps only
// inner: 0 = walk the outer list, 1 = walk the inner list, 2 = walk the outer list and derive the inner one
static int force_launch(Unit *u, int mode, double *fout, int part = 0, int inner = 0, double keep_len = 0.0) {
  const int tpb = FORCE_TPB;
  const int *worder = part ? u->warp_order.ptr : nullptr;
  const long long wbase = (part == 2) ? u->n_interior : 0;
  const long long wcount = (part == 0) ? u->nwarps : (part == 1 ? u->n_interior : u->nwarps - u->n_interior);
  if (wcount <= 0) return MDB200_OK;
  const unsigned grid = (unsigned)((wcount * 32 + tpb - 1) / tpb);
  ForceConsts fc = make_fc(u);
  fc.keep2 = keep_len * keep_len;
  const unsigned *lst = (inner == 1) ? u->list_in.ptr : u->list.ptr;
  const int *wlen = (inner == 1) ? u->warp_len_in.ptr : u->warp_len.ptr;
  if (inner == 2 && (mode == FORCE_IMPULSE || mode == FORCE_KICK)) { set_error("pruning launch needs a fused mode"); return MDB200_EINVAL; }
#define KLAUNCH(V, M, F, QOUT) k_force<V, M><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(F, QOUT, lst, wlen), worder, wbase, wcount, fc)
#define PLAUNCH(V, M) k_force_prune<V, M><<<grid, tpb, 0, u->stream>>>(FORCE_ARGS(nullptr, u->q_alt.ptr, lst, wlen), worder, wbase, wcount, fc)
#define LAUNCH(V)                                                                  \
  do {                                                                             \
    if (mode == FORCE_IMPULSE) KLAUNCH(V, FORCE_IMPULSE, fout, nullptr);           \
    else if (mode == FORCE_KICK) KLAUNCH(V, FORCE_KICK, nullptr, nullptr);         \
    else if (mode == FORCE_KICK_DRIFT && inner == 2) PLAUNCH(V, FORCE_KICK_DRIFT); \
    else if (mode == FORCE_KICK_DRIFT) KLAUNCH(V, FORCE_KICK_DRIFT, nullptr, u->q_alt.ptr); \
    else if (inner == 2) PLAUNCH(V, FORCE_KICK_DRIFT2);                            \
    else KLAUNCH(V, FORCE_KICK_DRIFT2, nullptr, u->q_alt.ptr);                     \
  } while (0)
  switch (force_variant()) {
    case 0: LAUNCH(0); break;
    case 2: LAUNCH(2); break;
    default: LAUNCH(1); break;
  }
#undef LAUNCH
#undef KLAUNCH
#undef PLAUNCH
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
// inner: 0 = walk the outer list, 1 = walk the inner list, 2 = walk the outer list and derive the inner one
// (entries within keep_len); the caller keeps u->inner_valid and the budget (api.cu).
// part 0 = every force-warp; 1 = interior warps only (no ghost in any row: may run before the ghost
// exchange has finished); 2 = boundary warps.  The buffers swap after part 0 or 2.
int force_kick_drift(Unit *u, bool predrift, int part, int inner, double keep_len) {
  if (u->pn == 0 || u->nwarps == 0) return MDB200_OK;
  if (inner == 2) {
    MDB_TRY(u->list_in.ensure(u->list.cap));  // same rows, same offsets
    MDB_TRY(u->warp_len_in.ensure((size_t)u->nwarps + 2));
  }
  static const bool dbg = getenv("MDB200_DEBUG_FORCE") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, u->stream); }
  MDB_TRY(force_launch(u, predrift ? FORCE_KICK_DRIFT2 : FORCE_KICK_DRIFT, nullptr, part, inner, keep_len));
  if (dbg) {
    cudaEventRecord(e1, u->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[mdb200] force launch: part %d list %s keep %.4f  %.3f ms\n", part,
            inner == 0 ? "outer" : (inner == 1 ? "inner" : "outer->inner"), keep_len, ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
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
