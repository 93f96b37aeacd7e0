// Elementwise integrators, thermostats and the Verlet-margin expiry test.
//
//   drift               ForceCalculator::UpdatePositionHalf   fcalculator.cc:22-32
//   Nose-Hoover scaling ForceCalculator::HeatbathMomenta      fcalculator.cc:752-762
//   Langevin            ForceCalculator::Langevin             fcalculator.cc:765-778
//   expiry test         PairList::IsPairListExpired / UpdatePairListValidity /
//                       CheckByDisplacement                   pairlist.cc:26-121
//
// Arithmetic that feeds decisions or is cheap to keep identical (drift, velocity budget,
// displacement test) is written with explicit round-to-nearest mul/add intrinsics so no FMA
// contraction happens: it then matches the scalar x86-64 reference bit for bit on equal inputs.
// All kernels are HBM-bound streams: 64 B read + 32 B written per particle for the drift.
#include <math.h>
#include "rng.cuh"
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "unit.h"

namespace mdb {

static const int TPB = 256;
static inline unsigned nblk(long long n, int tpb = TPB) {
  long long b = (n + tpb - 1) / tpb;
  return (unsigned)(b < 1 ? 1 : b);
}

__device__ __forceinline__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ double bits_d(unsigned long long b) { return __longlong_as_double((long long)b); }

__device__ __forceinline__ double block_max(double v, double *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = fmax(r, sh[w]);
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// drift
// ---------------------------------------------------------------------------------------------
__global__ void k_drift_half(double4 *__restrict__ q, const double4 *__restrict__ p, int pn, double dt2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  double4 qq = q[i];
  const double4 pp = p[i];
  qq.x = __dadd_rn(qq.x, __dmul_rn(pp.x, dt2));
  qq.y = __dadd_rn(qq.y, __dmul_rn(pp.y, dt2));
  qq.z = __dadd_rn(qq.z, __dmul_rn(pp.z, dt2));
  qq.w = qq.y;  // lane 3 of a position record repeats y (tile.cuh)
  q[i] = qq;
}

int drift_half_buffer(Unit *u, double4 *qbuf) {
  if (u->pn == 0) return MDB200_OK;
  k_drift_half<<<nblk(u->pn), TPB, 0, u->stream>>>(qbuf, u->p.ptr, (int)u->pn, u->c.dt2);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}
int drift_half(Unit *u) { return drift_half_buffer(u, u->q.ptr); }

// ---------------------------------------------------------------------------------------------
// thermostats
// ---------------------------------------------------------------------------------------------
__global__ void k_scale_p(double4 *__restrict__ p, int pn, double f) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  double4 v = p[i];
  v.x *= f; v.y *= f; v.z *= f; v.w *= f;  // all D = 4 lanes (fcalculator.cc:757-761)
  p[i] = v;
}

int heatbath_momenta(Unit *u) {
  u->vmax_ready = false;
  if (u->pn == 0) return MDB200_OK;
  // exp() evaluated on the host exactly where the reference evaluates it (fcalculator.cc:756)
  const double exp1 = exp(-u->c.dt2 * u->zeta);
  k_scale_p<<<nblk(u->pn), TPB, 0, u->stream>>>(u->p.ptr, (int)u->pn, exp1);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

// (per-particle counter RNG and the Langevin update: rng.cuh, shared with the fused force kernel)
__global__ void k_langevin(double4 *__restrict__ p, const int *__restrict__ ref_of_slot, int pn, double gamma,
                           double amp, double dt, unsigned key0, unsigned key1, unsigned step_lo,
                           unsigned step_hi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  double4 v = p[i];
  langevin_update(v, (unsigned)ref_of_slot[i], gamma, amp, dt, key0, key1, step_lo, step_hi);
  p[i] = v;
}

int langevin_kick(Unit *u) {
  u->vmax_ready = false;
  if (u->pn == 0) return MDB200_OK;
  const double dt = u->c.dt;
  const double g = u->prm.heatbath_gamma;
  const double T = u->prm.aimed_temperature;
  const double amp = sqrt(2.0 * g * T / dt);  // hb_D, fcalculator.cc:771
  unsigned long long sc = (unsigned long long)u->step_count;
  unsigned key0 = (unsigned)(u->seed + u->prm.unit_id);  // MDUnit::SetSeed: seed + id (mdunit.h:58)
  unsigned key1 = 0x6d646163u;                            // "mdac"
  k_langevin<<<nblk(u->pn), TPB, 0, u->stream>>>(u->p.ptr, u->ref_of_slot.ptr, (int)u->pn, g, amp, dt, key0, key1,
                                                  (unsigned)(sc & 0xffffffffu), (unsigned)(sc >> 32));
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

// ---------------------------------------------------------------------------------------------
// expiry test
// ---------------------------------------------------------------------------------------------
// Stage 1: velocity budget (pairlist.cc:34-36, 48-65).  max |p|^2 over locals; non-negative
// doubles order like their bit patterns, so one 64-bit atomicMax per block suffices.  The last
// block to finish does the scalar bookkeeping.
// scalar bookkeeping of PairList::UpdatePairListValidity (pairlist.cc:48-65) once max |p|^2 is known
__device__ __forceinline__ void expiry_budget(Ctrl *ctrl, double dt, double BL) {
  ctrl->last_vmax2_bits = atomicAdd(&ctrl->vmax2_bits, 0ull);
  double bl = ctrl->buffer_length;
  if (bl > 0.0) {
    double vmax = sqrt(bits_d(atomicAdd(&ctrl->vmax2_bits, 0ull)));
    double dec = __dmul_rn(__dmul_rn(vmax, 2.0), dt);
    if (dec > BL) atomicOr(&ctrl->err, (int)DE_TOOFAST);
    bl = __dadd_rn(bl, -dec);
    ctrl->buffer_length = bl;
  }
  ctrl->phase = (bl < 0.0) ? 1 : 0;
  ctrl->expired = 0;
  ctrl->vmax2_bits = 0ull;
  ctrl->dr2max_bits = 0ull;
  ctrl->cand_count = 0;
  ctrl->ticket[0] = 0u;
  ctrl->ticket[1] = 0u;
  ctrl->ticket[2] = 0u;
}

__global__ void k_expiry_vmax(const double4 *__restrict__ p, int pn, Ctrl *ctrl, double dt, double BL) {
  __shared__ double sh[TPB / 32];
  __shared__ bool last;
  double v = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pn; i += (long long)gridDim.x * blockDim.x) {
    const double4 pp = p[i];
    double s = __dadd_rn(__dadd_rn(__dmul_rn(pp.x, pp.x), __dmul_rn(pp.y, pp.y)), __dmul_rn(pp.z, pp.z));
    v = fmax(v, s);
  }
  v = block_max(v, sh);
  if (threadIdx.x == 0) {
    atomicMax(&ctrl->vmax2_bits, dbits(v));
    __threadfence();
    unsigned t = atomicAdd(&ctrl->ticket[0], 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    expiry_budget(ctrl, dt, BL);
  }
}

// The same bookkeeping when max |p|^2 is already in ctrl->vmax2_bits (left there by the fused kick + drift
// kernel, force.cu): one thread.
__global__ void k_expiry_budget(Ctrl *ctrl, double dt, double BL) { expiry_budget(ctrl, dt, BL); }

// Stage 2: displacement since the snapshot over locals AND ghosts (pairlist.cc:69-83).
// End-of-step position of item i.  Normally q[i]; in the pre-drifted state (unit.h) the locals' end-of-step
// positions are recomputed from the previous half-step buffer and p exactly as k_drift_half would have.
__device__ __forceinline__ double4 current_position(const double4 *__restrict__ q, const double4 *__restrict__ qprev,
                                                    const double4 *__restrict__ p, int i, int pn, double dt2) {
  if (qprev == nullptr || i >= pn) return q[i];
  double4 v = qprev[i];
  const double4 pp = p[i];
  v.x = __dadd_rn(v.x, __dmul_rn(pp.x, dt2));
  v.y = __dadd_rn(v.y, __dmul_rn(pp.y, dt2));
  v.z = __dadd_rn(v.z, __dmul_rn(pp.z, dt2));
  return v;
}

__global__ void k_expiry_disp(const double4 *__restrict__ q, const double4 *__restrict__ qprev,
                              const double4 *__restrict__ p, int pn, double dt2, const double4 *__restrict__ q0,
                              double *__restrict__ disp, int tn, Ctrl *ctrl, double BL) {
  if (ctrl->phase == 0) return;
  __shared__ double sh[TPB / 32];
  __shared__ bool last;
  double v = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tn; i += (long long)gridDim.x * blockDim.x) {
    const double4 a = q0[i], b = current_position(q, qprev, p, (int)i, pn, dt2);
    double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    disp[i] = d2;
    v = fmax(v, d2);
  }
  v = block_max(v, sh);
  if (threadIdx.x == 0) {
    atomicMax(&ctrl->dr2max_bits, dbits(v));
    __threadfence();
    unsigned t = atomicAdd(&ctrl->ticket[1], 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double dmax = sqrt(bits_d(atomicAdd(&ctrl->dr2max_bits, 0ull)));
    double cd = __dadd_rn(BL, -dmax);  // pairlist.cc:88
    ctrl->c_d2 = __dmul_rn(cd, cd);
  }
}

// Stage 3: candidates (pairlist.cc:90-97) and, by the last block, the candidate-pair test
// (pairlist.cc:104-119).  Both outcomes depend only on the candidate SET, not on its order.
__global__ void k_expiry_cand(const double4 *__restrict__ q, const double4 *__restrict__ qprev,
                              const double4 *__restrict__ p, int pn, double dt2, const double4 *__restrict__ q0,
                              const double *__restrict__ disp, int *__restrict__ cand, int tn, int Nl, Ctrl *ctrl,
                              double SL2, double CL2) {
  if (ctrl->phase == 0) return;
  __shared__ bool last;
  __shared__ int hit;
  const double c_d2 = ctrl->c_d2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tn; i += (long long)gridDim.x * blockDim.x) {
    if (disp[i] > c_d2) {
      int k = atomicAdd(&ctrl->cand_count, 1);
      if (k < Nl) cand[k] = (int)i;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned t = atomicAdd(&ctrl->ticket[2], 1u);
    last = (t == gridDim.x - 1);
    hit = 0;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int cnt = atomicAdd(&ctrl->cand_count, 0);
  if (cnt >= Nl) {
    // pairlist.cc:93-95 returns true from inside the loop, which needs at least one candidate
    if (threadIdx.x == 0) ctrl->expired = (cnt > 0) ? 1 : 0;
    return;
  }
  const long long npair = (long long)cnt * (cnt - 1) / 2;
  for (long long t = threadIdx.x; t < npair && !hit; t += blockDim.x) {
    // unrank t -> (a,b), a < b
    long long a = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (a * (a + 1) / 2 > t) a--;
    while ((a + 1) * (a + 2) / 2 <= t) a++;
    long long b2 = t - a * (a + 1) / 2;  // b2 <= a ; pair (b2, a+1)
    int i1 = cand[b2], i2 = cand[a + 1];
    const double4 o1 = q0[i1], o2 = q0[i2];
    double dx = o1.x - o2.x, dy = o1.y - o2.y, dz = o1.z - o2.z;
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (d2 < SL2) continue;
    const double4 n1 = current_position(q, qprev, p, i1, pn, dt2), n2 = current_position(q, qprev, p, i2, pn, dt2);
    dx = n1.x - n2.x; dy = n1.y - n2.y; dz = n1.z - n2.z;
    d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (d2 < CL2) hit = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) ctrl->expired = hit ? 1 : 0;
}

// the displacement test of the step (stages 2 and 3); both kernels return at once unless ctrl->phase is set
static int expiry_disp_launch(Unit *u) {
  const int grid_t = (int)((u->tn + TPB - 1) / TPB) < 1 ? 1 : (int)(((u->tn + TPB - 1) / TPB) > 1184 ? 1184 : ((u->tn + TPB - 1) / TPB));
  const int cll = u->prm.check_list_length;
  const int Nl = (u->tn > cll) ? cll : (int)u->tn;  // pairlist.cc:85
  const double4 *qprev = u->predrifted ? u->q_alt.ptr : nullptr;
  // pre-drifted after a Langevin step (predrifted == 2): the end-of-step positions are the previous buffer as it
  // is, i.e. "previous + p * 0" (adding +-0 is exact)
  const double dt2 = (u->predrifted == 2) ? 0.0 : u->c.dt2;
  k_expiry_disp<<<grid_t, TPB, 0, u->stream>>>(u->q.ptr, qprev, u->p.ptr, (int)u->pn, dt2, u->q0.ptr, u->disp.ptr,
                                               (int)u->tn, u->ctrl, u->c.BL);
  k_expiry_cand<<<grid_t, TPB, 0, u->stream>>>(u->q.ptr, qprev, u->p.ptr, (int)u->pn, dt2, u->q0.ptr, u->disp.ptr,
                                               u->cand.ptr, (int)u->tn, Nl, u->ctrl, u->c.SL2, u->c.CL2);
  u->launches += 2;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}
// buffer_length ... err of Ctrl (one copy): what the host needs of a check
static int expiry_pull(Unit *u) {
  static_assert(offsetof(Ctrl, err) + sizeof(int) - offsetof(Ctrl, buffer_length) <= 64, "Ctrl head");
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->buffer_length, &u->ctrl->buffer_length,
                           offsetof(Ctrl, err) + sizeof(int) - offsetof(Ctrl, buffer_length), cudaMemcpyDeviceToHost, u->stream));
  return MDB200_OK;
}

// The displacement kernels only act once the velocity budget is used up (ctrl->phase), a few steps before a rebuild.
// While the budget the host saw after the previous step still covers TWICE the decrement of that step they are not even
// launched (two launches of the five a small system's step has); should the device find the budget exhausted all the
// same, expiry_check_finish launches them after the synchronisation.
int expiry_check_launch(Unit *u) {
  const int grid_p = (int)((u->pn + TPB - 1) / TPB) < 1 ? 1 : (int)(((u->pn + TPB - 1) / TPB) > 1184 ? 1184 : ((u->pn + TPB - 1) / TPB));
  const int cll = u->prm.check_list_length;
  MDB_TRY(u->cand.ensure((size_t)(cll > 0 ? cll : 1) + 1));
  MDB_TRY(u->disp.ensure((size_t)u->cap_t + 1));
  u->lifetime++;
  if (u->vmax_ready)
    k_expiry_budget<<<1, 1, 0, u->stream>>>(u->ctrl, u->c.dt, u->c.BL);
  else {
    // a maximum a fused force launch left behind and nobody consumed (an external rebuild or rescale in between) must
    // not enter this test: pairlist.cc:48-65 looks at the current locals only
    MDB_CUDA(cudaMemsetAsync(&u->ctrl->vmax2_bits, 0, sizeof(unsigned long long), u->stream));
    k_expiry_vmax<<<grid_p, TPB, 0, u->stream>>>(u->p.ptr, (int)u->pn, u->ctrl, u->c.dt, u->c.BL);
  }
  u->launches++;
  u->vmax_ready = false;
  static const bool always = getenv("MDB200_EXPIRY_ALWAYS") != nullptr;
  double v2;
  memcpy(&v2, &u->ctrl_host->last_vmax2_bits, sizeof(double));
  const double dec = 2.0 * sqrt(v2) * u->c.dt;
  // lifetime 1 = first check of a list: ctrl_host still describes the previous one
  u->expiry_disp_skipped = !always && u->lifetime > 1 && u->ctrl_host->buffer_length > 0.0 && dec == dec &&
                           u->ctrl_host->buffer_length - 2.0 * dec > 0.0;
  if (!u->expiry_disp_skipped) MDB_TRY(expiry_disp_launch(u));
  MDB_CUDA(cudaGetLastError());
  return expiry_pull(u);
}
int expiry_check_finish(Unit *u) {
  if (!u->expiry_disp_skipped || u->ctrl_host->phase == 0) return MDB200_OK;
  u->expiry_disp_skipped = false;
  MDB_TRY(expiry_disp_launch(u));
  MDB_TRY(expiry_pull(u));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  return MDB200_OK;
}
int expiry_check_local(Unit *u, int *expired) {
  MDB_TRY(expiry_check_launch(u));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  MDB_TRY(expiry_check_finish(u));
  *expired = u->ctrl_host->expired;
  return MDB200_OK;
}

}  // namespace mdb
