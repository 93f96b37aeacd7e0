// C-ABI of libmdacp_b200.so (include/mdacp_b200.h) and the per-step orchestration that
// MDManager does in the reference (mdmanager.cc:116-349).  Host code only; kernels live in
// cells.cu / exchange.cu / integrate.cu / force.cu / observe.cu.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <vector>
#include "comm.h"
#include "unit.h"

namespace mdb {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int check_device_error(Unit *u) {
  const int e = u->ctrl_host->err;
  if (e == 0) return MDB200_OK;
  if (e & DE_TOOFAST) {
    set_error("Too fast particles exists. Try smaller time step (unit %d)", u->prm.unit_id);
    return MDB200_ETOOFAST;
  }
  if (e & DE_POSITION) {
    set_error("Invalid position (unit %d)", u->prm.unit_id);
    return MDB200_EPOSITION;
  }
  if (e & DE_CELL) {
    set_error("Invalid index (unit %d)", u->prm.unit_id);
    return MDB200_ECELL;
  }
  if (e & DE_COMM) {
    set_error("ghost exchange timed out waiting for a neighbour process (unit %d)", u->prm.unit_id);
    return MDB200_ECOMM;
  }
  set_error("capacity exceeded (unit %d)", u->prm.unit_id);
  return MDB200_ECAPACITY;
}

static int pull_error_word(Unit *u) {
  MDB_CUDA(cudaMemcpyAsync(&u->ctrl_host->err, &u->ctrl->err, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  return check_device_error(u);
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

// ---------------------------------------------------------------------------------------------
// phase timing (mirrors the StopWatch objects of mdmanager.cc:117-147)
// ---------------------------------------------------------------------------------------------
enum Phase { PH_ALL = 0, PH_FORCE, PH_COMM, PH_PAIR, PH_INTEG, PH_CHECK };
struct PhaseTimer {
  Unit *u;
  int ph;
  cudaEvent_t a, b;
  bool on;
  PhaseTimer(Unit *u_, int ph_) : u(u_), ph(ph_), on(u_->profiling >= 2) {
    if (!on) return;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, u->stream);
  }
  ~PhaseTimer() {
    if (!on) return;
    cudaEventRecord(b, u->stream);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    u->t_ms[ph] += ms;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
};

static void derive_timer_resolve(Unit *u) {
  if (!u->dev_pending) return;
  float ms = 0.f;
  if (cudaEventSynchronize(u->dev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, u->dev[0], u->dev[1]) == cudaSuccess) {
    u->derive_ms += ms;
    u->derive_launches++;
  }
  u->dev_pending = false;
}

// level-1 profiling: two events around the force kernel, resolved at the next host sync
static void force_timer_begin(Unit *u) {
  if (u->profiling != 1) return;
  if (!u->fev_ready) {
    cudaEventCreate(&u->fev[0]);
    cudaEventCreate(&u->fev[1]);
    u->fev_ready = true;
  }
  cudaEventRecord(u->fev[0], u->stream);
}
static void force_timer_end(Unit *u) {
  if (u->profiling != 1) return;
  cudaEventRecord(u->fev[1], u->stream);
  u->fev_pending = true;
}
static void force_timer_resolve(Unit *u) {
  if (!u->fev_pending) return;
  float ms = 0.f;
  if (cudaEventSynchronize(u->fev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, u->fev[0], u->fev[1]) == cudaSuccess) {
    u->t_ms[PH_FORCE] += ms;
    u->force_launches++;
  }
  u->fev_pending = false;
}

// the same around the inner-list derivation (reported by mdb200_list_stats)
static void derive_timer_begin(Unit *u) {
  if (u->profiling != 1) return;
  if (!u->dev_ready) {
    cudaEventCreate(&u->dev[0]);
    cudaEventCreate(&u->dev[1]);
    u->dev_ready = true;
  }
  derive_timer_resolve(u);
  cudaEventRecord(u->dev[0], u->stream);
}
static void derive_timer_end(Unit *u) {
  if (u->profiling != 1) { u->derive_launches++; return; }
  cudaEventRecord(u->dev[1], u->stream);
  u->dev_pending = true;
}

// make every stream of the group wait for everything queued so far on the others
static int stage_barrier(Unit **us, int n) {
  std::vector<cudaStream_t> streams;
  std::vector<int> devs;
  for (int i = 0; i < n; i++) {
    if (std::find(streams.begin(), streams.end(), us[i]->stream) == streams.end()) {
      streams.push_back(us[i]->stream);
      devs.push_back(us[i]->device);
    }
  }
  if (streams.size() <= 1) return MDB200_OK;
  std::vector<cudaEvent_t> evs(streams.size());
  for (size_t k = 0; k < streams.size(); k++) {
    DeviceGuard g(devs[k]);
    MDB_CUDA(cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming));
    MDB_CUDA(cudaEventRecord(evs[k], streams[k]));
  }
  for (size_t k = 0; k < streams.size(); k++)
    for (size_t j = 0; j < streams.size(); j++)
      if (j != k) {
        DeviceGuard g(devs[k]);
        MDB_CUDA(cudaStreamWaitEvent(streams[k], evs[j], 0));
      }
  for (size_t k = 0; k < streams.size(); k++) cudaEventDestroy(evs[k]);
  return MDB200_OK;
}

// momenta queued by mdb200_upload_begin: everything that reads p on the unit's stream comes after their arrival
static int settle_upload(Unit *u) {
  if (!u->p_pending) return MDB200_OK;
  MDB_CUDA(cudaStreamWaitEvent(u->stream, u->ev_p_up, 0));
  u->p_pending = false;
  return MDB200_OK;
}

// Undo the speculative pre-drift (unit.h, force.cu MODE 3): afterwards q holds the end-of-step positions.
static int materialize(Unit *u, bool settle = true) {
  if (settle) MDB_TRY(settle_upload(u));  // (the first build after mdb200_upload_begin passes false: it needs no momenta)
  if (!u->predrifted) return MDB200_OK;
  DeviceGuard g(u->device);
  // NVE: last half-step positions + p dt/2, bit-identical to the plain path; Langevin: they are the end-of-step
  // positions already (mdmanager.cc:186-192 has no second half drift)
  if (u->predrifted == 1) MDB_TRY(drift_half_buffer(u, u->q_alt.ptr));
  std::swap(u->q, u->q_alt);
  u->predrifted = 0;
  return MDB200_OK;
}

// cell grid and periodic shifts from the current box c.L / c.s / c.e (MeshList::ChangeScale, meshlist.cc:31-55)
static void setup_cells(Unit *u) {
  const mdb200_params &p = u->prm;
  Consts &c = u->c;
  c.ncell = 1;
  for (int d = 0; d < 3; d++) {
    // meshlist.cc:31-55: ms = int(w / SL) interior cells of edge w/ms >= SL, one halo layer
    const double w = c.e[d] - c.s[d];
    int ms = static_cast<int>(w / c.SL);
    if (ms < 1) ms = 1;
    c.ms[d] = ms;
    c.m[d] = ms + 2;
    c.inv_edge[d] = 1.0 / (w / static_cast<double>(ms));
    c.ncell *= c.m[d];
  }
  // periodic shifts applied when a ghost crosses the global box (mdunit.cc:141-163)
  for (int dir = 0; dir < 6; dir++) {
    for (int d = 0; d < 3; d++) u->shift[dir][d] = 0.0;
    const int axis = dir / 2;
    const bool low = (dir % 2 == 0);
    const bool over = low ? (p.grid_pos[axis] == 0) : (p.grid_pos[axis] == p.grid_size[axis] - 1);
    if (over) u->shift[dir][axis] = low ? c.L[axis] : -c.L[axis];  // ParaInfo::IsOverBoundary, parainfo.cc:256-275
  }
}

static void setup_consts(Unit *u) {
  const mdb200_params &p = u->prm;
  Consts &c = u->c;
  for (int d = 0; d < 3; d++) {
    c.L[d] = p.L[d];
    // mdunit.cc:19-24
    const double ul = p.L[d] / static_cast<double>(p.grid_size[d]);
    c.s[d] = ul * static_cast<double>(p.grid_pos[d]);
    c.e[d] = c.s[d] + ul;
  }
  c.dt = p.time_step;
  c.dt2 = p.time_step * 0.5;
  c.BL = p.buffer_length;
  c.SL = p.cutoff + p.buffer_length;  // simulationinfo.cc:7
  c.SL2 = c.SL * c.SL;
  c.CL2 = p.cutoff * p.cutoff;
  {
    // variables.cc:21-27
    const double s2 = 1.0 / (p.cutoff * p.cutoff);
    const double s6 = s2 * s2 * s2;
    const double s8 = s6 * s2;
    const double s12 = s6 * s6;
    const double s14 = s12 * s2;
    c.C2 = 6.0 * s14 - 3.0 * s8;
    c.C0 = -s12 + s6 - c.C2 / s2;
  }
  setup_cells(u);
}

static int reset_ctrl(Unit *u, bool keep_zeta) {
  Ctrl h;
  memset(&h, 0, sizeof(h));
  h.buffer_length = u->c.BL;
  (void)keep_zeta;
  MDB_CUDA(cudaMemcpyAsync(u->ctrl, &h, sizeof(Ctrl), cudaMemcpyHostToDevice, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  memset(u->ctrl_host, 0, sizeof(Ctrl));
  return MDB200_OK;
}

// late momenta (unit.h): slot s takes the momentum of upload row psrc[s]
__global__ void k_gather_momenta(const double4 *__restrict__ p_up, const int *__restrict__ psrc, double4 *__restrict__ p, int pn) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < pn) p[s] = p_up[psrc[s]];
}
__global__ void k_iota(int *a, int *b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = i; b[i] = i; }
}
__global__ void k_pack_host_rows(const double *__restrict__ src4, double4 *__restrict__ dst, int n, int zero_w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 v = reinterpret_cast<const double4 *>(src4)[i];
  if (zero_w) v.w = v.y;  // positions: lane 3 repeats y on the device (tile.cuh)
  dst[i] = v;
}
// zero_w: positions leave the device with lane 3 = 0 as the reference's Variables::q has it
__global__ void k_gather_ref_order(const double4 *__restrict__ src, const int *__restrict__ slot_of_ref,
                                   double4 *__restrict__ dst, int pn, int tn, int zero_w) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= tn) return;
  double4 v = src[(r < pn) ? slot_of_ref[r] : r];
  if (zero_w) v.w = 0.0;
  dst[r] = v;
}
__global__ void k_gather_ref_order_i(const int *__restrict__ src, const int *__restrict__ slot_of_ref,
                                     int *__restrict__ dst, int pn) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= pn) return;
  dst[r] = src[slot_of_ref[r]];
}

// ---------------------------------------------------------------------------------------------
// rebuild = MDManager::MakePairList (mdmanager.cc:262-280)
// ---------------------------------------------------------------------------------------------
struct DebugClock {
  bool on;
  double t0;
  cudaStream_t st;
  static double now() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  }
  explicit DebugClock(cudaStream_t s) : on(getenv("MDB200_DEBUG") != nullptr), t0(0), st(s) {
    if (on) { cudaStreamSynchronize(st); t0 = now(); }
  }
  void lap(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const double t = now();
    fprintf(stderr, "[mdb200] rebuild %-28s %9.3f ms\n", what, 1e3 * (t - t0));
    t0 = t;
  }
};

static int make_pairlist(Unit **us, int n) {
  std::vector<long long> cnt(n);
  // Momenta still on their way (mdb200_upload_begin).  One unit in one process: the build below needs positions only,
  // so it carries the upload row of every local through migration and cell sort and the momenta are gathered into slot
  // order on the upload stream when they are there (MDB200_LATE_P=0: wait for them first, like everybody else).
  {
    static const bool no_late = getenv("MDB200_LATE_P") && atoi(getenv("MDB200_LATE_P")) == 0;
    const bool late = n == 1 && !comm_multi_process() && us[0]->p_pending && !us[0]->list_initialized && !no_late;
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      us[i]->late_p = late;
      if (!late) MDB_TRY(settle_upload(us[i]));
    }
  }
  struct LateScope {   // whatever happens below, nobody is left reading momenta that have not arrived
    Unit *u;
    ~LateScope() {
      if (u->late_p) {
        cudaStreamSynchronize(u->ustream);
        u->late_p = false;
        u->p_pending = false;
      }
    }
  } late_scope{us[0]};
  for (int i = 0; i < n; i++) us[i]->vmax_ready = false;  // migration changes the set of locals
  for (int i = 0; i < n; i++) us[i]->inner_valid = false;  // the inner list is derived from the list being replaced
  DebugClock dbg(us[0]->stream);
  std::vector<PhaseTimer *> timers;
  struct TimerScope {
    std::vector<PhaseTimer *> &t;
    ~TimerScope() { for (auto *x : t) delete x; }
  } scope{timers};
  for (int i = 0; i < n; i++) timers.push_back(new PhaseTimer(us[i], PH_PAIR));

  // --- SendParticles (mdmanager.cc:251-259): six one-hop migrations, then the periodic wrap
  for (int i = 0; i < n; i++) {
    DeviceGuard g(us[i]->device);
    us[i]->scratch_tail = std::max(us[i]->pn, us[i]->tn);
    MDB_TRY(rebuild_begin_order(us[i]));
  }
  for (int dir = 0; dir < 6; dir++) {
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      MDB_TRY(migrate_flag_and_count(us[i], dir, &cnt[i]));
    }
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      MDB_TRY(migrate_compact_pack(us[i], dir, cnt[i]));
    }
    MDB_TRY(stage_barrier(us, n));
    MDB_TRY(comm_migrate(us, n, dir, cnt.data()));
    MDB_TRY(stage_barrier(us, n));
  }
  dbg.lap("migration");
  for (int i = 0; i < n; i++) {
    DeviceGuard g(us[i]->device);
    MDB_TRY(wrap_periodic(us[i]));
    // --- MeshList cell sort: locals become cell-ordered; tn := pn (mdmanager.cc:264-268)
    MDB_TRY(cell_sort_locals(us[i]));
    if (us[i]->late_p) {
      // p (before the sort's swap: now p_alt) receives the upload rows; slot s wants row psrc[s]
      Unit *u = us[i];
      MDB_CUDA(cudaEventRecord(u->ev_perm, u->stream));
      MDB_CUDA(cudaStreamWaitEvent(u->ustream, u->ev_perm, 0));
      if (u->pn > 0) {
        k_gather_momenta<<<(unsigned)((u->pn + 255) / 256), 256, 0, u->ustream>>>(u->p_alt.ptr, u->psrc.ptr, u->p.ptr, (int)u->pn);
        u->launches++;
        MDB_CUDA(cudaGetLastError());
      }
      MDB_CUDA(cudaEventRecord(u->ev_p_ready, u->ustream));
    }
  }
  dbg.lap("wrap + cell sort");
  // --- border lists + first ghost exchange (mdmanager.cc:269-275)
  comm_p2p_invalidate();
  for (int dir = 0; dir < 6; dir++) {
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      MDB_TRY(find_border(us[i], dir, &cnt[i]));
    }
    MDB_TRY(stage_barrier(us, n));
    MDB_TRY(comm_border_setup(us, n, dir));
    MDB_TRY(stage_barrier(us, n));
  }
  MDB_TRY(comm_p2p_update(us, n));  // ghost windows of the neighbour processes for the per-step peer-to-peer refresh
  for (int i = 0; i < n; i++) {
    us[i]->ghost_composed = false;
    if (n == 1) {
      DeviceGuard g(us[i]->device);
      MDB_TRY(ghost_compose(us[i]));  // a periodic box on one GPU: the per-step refresh becomes one launch
    }
  }
  dbg.lap("border lists + ghosts");
  // --- per unit MDUnit::MakePairList (mdunit.cc:194-198)
  int lerr = MDB200_OK;
  for (int i = 0; i < n; i++) {
    Unit *u = us[i];
    DeviceGuard g(u->device);
    MDB_TRY(bin_ghosts(u));
    MDB_TRY(snapshot_positions(u));  // PairList::Init
    MDB_TRY(build_pair_list(u));     // MeshList::MakeList
    // PairList::Init: buffer_length = BufferLength, lifetime = 0
    const double bl = u->c.BL;
    MDB_CUDA(cudaMemcpyAsync(&u->ctrl->buffer_length, &bl, sizeof(double), cudaMemcpyHostToDevice, u->stream));
    u->lifetime = 0;
    u->list_initialized = true;
    u->nbuilds++;
    const int rc = pull_error_word(u);   // not returned yet: the other ranks must learn of it first
    if (rc != MDB200_OK && lerr == MDB200_OK) lerr = rc;
  }
  dbg.lap("ghost bins + snapshot + search");
  if (us[0]->late_p) {
    MDB_CUDA(cudaStreamWaitEvent(us[0]->stream, us[0]->ev_p_ready, 0));
    us[0]->late_p = false;
    us[0]->p_pending = false;
  }
  // a rank whose rebuild failed must not leave the others waiting in the next collective
  int gerr = lerr;
  MDB_TRY(comm_allreduce_or(us, n, &gerr));
  if (lerr != MDB200_OK) return lerr;
  if (gerr != MDB200_OK) {
    set_error("another rank reported error %d while rebuilding the pair list", gerr);
    return gerr;
  }
  return MDB200_OK;
}

// MDManager::SendBorderParticles (mdmanager.cc:283-292)
static int send_border_particles(Unit **us, int n) {
  if (n == 1 && us[0]->ghost_composed) {
    DeviceGuard g(us[0]->device);
    return ghost_refresh_composed(us[0], us[0]->stream);
  }
  MDB_TRY(comm_refresh_begin(us, n, false));
  for (int dir = 0; dir < 6; dir++) {
    MDB_TRY(comm_border_refresh(us, n, dir));
    if (dir % 2 == 1) MDB_TRY(stage_barrier(us, n));
  }
  return MDB200_OK;
}

// Inner-list budget (force.cu PRUNE), shared by all units of the process group: how far any pair can still
// have closed since the inner list was derived.  us[0]->inner_left < 0 = derive again at the next force launch.
// (kept in the group's first unit: inner_left = remaining distance budget, inner_delta = the delta the next
// derivation will use, 0 = pruning off)
static int inner_steps() {           // steps one derivation should last (MDB200_INNER_STEPS overrides)
  static const int v = getenv("MDB200_INNER_STEPS") ? std::max(4, atoi(getenv("MDB200_INNER_STEPS"))) : 16;
  return v;
}

static int any_expired(Unit **us, int n, int *expired) {
  int ex = 0, lerr = MDB200_OK;
  bool uninit = false;
  unsigned long long vbits = 0ull;
  for (int i = 0; i < n; i++)
    if (!us[i]->list_initialized) uninit = true;
  if (uninit) {
    ex = 1;  // pairlist.cc:27-29
  } else {
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      PhaseTimer t(us[i], PH_CHECK);
      MDB_TRY(expiry_check_launch(us[i]));
    }
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      MDB_CUDA(cudaStreamSynchronize(us[i]->stream));
      MDB_TRY(expiry_check_finish(us[i]));
      force_timer_resolve(us[i]);
      const int rc = check_device_error(us[i]);   // not returned yet: the other ranks must learn of it first
      if (rc != MDB200_OK && lerr == MDB200_OK) lerr = rc;
      ex |= us[i]->ctrl_host->expired;
      vbits = std::max(vbits, us[i]->ctrl_host->last_vmax2_bits);
    }
  }
  // Communicator::AllReduceBoolean (mdmanager.cc:375); the global max |p|^2 and the error code ride along
  int gerr = lerr;
  MDB_TRY(comm_allreduce_flag_bits(us, n, &ex, &gerr, &vbits));
  if (lerr != MDB200_OK) return lerr;
  if (gerr != MDB200_OK) {
    set_error("another rank reported error %d in this step", gerr);
    return gerr;
  }
  *expired = ex;
  // every particle (also the neighbours' that we see as ghosts) moves at most vmax dt per step, a pair
  // distance changes by at most twice that
  double v2;
  memcpy(&v2, &vbits, sizeof(double));
  const double per_step = 2.0 * sqrt(v2) * us[0]->c.dt;
  if (uninit || !(per_step == per_step)) {
    us[0]->inner_left = -1.0;
    us[0]->inner_delta = 0.0;
  } else {
    us[0]->inner_left -= per_step;
    // next derivation: a delta that lasts INNER_STEPS steps, at most half the skin; not worth it below 4 steps
    static const bool no_prune = getenv("MDB200_NO_PRUNE") != nullptr;
    const double want = inner_steps() * per_step, cap = 0.5 * us[0]->c.BL;
    us[0]->inner_delta = (4.0 * per_step > cap || no_prune) ? 0.0 : std::min(want, cap);
  }
  return MDB200_OK;
}

static int kinetic_temperature(Unit **us, int n, double *T) {
  double tot[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; i++) {
    DeviceGuard g(us[i]->device);
    double o[4];
    MDB_TRY(observe_kinetic(us[i], o));
    tot[0] += o[0];
    tot[3] += o[3];
  }
  MDB_TRY(comm_allreduce_sum(us, n, tot, 4));
  *T = tot[0] / tot[3] / 1.5;  // mdmanager.cc:390-392, 403-407
  return MDB200_OK;
}

// MDManager::Calculate (mdmanager.cc:116-148)
static int calculate(Unit **us, int n) {
  int expired = 0;
  MDB_TRY(any_expired(us, n, &expired));
  const bool ct = us[0]->prm.control_temperature != 0;
  const bool langevin = ct && us[0]->prm.heatbath_type != MDB200_HT_NOSEHOOVER;
  static const bool no_predrift = getenv("MDB200_NO_PREDRIFT") != nullptr;
  if (expired || (ct && !langevin) || no_predrift)
    for (int i = 0; i < n; i++) MDB_TRY(materialize(us[i], !expired));
  if (expired) MDB_TRY(make_pairlist(us, n));
  // The force kernel walks the INNER list (force.cu): the entries of the pair list within cutoff + delta when it
  // was derived.  A new one is due when there is none (rebuild, momenta or positions changed outside the step) or
  // when the velocity budget of a pruned one is used up; it is derived below, once the ghosts of this step are in.
  bool derive = false;
  for (int i = 0; i < n; i++) derive = derive || !us[i]->inner_valid;
  if (!derive && us[0]->inner_keep < us[0]->c.SL && us[0]->inner_left < 0.0) derive = true;
  double keep_len = us[0]->c.SL;
  if (derive) {
    if (us[0]->inner_delta > 0.0) keep_len = std::min(us[0]->prm.cutoff + us[0]->inner_delta, us[0]->c.SL);
    us[0]->inner_left = us[0]->inner_delta;
  }
  for (int i = 0; i < n; i++) {
    DeviceGuard g(us[i]->device);
    PhaseTimer t(us[i], PH_INTEG);
    if (us[i]->predrifted) us[i]->predrifted = 0;  // the previous force kernel already did this half drift
    else MDB_TRY(drift_half(us[i]));
  }
  // plain NVE: let the force kernel also do the first half drift of the next step (undone on demand)
  const bool predrift = !ct && !no_predrift;
  // one unit per process, neighbours behind NCCL, plain NVE: run the ghost exchange on a second stream and
  // let the interior force-warps (no ghost in their rows) start without waiting for it
  const bool split_test = getenv("MDB200_FORCE_SPLIT") != nullptr;
  // Measured (profiles/): the split launch costs ~0.15 ms; the exchange it hides costs ~0.15 ms with 2 remote
  // directions (2x1x1 units: no gain) and ~0.5 ms with 6 (2x2x2: 4.61 -> 4.34 ms/step), so overlap from 4 up.
  const char *ov_env = getenv("MDB200_OVERLAP");  // 0 / 1 forces it off / on
  const bool ov_want = ov_env ? (ov_env[0] == '1') : (n == 1 && comm_remote_directions(us[0]) >= 4);
  const bool overlap = !ct && n == 1 && comm_multi_process() && us[0]->profiling < 2 && us[0]->ntiles > 0 && ov_want;
  if (overlap) {
    Unit *u = us[0];
    DeviceGuard g(u->device);
    if (!u->cstream) {
      // highest priority: the exchange kernels are tiny, but at equal priority their blocks would queue behind
      // the ~500k blocks of the interior force launch and the "overlap" would serialise
      int prio_lo = 0, prio_hi = 0;
      MDB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      MDB_CUDA(cudaStreamCreateWithPriority(&u->cstream, cudaStreamNonBlocking, prio_hi));
      MDB_CUDA(cudaEventCreateWithFlags(&u->ev_drift, cudaEventDisableTiming));
      MDB_CUDA(cudaEventCreateWithFlags(&u->ev_ghost, cudaEventDisableTiming));
    }
    MDB_CUDA(cudaEventRecord(u->ev_drift, u->stream));
    MDB_CUDA(cudaStreamWaitEvent(u->cstream, u->ev_drift, 0));
    MDB_TRY(comm_refresh_begin(us, n, true));
    for (int dir = 0; dir < 6; dir++) MDB_TRY(comm_border_refresh(us, n, dir, true));
    MDB_CUDA(cudaEventRecord(u->ev_ghost, u->cstream));
  } else {
    std::vector<PhaseTimer *> tm;
    for (int i = 0; i < n; i++) tm.push_back(new PhaseTimer(us[i], PH_COMM));
    int r = send_border_particles(us, n);
    for (auto *x : tm) delete x;
    MDB_TRY(r);
  }
  // the inner list for the velocity budget decided above, once the ghosts of this step are in.  Where the force
  // launch(es) of this step take the persistent kernel they derive it on the side (they compute every r2 anyway: the
  // outer list is walked once instead of a separate compaction pass); otherwise k_tile_compact does.
  std::vector<double> fused(n, 0.0);
  if (derive) {
    const bool fusable_step = !(ct && !langevin) && !(langevin && no_predrift);
    const bool split = overlap || (split_test && !ct);
    for (int i = 0; i < n; i++) {
      Unit *u = us[i];
      DeviceGuard g(u->device);
      derive_timer_begin(u);
      if (fusable_step && keep_len < u->c.SL) {
        MDB_TRY(ensure_rainbow(u));
        if (force_can_derive(u, keep_len, split)) fused[i] = keep_len;
      }
      if (fused[i] == 0.0) {
        if (overlap) MDB_CUDA(cudaStreamWaitEvent(u->stream, u->ev_ghost, 0));  // the derivation reads the ghosts
        MDB_TRY(derive_inner(u, keep_len));
      }
      derive_timer_end(u);
    }
  }
  if (ct && us[0]->prm.heatbath_type == MDB200_HT_NOSEHOOVER) {
    // MDManager::CalculateNoseHoover (mdmanager.cc:160-175)
    double t = 0.0;
    MDB_TRY(kinetic_temperature(us, n, &t));
    for (int i = 0; i < n; i++) {
      Unit *u = us[i];
      DeviceGuard g(u->device);
      const double tau = u->prm.heatbath_tau;
      u->zeta += ((t - u->prm.aimed_temperature) / (tau * tau)) * u->c.dt2;  // fcalculator.cc:744-750
      MDB_TRY(heatbath_momenta(u));
      MDB_TRY(force_kick(u));
      MDB_TRY(heatbath_momenta(u));
    }
    MDB_TRY(kinetic_temperature(us, n, &t));
    for (int i = 0; i < n; i++) {
      Unit *u = us[i];
      DeviceGuard g(u->device);
      const double tau = u->prm.heatbath_tau;
      u->zeta += ((t - u->prm.aimed_temperature) / (tau * tau)) * u->c.dt2;
      MDB_TRY(drift_half(u));
    }
  } else if (ct) {
    // MDManager::CalculateLangevin (mdmanager.cc:186-192): no second half drift (finding 5)
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      PhaseTimer t(us[i], PH_FORCE);
      force_timer_begin(us[i]);
      if (no_predrift) {
        MDB_TRY(force_kick(us[i]));
        MDB_TRY(langevin_kick(us[i]));
      } else {
        // kick, Langevin update and the next step's first half drift in one launch (force.cu)
        MDB_TRY(force_kick_langevin(us[i], fused[i]));
        us[i]->predrifted = 2;
      }
      force_timer_end(us[i]);
    }
  } else if (overlap) {
    // MDManager::CalculateForce (mdmanager.cc:151-157) with the ghost exchange (queued above on the second
    // stream) overlapped by the interior part of the force kernel
    Unit *u = us[0];
    DeviceGuard g(u->device);
    force_timer_begin(u);
    MDB_TRY(force_kick_drift(u, predrift, 1, fused[0]));
    MDB_CUDA(cudaStreamWaitEvent(u->stream, u->ev_ghost, 0));
    MDB_TRY(force_kick_drift(u, predrift, 2, fused[0]));
    force_timer_end(u);
    u->predrifted = predrift ? 1 : 0;
  } else {
    // MDManager::CalculateForce (mdmanager.cc:151-157)
    for (int i = 0; i < n; i++) {
      DeviceGuard g(us[i]->device);
      {
        // kick and second half drift in one kernel (force.cu, FORCE_KICK_DRIFT)
        PhaseTimer t(us[i], PH_FORCE);
        force_timer_begin(us[i]);
        if (split_test) {  // test hook: the interior / boundary launches of the overlapped path, back to back
          MDB_TRY(force_kick_drift(us[i], predrift, 1, fused[i]));
          MDB_TRY(force_kick_drift(us[i], predrift, 2, fused[i]));
        } else {
          MDB_TRY(force_kick_drift(us[i], predrift, 0, fused[i]));
        }
        force_timer_end(us[i]);
        us[i]->predrifted = predrift ? 1 : 0;
      }
    }
  }
  for (int i = 0; i < n; i++) us[i]->step_count++;
  return MDB200_OK;
}

// room for pn locals (+ their ghosts) in the particle store, contents kept: locals with head-room for migration, ghosts
// from the halo-shell volume ratio
int reserve_store(Unit *u, long long pn) {
  long long need_p = std::max<long long>(u->prm.pn_capacity, pn);
  double ratio = 1.0;
  for (int d = 0; d < 3; d++) {
    const double w = u->c.e[d] - u->c.s[d];
    ratio *= (w + 2.0 * u->c.SL) / w;
  }
  long long need_t = std::max<long long>(u->prm.tn_capacity, (long long)(need_p * std::min(ratio, 27.0) * 1.05) + 1024);
  // arrivals of a migration are appended past the ghosts before the cell sort compacts them (cells.cu:
  // migrate_unpack), so the per-local arrays must reach past tn as well -- matters for small boxes, where the
  // ghost shell is as large as the box; sized here so that no rebuild re-allocates
  need_p = std::max<long long>(need_p, need_t - need_p / 4);
  // test hook (tests/test_gpu_multiprocess.py): start from the bare minimum so that the rebuilds have to grow the
  // position buffers -- the path that re-exports the peer-to-peer windows and parks replaced allocations
  static const bool small_cap = getenv("MDB200_DEBUG_SMALL_CAPACITY") != nullptr;
  if (small_cap) { need_p = pn + 64; need_t = pn + 64; }
  MDB_TRY(ensure_capacity(u, need_p, need_t));
  // what mdb200_upload_begin and the late-momenta build need, created with the store rather than inside somebody's
  // timed first call (stream / event creation and cudaMalloc synchronise the device)
  if (!u->ustream) {
    MDB_CUDA(cudaStreamCreateWithFlags(&u->ustream, cudaStreamNonBlocking));
    MDB_CUDA(cudaEventCreateWithFlags(&u->ev_q_up, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&u->ev_p_up, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&u->ev_perm, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&u->ev_p_ready, cudaEventDisableTiming));
  }
  MDB_TRY(u->psrc.ensure((size_t)u->cap_p + 1));
  MDB_TRY(u->psrc_alt.ensure((size_t)u->cap_p + 1));
  MDB_TRY(u->mig_send_s.ensure((size_t)(need_p / 64) + 4096));
  return MDB200_OK;
}

// the store now holds pn particles in reference order (slot = reference index), no ghosts, no pair list
// (pairlist.h:25: the next step rebuilds first)
int store_replaced(Unit *u, long long pn) {
  u->pn = pn;
  u->tn = pn;
  u->scratch_tail = pn;
  u->list_initialized = false;
  u->ghost_composed = false;
  u->vmax_ready = false;
  u->predrifted = 0;
  u->inner_valid = false;
  u->max_row = 0;
  u->max_row_exact = 0;
  u->npairs = 0;
  u->list_entries = 0;
  for (int d = 0; d < 6; d++) { u->n_border[d] = 0; u->recv_cnt[d] = 0; u->recv_off[d] = pn; }
  return reset_ctrl(u, true);
}

}  // namespace mdb

using namespace mdb;

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

int mdb200_abi_version(void) { return MDB200_ABI_VERSION; }
const char *mdb200_last_error(void) { return g_err; }

int mdb200_device_count(int *count) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *count = 0;
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return MDB200_ECUDA;
  }
  *count = c;
  return MDB200_OK;
}

int mdb200_create(const mdb200_params *params, int device, mdb200_unit **out) {
  if (!params || !out) { set_error("null argument"); return MDB200_EINVAL; }
  if (params->cutoff <= 0 || params->buffer_length <= 0 || params->time_step <= 0) {
    set_error("cutoff, buffer_length and time_step must be positive");
    return MDB200_EINVAL;
  }
  for (int d = 0; d < 3; d++)
    if (params->grid_size[d] < 1 || params->grid_pos[d] < 0 || params->grid_pos[d] >= params->grid_size[d] ||
        !(params->L[d] > 0)) {
      set_error("bad grid or box along axis %d", d);
      return MDB200_EINVAL;
    }
  int ndev = 0;
  MDB_TRY(mdb200_device_count(&ndev));
  if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return MDB200_ECUDA; }
  mdb200_unit *h = new mdb200_unit();
  Unit *u = &h->u;
  u->prm = *params;
  u->device = device;
  setup_consts(u);
  for (int d = 0; d < 3; d++) {
    const double w = u->c.e[d] - u->c.s[d];
    if (static_cast<int>(w / u->c.SL) < 1) {
      delete h;
      set_error("sub-box narrower than the search length along axis %d", d);
      return MDB200_EINVAL;
    }
  }
  DeviceGuard g(device);
  // one stream per device for the whole process: units that share a device are then ordered
  // by the stream itself when they hand particles to each other
  {
    static std::map<int, cudaStream_t> dev_stream;
    static std::mutex dev_stream_mutex;  // mdb200_create may be called from one host thread per unit
    std::lock_guard<std::mutex> lock(dev_stream_mutex);
    auto it = dev_stream.find(device);
    if (it == dev_stream.end()) {
      cudaStream_t st;
      MDB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      it = dev_stream.emplace(device, st).first;
    }
    u->stream = it->second;
    u->own_stream = false;
  }
  // (no MDB_CUDA / MDB_TRY early returns from here on: what has been allocated is released on failure)
  int rc = MDB200_OK;
  if (cudaMalloc(&u->ctrl, sizeof(Ctrl)) != cudaSuccess || cudaMallocHost(&u->ctrl_host, sizeof(Ctrl)) != cudaSuccess) {
    set_error("CUDA: %s (control block of unit %d)", cudaGetErrorString(cudaGetLastError()), params->unit_id);
    rc = MDB200_ECUDA;
  }
  if (rc == MDB200_OK) rc = reset_ctrl(u, false);
  if (rc != MDB200_OK) {
    if (u->ctrl) cudaFree(u->ctrl);
    if (u->ctrl_host) cudaFreeHost(u->ctrl_host);
    delete h;
    return rc;
  }
  for (int d = 0; d < 6; d++) { u->nbr[d] = nullptr; u->src[d] = nullptr; }
  *out = h;
  return MDB200_OK;
}

int mdb200_destroy(mdb200_unit *h) {
  if (!h) return MDB200_OK;
  Unit *u = &h->u;
  DeviceGuard g(u->device);
  cudaStreamSynchronize(u->stream);
  comm_release(u);
  u->q.release(); u->q_alt.release(); u->p.release(); u->p_alt.release();
  u->type.release(); u->type_alt.release(); u->q0.release();
  u->ref_of_slot.release(); u->slot_of_ref.release(); u->ord.release(); u->ord_alt.release();
  u->flags.release(); u->scan.release(); u->scan_ws.release(); u->scan_ws64.release();
  u->cell_of.release(); u->cell_count.release(); u->cell_start.release(); u->cell_cursor.release();
  u->cell_items.release(); u->gcell_count.release(); u->gcell_start.release(); u->gcell_cursor.release();
  u->gcell_items.release();
  for (int d = 0; d < 6; d++) u->border[d].release();
  u->ghost_src.release();
  u->ghost_shift.release();
  u->send_buf.release(); u->mig_send_q.release(); u->mig_send_p.release(); u->mig_recv_q.release();
  u->mig_recv_p.release(); u->mig_send_t.release(); u->mig_recv_t.release();
  u->row_count.release(); u->list_in.release(); u->fw_len_in.release(); u->tile_fw0.release(); u->tile_flag.release(); u->tile_order.release(); u->tile_tmp.release();
  if (u->ustream) {
    cudaStreamSynchronize(u->ustream);
    cudaStreamDestroy(u->ustream);
    cudaEventDestroy(u->ev_q_up); cudaEventDestroy(u->ev_p_up); cudaEventDestroy(u->ev_perm); cudaEventDestroy(u->ev_p_ready);
  }
  u->psrc.release(); u->psrc_alt.release(); u->mig_send_s.release();
  if (u->cstream) { cudaStreamDestroy(u->cstream); cudaEventDestroy(u->ev_drift); cudaEventDestroy(u->ev_ghost); }
  u->fw_len.release(); u->fw_off.release(); u->fw_sent.release(); u->list.release();
  u->disp.release(); u->cand.release(); u->partial.release();
  u->snap_q.release(); u->snap_p.release(); u->snap_t.release();
  if (u->snap_stream) {
    cudaStreamSynchronize(u->snap_stream);
    cudaStreamDestroy(u->snap_stream); cudaEventDestroy(u->snap_ready); cudaEventDestroy(u->snap_done);
  }
  if (u->snap_hq) { cudaFreeHost(u->snap_hq); cudaFreeHost(u->snap_hp); cudaFreeHost(u->snap_ht); }
  if (u->ctrl) cudaFree(u->ctrl);
  if (u->ctrl_host) cudaFreeHost(u->ctrl_host);
  if (u->own_stream) cudaStreamDestroy(u->stream);
  delete h;
  return MDB200_OK;
}

int mdb200_get_rect(mdb200_unit *h, double s[3], double e[3]) {
  if (!h) return MDB200_EINVAL;
  for (int d = 0; d < 3; d++) { s[d] = h->u.c.s[d]; e[d] = h->u.c.e[d]; }
  return MDB200_OK;
}

// async_p: the momenta travel on the upload stream straight into p (no lane to repair: Variables::p has lane 3 = 0)
// and the call returns once the positions are there (mdb200_upload_begin)
static int upload_impl(mdb200_unit *h, const double (*q)[4], const double (*p)[4], const int *type, long long pn, bool async_p) {
  if (!h || pn < 0 || (pn > 0 && (!q || !p))) { set_error("bad upload arguments"); return MDB200_EINVAL; }
  if (pn > 2000000000LL) { set_error("pn exceeds 32-bit index range"); return MDB200_ECAPACITY; }
  Unit *u = &h->u;
  DeviceGuard g(u->device);
  MDB_TRY(settle_upload(u));  // an earlier mdb200_upload_begin whose momenta nobody has waited for
  MDB_TRY(reserve_store(u, pn));
  if (pn > 0) {
    // positions are staged through the alt buffer (the device-side repack sets lane 3 = y, tile.cuh)
    MDB_CUDA(cudaMemcpyAsync(u->q_alt.ptr, q, sizeof(double) * 4 * (size_t)pn, cudaMemcpyHostToDevice, u->stream));
    const unsigned nb = (unsigned)((pn + 255) / 256);
    if (async_p) {
      // behind the positions on the link (and behind whatever the unit's stream did with p before)
      MDB_CUDA(cudaEventRecord(u->ev_q_up, u->stream));
      MDB_CUDA(cudaStreamWaitEvent(u->ustream, u->ev_q_up, 0));
      MDB_CUDA(cudaMemcpyAsync(u->p.ptr, p, sizeof(double) * 4 * (size_t)pn, cudaMemcpyHostToDevice, u->ustream));
      MDB_CUDA(cudaEventRecord(u->ev_p_up, u->ustream));
    } else {
      MDB_CUDA(cudaMemcpyAsync(u->p_alt.ptr, p, sizeof(double) * 4 * (size_t)pn, cudaMemcpyHostToDevice, u->stream));
    }
    k_pack_host_rows<<<nb, 256, 0, u->stream>>>((const double *)u->q_alt.ptr, u->q.ptr, (int)pn, 1);
    if (!async_p) k_pack_host_rows<<<nb, 256, 0, u->stream>>>((const double *)u->p_alt.ptr, u->p.ptr, (int)pn, 0);
    if (type)
      MDB_CUDA(cudaMemcpyAsync(u->type.ptr, type, sizeof(int) * (size_t)pn, cudaMemcpyHostToDevice, u->stream));
    else
      MDB_CUDA(cudaMemsetAsync(u->type.ptr, 0, sizeof(int) * (size_t)pn, u->stream));
    k_iota<<<nb, 256, 0, u->stream>>>(u->ref_of_slot.ptr, u->slot_of_ref.ptr, (int)pn);
    u->launches += 3;
    MDB_CUDA(cudaGetLastError());
  }
  MDB_TRY(store_replaced(u, pn));
  u->p_pending = async_p && pn > 0;
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  return MDB200_OK;
}

int mdb200_upload(mdb200_unit *h, const double (*q)[4], const double (*p)[4], const int *type, long long pn) {
  return upload_impl(h, q, p, type, pn, false);
}
int mdb200_upload_begin(mdb200_unit *h, const double (*q)[4], const double (*p)[4], const int *type, long long pn) {
  static const bool off = getenv("MDB200_UPLOAD_ASYNC") && atoi(getenv("MDB200_UPLOAD_ASYNC")) == 0;
  return upload_impl(h, q, p, type, pn, !off);
}

// On-device initial states (generate.cu).  The makers append to the particle store like MDUnit::AddParticle; a store
// that has been stepped (slot order != reference order) must be cleared first.
static int gen_guard(mdb200_unit *h) {
  if (!h) { set_error("null unit"); return MDB200_EINVAL; }
  {
    DeviceGuard g(h->u.device);
    MDB_TRY(settle_upload(&h->u));
  }
  if (h->u.list_initialized) {
    set_error("the generators append to a freshly uploaded / generated store: call mdb200_clear first");
    return MDB200_EINVAL;
  }
  return MDB200_OK;
}
int mdb200_clear(mdb200_unit *h) {
  if (!h) return MDB200_EINVAL;
  DeviceGuard g(h->u.device);
  MDB_TRY(settle_upload(&h->u));
  return store_replaced(&h->u, 0);
}
int mdb200_generate_benchmark(mdb200_unit *h, double density, long long *added) {
  MDB_TRY(gen_guard(h));
  if (!(density > 0.0)) { set_error("density must be positive"); return MDB200_EINVAL; }
  DeviceGuard g(h->u.device);
  long long n = 0;
  MDB_TRY(generate_benchmark(&h->u, density, &n));
  if (added) *added = n;
  return MDB200_OK;
}
int mdb200_generate_lattice(mdb200_unit *h, double density, double radius, long long *added) {
  MDB_TRY(gen_guard(h));
  if (!(density > 0.0)) { set_error("density must be positive"); return MDB200_EINVAL; }
  DeviceGuard g(h->u.device);
  long long n = 0;
  MDB_TRY(generate_lattice(&h->u, density, radius, &n));
  if (added) *added = n;
  return MDB200_OK;
}
int mdb200_generate_ball(mdb200_unit *h, const double c[3], double r, double xv, double density, long long *added) {
  MDB_TRY(gen_guard(h));
  if (!c || !(density > 0.0) || !(r > 0.0)) { set_error("bad ball arguments"); return MDB200_EINVAL; }
  DeviceGuard g(h->u.device);
  long long n = 0;
  MDB_TRY(generate_ball(&h->u, c, r, xv, density, &n));
  if (added) *added = n;
  return MDB200_OK;
}
int mdb200_generate_initial_velocity(mdb200_unit *h, double v0, int seed) {
  MDB_TRY(gen_guard(h));
  DeviceGuard g(h->u.device);
  return generate_initial_velocity(&h->u, v0, seed);
}

int mdb200_download(mdb200_unit *h, double (*q)[4], double (*p)[4], int *type, long long cap_rows, long long *pn,
                    long long *tn) {
  if (!h) return MDB200_EINVAL;
  Unit *u = &h->u;
  DeviceGuard g(u->device);
  MDB_TRY(materialize(u));
  if (pn) *pn = u->pn;
  if (tn) *tn = u->tn;
  if (!q && !p && !type) return MDB200_OK;
  if (cap_rows < u->pn) {
    set_error("download buffers too small: %lld rows, need pn %lld (tn %lld)", cap_rows, u->pn, u->tn);
    return MDB200_ECAPACITY;
  }
  if (u->tn == 0) return MDB200_OK;
  const long long qrows = cap_rows < u->tn ? cap_rows : u->tn;  // locals always, ghosts as far as they fit
  const unsigned nbt = (unsigned)((qrows + 255) / 256), nbp = (unsigned)((u->pn + 255) / 256);
  if (q) {
    k_gather_ref_order<<<nbt, 256, 0, u->stream>>>(u->q.ptr, u->slot_of_ref.ptr, u->q_alt.ptr, (int)u->pn, (int)qrows, 1);
    u->launches++;
    MDB_CUDA(cudaMemcpyAsync(q, u->q_alt.ptr, sizeof(double) * 4 * (size_t)qrows, cudaMemcpyDeviceToHost, u->stream));
  }
  if (p && u->pn) {
    k_gather_ref_order<<<nbp, 256, 0, u->stream>>>(u->p.ptr, u->slot_of_ref.ptr, u->p_alt.ptr, (int)u->pn, (int)u->pn, 0);
    u->launches++;
    MDB_CUDA(cudaMemcpyAsync(p, u->p_alt.ptr, sizeof(double) * 4 * (size_t)u->pn, cudaMemcpyDeviceToHost, u->stream));
  }
  if (type && u->pn) {
    k_gather_ref_order_i<<<nbp, 256, 0, u->stream>>>(u->type.ptr, u->slot_of_ref.ptr, u->type_alt.ptr, (int)u->pn);
    u->launches++;
    MDB_CUDA(cudaMemcpyAsync(type, u->type_alt.ptr, sizeof(int) * (size_t)u->pn, cudaMemcpyDeviceToHost, u->stream));
  }
  MDB_CUDA(cudaGetLastError());
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  return MDB200_OK;
}

// Dump path (MDUnit::SaveAsCdview mdunit.cc:52-67, FileDumper mdmanager.cc:475-505 read Variables::q/p/type of
// the locals): the reference-ordered copy is gathered on the unit's stream into staging buffers nothing else
// uses, and travels to pinned host memory on a second stream, so the following steps overlap the transfer.
int mdb200_snapshot_begin(mdb200_unit *h) {
  if (!h) return MDB200_EINVAL;
  Unit *u = &h->u;
  DeviceGuard g(u->device);
  if (u->snap_pending) {  // a snapshot nobody waited for: finish it before its buffers are reused
    MDB_CUDA(cudaEventSynchronize(u->snap_done));
    u->snap_pending = false;
  }
  MDB_TRY(materialize(u));
  if (!u->snap_stream) {
    MDB_CUDA(cudaStreamCreateWithFlags(&u->snap_stream, cudaStreamNonBlocking));
    MDB_CUDA(cudaEventCreateWithFlags(&u->snap_ready, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&u->snap_done, cudaEventDisableTiming));
  }
  const long long pn = u->pn;
  u->snap_pn = pn;
  if (pn > u->snap_hcap) {
    if (u->snap_hq) { cudaFreeHost(u->snap_hq); cudaFreeHost(u->snap_hp); cudaFreeHost(u->snap_ht); }
    u->snap_hq = u->snap_hp = nullptr;
    u->snap_ht = nullptr;
    u->snap_hcap = 0;
    const long long cap = pn + pn / 16 + 64;
    MDB_CUDA(cudaMallocHost(&u->snap_hq, sizeof(double4) * (size_t)cap));
    MDB_CUDA(cudaMallocHost(&u->snap_hp, sizeof(double4) * (size_t)cap));
    MDB_CUDA(cudaMallocHost(&u->snap_ht, sizeof(int) * (size_t)cap));
    u->snap_hcap = cap;
  }
  if (pn > 0) {
    MDB_TRY(u->snap_q.ensure((size_t)pn));
    MDB_TRY(u->snap_p.ensure((size_t)pn));
    MDB_TRY(u->snap_t.ensure((size_t)pn));
    const unsigned nb = (unsigned)((pn + 255) / 256);
    k_gather_ref_order<<<nb, 256, 0, u->stream>>>(u->q.ptr, u->slot_of_ref.ptr, u->snap_q.ptr, (int)pn, (int)pn, 1);
    k_gather_ref_order<<<nb, 256, 0, u->stream>>>(u->p.ptr, u->slot_of_ref.ptr, u->snap_p.ptr, (int)pn, (int)pn, 0);
    k_gather_ref_order_i<<<nb, 256, 0, u->stream>>>(u->type.ptr, u->slot_of_ref.ptr, u->snap_t.ptr, (int)pn);
    u->launches += 3;
    MDB_CUDA(cudaGetLastError());
    MDB_CUDA(cudaEventRecord(u->snap_ready, u->stream));
    MDB_CUDA(cudaStreamWaitEvent(u->snap_stream, u->snap_ready, 0));
    MDB_CUDA(cudaMemcpyAsync(u->snap_hq, u->snap_q.ptr, sizeof(double4) * (size_t)pn, cudaMemcpyDeviceToHost, u->snap_stream));
    MDB_CUDA(cudaMemcpyAsync(u->snap_hp, u->snap_p.ptr, sizeof(double4) * (size_t)pn, cudaMemcpyDeviceToHost, u->snap_stream));
    MDB_CUDA(cudaMemcpyAsync(u->snap_ht, u->snap_t.ptr, sizeof(int) * (size_t)pn, cudaMemcpyDeviceToHost, u->snap_stream));
  }
  MDB_CUDA(cudaEventRecord(u->snap_done, u->snap_stream));
  u->snap_pending = true;
  return MDB200_OK;
}

int mdb200_snapshot_wait(mdb200_unit *h, const double (**q)[4], const double (**p)[4], const int **type,
                         long long *pn) {
  if (!h) return MDB200_EINVAL;
  Unit *u = &h->u;
  if (!u->snap_done) { set_error("no snapshot was begun"); return MDB200_EINVAL; }
  DeviceGuard g(u->device);
  MDB_CUDA(cudaEventSynchronize(u->snap_done));
  u->snap_pending = false;
  if (q) *q = reinterpret_cast<const double (*)[4]>(u->snap_hq);
  if (p) *p = reinterpret_cast<const double (*)[4]>(u->snap_hp);
  if (type) *type = u->snap_ht;
  if (pn) *pn = u->snap_pn;
  return MDB200_OK;
}

int mdb200_get_counts(mdb200_unit *h, long long *pn, long long *tn, long long *npairs, long long *nbuilds) {
  if (!h) return MDB200_EINVAL;
  if (pn) *pn = h->u.pn;
  if (tn) *tn = h->u.tn;
  if (npairs) *npairs = h->u.npairs;
  if (nbuilds) *nbuilds = h->u.nbuilds;
  return MDB200_OK;
}

int mdb200_comm_unique_id(void *out128) { return comm_unique_id(out128); }

int mdb200_comm_init(mdb200_unit **units, int n, int nranks, int rank, const void *nccl_unique_id) {
  if (!units || n < 1) { set_error("no units"); return MDB200_EINVAL; }
  std::vector<Unit *> us(n);
  for (int i = 0; i < n; i++) us[i] = &units[i]->u;
  return comm_init(us.data(), n, nranks, rank, nccl_unique_id);
}

static int unpack_units(mdb200_unit **units, int n, std::vector<Unit *> &us) {
  if (!units || n < 1) { set_error("no units"); return MDB200_EINVAL; }
  us.resize(n);
  for (int i = 0; i < n; i++) {
    if (!units[i]) { set_error("null unit"); return MDB200_EINVAL; }
    us[i] = &units[i]->u;
  }
  if (!us[0]->comm_ready) {
    // a lone unit that was never wired: periodic self-neighbour in every direction
    if (n == 1 && us[0]->prm.grid_size[0] * us[0]->prm.grid_size[1] * us[0]->prm.grid_size[2] == 1) {
      MDB_TRY(comm_init(us.data(), 1, 1, 0, nullptr));
    } else {
      set_error("mdb200_comm_init has not been called for these units");
      return MDB200_ECOMM;
    }
  }
  return MDB200_OK;
}

int mdb200_make_pairlist(mdb200_unit **units, int n) {
  std::vector<Unit *> us;
  MDB_TRY(unpack_units(units, n, us));
  for (int i = 0; i < n; i++) MDB_TRY(materialize(us[i], false));
  return make_pairlist(us.data(), n);
}

__global__ void k_scale_positions(double4 *__restrict__ q, int pn, double alpha) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  double4 v = q[i];
  v.x *= alpha; v.y *= alpha; v.z *= alpha;  // Variables::ChangeScale (variables.cc:133-139)
  v.w = v.y;
  q[i] = v;
}

// MDManager::ChangeScale (mdmanager.cc:432-443): rebuild, scale the box (SimulationInfo::L, every unit's rect
// by MDRect::ChangeScale mdrect.cc:55-60, i.e. s *= alpha, e *= alpha), the locals' positions and the cell grid
// (MeshList::ChangeScale), rebuild again.
int mdb200_change_scale(mdb200_unit **units, int n, double alpha) {
  std::vector<Unit *> us;
  MDB_TRY(unpack_units(units, n, us));
  if (!(alpha > 0.0)) { set_error("scale factor must be positive"); return MDB200_EINVAL; }
  // every unit first: a refusal must not leave some units scaled and others not
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) {
      const double w = us[i]->c.e[d] * alpha - us[i]->c.s[d] * alpha;
      if (static_cast<int>(w / us[i]->c.SL) < 1) {
        set_error("sub-box narrower than the search length along axis %d after scaling", d);
        return MDB200_EINVAL;
      }
    }
  for (int i = 0; i < n; i++) MDB_TRY(materialize(us[i]));
  MDB_TRY(make_pairlist(us.data(), n));
  for (int i = 0; i < n; i++) {
    Unit *u = us[i];
    DeviceGuard g(u->device);
    for (int d = 0; d < 3; d++) {
      u->prm.L[d] = u->prm.L[d] * alpha;
      u->c.L[d] = u->prm.L[d];
      u->c.s[d] *= alpha;
      u->c.e[d] *= alpha;
    }
    setup_cells(u);
    if (u->pn > 0) {
      k_scale_positions<<<(unsigned)((u->pn + 255) / 256), 256, 0, u->stream>>>(u->q.ptr, (int)u->pn, alpha);
      u->launches++;
      MDB_CUDA(cudaGetLastError());
    }
    u->inner_valid = false;
  }
  return make_pairlist(us.data(), n);
}

int mdb200_step(mdb200_unit **units, int n, int nsteps) {
  std::vector<Unit *> us;
  MDB_TRY(unpack_units(units, n, us));
  for (int s = 0; s < nsteps; s++) MDB_TRY(calculate(us.data(), n));
  return MDB200_OK;
}

int mdb200_is_pairlist_expired(mdb200_unit **units, int n, int *expired) {
  std::vector<Unit *> us;
  MDB_TRY(unpack_units(units, n, us));
  return any_expired(us.data(), n, expired);
}

int mdb200_update_position_half(mdb200_unit *h) {
  if (!h) return MDB200_EINVAL;
  DeviceGuard g(h->u.device);
  MDB_TRY(materialize(&h->u));
  h->u.inner_valid = false;  // positions change outside the budgeted step
  return drift_half(&h->u);
}

int mdb200_send_border_particles(mdb200_unit **units, int n) {
  std::vector<Unit *> us;
  MDB_TRY(unpack_units(units, n, us));
  for (int i = 0; i < n; i++) MDB_TRY(materialize(us[i]));
  return send_border_particles(us.data(), n);
}

int mdb200_calculate_force(mdb200_unit *h) {
  if (!h) return MDB200_EINVAL;
  if (!h->u.list_initialized) { set_error("pair list not built"); return MDB200_EINVAL; }
  DeviceGuard g(h->u.device);
  MDB_TRY(materialize(&h->u));
  h->u.inner_valid = false;  // momenta change outside the budgeted step
  return force_kick(&h->u);
}

int mdb200_heatbath_zeta(mdb200_unit *h, double t) {
  if (!h) return MDB200_EINVAL;
  Unit *u = &h->u;
  const double tau = u->prm.heatbath_tau;
  u->zeta += ((t - u->prm.aimed_temperature) / (tau * tau)) * u->c.dt2;
  return MDB200_OK;
}
int mdb200_heatbath_momenta(mdb200_unit *h) {
  if (!h) return MDB200_EINVAL;
  DeviceGuard g(h->u.device);
  MDB_TRY(materialize(&h->u));
  h->u.inner_valid = false;
  return heatbath_momenta(&h->u);
}
int mdb200_langevin(mdb200_unit *h) {
  if (!h) return MDB200_EINVAL;
  DeviceGuard g(h->u.device);
  MDB_TRY(materialize(&h->u));
  h->u.inner_valid = false;
  return langevin_kick(&h->u);
}
int mdb200_set_seed(mdb200_unit *h, int seed) {
  if (!h) return MDB200_EINVAL;
  h->u.seed = seed;
  return MDB200_OK;
}
int mdb200_get_zeta(mdb200_unit *h, double *z) {
  if (!h || !z) return MDB200_EINVAL;
  *z = h->u.zeta;
  return MDB200_OK;
}
int mdb200_set_zeta(mdb200_unit *h, double z) {
  if (!h) return MDB200_EINVAL;
  h->u.zeta = z;
  return MDB200_OK;
}
int mdb200_set_control_temperature(mdb200_unit *h, int on) {
  if (!h) return MDB200_EINVAL;
  h->u.prm.control_temperature = on ? 1 : 0;
  return MDB200_OK;
}
int mdb200_set_aimed_temperature(mdb200_unit *h, double t) {
  if (!h) return MDB200_EINVAL;
  h->u.prm.aimed_temperature = t;
  return MDB200_OK;
}

int mdb200_observe_unit(mdb200_unit *h, double out[4]) {
  if (!h || !out) return MDB200_EINVAL;
  DeviceGuard g(h->u.device);
  MDB_TRY(materialize(&h->u));
  return observe_unit(&h->u, out);
}

int mdb200_observe(mdb200_unit **units, int n, double out[4]) {
  std::vector<Unit *> us;
  MDB_TRY(unpack_units(units, n, us));
  double tot[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; i++) {
    DeviceGuard g(us[i]->device);
    double o[4];
    MDB_TRY(materialize(us[i]));
    MDB_TRY(observe_unit(us[i], o));
    for (int k = 0; k < 4; k++) tot[k] += o[k];
  }
  MDB_TRY(comm_allreduce_sum(us.data(), n, tot, 4));
  for (int k = 0; k < 4; k++) out[k] = tot[k];
  return MDB200_OK;
}

int mdb200_get_pairs(mdb200_unit *h, int *key, int *partner, long long cap, long long *npairs) {
  if (!h || !npairs) return MDB200_EINVAL;
  DeviceGuard g(h->u.device);
  return extract_pairs(&h->u, key, partner, cap, npairs);
}

int mdb200_get_impulse(mdb200_unit *h, double *f3, long long cap_rows) {
  if (!h || !f3) return MDB200_EINVAL;
  if (!h->u.list_initialized) { set_error("pair list not built"); return MDB200_EINVAL; }
  DeviceGuard g(h->u.device);
  MDB_TRY(materialize(&h->u));
  return force_impulse(&h->u, f3, cap_rows);
}

int mdb200_timers(mdb200_unit *h, double out_ms[6], long long out_counts[2], int reset) {
  if (!h) return MDB200_EINVAL;
  Unit *u = &h->u;
  {
    DeviceGuard g(u->device);
    force_timer_resolve(u);
  }
  if (out_ms) for (int k = 0; k < 6; k++) out_ms[k] = u->t_ms[k];
  if (out_counts) { out_counts[0] = u->launches; out_counts[1] = u->force_launches; }
  if (reset) {
    for (int k = 0; k < 6; k++) u->t_ms[k] = 0.0;
    u->launches = 0;
    u->force_launches = 0;
    derive_timer_resolve(u);
    u->derive_ms = 0.0;
    u->derive_launches = 0;
  }
  return MDB200_OK;
}
int mdb200_set_profiling(mdb200_unit *h, int on) {
  if (!h) return MDB200_EINVAL;
  h->u.profiling = on < 0 ? 0 : (on > 2 ? 2 : on);
  return MDB200_OK;
}
int mdb200_list_stats(mdb200_unit *h, double out[8]) {
  if (!h || !out) return MDB200_EINVAL;
  Unit *u = &h->u;
  DeviceGuard g(u->device);
  for (int k = 0; k < 8; k++) out[k] = 0.0;
  out[0] = (double)u->nfw;            // force-warps
  out[1] = (double)u->list_entries;   // stored 16-bit entries of the outer list (padded)
  out[2] = (double)u->max_row;
  out[4] = (double)kG;
  derive_timer_resolve(u);
  out[6] = u->derive_ms;              // inner-list derivations: device time (profiling level 1) and count
  out[7] = (double)u->derive_launches;
  if (u->pn > 0 && u->list_initialized) {
    std::vector<int> rc((size_t)u->pn);
    MDB_CUDA(cudaMemcpyAsync(rc.data(), u->row_count.ptr, sizeof(int) * rc.size(), cudaMemcpyDeviceToHost, u->stream));
    MDB_CUDA(cudaStreamSynchronize(u->stream));
    double s = 0;
    for (size_t k = 0; k < rc.size(); k++) s += rc[k];
    out[3] = s;                       // sum of outer row lengths (unpadded entries)
    std::vector<int> wl((size_t)u->nfw);
    const int *src = (u->inner_valid && !u->inner_is_outer) ? u->fw_len_in.ptr : u->fw_len.ptr;
    MDB_CUDA(cudaMemcpyAsync(wl.data(), src, sizeof(int) * wl.size(), cudaMemcpyDeviceToHost, u->stream));
    MDB_CUDA(cudaStreamSynchronize(u->stream));
    double w = 0;
    for (size_t k = 0; k < wl.size(); k++) w += wl[k];
    out[5] = 32.0 * w;                // entries the force kernel walks (inner list if there is one; rows padded)
  }
  return MDB200_OK;
}
int mdb200_event_record(mdb200_unit *h, int slot) {
  if (!h || slot < 0 || slot >= 8) return MDB200_EINVAL;
  Unit *u = &h->u;
  DeviceGuard g(u->device);
  if (!u->ev_ready) {
    for (int k = 0; k < 8; k++) MDB_CUDA(cudaEventCreate(&u->ev[k]));
    u->ev_ready = true;
  }
  MDB_CUDA(cudaEventRecord(u->ev[slot], u->stream));
  return MDB200_OK;
}
int mdb200_event_elapsed_ms(mdb200_unit *h, int a, int b, double *ms) {
  if (!h || !ms || a < 0 || a >= 8 || b < 0 || b >= 8 || !h->u.ev_ready) return MDB200_EINVAL;
  Unit *u = &h->u;
  DeviceGuard g(u->device);
  MDB_CUDA(cudaEventSynchronize(u->ev[b]));
  float f = 0.f;
  MDB_CUDA(cudaEventElapsedTime(&f, u->ev[a], u->ev[b]));
  *ms = f;
  return MDB200_OK;
}
// Measured FP64 peak of the device (the denominator of the force kernel's FP64 roofline fraction): 16 independent
// DFMA chains per thread, four 256-thread blocks per SM (64 registers per thread: the chains stay in registers).
__global__ void __launch_bounds__(256, 4) k_fp64_peak(double *out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}
int mdb200_measure_fp64_peak(int device, double *tflops) {
  if (!tflops) return MDB200_EINVAL;
  int ndev = 0;
  MDB_TRY(mdb200_device_count(&ndev));
  if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return MDB200_ECUDA; }
  DeviceGuard g(device);
  int sms = 0;
  MDB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  double *d = nullptr;
  MDB_CUDA(cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  MDB_CUDA(cudaEventCreate(&e0));
  MDB_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, grid = 16 * sms;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    MDB_CUDA(cudaEventRecord(e0, 0));
    k_fp64_peak<<<grid, 256>>>(d, iters, 0.999999, 1e-7);
    MDB_CUDA(cudaEventRecord(e1, 0));
    MDB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    MDB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return MDB200_OK;
}

int mdb200_synchronize(mdb200_unit *h) {
  if (!h) return MDB200_EINVAL;
  DeviceGuard g(h->u.device);
  MDB_TRY(settle_upload(&h->u));
  MDB_CUDA(cudaStreamSynchronize(h->u.stream));
  return MDB200_OK;
}

}  // extern "C"
