// Transport between units.  Three cases per (unit, direction):
//   1. neighbour is a unit of this process (incl. the unit itself through the periodic wrap):
//      the gather/put kernel stores straight into the neighbour's arrays -- on another device of
//      the same process that is a peer store over NVLink (peer access enabled at comm_init);
//   2. neighbour lives in another process: pack kernel -> ncclSend / ncclRecv directly into the
//      ghost window (one group per direction), NCCL resolved at run time so the library has no
//      link-time dependency on a particular libnccl (torch ships its own);
//      -- that is the path of the rebuild (migration, first ghost transfer).  The PER-STEP ghost refresh
//      between processes does not go through NCCL: after every rebuild the processes exchange CUDA IPC handles
//      of their position buffers, and the same gather/put kernel then stores over NVLink straight into the
//      neighbour process's ghost window (both position buffers), as in case 1.  Ordering is kept by two epoch
//      counters per direction in a small IPC-shared mailbox: the sender's last block publishes "arrived" after a
//      system-scope fence, a one-thread kernel on the receiver's stream waits for it, and the receiver
//      acknowledges a round at the start of the next one so that the sender never overwrites ghosts still in use;
//   3. scalar reductions: MPI_Allreduce(LOR/SUM) -> ncclAllReduce on a few words.
// The order of operations per direction is the reference's (mdmanager.cc:195-248, 295-349):
// all units pack, then all units receive.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "comm.h"

namespace mdb {

// ---------------------------------------------------------------------------------------------
// NCCL, loaded lazily
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  bool loaded = false;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.loaded) return MDB200_OK;
  void *h = nullptr;
  // prefer a libnccl the host application already mapped (e.g. torch's bundled copy)
  if (dlsym(RTLD_DEFAULT, "ncclCommInitRank")) h = RTLD_DEFAULT;
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("cannot load libnccl: %s", dlerror());
    return MDB200_ECOMM;
  }
#define LOADSYM(name)                                                      \
  g_nccl.name = (decltype(g_nccl.name))dlsym(h, "nccl" #name);             \
  if (!g_nccl.name) { set_error("libnccl lacks nccl" #name); return MDB200_ECOMM; }
  LOADSYM(GetUniqueId)
  LOADSYM(CommInitRank)
  LOADSYM(CommDestroy)
  LOADSYM(Send)
  LOADSYM(Recv)
  LOADSYM(GroupStart)
  LOADSYM(GroupEnd)
  LOADSYM(AllReduce)
  LOADSYM(AllGather)
  LOADSYM(GetErrorString)
#undef LOADSYM
  g_nccl.loaded = true;
  return MDB200_OK;
}

#define MDB_NCCL(call)                                                                           \
  do {                                                                                           \
    ncclResult_t r__ = (call);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      set_error("NCCL error %s at %s:%d", g_nccl.GetErrorString(r__), __FILE__, __LINE__);       \
      return MDB200_ECOMM;                                                                       \
    }                                                                                            \
  } while (0)

struct CommState {
  int nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  int *d_cnt = nullptr;  // device scratch: [0] send count, [1] recv count
  void *d_cnt64 = nullptr; // device scratch for the {flag, bits} all-reduce
  int users = 0;
};
static CommState g_comm;

// ---------------------------------------------------------------------------------------------
// peer-to-peer ghost refresh between processes (one unit per process)
// ---------------------------------------------------------------------------------------------
struct Mailbox {
  unsigned long long arrive[6];  // written by the sender of direction d: last refresh round whose ghosts have landed
  unsigned long long ack[6];     // written by the receiver of direction d: last round it has finished reading
  unsigned int ticket[6];        // last-block tickets of this unit's put kernels
  unsigned int pad[2];
};
struct Window {                  // what a receiver tells its sender after a rebuild
  cudaIpcMemHandle_t hq, hq2;    // the two position buffers
  long long off, cnt;            // ghost window of the direction
};
struct P2P {
  bool enabled = false, windows_valid = false;
  Mailbox *mine = nullptr;
  Mailbox *nbr_mail[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // unit we send to in dir
  Mailbox *src_mail[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // unit we receive from in dir
  double4 *dst_q[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double4 *dst_q2[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned long long epoch = 0;  // refresh rounds so far (identical on every process)
  std::map<std::string, void *> opened;  // imported allocations by handle bytes
  std::set<std::string> mail_keys;
  Window *d_win = nullptr;       // device scratch: [0] send, [1] recv
};
static P2P g_p2p;

static int p2p_open(const cudaIpcMemHandle_t &h, void **out, std::string *key_out = nullptr) {
  std::string key(reinterpret_cast<const char *>(&h), sizeof(h));
  if (key_out) *key_out = key;
  auto it = g_p2p.opened.find(key);
  if (it != g_p2p.opened.end()) { *out = it->second; return MDB200_OK; }
  void *p = nullptr;
  MDB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  g_p2p.opened[key] = p;
  *out = p;
  return MDB200_OK;
}

// A neighbour that never shows up raises DE_COMM instead of hanging the device (and once the flag is up nobody waits
// any more).  The limit is wall time (globaltimer, nanoseconds; MDB200_P2P_TIMEOUT_S, default 30 s, 0 = wait for
// ever: under a debugger, compute-sanitizer or an oversubscribed host a neighbour may be legitimately slow).
__device__ unsigned long long g_p2p_timeout_ns = 30000000000ull;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long want, const Ctrl *ctrl) {
  const unsigned long long t0 = global_ns(), limit = g_p2p_timeout_ns;
  while (*(const volatile unsigned long long *)flag < want) {
    if ((*(const volatile int *)&ctrl->err & (int)DE_COMM) || (limit && global_ns() - t0 > limit)) return false;
    __nanosleep(64);
  }
  return true;
}

struct MailPtrs { Mailbox *m[6]; };
// start of refresh round `epoch`: everything up to round epoch - 1 has been read here (stream order)
__global__ void k_p2p_ack(MailPtrs src, unsigned long long done_epoch) {
  for (int d = 0; d < 6; d++)
    if (src.m[d]) *(volatile unsigned long long *)&src.m[d]->ack[d] = done_epoch;
  __threadfence_system();
}

// k_ghost_put (exchange.cu) into the neighbour PROCESS's ghost window
__global__ void k_ghost_put_p2p(const double4 *__restrict__ sq, const int *__restrict__ border,
                                double4 *__restrict__ dq, double4 *__restrict__ dq2, int n, double sx, double sy,
                                double sz, Mailbox *mine, Mailbox *peer, int dir, unsigned long long epoch, Ctrl *ctrl) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = spin_until(&mine->ack[dir], epoch - 1, ctrl) ? 1 : 0;  // the window is free again
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (ok && i < n) {
    double4 v = sq[border[i]];
    v = make_double4(v.x + sx, v.y + sy, v.z + sz, v.y + sy);  // mdunit.cc:167; lane 3 repeats y (tile.cuh)
    dq[i] = v;
    dq2[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!ok) atomicOr(&ctrl->err, (int)DE_COMM);
    const unsigned t = atomicAdd(&mine->ticket[dir], 1u);
    if (t == gridDim.x - 1) {  // last block: every block's stores are fenced
      mine->ticket[dir] = 0u;
      __threadfence_system();
      *(volatile unsigned long long *)&peer->arrive[dir] = epoch;
    }
  }
}

__global__ void k_p2p_wait(Mailbox *mine, int dir, unsigned long long epoch, Ctrl *ctrl) {
  if (!spin_until(&mine->arrive[dir], epoch, ctrl)) atomicOr(&ctrl->err, (int)DE_COMM);
}

bool comm_p2p_enabled() { return g_p2p.enabled; }
void comm_p2p_invalidate() { g_p2p.windows_valid = false; }

int comm_unique_id(void *out128) {
  if (!out128) return MDB200_EINVAL;
  MDB_TRY(load_nccl());
  ncclUniqueId id;
  MDB_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  memcpy(out128, &id, 128);
  return MDB200_OK;
}

static bool is_remote(const Unit *u, int dir) { return u->nbr[dir] == nullptr; }
static bool src_remote(const Unit *u, int dir) { return u->src[dir] == nullptr; }
static const int OPP[6] = {1, 0, 3, 2, 5, 4};

// mailboxes: one per process, exported once; all-gathered handles, the neighbours' mailboxes mapped here.
// MDB200_P2P=0 keeps every refresh on NCCL; a process that cannot map a neighbour switches it off everywhere.
static int p2p_init(Unit *u) {
  const char *env = getenv("MDB200_P2P");
  int ok = (env && env[0] == '0') ? 0 : 1;
  cudaStream_t st = u->stream;
  const int nr = g_comm.nranks;
  MDB_CUDA(cudaMalloc(&g_p2p.mine, sizeof(Mailbox)));
  MDB_CUDA(cudaMemsetAsync(g_p2p.mine, 0, sizeof(Mailbox), st));
  MDB_CUDA(cudaMalloc(&g_p2p.d_win, 2 * sizeof(Window)));
  cudaIpcMemHandle_t mh;
  memset(&mh, 0, sizeof(mh));
  if (cudaIpcGetMemHandle(&mh, g_p2p.mine) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  cudaIpcMemHandle_t *d_all = nullptr;
  MDB_CUDA(cudaMalloc(&d_all, sizeof(cudaIpcMemHandle_t) * (size_t)nr));
  MDB_CUDA(cudaMemcpyAsync(d_all + g_comm.rank, &mh, sizeof(mh), cudaMemcpyHostToDevice, st));
  MDB_NCCL(g_nccl.AllGather(d_all + g_comm.rank, d_all, sizeof(mh), ncclChar, g_comm.comm, st));
  std::vector<cudaIpcMemHandle_t> all((size_t)nr);
  MDB_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(mh) * (size_t)nr, cudaMemcpyDeviceToHost, st));
  MDB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_all);
  if (ok) {
    for (int dir = 0; dir < 6 && ok; dir++) {
      if (!is_remote(u, dir)) continue;
      void *a = nullptr, *b = nullptr;
      std::string ka, kb;
      if (p2p_open(all[(size_t)u->prm.neighbor_rank[dir]], &a, &ka) != MDB200_OK ||
          p2p_open(all[(size_t)u->prm.neighbor_rank[OPP[dir]]], &b, &kb) != MDB200_OK) {
        ok = 0;
        cudaGetLastError();
        break;
      }
      g_p2p.nbr_mail[dir] = static_cast<Mailbox *>(a);
      g_p2p.src_mail[dir] = static_cast<Mailbox *>(b);
      g_p2p.mail_keys.insert(ka);
      g_p2p.mail_keys.insert(kb);
    }
  }
  // everybody or nobody
  MDB_CUDA(cudaMemcpyAsync(g_comm.d_cnt + 4, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
  MDB_NCCL(g_nccl.AllReduce(g_comm.d_cnt + 4, g_comm.d_cnt + 4, 1, ncclInt32, ncclMin, g_comm.comm, st));
  MDB_CUDA(cudaMemcpyAsync(&ok, g_comm.d_cnt + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
  MDB_CUDA(cudaStreamSynchronize(st));
  g_p2p.enabled = ok != 0;
  if (const char *ts = getenv("MDB200_P2P_TIMEOUT_S")) {
    const double sec = atof(ts);
    const unsigned long long ns = sec <= 0.0 ? 0ull : (unsigned long long)(sec * 1.0e9);
    MDB_CUDA(cudaMemcpyToSymbol(g_p2p_timeout_ns, &ns, sizeof(ns)));
  }
  if (g_p2p.enabled) {
    u->q.graveyard = &u->deferred_free;
    u->q_alt.graveyard = &u->deferred_free;
  }
  if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] rank %d: peer-to-peer ghost refresh %s\n", g_comm.rank, ok ? "on" : "off");
  return MDB200_OK;
}

// After the border set-up of a rebuild: every receiver tells its sender where its ghost window is now.
int comm_p2p_update(Unit **us, int n) {
  if (!g_p2p.enabled || n != 1) return MDB200_OK;
  Unit *u = us[0];
  cudaSetDevice(u->device);
  cudaStream_t st = u->stream;
  // A process that cannot export or map a window does not fail the run: it keeps taking part in the exchanges
  // below and reports ok = 0 through the closing all-reduce, after which EVERY process falls back to NCCL.
  int ok = 1;
  cudaIpcMemHandle_t hq, hq2;
  memset(&hq, 0, sizeof(hq));
  memset(&hq2, 0, sizeof(hq2));
  if (cudaIpcGetMemHandle(&hq, u->q.ptr) != cudaSuccess || cudaIpcGetMemHandle(&hq2, u->q_alt.ptr) != cudaSuccess) {
    ok = 0;
    cudaGetLastError();
  }
  std::set<std::string> used = g_p2p.mail_keys;
  for (int dir = 0; dir < 6; dir++) {
    g_p2p.dst_q[dir] = g_p2p.dst_q2[dir] = nullptr;
    if (!is_remote(u, dir)) continue;
    Window w[2];
    memset(w, 0, sizeof(w));
    w[0].hq = hq; w[0].hq2 = hq2; w[0].off = u->recv_off[dir]; w[0].cnt = u->recv_cnt[dir];
    MDB_CUDA(cudaMemcpyAsync(g_p2p.d_win, &w[0], sizeof(Window), cudaMemcpyHostToDevice, st));
    MDB_NCCL(g_nccl.GroupStart());
    // my window for direction dir is filled by the unit in the opposite direction; I fill the window of nbr[dir]
    MDB_NCCL(g_nccl.Send(g_p2p.d_win, sizeof(Window), ncclChar, u->prm.neighbor_rank[OPP[dir]], g_comm.comm, st));
    MDB_NCCL(g_nccl.Recv(g_p2p.d_win + 1, sizeof(Window), ncclChar, u->prm.neighbor_rank[dir], g_comm.comm, st));
    MDB_NCCL(g_nccl.GroupEnd());
    MDB_CUDA(cudaMemcpyAsync(&w[1], g_p2p.d_win + 1, sizeof(Window), cudaMemcpyDeviceToHost, st));
    MDB_CUDA(cudaStreamSynchronize(st));
    if (!ok) continue;
    void *a = nullptr, *b = nullptr;
    std::string ka, kb;
    if (w[1].cnt != u->n_border[dir] || p2p_open(w[1].hq, &a, &ka) != MDB200_OK ||
        p2p_open(w[1].hq2, &b, &kb) != MDB200_OK) {
      ok = 0;
      cudaGetLastError();
      continue;
    }
    used.insert(ka);
    used.insert(kb);
    g_p2p.dst_q[dir] = static_cast<double4 *>(a) + w[1].off;
    g_p2p.dst_q2[dir] = static_cast<double4 *>(b) + w[1].off;
  }
  // unmap what the neighbours have replaced, then -- once every process has done so -- free what was replaced here
  if (u->cstream) MDB_CUDA(cudaStreamSynchronize(u->cstream));
  for (auto it = g_p2p.opened.begin(); it != g_p2p.opened.end();) {
    if (used.count(it->first)) { ++it; continue; }
    cudaIpcCloseMemHandle(it->second);
    it = g_p2p.opened.erase(it);
  }
  MDB_CUDA(cudaMemcpyAsync(g_comm.d_cnt + 4, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
  MDB_NCCL(g_nccl.AllReduce(g_comm.d_cnt + 4, g_comm.d_cnt + 4, 1, ncclInt32, ncclMin, g_comm.comm, st));
  MDB_CUDA(cudaMemcpyAsync(&ok, g_comm.d_cnt + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
  MDB_CUDA(cudaStreamSynchronize(st));
  for (void *p : u->deferred_free) cudaFree(p);
  u->deferred_free.clear();
  g_p2p.windows_valid = ok != 0;
  if (!ok) {  // some process could not set its windows up: every process refreshes through NCCL from here on
    g_p2p.enabled = false;
    u->q.graveyard = u->q_alt.graveyard = nullptr;
    for (auto it = g_p2p.opened.begin(); it != g_p2p.opened.end();) {  // nothing of the neighbours stays mapped
      if (g_p2p.mail_keys.count(it->first)) { ++it; continue; }
      cudaIpcCloseMemHandle(it->second);
      it = g_p2p.opened.erase(it);
    }
    if (getenv("MDB200_DEBUG")) fprintf(stderr, "[mdb200] rank %d: peer-to-peer ghost refresh switched off\n", g_comm.rank);
  }
  return MDB200_OK;
}

// Opens one refresh round: tells every sender that the ghosts of the previous round are no longer needed.
int comm_refresh_begin(Unit **us, int n, bool on_comm_stream) {
  if (!g_p2p.enabled || !g_p2p.windows_valid || n != 1) return MDB200_OK;
  Unit *u = us[0];
  cudaSetDevice(u->device);
  cudaStream_t st = (on_comm_stream && u->cstream) ? u->cstream : u->stream;
  g_p2p.epoch++;
  MailPtrs mp;
  for (int d = 0; d < 6; d++) mp.m[d] = is_remote(u, d) ? g_p2p.src_mail[d] : nullptr;
  k_p2p_ack<<<1, 1, 0, st>>>(mp, g_p2p.epoch - 1);
  u->launches++;
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int comm_init(Unit **us, int n, int nranks, int rank, const void *unique_id) {
  std::map<int, Unit *> by_id;
  for (int i = 0; i < n; i++) by_id[us[i]->prm.unit_id] = us[i];
  bool any_remote = false;
  for (int i = 0; i < n; i++) {
    Unit *u = us[i];
    for (int dir = 0; dir < 6; dir++) {
      auto it = by_id.find(u->prm.neighbor_id[dir]);
      u->nbr[dir] = (it == by_id.end()) ? nullptr : it->second;
      auto is = by_id.find(u->prm.neighbor_id[OPP[dir]]);
      u->src[dir] = (is == by_id.end()) ? nullptr : is->second;
      if (!u->nbr[dir] || !u->src[dir]) any_remote = true;
    }
  }
  if (any_remote && nranks <= 1) {
    set_error("a neighbour unit is neither in this process nor reachable (nranks = 1)");
    return MDB200_ECOMM;
  }
  if (any_remote && n != 1) {
    set_error("multi-process runs take exactly one unit per process (got %d)", n);
    return MDB200_EINVAL;
  }
  // units of one process that share a device already share that device's stream (mdb200_create);
  // units on different devices get peer access to each other
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      if (us[i]->device != us[j]->device) {
        int can = 0;
        MDB_CUDA(cudaDeviceCanAccessPeer(&can, us[i]->device, us[j]->device));
        if (!can) {
          set_error("devices %d and %d cannot access each other", us[i]->device, us[j]->device);
          return MDB200_ECOMM;
        }
        MDB_CUDA(cudaSetDevice(us[i]->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(us[j]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) MDB_CUDA(e);
        cudaGetLastError();
      }
  if (nranks > 1) {
    if (!unique_id) { set_error("nranks > 1 needs the NCCL unique id"); return MDB200_EINVAL; }
    MDB_TRY(load_nccl());
    if (!g_comm.comm) {
      ncclUniqueId id;
      memcpy(&id, unique_id, 128);
      MDB_CUDA(cudaSetDevice(us[0]->device));
      MDB_NCCL(g_nccl.CommInitRank(&g_comm.comm, nranks, id, rank));
      MDB_CUDA(cudaMalloc(&g_comm.d_cnt, 8 * sizeof(int)));
      MDB_CUDA(cudaMalloc(&g_comm.d_cnt64, 4 * sizeof(long long)));
      g_comm.nranks = nranks;
      g_comm.rank = rank;
    }
    if (any_remote && !g_p2p.mine) {
      MDB_TRY(p2p_init(us[0]));
    } else if (any_remote && g_p2p.enabled) {
      // a later unit of the same process (the mailboxes and the neighbours stay): its buffers will be mapped too
      us[0]->q.graveyard = &us[0]->deferred_free;
      us[0]->q_alt.graveyard = &us[0]->deferred_free;
      g_p2p.windows_valid = false;
    }
  }
  for (int i = 0; i < n; i++) us[i]->comm_ready = true;
  return MDB200_OK;
}

int comm_release(Unit *u) {
  u->comm_ready = false;
  for (void *p : u->deferred_free) cudaFree(p);
  u->deferred_free.clear();
  u->q.graveyard = u->q_alt.graveyard = nullptr;
  if (g_p2p.enabled && g_comm.nranks > 1) g_p2p.windows_valid = false;  // the windows pointed into this unit's peers
  return MDB200_OK;
}

// exchange one int with the neighbours of direction dir (count hand-shake), host-synchronous
static int exchange_count(Unit *u, int dir, long long send_n, long long *recv_n) {
  int h[2] = {(int)send_n, 0};
  MDB_CUDA(cudaMemcpyAsync(g_comm.d_cnt, h, sizeof(int), cudaMemcpyHostToDevice, u->stream));
  MDB_NCCL(g_nccl.GroupStart());
  MDB_NCCL(g_nccl.Send(g_comm.d_cnt, 1, ncclInt32, u->prm.neighbor_rank[dir], g_comm.comm, u->stream));
  MDB_NCCL(g_nccl.Recv(g_comm.d_cnt + 1, 1, ncclInt32, u->prm.neighbor_rank[OPP[dir]], g_comm.comm, u->stream));
  MDB_NCCL(g_nccl.GroupEnd());
  MDB_CUDA(cudaMemcpyAsync(&h[1], g_comm.d_cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  *recv_n = h[1];
  return MDB200_OK;
}

int comm_migrate(Unit **us, int n, int dir, const long long *n_leave) {
  for (int i = 0; i < n; i++) {
    Unit *u = us[i];
    if (!is_remote(u, dir)) {
      Unit *d = u->nbr[dir];
      cudaSetDevice(d->device);
      MDB_TRY(migrate_unpack(d, u->mig_send_q.ptr, u->mig_send_p.ptr, u->mig_send_t.ptr, n_leave[i], u->late_p ? u->mig_send_s.ptr : nullptr));
      continue;
    }
    cudaSetDevice(u->device);
    long long nr = 0;
    MDB_TRY(exchange_count(u, dir, n_leave[i], &nr));
    MDB_TRY(u->mig_recv_q.ensure((size_t)nr + 1));
    MDB_TRY(u->mig_recv_p.ensure((size_t)nr + 1));
    MDB_TRY(u->mig_recv_t.ensure((size_t)nr + 1));
    const int to = u->prm.neighbor_rank[dir], from = u->prm.neighbor_rank[OPP[dir]];
    MDB_NCCL(g_nccl.GroupStart());
    if (n_leave[i] > 0) {
      MDB_NCCL(g_nccl.Send(u->mig_send_q.ptr, (size_t)n_leave[i] * 4, ncclDouble, to, g_comm.comm, u->stream));
      MDB_NCCL(g_nccl.Send(u->mig_send_p.ptr, (size_t)n_leave[i] * 4, ncclDouble, to, g_comm.comm, u->stream));
      MDB_NCCL(g_nccl.Send(u->mig_send_t.ptr, (size_t)n_leave[i], ncclInt32, to, g_comm.comm, u->stream));
    }
    if (nr > 0) {
      MDB_NCCL(g_nccl.Recv(u->mig_recv_q.ptr, (size_t)nr * 4, ncclDouble, from, g_comm.comm, u->stream));
      MDB_NCCL(g_nccl.Recv(u->mig_recv_p.ptr, (size_t)nr * 4, ncclDouble, from, g_comm.comm, u->stream));
      MDB_NCCL(g_nccl.Recv(u->mig_recv_t.ptr, (size_t)nr, ncclInt32, from, g_comm.comm, u->stream));
    }
    MDB_NCCL(g_nccl.GroupEnd());
    MDB_TRY(migrate_unpack(u, u->mig_recv_q.ptr, u->mig_recv_p.ptr, u->mig_recv_t.ptr, nr));
  }
  return MDB200_OK;
}

int comm_border_setup(Unit **us, int n, int dir) {
  // receive windows first (every unit receives from exactly one source per direction)
  for (int i = 0; i < n; i++) {
    Unit *u = us[i];  // u as RECEIVER
    cudaSetDevice(u->device);
    long long nr = 0;
    if (!src_remote(u, dir)) {
      nr = u->src[dir]->n_border[dir];
    } else {
      MDB_TRY(exchange_count(u, dir, u->n_border[dir], &nr));
    }
    u->recv_off[dir] = u->tn;
    u->recv_cnt[dir] = nr;
    if (u->tn + nr > 2000000000LL) { set_error("tn exceeds 32-bit index range"); return MDB200_ECAPACITY; }
    MDB_TRY(ensure_capacity(u, u->pn, u->tn + nr));
  }
  MDB_TRY(comm_border_refresh(us, n, dir));
  for (int i = 0; i < n; i++) us[i]->tn += us[i]->recv_cnt[dir];
  return MDB200_OK;
}

bool comm_multi_process() { return g_comm.nranks > 1 && g_comm.comm != nullptr; }
int comm_remote_directions(const Unit *u) {
  int k = 0;
  for (int dir = 0; dir < 6; dir++) k += is_remote(u, dir) ? 1 : 0;
  return k;
}

// on_comm_stream: issue everything on the unit's second stream (u->cstream) so that the exchange overlaps
// the interior force kernel running on u->stream (api.cu:calculate)
int comm_border_refresh(Unit **us, int n, int dir, bool on_comm_stream) {
  for (int i = 0; i < n; i++) {
    Unit *u = us[i];  // u as SENDER
    cudaSetDevice(u->device);
    cudaStream_t st = (on_comm_stream && u->cstream) ? u->cstream : u->stream;
    if (!is_remote(u, dir)) {
      Unit *d = u->nbr[dir];
      MDB_TRY(ghost_put_to(u, dir, d->q.ptr + d->recv_off[dir], d->q_alt.ptr + d->recv_off[dir], st));
      continue;
    }
    if (g_p2p.enabled && g_p2p.windows_valid && n == 1) {
      // straight into the neighbour process's ghost window; then wait for our own window of this direction
      const long long nb = u->n_border[dir];
      const unsigned grid = (unsigned)std::max<long long>(1, (nb + 255) / 256);
      k_ghost_put_p2p<<<grid, 256, 0, st>>>(u->q.ptr, u->border[dir].ptr, g_p2p.dst_q[dir], g_p2p.dst_q2[dir], (int)nb,
                                            u->shift[dir][0], u->shift[dir][1], u->shift[dir][2], g_p2p.mine,
                                            g_p2p.nbr_mail[dir], dir, g_p2p.epoch, u->ctrl);
      k_p2p_wait<<<1, 1, 0, st>>>(g_p2p.mine, dir, g_p2p.epoch, u->ctrl);
      u->launches += 2;
      MDB_CUDA(cudaGetLastError());
      continue;
    }
    MDB_TRY(u->send_buf.ensure((size_t)u->n_border[dir] + 1));
    MDB_TRY(ghost_put_to(u, dir, u->send_buf.ptr, nullptr, st));
    const int to = u->prm.neighbor_rank[dir], from = u->prm.neighbor_rank[OPP[dir]];
    MDB_NCCL(g_nccl.GroupStart());
    if (u->n_border[dir] > 0)
      MDB_NCCL(g_nccl.Send(u->send_buf.ptr, (size_t)u->n_border[dir] * 4, ncclDouble, to, g_comm.comm, st));
    if (u->recv_cnt[dir] > 0)
      MDB_NCCL(g_nccl.Recv(u->q.ptr + u->recv_off[dir], (size_t)u->recv_cnt[dir] * 4, ncclDouble, from, g_comm.comm,
                           st));
    MDB_NCCL(g_nccl.GroupEnd());
    // positions are double-buffered: the ghost window lives in both buffers
    if (u->recv_cnt[dir] > 0)
      MDB_CUDA(cudaMemcpyAsync(u->q_alt.ptr + u->recv_off[dir], u->q.ptr + u->recv_off[dir],
                               sizeof(double4) * (size_t)u->recv_cnt[dir], cudaMemcpyDeviceToDevice, st));
  }
  return MDB200_OK;
}

int comm_allreduce_or(Unit **us, int n, int *flag) {
  (void)n;
  if (g_comm.nranks <= 1 || !g_comm.comm) return MDB200_OK;
  Unit *u = us[0];
  cudaSetDevice(u->device);
  int h = *flag;
  MDB_CUDA(cudaMemcpyAsync(g_comm.d_cnt + 2, &h, sizeof(int), cudaMemcpyHostToDevice, u->stream));
  MDB_NCCL(g_nccl.AllReduce(g_comm.d_cnt + 2, g_comm.d_cnt + 2, 1, ncclInt32, ncclMax, g_comm.comm, u->stream));
  MDB_CUDA(cudaMemcpyAsync(&h, g_comm.d_cnt + 2, sizeof(int), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  *flag = h;
  return MDB200_OK;
}

// {expiry flag, error code, max |p|^2 bits}, MAX over the ranks in one collective.  The error code rides along so that a
// rank whose device reported an error (too fast, invalid position, ...) does not simply leave the step while the others
// block in the next collective: every rank sees the largest code and returns.
int comm_allreduce_flag_bits(Unit **us, int n, int *flag, int *err, unsigned long long *bits) {
  (void)n;
  if (g_comm.nranks <= 1 || !g_comm.comm) return MDB200_OK;
  Unit *u = us[0];
  cudaSetDevice(u->device);
  // bit patterns of non-negative doubles order like signed 64-bit integers
  long long h[3] = {(long long)*flag, (long long)*err, (long long)*bits};
  long long *d = reinterpret_cast<long long *>(g_comm.d_cnt64);
  MDB_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, u->stream));
  MDB_NCCL(g_nccl.AllReduce(d, d, 3, ncclInt64, ncclMax, g_comm.comm, u->stream));
  MDB_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  *flag = (int)h[0];
  *err = (int)h[1];
  *bits = (unsigned long long)h[2];
  return MDB200_OK;
}

int comm_allreduce_sum(Unit **us, int n, double *v, int cnt) {
  (void)n;
  if (g_comm.nranks <= 1 || !g_comm.comm) return MDB200_OK;
  if (cnt > 4) return MDB200_EINVAL;
  Unit *u = us[0];
  cudaSetDevice(u->device);
  MDB_CUDA(cudaMemcpyAsync(u->ctrl->obs, v, sizeof(double) * cnt, cudaMemcpyHostToDevice, u->stream));
  MDB_NCCL(g_nccl.AllReduce(u->ctrl->obs, u->ctrl->obs, cnt, ncclDouble, ncclSum, g_comm.comm, u->stream));
  MDB_CUDA(cudaMemcpyAsync(v, u->ctrl->obs, sizeof(double) * cnt, cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  return MDB200_OK;
}

}  // namespace mdb
