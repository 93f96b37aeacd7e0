// Transport between units.  Three cases per (unit, direction):
//   1. neighbour is a unit of this process (incl. the unit itself through the periodic wrap):
//      the gather/put kernel stores straight into the neighbour's arrays -- on another device of
//      the same process that is a peer store over NVLink (peer access enabled at comm_init);
//   2. neighbour lives in another process: pack kernel -> ncclSend / ncclRecv directly into the
//      ghost window (one group per direction), NCCL resolved at run time so the library has no
//      link-time dependency on a particular libnccl (torch ships its own);
//   3. scalar reductions: MPI_Allreduce(LOR/SUM) -> ncclAllReduce on a few words.
// The order of operations per direction is the reference's (mdmanager.cc:195-248, 295-349):
// all units pack, then all units receive.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <map>
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
  }
  for (int i = 0; i < n; i++) us[i]->comm_ready = true;
  return MDB200_OK;
}

int comm_release(Unit *u) {
  u->comm_ready = false;
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
      MDB_TRY(migrate_unpack(d, u->mig_send_q.ptr, u->mig_send_p.ptr, u->mig_send_t.ptr, n_leave[i]));
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

int comm_allreduce_flag_bits(Unit **us, int n, int *flag, unsigned long long *bits) {
  (void)n;
  if (g_comm.nranks <= 1 || !g_comm.comm) return MDB200_OK;
  Unit *u = us[0];
  cudaSetDevice(u->device);
  // bit patterns of non-negative doubles order like signed 64-bit integers
  long long h[2] = {(long long)*flag, (long long)*bits};
  long long *d = reinterpret_cast<long long *>(g_comm.d_cnt64);
  MDB_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, u->stream));
  MDB_NCCL(g_nccl.AllReduce(d, d, 2, ncclInt64, ncclMax, g_comm.comm, u->stream));
  MDB_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, u->stream));
  MDB_CUDA(cudaStreamSynchronize(u->stream));
  *flag = (int)h[0];
  *bits = (unsigned long long)h[1];
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
