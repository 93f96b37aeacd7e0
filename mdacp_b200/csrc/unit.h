// Internal state of one spatial unit (one reference MDUnit) on one B200.
// Not part of the C-ABI (include/mdacp_b200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../include/mdacp_b200.h"

namespace mdb {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define MDB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      mdb::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__,      \
                     __LINE__, #call);                                                      \
      return MDB200_ECUDA;                                                                  \
    }                                                                                       \
  } while (0)
#define MDB_TRY(call)                  \
  do {                                 \
    int r__ = (call);                  \
    if (r__ != MDB200_OK) return r__;  \
  } while (0)

// device-side error word (bit set by kernels, read back at sync points)
enum DevErr : int { DE_TOOFAST = 1, DE_POSITION = 2, DE_CELL = 4, DE_CAPACITY = 8, DE_COMM = 16 };

// ---------------------------------------------------------------------------------------------
// growable device buffer
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T *ptr = nullptr;
  size_t cap = 0;
  // set for buffers other processes have mapped (comm.cu, peer-to-peer ghost exchange): a replaced allocation
  // is parked here and freed once the importers have closed it
  std::vector<void *> *graveyard = nullptr;
  __attribute__((noinline)) int ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return MDB200_OK;
    // generous head-room: re-allocation (cudaMalloc + cudaFree, both device-synchronising) inside the step
    // loop costs far more than the memory
    size_t ncap = n + n / 4 + 64;
    T *np = nullptr;
    if (cap && getenv("MDB200_DEBUG"))
      fprintf(stderr, "[mdb200] realloc %zu -> %zu elements of %zu bytes (requested %zu, caller %p)\n", cap, ncap,
              sizeof(T), n, __builtin_return_address(0));
    MDB_CUDA(cudaMalloc(&np, ncap * sizeof(T)));
    if (keep && ptr && cap) {
      MDB_CUDA(cudaMemcpyAsync(np, ptr, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
      MDB_CUDA(cudaStreamSynchronize(st));
    }
    if (ptr) {
      if (graveyard) graveyard->push_back(ptr);
      else cudaFree(ptr);
    }
    ptr = np;
    cap = ncap;
    return MDB200_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// constants every kernel needs, passed by value
struct Consts {
  double L[3];
  double s[3], e[3];     // sub-box (mdunit.cc:19-24)
  double dt, dt2;        // TimeStep, TimeStep*0.5 (fcalculator.cc:23)
  double CL2, SL2, SL;   // cutoff^2, search^2, search length
  double C2, C0;         // variables.cc:21-27
  double BL;             // BufferLength
  // cell grid: interior ms[d] cells of edge w/ms >= SL plus one halo layer (meshlist.cc:31-55)
  int ms[3];
  int m[3];              // ms + 2
  double inv_edge[3];
  int ncell;
};

// scalar device-resident control block (one per unit)
struct Ctrl {
  double buffer_length;          // PairList::buffer_length (pairlist.h:15)
  double c_d2;                   // (BufferLength - sqrt(dr2_max))^2, pairlist.cc:88-89
  unsigned long long vmax2_bits; // max |p|^2 as ordered bits
  unsigned long long dr2max_bits;
  unsigned long long last_vmax2_bits;  // max |p|^2 the last expiry test saw (host: inner-list budget)
  int phase;                     // 1 = displacement check active in this call
  int cand_count;
  int expired;
  int err;                       // DevErr bits
  unsigned int ticket[4];        // last-block tickets
  int too_fast_index;
  unsigned long long npairs;     // reference-style pair count of the current list
  int max_row;
  double obs[4];                 // K,U,W scratch
  double zeta;                   // Variables::Zeta (variables.h:19)
  int scratch[8];
};

struct Unit {
  mdb200_params prm;
  Consts c;
  int device = 0;
  cudaStream_t stream = 0;
  bool own_stream = false;

  long long pn = 0, tn = 0;       // Variables::particle_number / total_particle_number
  long long cap_p = 0, cap_t = 0; // capacities (rows); sentinel particle lives at q[cap_t]
  bool list_initialized = false;  // PairList::initialized (pairlist.h:25)
  long long nbuilds = 0;          // MeshList::number_of_constructions
  long long lifetime = 0;         // PairList::lifetime
  long long step_count = 0;       // number of completed Calculate() calls
  int seed = 0;                   // MDUnit::SetSeed
  double zeta = 0.0;              // Variables::Zeta (variables.h:19)
  bool vmax_ready = false;        // ctrl->vmax2_bits already holds max |p|^2 of the current momenta
  // speculative pre-drift (force.cu MODE 3): q holds the locals ALREADY advanced by the next step's first half
  // drift; q_alt still holds them at the last half step, from which the end-of-step positions are
  // q_alt + p dt/2 (bit-identical).  api.cu:materialize() undoes it on demand.
  // 1 = NVE (end-of-step positions = q_alt + p dt/2), 2 = Langevin (the step has no second half drift: the
  // end-of-step positions are q_alt itself)
  int predrifted = 0;

  // particle store, SLOT order: locals [0,pn) cell-sorted, ghosts [pn,tn) in arrival order
  DevBuf<double4> q, q_alt;       // positions (x,y,z,0)
  DevBuf<double4> p, p_alt;       // momenta of locals (px,py,pz,0)
  DevBuf<int> type, type_alt;     // Variables::type of locals
  DevBuf<double4> q0;             // PairList::qb_old snapshot (slot order)
  DevBuf<int> ref_of_slot;        // reference index of the local in slot s
  DevBuf<int> slot_of_ref;        // inverse
  DevBuf<int> ord, ord_alt;       // rebuild scratch: handles (slots) in reference order
  DevBuf<int> flags, scan;        // rebuild scratch
  DevBuf<int> scan_ws;            // scan workspace
  DevBuf<long long> scan_ws64;    // scan workspace (64-bit offsets)
  long long scratch_tail = 0;     // first free slot while arrivals are appended during a rebuild

  // cells
  DevBuf<int> cell_of;            // per item
  DevBuf<int> cell_count, cell_start, cell_cursor, cell_items;     // locals (items = ref idx)
  DevBuf<int> gcell_count, gcell_start, gcell_cursor, gcell_items; // ghosts (items = slot)

  // ghosts: per direction send list (slots) and receive window
  DevBuf<int> border[6];
  long long n_border[6] = {0, 0, 0, 0, 0, 0};
  long long recv_off[6] = {0, 0, 0, 0, 0, 0}, recv_cnt[6] = {0, 0, 0, 0, 0, 0};
  // one-unit periodic box: every ghost traced back to the LOCAL it is an image of (exchange.cu:ghost_compose), so that
  // the per-step refresh is one launch instead of six dependent ones
  DevBuf<int> ghost_src;          // per ghost (index - pn): source local slot
  DevBuf<double4> ghost_shift;    // per ghost: the sum of the periodic shifts on its way (one per axis at most)
  bool ghost_composed = false;
  bool expiry_disp_skipped = false;  // the last expiry_check_launch left the displacement kernels out (integrate.cu)
  double shift[6][3];             // mdunit.cc:141-163
  DevBuf<double4> send_buf;       // pack buffer for off-device neighbours
  DevBuf<double4> mig_send_q, mig_send_p, mig_recv_q, mig_recv_p;
  DevBuf<int> mig_send_t, mig_recv_t;

  // ---- tiles (tile.cuh) ------------------------------------------------------------------------------------
  // The interior cells of the unit are grouped into TILES of tdim[0] x tdim[1] x tdim[2] cells (at most 3 per
  // axis).  A force / search / observe block owns one tile: it stages the positions of the tile's cells and of
  // the one-cell halo around them in shared memory (32-byte records, ~3,150 of them for 3x3x3 cells at density
  // 0.7) and all partner indices are 16-bit indices into that STAGED array.  The staged order -- interior cells,
  // then the halo's local cells, then the halo's ghost cells -- follows from cell_start / gcell_start alone, so
  // it is the same in every kernel between two rebuilds.
  int tdim[3] = {1, 1, 1};        // cells per tile edge
  int tnum[3] = {0, 0, 0};        // tiles per axis
  int ntiles = 0;
  int tile_cap = 0;               // staged records of the largest tile (sizes the kernels' shared memory)
  int tile_maxint = 0;            // interior locals of the largest tile
  DevBuf<int> tile_fw0;           // first force-warp of tile t (ntiles + 1 entries)
  DevBuf<int> tile_flag;          // 1 = the tile stages ghosts (boundary tile)
  DevBuf<int> tile_order;         // the tiles every launch goes through: those with interior locals ("active"), the ones
                                  // that stage no ghost first, then the boundary tiles (overlapped exchange, api.cu)
  DevBuf<int> tile_tmp;           // rebuild scratch of the ordering
  int n_interior_tiles = 0;       // active tiles that stage no ghost
  int n_active_tiles = 0;         // tiles with rows (a droplet in vacuum: a few per cent of ntiles)
  // FULL neighbour lists (every local's row holds all its partners), kG = 4 lanes per local: force-warp w of
  // a tile owns the tile's interior locals with staged index [8 w', 8 w' + 8); entry e of the local in group m
  // is handled by lane 4 m + (e & 3) at step e >> 2.  Storage: 16-bit entries, four steps of a lane in one
  // 8-byte word, word (c, lane) of the row block at uint2 index fw_off[w] + 32 c + lane -- so a local's 16
  // entries of chunk c are one aligned 32-byte sector.  Rows are padded with the tile's sentinel record to
  // the force-warp's step count (even), storage to a multiple of 4 steps.
  //   list     OUTER list: every partner within the search length, in ARRIVAL order (tile_search.cu); walked
  //            by the observers and by the derivation below
  //   list_in  INNER list: the entries of `list` within cutoff + inner_delta when it was last derived
  //            (force.cu: k_tile_derive), re-dealt in "rainbow" order (partner index mod 4 = lane in group, so
  //            that a quarter-warp's shared-memory reads are bank-conflict free), same row offsets, own step
  //            counts.  The force kernel always walks this list; api.cu keeps its velocity budget.
  long long nfw = 0;              // force-warps of the unit = sum over tiles of ceil(interior locals / 8)
  DevBuf<int> row_count;          // outer entries per local slot
  DevBuf<int> fw_len, fw_len_in;  // steps per force-warp (even)
  DevBuf<int> fw_sent;            // staged index of the sentinel record of the force-warp's tile (k_redeal)
  DevBuf<long long> fw_off;       // first uint2 word of the force-warp's row block (both lists)
  DevBuf<uint2> list, list_in;
  bool inner_valid = false;
  bool outer_rainbow = false;     // the outer list has been re-dealt into rainbow order since the last rebuild
  bool inner_is_outer = false;    // keep >= search length: the force kernel walks the (re-dealt) outer list itself
  double inner_left = -1.0, inner_delta = 0.0;  // budget of the unit GROUP, kept in its first unit (api.cu)
  double inner_keep = 0.0;        // radius the present inner list was derived with
  long long list_entries = 0;     // stored 16-bit entries (padded)
  long long npairs = 0;           // reference-style count of listed pairs (meshlist.cc:309)
  int max_row = 0;                // longest row (upper bound after a two-pass build)
  int max_row_exact = 0;          // longest row seen so far (sizes the fixed-capacity row blocks of the next build)

  // expiry check
  DevBuf<double> disp;            // PairList::disp
  DevBuf<int> cand;               // candidate list (pairlist.cc:86-97)

  // reductions
  DevBuf<double> partial;         // per-block partial sums

  Ctrl *ctrl = nullptr;           // device
  Ctrl *ctrl_host = nullptr;      // pinned mirror

  // neighbours resolved by mdb200_comm_init: pointer if the unit lives in this process
  Unit *nbr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // destination when sending in dir
  Unit *src[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // source when receiving in dir (neighbour in opposite dir)
  bool comm_ready = false;
  std::vector<void *> deferred_free;  // replaced q / q_alt allocations still mapped by neighbour processes
  // Momenta still on their way from the host (mdb200_upload_begin): they travel on `ustream` into p while the first
  // pair-list build -- which needs positions only -- runs on `stream`.  A single unit of a single process carries the
  // upload row of every local through migration and cell sort (psrc) and gathers the momenta into slot order once
  // they have arrived; everybody else just waits for the copy first (api.cu: settle_upload / make_pairlist).
  cudaStream_t ustream = 0;
  cudaEvent_t ev_q_up = 0, ev_p_up = 0, ev_perm = 0, ev_p_ready = 0;
  bool p_pending = false;         // the copy of the momenta has been queued and nobody has waited for it yet
  bool late_p = false;            // this rebuild carries psrc instead of p
  DevBuf<int> psrc, psrc_alt, mig_send_s;
  cudaStream_t cstream = 0;       // second stream for the overlapped ghost exchange (created on demand)
  cudaEvent_t ev_drift = 0, ev_ghost = 0;

  // asynchronous snapshot for the dump path (api.cu: mdb200_snapshot_begin / _wait): reference-ordered copies
  // of q, p, type staged on the device, then copied to pinned host memory on a second stream while the step
  // loop goes on
  DevBuf<double4> snap_q, snap_p;
  DevBuf<int> snap_t;
  double4 *snap_hq = nullptr, *snap_hp = nullptr;   // pinned host
  int *snap_ht = nullptr;
  long long snap_hcap = 0, snap_pn = 0;
  cudaStream_t snap_stream = 0;
  cudaEvent_t snap_ready = 0, snap_done = 0;
  bool snap_pending = false;

  // profiling
  int profiling = 0;               // 0 off, 1 force kernel only (no extra syncs), 2 every phase
  double t_ms[6] = {0, 0, 0, 0, 0, 0};
  long long launches = 0;
  long long force_launches = 0;    // force kernels timed so far (level >= 1)
  cudaEvent_t ev[8];               // user stopwatch slots
  bool ev_ready = false;
  cudaEvent_t fev[2];              // light force timer
  bool fev_ready = false, fev_pending = false;
  cudaEvent_t dev[2];              // the same for the inner-list derivation
  bool dev_ready = false, dev_pending = false;
  double derive_ms = 0.0;
  long long derive_launches = 0;
};


// lanes that share one local in the force / observe kernels, locals per force-warp
constexpr int kG = 4;
constexpr int kPW = 32 / kG;

// ---------------------------------------------------------------------------------------------
// kernels / phases (host launchers; all asynchronous on u->stream unless stated)
// ---------------------------------------------------------------------------------------------
// scan.cu
size_t scan_workspace_ints(long long n);
int exclusive_scan_i32(const int *in, int *out, long long n, int *ws, cudaStream_t st, long long *launches);
int exclusive_scan_i32_to_i64(const int *in, long long *out, long long n, long long *ws, cudaStream_t st,
                              long long *launches, int scale);

// cells.cu
int ensure_capacity(Unit *u, long long need_p, long long need_t);
// api.cu: room for pn locals in the particle store (contents kept) / bookkeeping after the store was replaced
int reserve_store(Unit *u, long long pn);
int store_replaced(Unit *u, long long pn);
// generate.cu: on-device initial states (confmaker.cc, collision.cc, burst.cc, variables.cc:82-108)
int generate_benchmark(Unit *u, double density, long long *added);
int generate_lattice(Unit *u, double density, double radius, long long *added);
int generate_ball(Unit *u, const double c[3], double r, double xv, double density, long long *added);
int generate_initial_velocity(Unit *u, double v0, int seed);
int rebuild_begin_order(Unit *u);                                        // ord[r] = slot_of_ref[r]
int migrate_flag_and_count(Unit *u, int dir, long long *n_leave);        // sync
int migrate_compact_pack(Unit *u, int dir, long long n_leave);           // keepers compacted, leavers packed
int migrate_unpack(Unit *dst, const double4 *rq, const double4 *rp, const int *rt, long long n, const int *rs = nullptr);
int wrap_periodic(Unit *u);
int cell_sort_locals(Unit *u);
int ghost_compose(Unit *u);          // after the six border set-ups of a rebuild (one unit that is all its own neighbours)
int ghost_refresh_composed(Unit *u, cudaStream_t st);
int find_border(Unit *u, int dir, long long *n_border);                  // sync
int bin_ghosts(Unit *u);
int snapshot_positions(Unit *u);

// tile_search.cu
int build_pair_list(Unit *u);                                            // sync (tile census, list size)
int extract_pairs(Unit *u, int *key, int *partner, long long cap, long long *npairs); // sync

// exchange.cu
int ghost_put(Unit *srcu, int dir, Unit *dstu, long long dst_off);       // dst.q[off+i] = src.q[border[i]] + shift
int ghost_pack(Unit *srcu, int dir);                                     // into send_buf

// integrate.cu
int drift_half(Unit *u);
int drift_half_buffer(Unit *u, double4 *qbuf);                           // the same on a given position buffer
int heatbath_momenta(Unit *u);
int langevin_kick(Unit *u);
int expiry_check_finish(Unit *u);   // after the stream synchronisation that follows expiry_check_launch
int expiry_check_local(Unit *u, int *expired);                           // sync

// force.cu
// inner list = the outer entries within keep_len of the CURRENT positions (rainbow order); sets inner_valid
int derive_inner(Unit *u, double keep_len);
int force_kick(Unit *u);                                                 // derives a full inner list if there is none
// fused kick + second half drift (+ the next step's first half drift if predrift); swaps q / q_alt.
// part: 0 every tile, 1 interior tiles, 2 boundary tiles.  Walks the inner list (derived on demand).
int force_kick_drift(Unit *u, bool predrift, int part, double derive_keep = 0.0);
int ensure_rainbow(Unit *u);                          // outer list into rainbow order (once per rebuild)
bool force_can_derive(const Unit *u, double keep_len, bool split);  // the step's force launch(es) can write the inner list on the side
int force_kick_langevin(Unit *u, double derive_keep = 0.0);                   // fused Langevin step (+ next first half drift)
int force_impulse(Unit *u, double *f3_host, long long cap_rows);         // sync

// observe.cu
int observe_unit(Unit *u, double out[4]);                                // sync

}  // namespace mdb

struct mdb200_unit {
  mdb::Unit u;
};
