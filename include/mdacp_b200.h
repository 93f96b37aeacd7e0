/*
 * mdacp_b200.h -- C-ABI of libmdacp_b200.so, the B200 (sm_100a) replacement for the
 * data-parallel hot path of kaityo256/mdacp.
 *
 * The reference has no plugin/FFI layer; its seam is the set of calls that
 * MDManager / MDUnit make into the per-unit numerics ("L2" in SURVEY.md section 1).
 * Every entry point below names the reference interface it replaces
 * (file:line under the reference tree).  INTEGRATION.md shows the one-line forwards
 * a maintainer adds to MDUnit / MDManager to bind them.
 *
 * Conventions
 *  - one opaque handle (mdb200_unit) per reference MDUnit, i.e. per spatial sub-box;
 *    a unit is bound to one CUDA device; several units may share a device;
 *  - plain pointers and sizes only; host arrays use the reference's Variables layout
 *    `double q[][4]`, `double p[][4]`, `int type[]` (variables.h:16-18), lane 3 = padding;
 *  - every call returns 0 on success or an MDB200_E* code; mdb200_last_error() gives text.
 *    The reference prints + exit(1) on the same conditions (mdconfig.h:21); the driver shim
 *    is expected to do that on a non-zero return, so user-visible behaviour is unchanged;
 *  - there is NO CPU fallback: without a CUDA device every compute entry returns MDB200_ECUDA;
 *  - calls on distinct handles are thread-safe; the multi-unit calls (those taking
 *    `mdb200_unit **units, int n`) are issued once by the master thread, like the serial
 *    sections of MDManager (mdmanager.cc:195-349).
 */
#ifndef MDACP_B200_H
#define MDACP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MDB200_ABI_VERSION 1

enum {
  MDB200_OK = 0,
  MDB200_ECUDA = 1,       /* CUDA runtime failure / no device / library built without it   */
  MDB200_ECAPACITY = 2,   /* particle, ghost or pair capacity exceeded (variables.cc:41,
                             meshlist.cc:309 are asserts in the reference)                  */
  MDB200_ETOOFAST = 3,    /* "Too fast particles exists" (pairlist.cc:59-63)                */
  MDB200_EPOSITION = 4,   /* "Invalid position" after periodic wrap (variables.cc:70-77)    */
  MDB200_ECELL = 5,       /* "Invalid index" in cell assignment (meshlist.cc:179-184)       */
  MDB200_EINVAL = 6,      /* bad argument                                                   */
  MDB200_ECOMM = 7        /* NCCL failure / communicator not initialised                    */
};

enum { MDB200_LEFT = 0, MDB200_RIGHT, MDB200_BACK, MDB200_FORWARD, MDB200_DOWN, MDB200_UP };
enum { MDB200_HT_NOSEHOOVER = 0, MDB200_HT_LANGEVIN = 1 }; /* mdconfig.h:28 */

/* POD snapshot of SimulationInfo (simulationinfo.h:20-32) + the compile-time constants of
 * mdconfig.h:15 / simulationinfo.cc:6-7 made run-time + the unit's place in ParaInfo's grid
 * (parainfo.h:28-33).  Zero-initialise, then fill. */
typedef struct mdb200_params {
  double L[3];              /* SimulationInfo::L -- global box                              */
  double time_step;         /* TimeStep                                                      */
  double cutoff;            /* CUTOFF_LENGTH, mdconfig.h:15 (3.0 as shipped)                */
  double buffer_length;     /* BufferLength, simulationinfo.cc:6 (0.3); SearchLength = sum  */
  double aimed_temperature; /* AimedTemperature                                              */
  double heatbath_tau;      /* HeatbathTau                                                   */
  double heatbath_gamma;    /* HeatbathGamma                                                 */
  int control_temperature;  /* ControlTemperature (0/1)                                      */
  int heatbath_type;        /* MDB200_HT_*                                                   */
  int check_list_length;    /* CheckListLength (500)                                         */
  int grid_size[3];         /* ParaInfo::GetGridSize  -- units per axis                      */
  int grid_pos[3];          /* ParaInfo::GetGridPosition(id)                                 */
  int unit_id;              /* MDUnit::GetID                                                 */
  int neighbor_id[6];       /* ParaInfo::GetNeighborID(id, dir), dir = MDB200_LEFT..UP       */
  int neighbor_rank[6];     /* process rank owning that neighbour (ignored for local units)  */
  long long pn_capacity;    /* max local particles of this unit; 0 = sized at first upload   */
  long long tn_capacity;    /* max local+ghost particles; 0 = automatic                      */
} mdb200_params;

typedef struct mdb200_unit mdb200_unit;

/* Library / build information. */
int mdb200_abi_version(void);
const char *mdb200_last_error(void);
int mdb200_device_count(int *count);

/* MDUnit::MDUnit (mdunit.cc:8-27): computes the sub-box from grid_size/grid_pos exactly as
 * mdunit.cc:19-24 and allocates device state on `device`. */
int mdb200_create(const mdb200_params *params, int device, mdb200_unit **out);
/* MDUnit::~MDUnit (mdunit.cc:29-33) */
int mdb200_destroy(mdb200_unit *u);

/* Sub-box of the unit, MDUnit::GetRect (mdunit.h:41): s[3], e[3]. */
int mdb200_get_rect(mdb200_unit *u, double s[3], double e[3]);

/* Host Variables -> device.  Replaces the role of Variables::AddParticle (variables.cc:31-42)
 * after the (unchanged, host-side) initial-state generators ran: particles keep the caller's
 * index order as their "reference index".  pn particles, no ghosts; the pair list becomes
 * uninitialised (pairlist.h:25), so the next mdb200_step rebuilds first, as collision.cc relies on. */
int mdb200_upload(mdb200_unit *u, const double (*q)[4], const double (*p)[4], const int *type,
                  long long pn);
/* As mdb200_upload, but returns as soon as the POSITIONS have arrived: the momenta keep travelling (second stream)
 * while the caller's next mdb200_make_pairlist / mdb200_step builds the pair list, which needs positions only
 * (mdmanager.cc:262-280 reads q alone until the first kick).  Any other entry waits for them first, so results never
 * differ from mdb200_upload; what changes is the contract on the host arrays: `p` must stay valid and unmodified until
 * the next call on this unit has returned.  Worth it for page-locked host arrays only. */
int mdb200_upload_begin(mdb200_unit *u, const double (*q)[4], const double (*p)[4], const int *type,
                        long long pn);
/* Device -> host Variables, in REFERENCE INDEX ORDER: exactly the order the reference's arrays
 * would hold after the same history of migrations (mdunit.cc:78-122).  q receives min(tn, cap_rows)
 * rows (locals, then as many ghosts as fit), p and type receive pn rows.  Any of q/p/type may be
 * NULL.  cap_rows is the row capacity of every non-NULL buffer; MDB200_ECAPACITY if cap_rows < pn. */
int mdb200_download(mdb200_unit *u, double (*q)[4], double (*p)[4], int *type, long long cap_rows,
                    long long *pn, long long *tn);
int mdb200_get_counts(mdb200_unit *u, long long *pn, long long *tn, long long *npairs,
                      long long *nbuilds);
/* Dump path.  The reference's writers -- MDUnit::SaveAsCdview (mdunit.cc:52-67), FileDumper
 * (mdmanager.cc:475-505), Variables::SaveToStream (variables.cc:111-118) -- read q, p, type of the
 * locals from the host Variables in index order.  mdb200_snapshot_begin queues a copy of exactly
 * those arrays (reference index order, end-of-step positions) to library-owned pinned host memory on
 * a second stream and returns at once; steps issued afterwards overlap the transfer and do not
 * change what was captured.  mdb200_snapshot_wait blocks until the copy has landed and hands out
 * read-only views (rows of 4 doubles / ints, pn rows) that stay valid until the next
 * mdb200_snapshot_begin or mdb200_destroy on the unit. */
int mdb200_snapshot_begin(mdb200_unit *u);
int mdb200_snapshot_wait(mdb200_unit *u, const double (**q)[4], const double (**p)[4],
                         const int **type, long long *pn);

/* Join the units of this process (and, if nranks > 1, of the other processes through NCCL)
 * into one simulation.  Replaces MPI_Init + ParaInfo wiring (mdmanager.cc:11-47).
 * nccl_unique_id: 128 bytes obtained from mdb200_comm_unique_id on rank 0 and broadcast by the
 * caller (torch.distributed / MPI / a file); NULL when nranks == 1. */
int mdb200_comm_unique_id(void *out128);
int mdb200_comm_init(mdb200_unit **units, int n, int nranks, int rank, const void *nccl_unique_id);

/* MDManager::MakePairList (mdmanager.cc:262-280): migration (SendParticles, :251-259 with
 * mdunit.cc:78-122 and variables.cc:59-79), ghost set-up (FindBorderParticles mdunit.cc:125-134 +
 * SendBorderParticlesSub mdmanager.cc:295-349), then per unit MDUnit::MakePairList
 * (mdunit.cc:194-198): cell sort (meshlist.cc:153-205), PairList::Init (pairlist.cc:11-23),
 * MeshList::MakeList (meshlist.cc:58-87, 220-271). */
int mdb200_make_pairlist(mdb200_unit **units, int n);

/* MDManager::Calculate x nsteps (mdmanager.cc:116-148): expiry test (pairlist.cc:26-121 +
 * mdmanager.cc:369-377) and rebuild, half drift (fcalculator.cc:22-32), ghost refresh
 * (mdmanager.cc:283-349), then by SimulationInfo: plain kick + half drift (mdmanager.cc:151-157),
 * Nose-Hoover (:160-175) or Langevin (:186-192). */
int mdb200_step(mdb200_unit **units, int n, int nsteps);

/* MDManager::ChangeScale (mdmanager.cc:432-443): MakePairList, scale the global box, every unit's rect
 * (MDRect::ChangeScale, mdrect.cc:55-60), the locals' positions (Variables::ChangeScale,
 * variables.cc:133-139) and the cell grid (MeshList::ChangeScale, meshlist.cc:31-55) by alpha,
 * MakePairList again.  mdb200_get_rect returns the scaled rect afterwards. */
int mdb200_change_scale(mdb200_unit **units, int n, double alpha);

/* The pieces of one step, exposed for parity tests against the matching reference functions. */
int mdb200_is_pairlist_expired(mdb200_unit **units, int n, int *expired); /* mdmanager.cc:369-377 */
int mdb200_update_position_half(mdb200_unit *u);  /* ForceCalculator::UpdatePositionHalf, fcalculator.cc:22-32 */
int mdb200_send_border_particles(mdb200_unit **units, int n); /* MDManager::SendBorderParticles, mdmanager.cc:283-292 */
int mdb200_calculate_force(mdb200_unit *u);       /* ForceCalculator::CalculateForce, fcalculator.cc:35-50 (kick p += F dt) */
int mdb200_heatbath_zeta(mdb200_unit *u, double current_temperature); /* fcalculator.cc:744-750 */
int mdb200_heatbath_momenta(mdb200_unit *u);      /* fcalculator.cc:752-762 */
int mdb200_langevin(mdb200_unit *u);              /* fcalculator.cc:765-778 (per-particle counter RNG) */
int mdb200_set_seed(mdb200_unit *u, int seed);    /* MDUnit::SetSeed, mdunit.h:58 (seed + unit id) */
int mdb200_get_zeta(mdb200_unit *u, double *zeta);
int mdb200_set_zeta(mdb200_unit *u, double zeta);

/* Runtime setters mirroring MDManager::SetControlTemperature / SetAimedTemperature (mdmanager.h:79-80). */
int mdb200_set_control_temperature(mdb200_unit *u, int on);
int mdb200_set_aimed_temperature(mdb200_unit *u, double t);

/* MDManager::ObserveDouble over the three observers (mdmanager.cc:380-387, observer.cc:5-79)
 * plus GetTotalParticleNumber (mdmanager.cc:359-366), fused: out = {K, U, W, N} summed over
 * the given units and all ranks, with the reference's raw definitions:
 * K = sum 0.5 p^2 (all 4 lanes), U and W over the pair list with weight 0.5 for pairs that
 * touch a ghost, stale (half-step) ghost positions (SURVEY.md finding 2).
 * Then T = K/N/1.5, P = (T - W/N) N / V, E = (K+U)/N as mdmanager.cc:390-429. */
int mdb200_observe(mdb200_unit **units, int n, double out[4]);
/* Per-unit raw sums without the cross-unit reduction: out = {K, U, W, pn}. */
int mdb200_observe_unit(mdb200_unit *u, double out[4]);

/* Parity only: the unit's pair list as the reference's raw arrays would hold it as a SET --
 * MeshList::GetKeyParticles / GetPartnerParticles (meshlist.h:46-48), reference indices,
 * key < partner, one entry per listed pair; order is unspecified (sort before comparing). */
int mdb200_get_pairs(mdb200_unit *u, int *key, int *partner, long long cap, long long *npairs);

/* Kick impulse of the current list at the current positions without touching p
 * (what ForceDumper in oracle/ref_harness measures): f[pn][3] in reference index order. */
int mdb200_get_impulse(mdb200_unit *u, double *f3, long long cap_rows);

/* Device-side timing of the phases MDManager::Calculate wraps in StopWatch objects
 * (mdmanager.cc:117-147): accumulated milliseconds since the last reset and launch counts.
 * out_ms = {all, force, comm, pair, integrate, check}; out_counts = {kernels launched, force
 * kernels timed}.
 * Profiling level 1 times only the force kernel, with events resolved at the step's existing
 * host synchronisation (no extra syncs); level 2 times every phase (adds a sync per phase). */
int mdb200_timers(mdb200_unit *u, double out_ms[6], long long out_counts[2], int reset);
/* Enable per-phase CUDA-event timing (adds event records to the stream; default off). */
int mdb200_set_profiling(mdb200_unit *u, int on);
/* Diagnostics of the device neighbour lists: out = {force warps, stored 16-bit entries of the outer list (padded),
 * longest row, unpadded outer entries, lanes per local, entries the force kernel walks (inner list, padded),
 * device ms spent deriving inner lists (profiling level 1), number of derivations}; the last two are reset by
 * mdb200_timers(reset). */
int mdb200_list_stats(mdb200_unit *u, double out[8]);
/* Measured FP64 throughput of a device in TFLOP/s (a DFMA loop; ~10 ms): the denominator of the force kernel's
 * FP64 roofline fraction, which BASELINE.json only quotes as a nominal ~37. */
int mdb200_measure_fp64_peak(int device, double *tflops);
/* CUDA-event stopwatch on the unit's stream (the stream every kernel of the unit is launched on):
 * record slot 0..7, then elapsed milliseconds between two recorded slots (synchronises on b). */
int mdb200_event_record(mdb200_unit *u, int slot);
int mdb200_event_elapsed_ms(mdb200_unit *u, int slot_a, int slot_b, double *ms);
/* Block until all work queued for the unit has finished. */
int mdb200_synchronize(mdb200_unit *u);

/* ---- On-device initial states (mdacp_b200/csrc/generate.cu) -------------------
 * The reference's configuration makers run on the host, serially per unit, and hand every site to
 * MDUnit::AddParticle (mdunit.cc:36-49).  These entries do the same on the device: each APPENDS the particles the
 * maker would add to this unit (the sites inside its box, in the maker's loop order, type 0) to the unit's particle
 * store, exactly as if the host maker had run and mdb200_upload had followed.  Positions, types and drift
 * velocities are bit-identical to the host makers'; mdb200_generate_initial_velocity reproduces the reference's
 * std::mt19937 draws exactly but takes cos / sin from CUDA's libm (a few ulp from glibc's), so it is an ALTERNATIVE
 * provenance of the initial momenta, not a bit-for-bit one.  A unit that has stepped must be cleared first
 * (MDB200_EINVAL otherwise).  `added` (may be NULL) receives the number of particles appended. */
/* Variables::Variables (variables.cc:8-17): empty store, particle_number = 0. */
int mdb200_clear(mdb200_unit *u);
/* ExtractConfigurationMaker::Execute (confmaker.cc:174-252), Mode=Benchmark: the unit's own FCC lattice minus
 * n_full - int(volume * density) vacancies drawn from std::mt19937(unit id). */
int mdb200_generate_benchmark(mdb200_unit *u, double density, long long *added);
/* radius <= 0: ConfigurationMaker::Execute (confmaker.cc:9-44), the global FCC lattice of the system box L;
 * radius > 0: DropletMaker::Execute (burst.cc:30-64), its sites strictly inside the sphere around L[X]/2. */
int mdb200_generate_lattice(mdb200_unit *u, double density, double radius, long long *added);
/* BallAdder::Execute (collision.cc:8-51): FCC ball of radius r around c with momentum (xv, 0, 0). */
int mdb200_generate_ball(mdb200_unit *u, const double c[3], double r, double xv, double density, long long *added);
/* Variables::SetInitialVelocity (variables.cc:82-108) via MDUnit::SetInitialVelocity (mdunit.h:49, seed = unit id):
 * |p| = v0, directions uniform on the sphere, centre-of-mass momentum of the unit removed. */
int mdb200_generate_initial_velocity(mdb200_unit *u, double v0, int seed);

#ifdef __cplusplus
}
#endif
#endif /* MDACP_B200_H */
