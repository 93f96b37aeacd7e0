// On-device initial states: what the reference's configuration makers do on the host --
//   ExtractConfigurationMaker::Execute   confmaker.cc:174-252   (Mode=Benchmark: unit-local FCC lattice minus random vacancies)
//   ConfigurationMaker::Execute          confmaker.cc:9-44      (global FCC lattice; langevin.cc:24)
//   BallAdder::Execute                   collision.cc:8-51      (FCC ball with a drift velocity)
//   DropletMaker::Execute                burst.cc:30-64         (global FCC lattice inside a sphere)
//   Variables::SetInitialVelocity        variables.cc:82-108    (directions uniform on the sphere from std::mt19937)
// -- as kernels that leave the particles where mdb200_upload would have put them.  The host generators are serial per
// unit (and the vacancy search is O(n_delete^2) in the reference): 1.4-9 s for the 10-16 M particle states of
// BASELINE.json, against a few ms here.
//
// Order and arithmetic.  The makers walk lattice sites in a fixed loop order and MDUnit::AddParticle (mdunit.cc:44-49)
// appends the sites that lie inside the unit's box; here every site of the walked index range gets a keep flag, an
// exclusive scan of the flags is the particle's index, and a second kernel writes it: the same particles in the same
// order.  Coordinates use the reference's expressions with separately rounded multiplications and additions (the
// parity build has no FMA), so POSITIONS, TYPES AND DRIFT VELOCITIES ARE BIT-IDENTICAL to the host generators'
// (tests/test_gpu_generate.py).  The random velocities reproduce the reference's std::mt19937 stream exactly (one block
// regenerates the 624-word state in three dependent phases, 227 + 227 + 170 words wide) and libstdc++'s
// generate_canonical; cos / sin come from CUDA's libm instead of glibc's, so momenta agree to a few ulp, not bit for
// bit, and the centre-of-mass average is a fixed-order tree sum instead of a serial one.  That is the change of
// provenance SURVEY.md announces for this row; the reference's own generators (host/hostlib.cpp, or confmaker.cc through
// integration/) remain the default of tests and bench.
#include <math.h>
#include <algorithm>
#include <random>
#include <unordered_set>
#include <vector>

#include "unit.h"

namespace mdb {

enum { GEN_BENCHMARK = 0, GEN_LATTICE = 1, GEN_BALL = 2, GEN_DROPLET = 3 };

struct GenSpec {
  int kind;
  int mx, my, mz;        // index range walked on each axis ...
  int ix0, iy0, iz0;     // ... and its first index (sites outside the unit's box on purpose: the flag filters them)
  int nx, ny;            // full loop extents (site number of the benchmark maker's deletion list)
  int n_delete;
  double sp, hs, eps;
  double org[3];         // benchmark: the box's lower corner; ball: its centre; droplet: the sphere's centre
  double r;              // ball / droplet radius
  double s[3], e[3];     // MDRect::IsInside (mdrect.h:33-36): s <= x < e
  double v[3];           // drift velocity of a ball
};

// site t of the walked range: basis b = t & 3 of lattice cell (t >> 2) -> coordinates and whether AddParticle keeps it
__device__ __forceinline__ bool gen_site(const GenSpec &g, long long t, const int *__restrict__ del, double x[3]) {
  const int b = (int)(t & 3);
  long long cell = t >> 2;
  const int ix = g.ix0 + (int)(cell % g.mx);
  cell /= g.mx;
  const int iy = g.iy0 + (int)(cell % g.my);
  const int iz = g.iz0 + (int)(cell / g.my);
  const double fx = __dmul_rn((double)ix, g.sp), fy = __dmul_rn((double)iy, g.sp), fz = __dmul_rn((double)iz, g.sp);
  bool keep = true;
  if (g.kind == GEN_BENCHMARK) {
    // confmaker.cc:209-246: s + i * sp (+ hs); basis (0,0,0) (0,h,h) (h,0,h) (h,h,0)
    x[0] = __dadd_rn(g.org[0], fx); x[1] = __dadd_rn(g.org[1], fy); x[2] = __dadd_rn(g.org[2], fz);
    if (b == 2 || b == 3) x[0] = __dadd_rn(x[0], g.hs);
    if (b == 1 || b == 3) x[1] = __dadd_rn(x[1], g.hs);
    if (b == 1 || b == 2) x[2] = __dadd_rn(x[2], g.hs);
    // confmaker.cc:193-200, 212-216: the site numbers drawn for deletion (sorted; binary search)
    const int n = (int)((((long long)iz * g.ny + iy) * g.nx + ix) * 4 + b);
    int lo = 0, hi = g.n_delete;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (del[mid] < n) lo = mid + 1;
      else hi = mid;
    }
    if (lo < g.n_delete && del[lo] == n) keep = false;
  } else if (g.kind == GEN_LATTICE || g.kind == GEN_DROPLET) {
    // confmaker.cc:22-40 / burst.cc:44-61: i * sp (+ hs) + eps, same basis order
    x[0] = fx; x[1] = fy; x[2] = fz;
    if (b == 2 || b == 3) x[0] = __dadd_rn(x[0], g.hs);
    if (b == 1 || b == 3) x[1] = __dadd_rn(x[1], g.hs);
    if (b == 1 || b == 2) x[2] = __dadd_rn(x[2], g.hs);
    x[0] = __dadd_rn(x[0], g.eps); x[1] = __dadd_rn(x[1], g.eps); x[2] = __dadd_rn(x[2], g.eps);
    if (g.kind == GEN_DROPLET) {
      // burst.cc:17-29 (AddSub): strictly inside the sphere
      const double dx = __dsub_rn(x[0], g.org[0]), dy = __dsub_rn(x[1], g.org[1]), dz = __dsub_rn(x[2], g.org[2]);
      const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (!(r2 < __dmul_rn(g.r, g.r))) keep = false;
    }
  } else {
    // collision.cc:27-47: i * sp - r + c; basis (0,0,0) (h,0,0) (0,h,0) (0,0,h); collision.cc:8-20: r2 > r * r is dropped
    x[0] = __dadd_rn(__dsub_rn(fx, g.r), g.org[0]);
    x[1] = __dadd_rn(__dsub_rn(fy, g.r), g.org[1]);
    x[2] = __dadd_rn(__dsub_rn(fz, g.r), g.org[2]);
    if (b == 1) x[0] = __dadd_rn(x[0], g.hs);
    if (b == 2) x[1] = __dadd_rn(x[1], g.hs);
    if (b == 3) x[2] = __dadd_rn(x[2], g.hs);
    const double dx = __dsub_rn(x[0], g.org[0]), dy = __dsub_rn(x[1], g.org[1]), dz = __dsub_rn(x[2], g.org[2]);
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (r2 > __dmul_rn(g.r, g.r)) keep = false;
  }
  // MDUnit::AddParticle (mdunit.cc:44-49)
  if (!(g.s[0] <= x[0] && x[0] < g.e[0] && g.s[1] <= x[1] && x[1] < g.e[1] && g.s[2] <= x[2] && x[2] < g.e[2])) keep = false;
  return keep;
}

__global__ void __launch_bounds__(256) k_gen_flags(GenSpec g, const int *__restrict__ del, int *__restrict__ flags, long long nsites) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsites) return;
  double x[3];
  flags[t] = gen_site(g, t, del, x) ? 1 : 0;
}

// Variables::AddParticle (variables.cc:31-42): q, p = v, type = 0, appended at pn0 + (number of kept sites before t)
__global__ void __launch_bounds__(256)
k_gen_write(GenSpec g, const int *__restrict__ del, const int *__restrict__ flags, const int *__restrict__ scanv, double4 *__restrict__ q,
            double4 *__restrict__ p, int *__restrict__ type, int *__restrict__ ref_of_slot, int *__restrict__ slot_of_ref, long long nsites,
            long long pn0) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsites || !flags[t]) return;
  double x[3];
  gen_site(g, t, del, x);
  const long long i = pn0 + scanv[t];
  q[i] = make_double4(x[0], x[1], x[2], x[1]);  // lane 3 repeats y on the device (tile.cuh)
  p[i] = make_double4(g.v[0], g.v[1], g.v[2], 0.0);
  type[i] = 0;
  ref_of_slot[i] = (int)i;
  slot_of_ref[i] = (int)i;
}

// ---------------------------------------------------------------------------------------------------------
// std::mt19937(seed): the tempered output words 0 .. nwords-1, in order.  One block: the 624-word state is regenerated
// in three dependent phases (words [0,227) need only the old state, [227,454) the new words 227 places back, [454,624)
// likewise, and word 623 the new word 0), ping-ponging between two copies in shared memory -- three barriers per 624
// words, ~10 ms for the 64 M words of 16 M particles.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned mt_twist(unsigned hi, unsigned lo, unsigned far) {
  const unsigned y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ unsigned mt_temper(unsigned y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__global__ void __launch_bounds__(256) k_mt19937_stream(unsigned seed, unsigned *__restrict__ out, long long nwords) {
  __shared__ unsigned st[2][624];
  const int i = threadIdx.x;
  if (i == 0) {
    unsigned x = seed;
    st[0][0] = x;
    for (unsigned k = 1; k < 624; k++) {
      x = 1812433253u * (x ^ (x >> 30)) + k;
      st[0][k] = x;
    }
  }
  __syncthreads();
  int cur = 0;
  for (long long base = 0; base < nwords; base += 624) {
    const unsigned *o = st[cur];
    unsigned *n = st[cur ^ 1];
    if (i < 227) {
      const unsigned v = mt_twist(o[i], o[i + 1], o[i + 397]);
      n[i] = v;
      if (base + i < nwords) out[base + i] = mt_temper(v);
    }
    __syncthreads();
    if (i < 227) {
      const int j = i + 227;
      const unsigned v = mt_twist(o[j], o[j + 1], n[j - 227]);
      n[j] = v;
      if (base + j < nwords) out[base + j] = mt_temper(v);
    }
    __syncthreads();
    if (i < 170) {
      const int j = i + 454;
      const unsigned v = (j < 623) ? mt_twist(o[j], o[j + 1], n[j - 227]) : mt_twist(o[623], n[0], n[396]);
      n[j] = v;
      if (base + j < nwords) out[base + j] = mt_temper(v);
    }
    __syncthreads();
    cur ^= 1;
  }
}

// libstdc++ std::generate_canonical<double, 53>(std::mt19937): two draws, (w0 + w1 2^32) / 2^64, never 1
__device__ __forceinline__ double gen_canonical(unsigned w0, unsigned w1) {
  const double sum = __dadd_rn((double)w0, __dmul_rn((double)w1, 4294967296.0));
  const double ret = __dmul_rn(sum, 0x1p-64);                  // / 2^64, exact
  return ret >= 1.0 ? 0x1.fffffffffffffp-1 : ret;             // nextafter(1, 0)
}

constexpr int kVelTpb = 256;
// Variables::SetInitialVelocity (variables.cc:84-91): particle i takes draws 2 i and 2 i + 1; per-block sums for the
// centre-of-mass average (fixed order)
__global__ void __launch_bounds__(kVelTpb)
k_gen_velocity(const unsigned *__restrict__ words, double4 *__restrict__ p, long long pn, double v0, double *__restrict__ partial) {
  const long long i = (long long)blockIdx.x * kVelTpb + threadIdx.x;
  double px = 0.0, py = 0.0, pz = 0.0;
  if (i < pn) {
    const uint4 w = reinterpret_cast<const uint4 *>(words)[i];
    const double z = __dsub_rn(__dmul_rn(gen_canonical(w.x, w.y), 2.0), 1.0);
    const double phi = __dmul_rn(__dmul_rn(gen_canonical(w.z, w.w), 2.0), M_PI);
    const double a = __dmul_rn(v0, sqrt(__dsub_rn(1.0, __dmul_rn(z, z))));
    double sn, cs;
    sincos(phi, &sn, &cs);
    px = __dmul_rn(a, cs);
    py = __dmul_rn(a, sn);
    pz = __dmul_rn(v0, z);
    p[i] = make_double4(px, py, pz, 0.0);
  }
  __shared__ double sh[3][kVelTpb / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    px += __shfl_down_sync(0xffffffffu, px, o);
    py += __shfl_down_sync(0xffffffffu, py, o);
    pz += __shfl_down_sync(0xffffffffu, pz, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = px; sh[1][threadIdx.x >> 5] = py; sh[2][threadIdx.x >> 5] = pz; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int k = 0; k < kVelTpb / 32; k++) s += sh[threadIdx.x][k];
    partial[3 * (long long)blockIdx.x + threadIdx.x] = s;
  }
}
// variables.cc:92-102: the average momentum (one block, fixed order)
__global__ void __launch_bounds__(1024) k_gen_velocity_mean(const double *__restrict__ partial, long long nblocks, long long pn, double *__restrict__ mean) {
  __shared__ double sh[3][1024];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long b = threadIdx.x; b < nblocks; b += 1024) { s0 += partial[3 * b]; s1 += partial[3 * b + 1]; s2 += partial[3 * b + 2]; }
  sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) mean[threadIdx.x] = sh[threadIdx.x][0] / (double)pn;
}
// variables.cc:103-107
__global__ void __launch_bounds__(256) k_gen_velocity_centre(double4 *__restrict__ p, long long pn, const double *__restrict__ mean) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  double4 v = p[i];
  v.x = __dsub_rn(v.x, mean[0]); v.y = __dsub_rn(v.y, mean[1]); v.z = __dsub_rn(v.z, mean[2]);
  p[i] = v;
}

// index range [i0, i1) of an axis whose site coordinates a * i + b0 (+ up to hs) can fall into [s, e), with a margin
// (the exact test is the flag's); clipped to [0, n)
static void axis_range(double sp, double b0, double hs, double s, double e, int n, int *i0, int *i1) {
  long long lo = (long long)floor((s - b0 - hs) / sp) - 2, hi = (long long)ceil((e - b0) / sp) + 2;
  lo = std::max<long long>(lo, 0);
  hi = std::min<long long>(hi, n);
  if (hi < lo) hi = lo;
  *i0 = (int)lo;
  *i1 = (int)hi;
}

// appends the kept sites of `g` to the unit's particle store (reference order), returns how many
static int gen_append(Unit *u, GenSpec &g, const std::vector<int> *del_host, long long *added) {
  for (int d = 0; d < 3; d++) { g.s[d] = u->c.s[d]; g.e[d] = u->c.e[d]; }
  *added = 0;
  const long long nsites = 4LL * g.mx * g.my * g.mz;
  if (nsites <= 0) return MDB200_OK;
  if (nsites > 2000000000LL) { set_error("generator: %lld lattice sites exceed the 32-bit index range", nsites); return MDB200_ECAPACITY; }
  cudaStream_t st = u->stream;
  DevBuf<int> flags, scanv, ws, del;
  struct Release {
    DevBuf<int> &a, &b, &c, &d;
    ~Release() { a.release(); b.release(); c.release(); d.release(); }
  } rel{flags, scanv, ws, del};
  MDB_TRY(flags.ensure((size_t)nsites + 2));
  MDB_TRY(scanv.ensure((size_t)nsites + 2));
  MDB_TRY(ws.ensure(2 * scan_workspace_ints(nsites + 2)));
  g.n_delete = del_host ? (int)del_host->size() : 0;
  MDB_TRY(del.ensure((size_t)g.n_delete + 1));
  if (g.n_delete)
    MDB_CUDA(cudaMemcpyAsync(del.ptr, del_host->data(), sizeof(int) * (size_t)g.n_delete, cudaMemcpyHostToDevice, st));
  const unsigned nb = (unsigned)((nsites + 255) / 256);
  k_gen_flags<<<nb, 256, 0, st>>>(g, del.ptr, flags.ptr, nsites);
  u->launches++;
  MDB_TRY(exclusive_scan_i32(flags.ptr, scanv.ptr, nsites, ws.ptr, st, &u->launches));
  int total = 0;
  MDB_CUDA(cudaMemcpyAsync(&total, scanv.ptr + nsites, sizeof(int), cudaMemcpyDeviceToHost, st));
  MDB_CUDA(cudaStreamSynchronize(st));
  if (total > 0) {
    MDB_TRY(reserve_store(u, u->pn + total));
    k_gen_write<<<nb, 256, 0, st>>>(g, del.ptr, flags.ptr, scanv.ptr, u->q.ptr, u->p.ptr, u->type.ptr, u->ref_of_slot.ptr,
                                    u->slot_of_ref.ptr, nsites, u->pn);
    u->launches++;
    MDB_CUDA(cudaGetLastError());
  }
  MDB_CUDA(cudaStreamSynchronize(st));  // the scratch buffers are released on return
  *added = total;
  return store_replaced(u, u->pn + total);
}

// ExtractConfigurationMaker::Execute (confmaker.cc:174-252)
int generate_benchmark(Unit *u, double density, long long *added) {
  const double w[3] = {u->c.e[0] - u->c.s[0], u->c.e[1] - u->c.s[1], u->c.e[2] - u->c.s[2]};
  const double eps = 0.00000001;
  if (!((w[0] - w[1]) / w[0] < eps && (w[0] - w[1]) / w[0] > -eps && (w[0] - w[2]) / w[0] < eps && (w[0] - w[2]) / w[0] > -eps)) {
    set_error("the benchmark configuration needs a cubic unit (confmaker.cc:181-182)");
    return MDB200_EINVAL;
  }
  double sp = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  GenSpec g{};
  g.kind = GEN_BENCHMARK;
  g.mx = g.nx = static_cast<unsigned long int>(w[0] / sp) + 1;
  g.my = g.ny = static_cast<unsigned long int>(w[1] / sp) + 1;
  g.mz = static_cast<unsigned long int>(w[2] / sp) + 1;
  const long long n_full_ll = 4LL * g.mx * g.my * g.mz;
  if (n_full_ll > 2000000000LL) { set_error("generator: %lld lattice sites exceed the 32-bit index range", n_full_ll); return MDB200_ECAPACITY; }
  const int n_full = (int)n_full_ll;
  const int n_exact = static_cast<int>(w[0] * w[1] * w[2] * density);
  const int n_delete = n_full - n_exact;
  // confmaker.cc:190-200: the vacancies.  A serial rejection loop over a few thousand draws of the unit's own
  // std::mt19937 -- host work of microseconds with a hash set (the reference scans a vector: O(n_delete^2))
  std::vector<int> vec_delete;
  if (n_delete > 0) {
    std::unordered_set<int> seen;
    std::mt19937 mt(u->prm.unit_id);
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    vec_delete.reserve((size_t)n_delete);
    while ((int)vec_delete.size() < n_delete) {
      const int v = static_cast<unsigned long int>(n_full * ud(mt));
      if (seen.insert(v).second) vec_delete.push_back(v);
    }
    std::sort(vec_delete.begin(), vec_delete.end());
  }
  sp = w[0] / g.mx;
  g.sp = sp;
  g.hs = sp * 0.5;
  for (int d = 0; d < 3; d++) g.org[d] = u->c.s[d];
  return gen_append(u, g, &vec_delete, added);
}

// ConfigurationMaker::Execute (confmaker.cc:9-44); radius > 0: DropletMaker::Execute (burst.cc:30-64)
int generate_lattice(Unit *u, double density, double radius, long long *added) {
  GenSpec g{};
  g.kind = radius > 0.0 ? GEN_DROPLET : GEN_LATTICE;
  g.sp = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  g.hs = g.sp * 0.5;
  g.eps = 0.0000001;
  g.r = radius;
  const int n[3] = {static_cast<int>(u->prm.L[0] / g.sp), static_cast<int>(u->prm.L[1] / g.sp), static_cast<int>(u->prm.L[2] / g.sp)};
  g.nx = n[0];
  g.ny = n[1];
  int i0[3], i1[3];
  for (int d = 0; d < 3; d++) {
    g.org[d] = u->prm.L[0] * 0.5;  // burst.cc:37: the sphere's centre is L[X] / 2 on all three axes
    axis_range(g.sp, g.eps, g.hs, u->c.s[d], u->c.e[d], n[d], &i0[d], &i1[d]);
  }
  g.ix0 = i0[0]; g.iy0 = i0[1]; g.iz0 = i0[2];
  g.mx = i1[0] - i0[0]; g.my = i1[1] - i0[1]; g.mz = i1[2] - i0[2];
  return gen_append(u, g, nullptr, added);
}

// BallAdder::Execute (collision.cc:22-51)
int generate_ball(Unit *u, const double c[3], double r, double xv, double density, long long *added) {
  GenSpec g{};
  g.kind = GEN_BALL;
  g.sp = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  g.hs = g.sp * 0.5;
  g.r = r;
  g.v[0] = xv;
  const int is = static_cast<int>(2 * r / g.sp);
  g.nx = g.ny = is;
  int i0[3], i1[3];
  for (int d = 0; d < 3; d++) {
    g.org[d] = c[d];
    axis_range(g.sp, c[d] - r, g.hs, u->c.s[d], u->c.e[d], is, &i0[d], &i1[d]);
  }
  g.ix0 = i0[0]; g.iy0 = i0[1]; g.iz0 = i0[2];
  g.mx = i1[0] - i0[0]; g.my = i1[1] - i0[1]; g.mz = i1[2] - i0[2];
  return gen_append(u, g, nullptr, added);
}

// Variables::SetInitialVelocity (variables.cc:82-108)
int generate_initial_velocity(Unit *u, double v0, int seed) {
  const long long pn = u->pn;
  if (pn == 0) return MDB200_OK;
  cudaStream_t st = u->stream;
  // the word stream (16 bytes per particle) is staged in the spare momentum buffer (32 bytes per particle)
  unsigned *words = reinterpret_cast<unsigned *>(u->p_alt.ptr);
  const long long nblocks = (pn + kVelTpb - 1) / kVelTpb;
  DevBuf<double> partial;
  MDB_TRY(partial.ensure((size_t)(3 * nblocks + 3)));
  k_mt19937_stream<<<1, 256, 0, st>>>((unsigned)seed, words, 4 * pn);
  k_gen_velocity<<<(unsigned)nblocks, kVelTpb, 0, st>>>(words, u->p.ptr, pn, v0, partial.ptr);
  k_gen_velocity_mean<<<1, 1024, 0, st>>>(partial.ptr, nblocks, pn, partial.ptr + 3 * nblocks);
  k_gen_velocity_centre<<<(unsigned)((pn + 255) / 256), 256, 0, st>>>(u->p.ptr, pn, partial.ptr + 3 * nblocks);
  u->launches += 4;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  partial.release();
  MDB_CUDA(e);
  u->vmax_ready = false;
  u->inner_valid = false;  // momenta change outside the budgeted step
  return MDB200_OK;
}

}  // namespace mdb
