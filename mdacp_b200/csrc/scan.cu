// Exclusive prefix sums (hand-written; no CUB/Thrust): the building block of the counting
// sort into cells (meshlist.cc:153-205 keeps `mesh_index` as such a scan), of the stable
// compactions used for migration / border lists, and of the CSR offsets (meshlist.cc:67-72).
//
// out has n+1 entries: out[i] = scale * sum_{k<i} in[k], out[n] = scale * total.
// Three-phase block scan: per-block sums -> recursive scan of the sums -> per-block scan.
// HBM traffic: 2 reads + 1 write of the array, coalesced; grid = ceil(n/2048) blocks.
#include "unit.h"

namespace mdb {

static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 8;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TAcc>
__device__ __forceinline__ TAcc warp_incl_scan(TAcc v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    TAcc t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

template <typename TIn, typename TAcc>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums(const TIn *__restrict__ in, TAcc *__restrict__ sums, long long n) {
  __shared__ TAcc wsum[SCAN_THREADS / 32];
  long long base = (long long)blockIdx.x * SCAN_TILE;
  TAcc s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = base + (long long)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += (TAcc)in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    TAcc t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; w++) t += wsum[w];
    sums[blockIdx.x] = t;
  }
}

// in-place friendly: thread t owns the SCAN_ITEMS consecutive items [base + t*ITEMS, ...)
template <typename TIn, typename TAcc, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_apply(const TIn *in, TOut *out, const TAcc *__restrict__ block_off,
                 long long n, long long n_out, TOut scale) {
  __shared__ TAcc wsum[SCAN_THREADS / 32];
  long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  TAcc v[SCAN_ITEMS];
  TAcc s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = base + k;
    v[k] = (i < n) ? (TAcc)in[i] : (TAcc)0;
    s += v[k];
  }
  TAcc incl = warp_incl_scan<TAcc>(s);
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  TAcc woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += wsum[w];
  TAcc run = (block_off ? block_off[blockIdx.x] : (TAcc)0) + woff + incl - s;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = base + k;
    if (i < n_out) out[i] = (TOut)run * scale;
    run += v[k];
  }
}

size_t scan_workspace_ints(long long n) {
  size_t tot = 0;
  long long m = n + 1;
  while (m > 1) {
    m = (m + SCAN_TILE - 1) / SCAN_TILE;
    tot += (size_t)m + 1;
    if (m == 1) break;
  }
  return tot + 8;
}

template <typename TIn, typename TAcc, typename TOut>
static int scan_rec(const TIn *in, TOut *out, long long n, long long n_out, TAcc *ws, cudaStream_t st,
                    long long *launches, TOut scale) {
  long long nb = (n_out + SCAN_TILE - 1) / SCAN_TILE;
  if (nb <= 0) return MDB200_OK;
  if (nb == 1) {
    scan_block_apply<TIn, TAcc, TOut><<<1, SCAN_THREADS, 0, st>>>(in, out, (const TAcc *)nullptr, n, n_out, scale);
    if (launches) (*launches)++;
  } else {
    TAcc *sums = ws;
    scan_block_sums<TIn, TAcc><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, sums, n);
    if (launches) (*launches)++;
    // exclusive scan of the nb block sums, in place (nb outputs suffice)
    MDB_TRY((scan_rec<TAcc, TAcc, TAcc>(sums, sums, nb, nb, ws + nb + 1, st, launches, (TAcc)1)));
    scan_block_apply<TIn, TAcc, TOut><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, sums, n, n_out, scale);
    if (launches) (*launches)++;
  }
  MDB_CUDA(cudaGetLastError());
  return MDB200_OK;
}

int exclusive_scan_i32(const int *in, int *out, long long n, int *ws, cudaStream_t st, long long *launches) {
  return scan_rec<int, int, int>(in, out, n, n + 1, ws, st, launches, 1);
}

int exclusive_scan_i32_to_i64(const int *in, long long *out, long long n, long long *ws, cudaStream_t st,
                              long long *launches, int scale) {
  return scan_rec<int, long long, long long>(in, out, n, n + 1, ws, st, launches, (long long)scale);
}

}  // namespace mdb
