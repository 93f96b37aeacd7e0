/*
 * Single-process MPI shim -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * The reference (kaityo256/mdacp) needs <mpi.h>; no MPI is installed in this
 * image.  This header provides the twelve MPI entry points the reference
 * uses for the one-rank case, so the unmodified
 * reference sources under /root/reference compile and run as one process
 * whose OpenMP threads are the spatial units.  It is not shipped with, nor
 * linked into, the product library.
 *
 * Semantics that matter:
 *  - size = 1, rank = 0;
 *  - MPI_Sendrecv copies send -> recv (the only peer is the rank itself);
 *  - MPI_Allreduce is a copy of count * sizeof(type) bytes;
 *  - MPI_Dims_create returns a balanced NON-INCREASING factorisation
 *    (parainfo.cc:32-34 reverses the thread dims assuming that order).
 */
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H

#include <string.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
/* datatype handle = element size in bytes (enough for memcpy semantics) */
#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_UNSIGNED_LONG 8
#define MPI_SUM 1
#define MPI_LOR 2
#define MPI_MAX 3

static inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }

static inline double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype stype, int, int,
                               void *rbuf, int rcount, MPI_Datatype rtype, int, int,
                               MPI_Comm, MPI_Status *) {
  size_t sb = (size_t)scount * (size_t)stype;
  size_t rb = (size_t)rcount * (size_t)rtype;
  size_t n = sb < rb ? sb : rb;
  if (n > 0 && sbuf && rbuf) memcpy(rbuf, sbuf, n);
  return MPI_SUCCESS;
}

static inline int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type,
                                MPI_Op, MPI_Comm) {
  memcpy(rbuf, sbuf, (size_t)count * (size_t)type);
  return MPI_SUCCESS;
}

static inline int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf,
                                int, MPI_Datatype, MPI_Comm) {
  if (scount > 0) memcpy(rbuf, sbuf, (size_t)scount * (size_t)stype);
  return MPI_SUCCESS;
}

static inline int MPI_Gather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf, int,
                             MPI_Datatype, int, MPI_Comm) {
  if (scount > 0) memcpy(rbuf, sbuf, (size_t)scount * (size_t)stype);
  return MPI_SUCCESS;
}

static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }

/* Balanced, non-increasing factorisation of nnodes into ndims factors
 * (8 -> 2,2,2 ; 4 -> 2,2,1 ; 2 -> 2,1,1 ; 27 -> 3,3,3 ; 12 -> 3,2,2). */
static inline int MPI_Dims_create(int nnodes, int ndims, int *dims) {
  for (int i = 0; i < ndims; i++) dims[i] = 1;
  int n = nnodes;
  /* peel prime factors from the largest down, always onto the smallest dim */
  int f[64], nf = 0;
  for (int p = 2; p * p <= n; p++)
    while (n % p == 0) { f[nf++] = p; n /= p; }
  if (n > 1) f[nf++] = n;
  for (int k = nf - 1; k >= 0; k--) {
    int m = 0;
    for (int i = 1; i < ndims; i++) if (dims[i] < dims[m]) m = i;
    dims[m] *= f[k];
  }
  /* sort non-increasing */
  for (int i = 0; i < ndims; i++)
    for (int j = i + 1; j < ndims; j++)
      if (dims[j] > dims[i]) { int t = dims[i]; dims[i] = dims[j]; dims[j] = t; }
  return MPI_SUCCESS;
}

#endif
