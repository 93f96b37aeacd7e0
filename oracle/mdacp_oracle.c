/*
 * mdacp_oracle.c -- plain-C restatement of the reference hot path (kaityo256/mdacp).
 *
 * TEST INFRASTRUCTURE ONLY (see mdacp_oracle.h).  Every function cites the reference
 * file:line it follows.  The loops keep the reference's operation ORDER (pair generation
 * order, software-pipelined force loop, six-direction exchange) so that results are
 * bit-identical with the scalar x86-64 build of the reference (no FMA contraction:
 * compile with -ffp-contract=off).
 */
#include "mdacp_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ND 4
enum { D_LEFT, D_RIGHT, D_BACK, D_FORWARD, D_DOWN, D_UP, MAX_DIR };
static const int OPP[MAX_DIR] = {D_RIGHT, D_LEFT, D_FORWARD, D_BACK, D_UP, D_DOWN}; /* mdconfig.h:27 */
static const int DIFF[MAX_DIR][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};

static char g_err[256] = "";
const char *orc_last_error(void) { return g_err; }
enum { ORC_OK = 0, ORC_ETOOFAST = 3, ORC_EPOSITION = 4, ORC_ECELL = 5, ORC_ECAPACITY = 2 };

typedef struct { double q[ND], p[ND]; int type; } pinfo_t; /* ParticleInfo, mdconfig.h:38-60 */

/* std::mt19937 (32-bit Mersenne twister, standard parameters) */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;
static void mt_seed(mt19937_t *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt19937_t *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
/* libstdc++ std::generate_canonical<double,53>(mt19937): two 32-bit draws */
static double mt_canonical(mt19937_t *g) {
  const double r = 4294967296.0;
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; k++) {
    sum += (double)mt_next(g) * tmp;
    tmp *= r;
  }
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}
/* libstdc++ std::normal_distribution<double>: Marsaglia polar, second variate cached */
typedef struct { double mean, stddev, saved; int saved_available; } normal_t;
static double normal_draw(normal_t *nd, mt19937_t *g) {
  double ret;
  if (nd->saved_available) {
    nd->saved_available = 0;
    ret = nd->saved;
  } else {
    double x, y, r2;
    do {
      x = 2.0 * mt_canonical(g) - 1.0;
      y = 2.0 * mt_canonical(g) - 1.0;
      r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    const double mult = sqrt(-2 * log(r2) / r2);
    nd->saved = x * mult;
    nd->saved_available = 1;
    ret = y * mult;
  }
  return ret * nd->stddev + nd->mean;
}

typedef struct {
  int id, gpos[3];
  double s[ND], e[ND]; /* MDRect */
  /* Variables (variables.h:10-38) */
  int pn, tn;
  int *type;
  double (*q)[ND], (*p)[ND];
  double zeta;
  /* MeshList (meshlist.h:10-60) */
  long number_of_pairs;
  double mesh_size[3];
  int mx, my, mz, number_of_mesh;
  int *particle_position, *key_particles, *partner_particles, *mesh_index, *mesh_index2, *mesh_particle_number;
  int *sortbuf, *key_pointer, *key_pointer2, *number_of_partners, *sorted_list;
  int number_of_constructions;
  int *vbuf; /* SearchMesh's gathered id vector */
  /* PairList (pairlist.h:10-30) */
  int lifetime, initialized;
  double buffer_length;
  double (*qb_old)[ND];
  double *disp;
  int *cand;
  /* MDUnit */
  int *border[MAX_DIR];
  int nborder[MAX_DIR];
  pinfo_t *send_buffer;
  int nsend;
  mt19937_t mt;
} unit_t;

struct orc_system {
  orc_params prm;
  double SL, C0, C2;
  int nunits;
  unit_t *u;
  double s_time;
  long pair_cap;
  int err;
};

/* ---- ParaInfo for one process (parainfo.cc:124-275): id = ix + iy*gx + iz*gx*gy ------------- */
static int neighbor_id(const orc_system *S, int id, int dir) {
  const int *g = S->prm.grid;
  int pos[3] = {id % g[0], (id / g[0]) % g[1], id / g[0] / g[1]};
  for (int d = 0; d < 3; d++) {
    pos[d] += DIFF[dir][d];
    if (pos[d] < 0) pos[d] += g[d];
    else if (pos[d] >= g[d]) pos[d] -= g[d];
  }
  return pos[0] + pos[1] * g[0] + pos[2] * g[0] * g[1];
}
static int is_over_boundary_grid(const orc_system *S, int id, int dir) {
  const int *g = S->prm.grid;
  int pos[3] = {id % g[0], (id / g[0]) % g[1], id / g[0] / g[1]};
  for (int d = 0; d < 3; d++) {
    pos[d] += DIFF[dir][d];
    if (pos[d] < 0 || pos[d] >= g[d]) return 1;
  }
  return 0;
}

/* ---- MeshList::ChangeScale (meshlist.cc:31-55) ---------------------------------------------- */
static void mesh_change_scale(orc_system *S, unit_t *u) {
  const double SL = S->SL;
  double w[3];
  int ms[3];
  for (int d = 0; d < 3; d++) {
    w[d] = u->e[d] - u->s[d];
    ms[d] = (int)(w[d] / SL);
    u->mesh_size[d] = w[d] / (double)ms[d];
  }
  u->mx = (int)(w[0] / u->mesh_size[0]) + 2;
  u->my = (int)(w[1] / u->mesh_size[1]) + 2;
  u->mz = (int)(w[2] / u->mesh_size[2]) + 2;
  u->number_of_mesh = u->mx * u->my * u->mz;
  u->mesh_index = (int *)calloc((size_t)u->number_of_mesh, sizeof(int));
  u->mesh_index2 = (int *)calloc((size_t)u->number_of_mesh, sizeof(int));
  u->mesh_particle_number = (int *)calloc((size_t)u->number_of_mesh, sizeof(int));
}

orc_system *orc_create(const orc_params *prm) {
  orc_system *S = (orc_system *)calloc(1, sizeof(orc_system));
  S->prm = *prm;
  if (S->prm.capacity <= 0) S->prm.capacity = 1000000;
  if (S->prm.pairs_per_particle <= 0) S->prm.pairs_per_particle = 80;
  S->SL = prm->cutoff + prm->buffer_length; /* simulationinfo.cc:7 */
  { /* variables.cc:21-27 */
    const double s2 = 1.0 / (prm->cutoff * prm->cutoff);
    const double s6 = s2 * s2 * s2;
    const double s8 = s6 * s2;
    const double s12 = s6 * s6;
    const double s14 = s12 * s2;
    S->C2 = 6.0 * s14 - 3.0 * s8;
    S->C0 = -s12 + s6 - S->C2 / s2;
  }
  S->nunits = prm->grid[0] * prm->grid[1] * prm->grid[2];
  S->u = (unit_t *)calloc((size_t)S->nunits, sizeof(unit_t));
  const size_t N = (size_t)S->prm.capacity;
  S->pair_cap = (long)N * S->prm.pairs_per_particle;
  for (int id = 0; id < S->nunits; id++) {
    unit_t *u = &S->u[id];
    u->id = id;
    u->gpos[0] = id % prm->grid[0];
    u->gpos[1] = (id / prm->grid[0]) % prm->grid[1];
    u->gpos[2] = id / prm->grid[0] / prm->grid[1];
    for (int d = 0; d < 3; d++) { /* mdunit.cc:19-24 */
      double ul = prm->L[d] / (double)prm->grid[d];
      u->s[d] = ul * (double)u->gpos[d];
      u->e[d] = u->s[d] + ul;
    }
    u->type = (int *)calloc(N, sizeof(int));
    u->q = (double(*)[ND])calloc(N, sizeof(double[ND]));
    u->p = (double(*)[ND])calloc(N, sizeof(double[ND]));
    u->qb_old = (double(*)[ND])calloc(N, sizeof(double[ND]));
    u->disp = (double *)calloc(N, sizeof(double));
    u->cand = (int *)calloc(N, sizeof(int));
    u->particle_position = (int *)calloc(N, sizeof(int));
    u->sortbuf = (int *)calloc(N, sizeof(int));
    u->key_pointer = (int *)calloc(N, sizeof(int));
    u->key_pointer2 = (int *)calloc(N, sizeof(int));
    u->number_of_partners = (int *)calloc(N, sizeof(int));
    u->vbuf = (int *)calloc(N, sizeof(int));
    u->key_particles = (int *)calloc((size_t)S->pair_cap, sizeof(int));
    u->partner_particles = (int *)calloc((size_t)S->pair_cap, sizeof(int));
    u->sorted_list = (int *)calloc((size_t)S->pair_cap + 16, sizeof(int));
    for (int d = 0; d < MAX_DIR; d++) u->border[d] = (int *)calloc(N, sizeof(int));
    u->send_buffer = (pinfo_t *)calloc(N, sizeof(pinfo_t));
    mesh_change_scale(S, u);
    mt_seed(&u->mt, (uint32_t)id); /* mdunit.cc:26 */
  }
  return S;
}

void orc_destroy(orc_system *S) {
  if (!S) return;
  for (int id = 0; id < S->nunits; id++) {
    unit_t *u = &S->u[id];
    free(u->type); free(u->q); free(u->p); free(u->qb_old); free(u->disp); free(u->cand);
    free(u->particle_position); free(u->sortbuf); free(u->key_pointer); free(u->key_pointer2);
    free(u->number_of_partners); free(u->vbuf); free(u->key_particles); free(u->partner_particles);
    free(u->sorted_list); free(u->send_buffer);
    free(u->mesh_index); free(u->mesh_index2); free(u->mesh_particle_number);
    for (int d = 0; d < MAX_DIR; d++) free(u->border[d]);
  }
  free(S->u);
  free(S);
}

int orc_num_units(orc_system *S) { return S->nunits; }
void orc_get_rect(orc_system *S, int unit, double lo[3], double hi[3]) {
  for (int d = 0; d < 3; d++) { lo[d] = S->u[unit].s[d]; hi[d] = S->u[unit].e[d]; }
}

int orc_set_unit(orc_system *S, int unit, const double *q4, const double *p4, const int *type, int n) {
  unit_t *u = &S->u[unit];
  if (n >= S->prm.capacity) { snprintf(g_err, sizeof g_err, "capacity"); return ORC_ECAPACITY; }
  for (int i = 0; i < n; i++) {
    for (int d = 0; d < ND; d++) { u->q[i][d] = q4[4 * (size_t)i + d]; u->p[i][d] = p4[4 * (size_t)i + d]; }
    u->type[i] = type ? type[i] : 0;
  }
  u->pn = u->tn = n;
  u->initialized = 0;
  return ORC_OK;
}

/* MDUnit::AddParticle (mdunit.cc:44-49) + MDRect::IsInside (mdrect.h:33-36) + Variables::AddParticle */
int orc_add_particle(orc_system *S, const double x[3], const double v[3], int type) {
  for (int id = 0; id < S->nunits; id++) {
    unit_t *u = &S->u[id];
    if (u->s[0] <= x[0] && x[0] < u->e[0] && u->s[1] <= x[1] && x[1] < u->e[1] && u->s[2] <= x[2] && x[2] < u->e[2]) {
      int k = u->pn;
      if (k + 1 >= S->prm.capacity) return ORC_ECAPACITY;
      for (int d = 0; d < 3; d++) { u->q[k][d] = x[d]; u->p[k][d] = v ? v[d] : 0.0; }
      u->type[k] = type;
      u->pn = u->tn = k + 1;
    }
  }
  return ORC_OK;
}

void orc_set_seed(orc_system *S, int seed) { /* mdunit.h:58 */
  for (int id = 0; id < S->nunits; id++) mt_seed(&S->u[id].mt, (uint32_t)(seed + id));
}
void orc_set_control_temperature(orc_system *S, int on) { S->prm.control_temperature = on; }

/* ---- ForceCalculator::UpdatePositionHalf (fcalculator.cc:22-32) ----------------------------- */
static void update_position_half(orc_system *S, unit_t *u) {
  const double dt2 = S->prm.time_step * 0.5;
  for (int i = 0; i < u->pn; i++) {
    u->q[i][0] += u->p[i][0] * dt2;
    u->q[i][1] += u->p[i][1] * dt2;
    u->q[i][2] += u->p[i][2] * dt2;
  }
}

/* ---- ForceCalculator::CalculateForceNext (fcalculator.cc:229-290), order kept literally ------ */
static void calculate_force(orc_system *S, unit_t *u) {
  const double CL2 = S->prm.cutoff * S->prm.cutoff;
  const double C2 = S->C2 * 8.0;
  const double dt = S->prm.time_step;
  double(*q)[ND] = u->q;
  double(*p)[ND] = u->p;
  const int pn = u->pn;
  const int *sorted_list = u->sorted_list;
  for (int i = 0; i < pn; i++) {
    const double qx_key = q[i][0], qy_key = q[i][1], qz_key = q[i][2];
    double pfx = 0, pfy = 0, pfz = 0;
    const int kp = u->key_pointer[i];
    int ja = sorted_list[kp];
    double dxa = q[ja][0] - qx_key, dya = q[ja][1] - qy_key, dza = q[ja][2] - qz_key;
    double df = 0.0;
    double dxb = 0.0, dyb = 0.0, dzb = 0.0;
    int jb = 0;
    const int np = u->number_of_partners[i];
    for (int k = kp; k < np + kp; k++) {
      const double dx = dxa, dy = dya, dz = dza;
      double r2 = (dx * dx + dy * dy + dz * dz);
      const int j = ja;
      ja = sorted_list[k + 1];
      dxa = q[ja][0] - qx_key;
      dya = q[ja][1] - qy_key;
      dza = q[ja][2] - qz_key;
      if (r2 > CL2) continue;
      pfx += df * dxb;
      pfy += df * dyb;
      pfz += df * dzb;
      p[jb][0] -= df * dxb;
      p[jb][1] -= df * dyb;
      p[jb][2] -= df * dzb;
      const double r6 = r2 * r2 * r2;
      df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2) + C2) * dt;
      jb = j;
      dxb = dx;
      dyb = dy;
      dzb = dz;
    }
    p[jb][0] -= df * dxb;
    p[jb][1] -= df * dyb;
    p[jb][2] -= df * dzb;
    p[i][0] += pfx + df * dxb;
    p[i][1] += pfy + df * dyb;
    p[i][2] += pfz + df * dzb;
  }
}

/* ---- thermostats (fcalculator.cc:744-778) --------------------------------------------------- */
static void heatbath_zeta(orc_system *S, unit_t *u, double ct) {
  const double dt2 = S->prm.time_step * 0.5;
  const double tau = S->prm.heatbath_tau;
  double t1 = (ct - S->prm.aimed_temperature) / (tau * tau);
  u->zeta += t1 * dt2;
}
static void heatbath_momenta(orc_system *S, unit_t *u) {
  const double dt2 = S->prm.time_step * 0.5;
  const double exp1 = exp(-dt2 * u->zeta);
  for (int i = 0; i < u->pn; i++)
    for (int d = 0; d < ND; d++) u->p[i][d] *= exp1;
}
static void langevin(orc_system *S, unit_t *u) {
  const double dt = S->prm.time_step;
  const double hb_gamma = S->prm.heatbath_gamma;
  const double T = S->prm.aimed_temperature;
  const double hb_D = sqrt(2.0 * hb_gamma * T / dt);
  normal_t nd = {0.0, hb_D, 0.0, 0}; /* a fresh distribution each call (fcalculator.cc:772) */
  for (int i = 0; i < u->pn; i++) {
    u->p[i][0] += (-hb_gamma * u->p[i][0] + normal_draw(&nd, &u->mt)) * dt;
    u->p[i][1] += (-hb_gamma * u->p[i][1] + normal_draw(&nd, &u->mt)) * dt;
    u->p[i][2] += (-hb_gamma * u->p[i][2] + normal_draw(&nd, &u->mt)) * dt;
  }
}

/* ---- MeshList::MakeMesh (meshlist.cc:153-205) ----------------------------------------------- */
static int make_mesh(orc_system *S, unit_t *u) {
  const int pn = u->tn;
  double(*q)[ND] = u->q;
  double imx = 1.0 / u->mesh_size[0], imy = 1.0 / u->mesh_size[1], imz = 1.0 / u->mesh_size[2];
  const double *s = u->s;
  const int mx = u->mx, my = u->my, mz = u->mz;
  for (int i = 0; i < u->number_of_mesh; i++) u->mesh_particle_number[i] = 0;
  for (int i = 0; i < pn; i++) {
    int ix = (int)((q[i][0] - s[0]) * imx) + 1;
    int iy = (int)((q[i][1] - s[1]) * imy) + 1;
    int iz = (int)((q[i][2] - s[2]) * imz) + 1;
    if (ix < 0) ix = mx - 1; else if (ix >= mx) ix = 0;
    if (iy < 0) iy = my - 1; else if (iy >= my) iy = 0;
    if (iz < 0) iz = mz - 1; else if (iz >= mz) iz = 0;
    int index = mx * my * iz + mx * iy + ix;
    if (index >= u->number_of_mesh || index < 0) {
      snprintf(g_err, sizeof g_err, "Invalid index");
      S->err = ORC_ECELL;
      return ORC_ECELL;
    }
    u->particle_position[i] = index;
    u->mesh_particle_number[index]++;
  }
  u->mesh_index[0] = 0;
  int sum = 0;
  for (int i = 0; i < u->number_of_mesh - 1; i++) {
    sum += u->mesh_particle_number[i];
    u->mesh_index[i + 1] = sum;
  }
  for (int i = 0; i < u->number_of_mesh; i++) u->mesh_index2[i] = 0;
  for (int i = 0; i < pn; i++) {
    int index = u->particle_position[i];
    int j = u->mesh_index[index] + u->mesh_index2[index];
    u->sortbuf[j] = i;
    u->mesh_index2[index]++;
  }
  return ORC_OK;
}

/* MeshList::RegisterPair (meshlist.cc:288-310) */
static inline void register_pair(orc_system *S, unit_t *u, int index1, int index2) {
  int i1, i2;
  if (index1 < index2) { i1 = index1; i2 = index2; } else { i1 = index2; i2 = index1; }
  if (u->number_of_pairs + 1 >= S->pair_cap) { S->err = ORC_ECAPACITY; return; }
  u->key_particles[u->number_of_pairs] = i1;
  u->partner_particles[u->number_of_pairs] = i2;
  u->number_of_partners[i1]++;
  u->number_of_pairs++;
}

/* MeshList::AppendList (meshlist.cc:208-217) */
static inline int append_list(unit_t *u, int ix, int iy, int iz, int *v, int n) {
  if (ix < 0 || ix >= u->mx) return n;
  if (iy < 0 || iy >= u->my) return n;
  if (iz < 0 || iz >= u->mz) return n;
  const int index = u->mx * u->my * iz + u->mx * iy + ix;
  const int in = u->mesh_particle_number[index];
  const int mi = u->mesh_index[index];
  memcpy(v + n, u->sortbuf + mi, sizeof(int) * (size_t)in);
  return n + in;
}

/* MeshList::SearchMesh (meshlist.cc:220-271) */
static void search_mesh(orc_system *S, unit_t *u, int index) {
  int ix = index % u->mx;
  int t = index / u->mx;
  int iy = t % u->my;
  int iz = t / u->my;
  int *v = u->vbuf;
  int n = 0;
  n = append_list(u, ix, iy, iz, v, n);
  n = append_list(u, ix + 1, iy, iz, v, n);
  n = append_list(u, ix - 1, iy + 1, iz, v, n);
  n = append_list(u, ix, iy + 1, iz, v, n);
  n = append_list(u, ix + 1, iy + 1, iz, v, n);
  n = append_list(u, ix - 1, iy, iz + 1, v, n);
  n = append_list(u, ix, iy, iz + 1, v, n);
  n = append_list(u, ix + 1, iy, iz + 1, v, n);
  n = append_list(u, ix - 1, iy - 1, iz + 1, v, n);
  n = append_list(u, ix, iy - 1, iz + 1, v, n);
  n = append_list(u, ix + 1, iy - 1, iz + 1, v, n);
  n = append_list(u, ix - 1, iy + 1, iz + 1, v, n);
  n = append_list(u, ix, iy + 1, iz + 1, v, n);
  n = append_list(u, ix + 1, iy + 1, iz + 1, v, n);
  const double S2 = S->SL * S->SL;
  const int pn = u->pn;
  double(*q)[ND] = u->q;
  const int in = u->mesh_particle_number[index];
  const int ln = n;
  for (int i = 0; i < in; i++) {
    const int i1 = v[i];
    const double x1 = q[i1][0], y1 = q[i1][1], z1 = q[i1][2];
    for (int j = i + 1; j < ln; j++) {
      const int i2 = v[j];
      if (i1 >= pn && i2 >= pn) continue;
      const double dx = x1 - q[i2][0];
      const double dy = y1 - q[i2][1];
      const double dz = z1 - q[i2][2];
      const double r2 = (dx * dx + dy * dy + dz * dz);
      if (r2 > S2) continue;
      register_pair(S, u, i1, i2);
    }
  }
}

/* MeshList::MakeList (meshlist.cc:58-87) */
static void make_list(orc_system *S, unit_t *u) {
  u->number_of_pairs = 0;
  const int pn = u->tn;
  for (int i = 0; i < pn; i++) u->number_of_partners[i] = 0;
  if (make_mesh(S, u) != ORC_OK) return;
  for (int i = 0; i < u->number_of_mesh; i++) search_mesh(S, u, i);
  int pos = 0;
  u->key_pointer[0] = 0;
  for (int i = 0; i < pn - 1; i++) {
    pos += u->number_of_partners[i];
    u->key_pointer[i + 1] = pos;
  }
  for (int i = 0; i < pn; i++) u->key_pointer2[i] = 0;
  const long s = u->number_of_pairs;
  for (long k = 0; k < s; k++) {
    int i = u->key_particles[k];
    int j = u->partner_particles[k];
    int index = u->key_pointer[i] + u->key_pointer2[i];
    u->sorted_list[index] = j;
    u->key_pointer2[i]++;
  }
  u->number_of_constructions++;
}

/* ---- PairList (pairlist.cc:11-121) ---------------------------------------------------------- */
static void pairlist_init(orc_system *S, unit_t *u) {
  u->initialized = 1;
  u->buffer_length = S->prm.buffer_length;
  u->lifetime = 0;
  for (int i = 0; i < u->tn; i++) {
    u->qb_old[i][0] = u->q[i][0];
    u->qb_old[i][1] = u->q[i][1];
    u->qb_old[i][2] = u->q[i][2];
  }
}
static void update_pairlist_validity(orc_system *S, unit_t *u) {
  double(*p)[ND] = u->p;
  double max_velocity = p[0][0] * p[0][0] + p[0][1] * p[0][1] + p[0][2] * p[0][2];
  for (int i = 1; i < u->pn; i++) {
    double v = p[i][0] * p[i][0] + p[i][1] * p[i][1] + p[i][2] * p[i][2];
    if (v > max_velocity) max_velocity = v;
  }
  max_velocity = sqrt(max_velocity);
  if (max_velocity * 2.0 * S->prm.time_step > S->prm.buffer_length) {
    snprintf(g_err, sizeof g_err, "Too fast particles exists. Try smaller time step");
    S->err = ORC_ETOOFAST;
  }
  u->buffer_length = u->buffer_length - max_velocity * 2.0 * S->prm.time_step;
}
static int check_by_displacement(orc_system *S, unit_t *u) {
  const int pn = u->tn;
  double(*q)[ND] = u->q;
  double(*qb_old)[ND] = u->qb_old;
  double dr2_max = 0.0;
  for (int i = 0; i < pn; i++) {
    double dx = qb_old[i][0] - q[i][0];
    double dy = qb_old[i][1] - q[i][1];
    double dz = qb_old[i][2] - q[i][2];
    double dr2 = dx * dx + dy * dy + dz * dz;
    u->disp[i] = dr2;
    if (dr2_max < dr2) dr2_max = dr2;
  }
  const int Nl = (pn > S->prm.check_list_length) ? S->prm.check_list_length : pn;
  int l = 0;
  const double c_distance = S->prm.buffer_length - sqrt(dr2_max);
  const double c_d2 = c_distance * c_distance;
  for (int i = 0; i < pn; i++) {
    if (u->disp[i] > c_d2) {
      u->cand[l++] = i;
      if (l >= Nl) return 1;
    }
  }
  const double SL2 = S->SL * S->SL;
  const double CL2 = S->prm.cutoff * S->prm.cutoff;
  for (int i = 0; i < l - 1; i++) {
    for (int j = i + 1; j < l; j++) {
      int i1 = u->cand[i], i2 = u->cand[j];
      double dx = qb_old[i1][0] - qb_old[i2][0];
      double dy = qb_old[i1][1] - qb_old[i2][1];
      double dz = qb_old[i1][2] - qb_old[i2][2];
      double dr2 = dx * dx + dy * dy + dz * dz;
      if (dr2 < SL2) continue;
      dx = q[i1][0] - q[i2][0];
      dy = q[i1][1] - q[i2][1];
      dz = q[i1][2] - q[i2][2];
      dr2 = dx * dx + dy * dy + dz * dz;
      if (dr2 < CL2) return 1;
    }
  }
  return 0;
}
static int is_pairlist_expired_unit(orc_system *S, unit_t *u) {
  if (!u->initialized) return 1;
  u->lifetime++;
  if (u->buffer_length > 0.0) update_pairlist_validity(S, u);
  int expired = 0;
  if (u->buffer_length < 0.0) expired = check_by_displacement(S, u);
  return expired;
}

/* ---- migration (mdunit.cc:78-122, mdrect.cc:5-21, variables.cc:59-79) ----------------------- */
static int is_over_boundary(const unit_t *u, int dir, const double *q) {
  switch (dir) {
    case D_LEFT: return (q[0] < u->s[0]);
    case D_RIGHT: return (q[0] >= u->e[0]);
    case D_BACK: return (q[1] < u->s[1]);
    case D_FORWARD: return (q[1] >= u->e[1]);
    case D_DOWN: return (q[2] < u->s[2]);
    case D_UP: return (q[2] >= u->e[2]);
  }
  return 0;
}
static void make_buffer_for_sending_particle(unit_t *u, int dir) {
  u->nsend = 0;
  int pn = u->pn;
  int index = 0;
  for (int i = 0; i < pn; i++) {
    if (is_over_boundary(u, dir, u->q[i])) {
      pinfo_t *pi = &u->send_buffer[u->nsend++];
      memset(pi, 0, sizeof(*pi));
      for (int d = 0; d < 3; d++) { pi->q[d] = u->q[i][d]; pi->p[d] = u->p[i][d]; }
      pi->type = u->type[i];
    } else {
      for (int d = 0; d < 3; d++) { u->q[index][d] = u->q[i][d]; u->p[index][d] = u->p[i][d]; }
      u->type[index] = u->type[i];
      index++;
    }
  }
  u->pn = index;
}
static void receive_particles(orc_system *S, unit_t *u, const pinfo_t *buf, int n) {
  int index = u->pn;
  if (index + n >= S->prm.capacity) { S->err = ORC_ECAPACITY; return; }
  for (int i = 0; i < n; i++) {
    for (int d = 0; d < 3; d++) { u->q[index][d] = buf[i].q[d]; u->p[index][d] = buf[i].p[d]; }
    u->type[index] = buf[i].type;
    index++;
  }
  u->pn = index;
}
static void adjust_periodic_boundary(orc_system *S, unit_t *u) {
  for (int i = 0; i < u->pn; i++)
    for (int d = 0; d < 3; d++) { /* lane 3 is compared with an uninitialised L[3] in the reference (Q16); skipped */
      if (u->q[i][d] < 0.0) u->q[i][d] += S->prm.L[d];
      else if (u->q[i][d] > S->prm.L[d]) u->q[i][d] -= S->prm.L[d];
    }
  for (int i = 0; i < u->pn; i++)
    for (int d = 0; d < 3; d++)
      if (u->q[i][d] < 0.0 || u->q[i][d] > S->prm.L[d]) {
        snprintf(g_err, sizeof g_err, "Invalid position");
        S->err = ORC_EPOSITION;
      }
}
/* MDManager::SendParticlesSub (mdmanager.cc:195-248), single process: hand-off by reference */
static void send_particles(orc_system *S) {
  for (int dir = 0; dir < MAX_DIR; dir++) {
    const int o_dir = OPP[dir];
    for (int i = 0; i < S->nunits; i++) make_buffer_for_sending_particle(&S->u[i], dir);
    for (int i = 0; i < S->nunits; i++) {
      const int id_src = neighbor_id(S, i, o_dir);
      receive_particles(S, &S->u[i], S->u[id_src].send_buffer, S->u[id_src].nsend);
    }
  }
  for (int i = 0; i < S->nunits; i++) adjust_periodic_boundary(S, &S->u[i]);
}

/* ---- ghosts (mdunit.cc:125-191, mdrect.cc:24-47, mdmanager.cc:283-349) ---------------------- */
static int is_inside_edge(orc_system *S, const unit_t *u, int dir, const double *q) {
  const double SL = S->SL;
  switch (dir) {
    case D_LEFT: return (q[0] < u->s[0] + SL && q[0] > u->s[0]);
    case D_RIGHT: return (q[0] > u->e[0] - SL && q[0] < u->e[0]);
    case D_BACK: return (q[1] < u->s[1] + SL && q[1] > u->s[1]);
    case D_FORWARD: return (q[1] > u->e[1] - SL && q[1] < u->e[1]);
    case D_DOWN: return (q[2] < u->s[2] + SL && q[2] > u->s[2]);
    case D_UP: return (q[2] > u->e[2] - SL && q[2] < u->e[2]);
  }
  return 0;
}
static void find_border_particles(orc_system *S, unit_t *u, int dir) {
  u->nborder[dir] = 0;
  for (int i = 0; i < u->tn; i++)
    if (is_inside_edge(S, u, dir, u->q[i])) u->border[dir][u->nborder[dir]++] = i;
}
static void make_buffer_for_border_particles(orc_system *S, unit_t *u, int dir) {
  u->nsend = 0;
  const double *L = S->prm.L;
  double diff[ND] = {0.0, 0.0, 0.0};
  if (is_over_boundary_grid(S, u->id, dir)) {
    switch (dir) {
      case D_LEFT: diff[0] = L[0]; break;
      case D_RIGHT: diff[0] -= L[0]; break;
      case D_BACK: diff[1] = L[1]; break;
      case D_FORWARD: diff[1] = -L[1]; break;
      case D_DOWN: diff[2] = L[2]; break;
      case D_UP: diff[2] = -L[2]; break;
    }
  }
  for (int i = 0; i < u->nborder[dir]; i++) {
    const int j = u->border[dir][i];
    pinfo_t *pi = &u->send_buffer[u->nsend++];
    memset(pi, 0, sizeof(*pi));
    pi->q[0] = u->q[j][0] + diff[0];
    pi->q[1] = u->q[j][1] + diff[1];
    pi->q[2] = u->q[j][2] + diff[2];
  }
}
static void receive_border_particles(orc_system *S, unit_t *u, const pinfo_t *buf, int n) {
  int index = u->tn;
  if (index + n >= S->prm.capacity) { S->err = ORC_ECAPACITY; return; }
  for (int i = 0; i < n; i++) {
    u->q[i + index][0] = buf[i].q[0];
    u->q[i + index][1] = buf[i].q[1];
    u->q[i + index][2] = buf[i].q[2];
  }
  u->tn = index + n;
}
static void send_border_particles_sub(orc_system *S, int dir) {
  const int o_dir = OPP[dir];
  for (int i = 0; i < S->nunits; i++) make_buffer_for_border_particles(S, &S->u[i], dir);
  for (int i = 0; i < S->nunits; i++) {
    const int id_src = neighbor_id(S, i, o_dir);
    receive_border_particles(S, &S->u[i], S->u[id_src].send_buffer, S->u[id_src].nsend);
  }
}
static void send_border_particles(orc_system *S) {
  for (int i = 0; i < S->nunits; i++) S->u[i].tn = S->u[i].pn;
  for (int dir = 0; dir < MAX_DIR; dir++) send_border_particles_sub(S, dir);
}

/* ---- MDManager::MakePairList (mdmanager.cc:262-280) ----------------------------------------- */
void orc_make_pair_list(orc_system *S) {
  send_particles(S);
  for (int i = 0; i < S->nunits; i++) S->u[i].tn = S->u[i].pn;
  for (int dir = 0; dir < MAX_DIR; dir++) {
    for (int i = 0; i < S->nunits; i++) find_border_particles(S, &S->u[i], dir);
    send_border_particles_sub(S, dir);
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < S->nunits; i++) { /* MDUnit::MakePairList (mdunit.cc:194-198), SortParticle = no */
    pairlist_init(S, &S->u[i]);
    make_list(S, &S->u[i]);
  }
}

/* ---- observers (observer.cc:5-79) ----------------------------------------------------------- */
static double kinetic(unit_t *u) {
  double e = 0;
  for (int i = 0; i < u->pn; i++)
    for (int d = 0; d < ND; d++) e += 0.5 * u->p[i][d] * u->p[i][d];
  return e;
}
static double potential(orc_system *S, unit_t *u) {
  const double CL2 = S->prm.cutoff * S->prm.cutoff;
  double(*q)[ND] = u->q;
  const double C2 = S->C2, C0 = S->C0;
  const int pn = u->pn;
  double energy = 0.0;
  for (long k = 0; k < u->number_of_pairs; k++) {
    int i = u->key_particles[k], j = u->partner_particles[k];
    double dx = q[i][0] - q[j][0], dy = q[i][1] - q[j][1], dz = q[i][2] - q[j][2];
    const double r2 = (dx * dx + dy * dy + dz * dz);
    if (r2 > CL2) continue;
    double e = 4.0 * (1.0 / (r2 * r2 * r2 * r2 * r2 * r2) - 1.0 / (r2 * r2 * r2) + C2 * r2 + C0);
    if (i >= pn || j >= pn) e *= 0.5;
    energy += e;
  }
  return energy;
}
static double virial(orc_system *S, unit_t *u) {
  double phi = 0.0;
  double(*q)[ND] = u->q;
  const int pn = u->pn;
  const double C_2 = S->C2;
  const double CL2 = S->prm.cutoff * S->prm.cutoff;
  for (long k = 0; k < u->number_of_pairs; k++) {
    int i = u->key_particles[k], j = u->partner_particles[k];
    double dx = q[j][0] - q[i][0], dy = q[j][1] - q[i][1], dz = q[j][2] - q[i][2];
    const double r2 = (dx * dx + dy * dy + dz * dz);
    if (r2 > CL2) continue;
    const double r6 = r2 * r2 * r2;
    double df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2) + C_2 * 8.0) * r2;
    if (i >= pn || j >= pn) df *= 0.5;
    phi += df;
  }
  return phi / 3.0;
}
void orc_observe(orc_system *S, int unit, double out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.0;
  for (int i = 0; i < S->nunits; i++) {
    if (unit >= 0 && i != unit) continue;
    unit_t *u = &S->u[i];
    out[0] += kinetic(u);
    if (u->number_of_constructions > 0) {
      out[1] += potential(S, u);
      out[2] += virial(S, u);
    }
    out[3] += (double)u->pn;
  }
}
static double temperature(orc_system *S) { /* mdmanager.cc:390-407 */
  double e = 0.0;
  unsigned long pn = 0;
  for (int i = 0; i < S->nunits; i++) { e += kinetic(&S->u[i]); pn += (unsigned long)S->u[i].pn; }
  return e / (double)pn / 1.5;
}

/* ---- MDManager::Calculate (mdmanager.cc:116-192) -------------------------------------------- */
/* MDManager::ChangeScale (mdmanager.cc:432-443) with MDUnit::ChangeScale (mdunit.cc:200-205):
 * MDRect::ChangeScale (mdrect.cc:55-60), Variables::ChangeScale (variables.cc:133-139, locals only),
 * MeshList::ChangeScale (meshlist.cc:31-55) */
void orc_change_scale(orc_system *S, double alpha) {
  orc_make_pair_list(S);
  for (int d = 0; d < 3; d++) S->prm.L[d] = S->prm.L[d] * alpha;
  for (int id = 0; id < S->nunits; id++) {
    unit_t *u = &S->u[id];
    for (int d = 0; d < 3; d++) {
      u->s[d] *= alpha;
      u->e[d] *= alpha;
    }
    for (int i = 0; i < u->pn; i++) {
      u->q[i][0] *= alpha;
      u->q[i][1] *= alpha;
      u->q[i][2] *= alpha;
    }
    free(u->mesh_index); free(u->mesh_index2); free(u->mesh_particle_number);
    mesh_change_scale(S, u);
  }
  orc_make_pair_list(S);
}

int orc_is_pairlist_expired(orc_system *S) {
  int expired = 0;
  for (int i = 0; i < S->nunits; i++) expired |= is_pairlist_expired_unit(S, &S->u[i]);
  return expired;
}
void orc_update_position_half(orc_system *S) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < S->nunits; i++) update_position_half(S, &S->u[i]);
}
void orc_send_border_particles(orc_system *S) { send_border_particles(S); }
void orc_calculate_force(orc_system *S) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < S->nunits; i++) calculate_force(S, &S->u[i]);
}
int orc_calculate(orc_system *S, int nsteps) {
  for (int step = 0; step < nsteps; step++) {
    if (orc_is_pairlist_expired(S)) orc_make_pair_list(S);
    orc_update_position_half(S);
    send_border_particles(S);
    if (S->prm.control_temperature) {
      if (S->prm.heatbath_type == 0) { /* CalculateNoseHoover, mdmanager.cc:160-175 */
        double t = temperature(S);
        for (int i = 0; i < S->nunits; i++) {
          heatbath_zeta(S, &S->u[i], t);
          heatbath_momenta(S, &S->u[i]);
          calculate_force(S, &S->u[i]);
          heatbath_momenta(S, &S->u[i]);
        }
        t = temperature(S);
        for (int i = 0; i < S->nunits; i++) {
          heatbath_zeta(S, &S->u[i], t);
          update_position_half(S, &S->u[i]);
        }
      } else { /* CalculateLangevin, mdmanager.cc:186-192: no second half drift */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < S->nunits; i++) {
          calculate_force(S, &S->u[i]);
          langevin(S, &S->u[i]);
        }
      }
    } else { /* CalculateForce, mdmanager.cc:151-157 */
#pragma omp parallel for schedule(static)
      for (int i = 0; i < S->nunits; i++) {
        calculate_force(S, &S->u[i]);
        update_position_half(S, &S->u[i]);
      }
    }
    S->s_time += S->prm.time_step;
    if (S->err) return S->err;
  }
  return S->err;
}

/* ---- accessors ------------------------------------------------------------------------------ */
void orc_get_counts(orc_system *S, int unit, int *pn, int *tn, long *npairs, int *nbuilds) {
  unit_t *u = &S->u[unit];
  if (pn) *pn = u->pn;
  if (tn) *tn = u->tn;
  if (npairs) *npairs = u->number_of_pairs;
  if (nbuilds) *nbuilds = u->number_of_constructions;
}
void orc_get_state(orc_system *S, int unit, double *q4, double *p4, int *type) {
  unit_t *u = &S->u[unit];
  if (q4) memcpy(q4, u->q, sizeof(double[ND]) * (size_t)u->tn);
  if (p4) memcpy(p4, u->p, sizeof(double[ND]) * (size_t)u->pn);
  if (type) memcpy(type, u->type, sizeof(int) * (size_t)u->pn);
}
void orc_get_pairs(orc_system *S, int unit, int *key, int *partner) {
  unit_t *u = &S->u[unit];
  memcpy(key, u->key_particles, sizeof(int) * (size_t)u->number_of_pairs);
  memcpy(partner, u->partner_particles, sizeof(int) * (size_t)u->number_of_pairs);
}
void orc_get_csr(orc_system *S, int unit, int *pointer, int *count, int *list) {
  unit_t *u = &S->u[unit];
  memcpy(pointer, u->key_pointer, sizeof(int) * (size_t)u->pn);
  memcpy(count, u->number_of_partners, sizeof(int) * (size_t)u->pn);
  memcpy(list, u->sorted_list, sizeof(int) * (size_t)u->number_of_pairs);
}
void orc_get_impulse(orc_system *S, int unit, double *f3) {
  unit_t *u = &S->u[unit];
  const size_t tn = (size_t)u->tn;
  double(*save)[ND] = (double(*)[ND])malloc(sizeof(double[ND]) * (tn + 1));
  memcpy(save, u->p, sizeof(double[ND]) * tn);
  memset(u->p, 0, sizeof(double[ND]) * tn);
  calculate_force(S, u);
  for (int i = 0; i < u->pn; i++)
    for (int d = 0; d < 3; d++) f3[3 * (size_t)i + d] = u->p[i][d];
  memcpy(u->p, save, sizeof(double[ND]) * tn);
  free(save);
}
double orc_get_zeta(orc_system *S, int unit) { return S->u[unit].zeta; }
double orc_simulation_time(orc_system *S) { return S->s_time; }
