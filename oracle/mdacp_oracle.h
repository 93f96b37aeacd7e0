/*
 * mdacp_oracle.h -- CPU restatement of the reference hot path (kaityo256/mdacp).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * load this library, and only as the checker / CPU baseline -- never the product path.
 *
 * Pinned: tests/test_oracle_pin.py checks this restatement BIT FOR BIT (q, p, type, pair arrays,
 * CSR, K/U/W, rebuild schedule) against dumps of the unmodified reference built by
 * oracle/Makefile (oracle/_ref/mdacp_scalar), and against the committed fixtures in
 * tests/golden/ generated from those dumps (oracle/make_golden.py).
 */
#ifndef MDACP_ORACLE_H
#define MDACP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  double L[3];
  double time_step;
  double cutoff;            /* CUTOFF_LENGTH (mdconfig.h:15) */
  double buffer_length;     /* simulationinfo.cc:6 */
  double aimed_temperature;
  double heatbath_tau;
  double heatbath_gamma;
  int control_temperature;
  int heatbath_type;        /* 0 Nose-Hoover, 1 Langevin (mdconfig.h:28) */
  int check_list_length;
  int grid[3];              /* unit grid; unit id = ix + iy*gx + iz*gx*gy (parainfo.cc:124-150, one process) */
  int capacity;             /* per-unit particle capacity (the reference's N, mdconfig.h:11) */
  int pairs_per_particle;   /* pair capacity = capacity * this (80 in mdconfig.h:12) */
} orc_params;

typedef struct orc_system orc_system;

orc_system *orc_create(const orc_params *prm);
void orc_destroy(orc_system *s);
int orc_num_units(orc_system *s);
void orc_get_rect(orc_system *s, int unit, double lo[3], double hi[3]);
/* load a unit's particles directly (reference index order), rows of 4 doubles */
int orc_set_unit(orc_system *s, int unit, const double *q4, const double *p4, const int *type, int n);
/* MDUnit::AddParticle semantics (mdunit.cc:44-49): every unit keeps the particles inside its box */
int orc_add_particle(orc_system *s, const double x[3], const double v[3], int type);
void orc_set_seed(orc_system *s, int seed);                  /* MDManager::SetSeed */
void orc_set_control_temperature(orc_system *s, int on);
void orc_make_pair_list(orc_system *s);                      /* MDManager::MakePairList */
void orc_change_scale(orc_system *s, double alpha);          /* MDManager::ChangeScale */
int orc_calculate(orc_system *s, int nsteps);                /* MDManager::Calculate; returns 0 or error code */
void orc_update_position_half(orc_system *s);
void orc_send_border_particles(orc_system *s);
void orc_calculate_force(orc_system *s);
int orc_is_pairlist_expired(orc_system *s);
void orc_observe(orc_system *s, int unit, double out[4]);    /* raw K, U, W, pn of one unit (unit<0: all) */
void orc_get_counts(orc_system *s, int unit, int *pn, int *tn, long *npairs, int *nbuilds);
void orc_get_state(orc_system *s, int unit, double *q4, double *p4, int *type);
void orc_get_pairs(orc_system *s, int unit, int *key, int *partner);
void orc_get_csr(orc_system *s, int unit, int *pointer, int *count, int *list);
void orc_get_impulse(orc_system *s, int unit, double *f3);   /* kick of the current list, p untouched */
double orc_get_zeta(orc_system *s, int unit);
double orc_simulation_time(orc_system *s);
const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
