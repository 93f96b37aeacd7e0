// Glue between the reference driver and libmdacp_b200.so -- see mdb200_glue.h.  Compiled together with a scratch
// copy of the reference sources (integration/Makefile); uses only the reference's public headers.
#include <mpi.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <iostream>
#include "mdb200_glue.h"
#include "mdconfig.h"
#include "mdunit.h"
#include "mpistream.h"
#include "parainfo.h"
#include "simulationinfo.h"
#include "variables.h"

namespace {

// 64-bit checksum of the host copy (q, p, type of the locals): has host code modified it since the last transfer?
unsigned long long host_checksum(Variables *v) {
  const int pn = v->GetParticleNumber();
  unsigned long long h = 1469598103934665603ull ^ (unsigned long long)pn;
  const unsigned long long *wq = reinterpret_cast<const unsigned long long *>(&v->q[0][0]);
  const unsigned long long *wp = reinterpret_cast<const unsigned long long *>(&v->p[0][0]);
  unsigned long long a = 0, b = 0;
  for (long i = 0; i < (long)pn * D; i++) {
    a += wq[i] * (2ull * (unsigned long long)i + 1ull);
    b += wp[i] * (2ull * (unsigned long long)i + 3ull);
  }
  unsigned long long c = 0;
  for (int i = 0; i < pn; i++) c += (unsigned long long)(unsigned)v->type[i] * (2ull * i + 5ull);
  return h ^ (a * 1099511628211ull) ^ (b * 14029467366897019727ull) ^ (c * 1609587929392839161ull);
}

unsigned long long g_state = 1;     // bumped whenever the simulation state changes
unsigned long long g_observed = 0;  // state the cached observation belongs to
double g_obs[4];

std::vector<mdb200_unit *> devices(std::vector<MDUnit *> &mdv) {
  std::vector<mdb200_unit *> d(mdv.size());
  for (size_t i = 0; i < mdv.size(); i++) {
    mdv[i]->GetGlue()->ToDevice(mdv[i]->HostVariables());
    d[i] = mdv[i]->GetGlue()->dev;
  }
  return d;
}

void stepped(std::vector<MDUnit *> &mdv) {
  for (size_t i = 0; i < mdv.size(); i++) mdv[i]->GetGlue()->dev_newer = true;
  g_state++;
}

}  // namespace

namespace mdb200_glue {

void Check(int rc) {
  if (rc == MDB200_OK) return;
  show_error(mdb200_last_error());
  exit(1);
}

void CommInit(std::vector<MDUnit *> &mdv, int num_procs, int rank) {
  std::vector<mdb200_unit *> d(mdv.size());
  for (size_t i = 0; i < mdv.size(); i++) d[i] = mdv[i]->GetGlue()->dev;
  unsigned char uid[128];
  memset(uid, 0, sizeof(uid));
  if (num_procs > 1) {
    // the one MPI call that remains on the hot path's set-up: rank 0's NCCL id to everybody
    if (rank == 0) Check(mdb200_comm_unique_id(uid));
    MPI_Bcast(uid, 128, MPI_BYTE, 0, MPI_COMM_WORLD);
  }
  Check(mdb200_comm_init(d.data(), (int)d.size(), num_procs, rank, num_procs > 1 ? uid : NULL));
}

void MakePairList(std::vector<MDUnit *> &mdv) {
  std::vector<mdb200_unit *> d = devices(mdv);
  Check(mdb200_make_pairlist(d.data(), (int)d.size()));
  stepped(mdv);
}

void Calculate(std::vector<MDUnit *> &mdv) {
  std::vector<mdb200_unit *> d = devices(mdv);
  Check(mdb200_step(d.data(), (int)d.size(), 1));
  stepped(mdv);
}

void SendBorderParticles(std::vector<MDUnit *> &mdv) {
  std::vector<mdb200_unit *> d = devices(mdv);
  Check(mdb200_send_border_particles(d.data(), (int)d.size()));
  stepped(mdv);
}

bool IsPairListExpired(std::vector<MDUnit *> &mdv) {
  std::vector<mdb200_unit *> d = devices(mdv);
  int e = 0;
  Check(mdb200_is_pairlist_expired(d.data(), (int)d.size(), &e));
  return e != 0;
}

void ChangeScale(std::vector<MDUnit *> &mdv, double alpha) {
  std::vector<mdb200_unit *> d = devices(mdv);
  Check(mdb200_change_scale(d.data(), (int)d.size(), alpha));
  stepped(mdv);
}

const double *Observe(std::vector<MDUnit *> &mdv) {
  std::vector<mdb200_unit *> d = devices(mdv);  // may upload, which bumps g_state
  if (g_observed != g_state) {
    Check(mdb200_observe(d.data(), (int)d.size(), g_obs));
    g_observed = g_state;
  }
  return g_obs;
}

void SetControlTemperature(std::vector<MDUnit *> &mdv, bool b) {
  for (size_t i = 0; i < mdv.size(); i++) Check(mdb200_set_control_temperature(mdv[i]->GetGlue()->dev, b ? 1 : 0));
}
void SetAimedTemperature(std::vector<MDUnit *> &mdv, double t) {
  for (size_t i = 0; i < mdv.size(); i++) Check(mdb200_set_aimed_temperature(mdv[i]->GetGlue()->dev, t));
}

}  // namespace mdb200_glue

// MDUnit::MDUnit (mdunit.cc:8-27): the device twin of the unit
void Mdb200UnitGlue::Create(int id, SimulationInfo *si, ParaInfo *pi) {
  mdb200_params P;
  memset(&P, 0, sizeof(P));
  for (int d = 0; d < 3; d++) P.L[d] = si->L[d];
  P.time_step = si->TimeStep;
  P.cutoff = CUTOFF_LENGTH;
  P.buffer_length = si->BufferLength;
  P.aimed_temperature = si->AimedTemperature;
  P.heatbath_tau = si->HeatbathTau;
  P.heatbath_gamma = si->HeatbathGamma;
  P.control_temperature = si->ControlTemperature ? 1 : 0;
  P.heatbath_type = si->HeatbathType;
  P.check_list_length = si->CheckListLength;
  int g[D], pos[D];
  pi->GetGridSize(g);
  pi->GetGridPosition(id, pos);
  for (int d = 0; d < 3; d++) { P.grid_size[d] = g[d]; P.grid_pos[d] = pos[d]; }
  P.unit_id = id;
  const int nt = omp_get_max_threads();
  for (int dir = 0; dir < MAX_DIR; dir++) {
    P.neighbor_id[dir] = pi->GetNeighborID(id, dir);
    P.neighbor_rank[dir] = P.neighbor_id[dir] / nt;  // MDManager::IsMyUnit's rule (mdmanager.cc:64-66)
  }
  int ndev = 0;
  mdb200_glue::Check(mdb200_device_count(&ndev));
  mdb200_glue::Check(mdb200_create(&P, (id % nt) % ndev, &dev));
}

void Mdb200UnitGlue::Destroy(void) {
  if (dev) mdb200_destroy(dev);
  dev = 0;
}

void Mdb200UnitGlue::ToHost(Variables *vars) {
  host_valid = true;
  if (!dev_newer) return;
  dev_newer = false;  // first, so that the accessors used below do not recurse
  long long pn = 0, tn = 0;
  mdb200_glue::Check(mdb200_download(dev, vars->q, vars->p, vars->type, N, &pn, &tn));
  vars->SetParticleNumber((int)pn);
  vars->SetTotalParticleNumber((int)(tn < N ? tn : N));
  double z = 0.0;
  mdb200_glue::Check(mdb200_get_zeta(dev, &z));
  vars->Zeta = z;
  host_sum = host_checksum(vars);
  host_pn = (int)pn;
}

void Mdb200UnitGlue::ToDevice(Variables *vars) {
  if (!host_valid) return;
  host_valid = false;
  if (dev_newer) return;  // cannot happen: a host access clears dev_newer first
  const unsigned long long s = host_checksum(vars);
  if (vars->GetParticleNumber() == host_pn && s == host_sum) return;  // host code only read the copy
  mdb200_glue::Check(mdb200_upload(dev, vars->q, vars->p, vars->type, vars->GetParticleNumber()));
  mdb200_glue::Check(mdb200_set_zeta(dev, vars->Zeta));
  host_sum = s;
  host_pn = vars->GetParticleNumber();
  g_state++;
}

void Mdb200UnitGlue::Stepped(void) {
  dev_newer = true;
  g_state++;
}

int Mdb200UnitGlue::ParticleNumber(Variables *vars) {
  if (!dev_newer) return vars->GetParticleNumber();
  long long pn = 0;
  mdb200_glue::Check(mdb200_get_counts(dev, &pn, NULL, NULL, NULL));
  return (int)pn;
}
int Mdb200UnitGlue::TotalParticleNumber(Variables *vars) {
  if (!dev_newer) return vars->GetTotalParticleNumber();
  long long tn = 0;
  mdb200_glue::Check(mdb200_get_counts(dev, NULL, &tn, NULL, NULL));
  return (int)tn;
}
