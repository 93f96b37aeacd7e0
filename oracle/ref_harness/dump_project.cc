// Oracle dump harness -- TEST INFRASTRUCTURE ONLY (oracle/).
//
// An extra "project" (Mode=OracleDump) linked next to the UNMODIFIED reference
// translation units from /root/reference.  It drives the reference through its
// public API only (MDManager / Executor / DoubleObserver)
// and writes full-precision binary records that tests/golden fixtures are
// generated from.  Nothing here is part of the product.
//
// cfg keys (besides the reference's own):
//   DumpInit=Benchmark|Langevin|Collision   initial state + entry behaviour of that project
//   DumpSteps=<n>            number of MDManager::Calculate() calls
//   DumpFile=<path>          output file
//   DumpStateAt=a,b,c        steps (count of completed Calculate calls; 0 = initial) to dump q,p,type
//   DumpPairsAt=a,b,c        steps at which the pair list (raw + CSR) is dumped
//   DumpForceAt=a,b,c        steps at which the kick impulse of the current list/positions is dumped
//   DumpPairsAtBuild=yes     additionally dump raw pairs after every rebuild
//   DumpChangeScaleAt=<k>, DumpChangeScaleAlpha=<a>   after the records of step k: MDManager::ChangeScale(a)
//   DumpSaveFiles=yes        after the last step call the reference's own writers in the working directory:
//                            MDManager::SaveAsCdview("final.cd"), SaveAsDumpFile("final.dump"), SaveConfiguration()
//
// Record format (little endian): char name[48]; int32 dtype (0=f64,1=i32,2=i64);
// int32 ndim; int64 shape[4]; payload.
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <set>
#include <sstream>
#include "projectmanager.h"
#include "mdmanager.h"
#include "confmaker.h"
#include "observer.h"
#include "fcalculator.h"
#include "simulationinfo.h"

namespace {

FILE *g_fp = NULL;

void put_record(const std::string &name, int dtype, int ndim, const int64_t shape[4],
                const void *data) {
  char nm[48];
  memset(nm, 0, sizeof(nm));
  strncpy(nm, name.c_str(), 47);
  fwrite(nm, 1, 48, g_fp);
  int32_t hd[2] = {dtype, ndim};
  fwrite(hd, sizeof(int32_t), 2, g_fp);
  fwrite(shape, sizeof(int64_t), 4, g_fp);
  size_t n = 1;
  for (int i = 0; i < ndim; i++) n *= (size_t)shape[i];
  size_t es = (dtype == 1) ? 4 : 8;
  if (n) fwrite(data, es, n, g_fp);
}
void put_f64(const std::string &name, const double *d, int64_t n0, int64_t n1 = -1) {
  int64_t shape[4] = {n0, n1 < 0 ? 1 : n1, 1, 1};
  put_record(name, 0, n1 < 0 ? 1 : 2, shape, d);
}
void put_i32(const std::string &name, const int *d, int64_t n0) {
  int64_t shape[4] = {n0, 1, 1, 1};
  put_record(name, 1, 1, shape, d);
}
std::string tag(const char *base, int step, int unit) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%s.s%d.u%d", base, step, unit);
  return buf;
}
std::set<int> parse_list(const std::string &s) {
  std::set<int> r;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    if (!tok.empty()) r.insert(atoi(tok.c_str()));
  }
  return r;
}

// q (tn x 3), p (pn x 3), type (pn)
void dump_state(MDManager *mdm, int step) {
  const int nu = mdm->GetTotalThreads();
  for (int u = 0; u < nu; u++) {
    Variables *v = mdm->GetMDUnit(u)->GetVariables();
    const int pn = v->GetParticleNumber();
    const int tn = v->GetTotalParticleNumber();
    std::vector<double> q(3 * (size_t)tn), p(3 * (size_t)pn);
    for (int i = 0; i < tn; i++)
      for (int d = 0; d < 3; d++) q[3 * (size_t)i + d] = v->q[i][d];
    for (int i = 0; i < pn; i++)
      for (int d = 0; d < 3; d++) p[3 * (size_t)i + d] = v->p[i][d];
    put_f64(tag("q", step, u), q.data(), tn, 3);
    put_f64(tag("p", step, u), p.data(), pn, 3);
    put_i32(tag("type", step, u), v->type, pn);
    double z = v->Zeta;
    put_f64(tag("zeta", step, u), &z, 1);
  }
}

class PairDumper : public DoubleObserver {
public:
  int step, unit;
  bool csr;
  double Observe(Variables *vars, MeshList *mesh) {
    const int np = mesh->GetPairNumber();
    put_i32(tag("pair_key", step, unit), mesh->GetKeyParticles(), np);
    put_i32(tag("pair_partner", step, unit), mesh->GetPartnerParticles(), np);
    if (csr) {
      const int pn = vars->GetParticleNumber();
      put_i32(tag("csr_pointer", step, unit), mesh->GetKeyPointerP(), pn);
      put_i32(tag("csr_count", step, unit), mesh->GetNumberOfPartners(), pn);
      put_i32(tag("csr_list", step, unit), mesh->GetSortedList(), np);
    }
    return 0.0;
  }
};

class BuildCounter : public DoubleObserver {
public:
  double Observe(Variables *, MeshList *mesh) { return (double)mesh->GetNumberOfConstructions(); }
};

// Kick impulse of the dispatcher ForceCalculator::CalculateForce (fcalculator.cc:35-50)
// on the current positions and list: zero p, kick, record, restore.
class ForceDumper : public DoubleObserver {
public:
  int step, unit;
  SimulationInfo *sinfo;
  double Observe(Variables *vars, MeshList *mesh) {
    const int pn = vars->GetParticleNumber();
    const int tn = vars->GetTotalParticleNumber();
    std::vector<double> save(4 * (size_t)tn);
    for (int i = 0; i < tn; i++)
      for (int d = 0; d < 4; d++) {
        save[4 * (size_t)i + d] = vars->p[i][d];
        vars->p[i][d] = 0.0;
      }
    ForceCalculator::CalculateForce(vars, mesh, sinfo);
    std::vector<double> f(3 * (size_t)pn);
    for (int i = 0; i < pn; i++)
      for (int d = 0; d < 3; d++) f[3 * (size_t)i + d] = vars->p[i][d];
    for (int i = 0; i < tn; i++)
      for (int d = 0; d < 4; d++) vars->p[i][d] = save[4 * (size_t)i + d];
    put_f64(tag("impulse", step, unit), f.data(), pn, 3);
    return 0.0;
  }
};

class GidSetter : public Executor {
public:
  int offset;
  void Execute(MDUnit *mdu) {
    Variables *v = mdu->GetVariables();
    const int pn = v->GetParticleNumber();
    for (int i = 0; i < pn; i++) v->type[i] = offset + i;
    offset += pn;
  }
};

// Same balls as collision.cc:54-61 (public API only: MDUnit::AddParticle).
class Ball : public Executor {
  double c[D], r, xv, density;
public:
  Ball(const double *c_, double r_, double xv_, double dens) : r(r_), xv(xv_), density(dens) {
    c[X] = c_[X]; c[Y] = c_[Y]; c[Z] = c_[Z];
  }
  void Add(double x[D], MDUnit *mdu) {
    const double dx = x[X] - c[X], dy = x[Y] - c[Y], dz = x[Z] - c[Z];
    const double r2 = dx * dx + dy * dy + dz * dz;
    double v[D] = {xv, 0, 0};
    if (r2 > r * r) return;
    mdu->AddParticle(x, v);
  }
  void Execute(MDUnit *mdu) {
    const double s = 1.0 / pow(density * 0.25, 1.0 / 3.0);
    const int is = static_cast<int>(2 * r / s);
    const double hs = s * 0.5;
    for (int iz = 0; iz < is; iz++)
      for (int iy = 0; iy < is; iy++)
        for (int ix = 0; ix < is; ix++) {
          double x1[D] = {ix * s - r + c[X], iy * s - r + c[Y], iz * s - r + c[Z]};
          double x2[D] = {x1[X] + hs, x1[Y], x1[Z]};
          double x3[D] = {x1[X], x1[Y] + hs, x1[Z]};
          double x4[D] = {x1[X], x1[Y], x1[Z] + hs};
          Add(x1, mdu); Add(x2, mdu); Add(x3, mdu); Add(x4, mdu);
        }
  }
};

class OracleDump : public Project {
public:
  OracleDump(void) { ProjectManager::GetInstance().AddProject("OracleDump", this); }
  void Run(MDManager *mdm);
};

void OracleDump::Run(MDManager *mdm) {
  Parameter *param = mdm->GetParameter();
  const std::string init = param->GetStringDef("DumpInit", "Benchmark");
  const int steps = param->GetIntegerDef("DumpSteps", 100);
  const std::string fname = param->GetStringDef("DumpFile", "oracle_dump.bin");
  const std::set<int> state_at = parse_list(param->GetStringDef("DumpStateAt", "0"));
  const std::set<int> pairs_at = parse_list(param->GetStringDef("DumpPairsAt", "0"));
  const std::set<int> force_at = parse_list(param->GetStringDef("DumpForceAt", ""));
  const bool pairs_at_build = param->GetBooleanDef("DumpPairsAtBuild", false);
  const int cs_at = param->GetIntegerDef("DumpChangeScaleAt", -1);
  const double cs_alpha = param->GetDoubleDef("DumpChangeScaleAlpha", 1.0);
  const int nu = mdm->GetTotalThreads();

  g_fp = fopen(fname.c_str(), "wb");
  if (!g_fp) { printf("cannot open %s\n", fname.c_str()); return; }

  bool initial_build = true;
  if (init == "Benchmark") {
    ExtractConfigurationMaker c(param);
    mdm->ExecuteAll(&c);
    mdm->SetInitialVelocity(param->GetDoubleDef("InitialVelocity", 1.0));
  } else if (init == "Langevin") {
    ConfigurationMaker c(param);
    mdm->ExecuteAll(&c);
    mdm->SetInitialVelocity(param->GetDoubleDef("InitialVelocity", 1.0));
    mdm->SetSeed(param->GetIntegerDef("Seed", 1));
    mdm->SetControlTemperature(true);
  } else if (init == "Collision") {
    const double *L = mdm->GetSystemSize();
    double c1[D] = {L[X] * 0.25, L[Y] * 0.5, L[Z] * 0.5};
    double c2[D] = {L[X] * 0.75, L[Y] * 0.5, L[Z] * 0.5};
    Ball b1(c1, L[Y] * 0.25, 1.0, 0.374), b2(c2, L[Y] * 0.25, -1.0, 0.374);
    mdm->ExecuteAll(&b1);
    mdm->ExecuteAll(&b2);
    initial_build = false;  // collision.cc never calls MakePairList before the loop (Q22)
  } else {
    printf("unknown DumpInit %s\n", init.c_str());
    return;
  }
  GidSetter gs;
  gs.offset = 0;
  for (int u = 0; u < nu; u++) mdm->GetMDUnit(u)->Execute(&gs);
  mdm->ShowSystemInformation();

  int gsz[D];
  mdm->GetGridSize(gsz);
  SimulationInfo sinfo(*param, gsz);
  sinfo.ControlTemperature = (init == "Langevin") ? true : sinfo.ControlTemperature;
  {
    double meta[16] = {0};
    meta[0] = nu; meta[1] = gsz[X]; meta[2] = gsz[Y]; meta[3] = gsz[Z];
    meta[4] = sinfo.L[X]; meta[5] = sinfo.L[Y]; meta[6] = sinfo.L[Z];
    meta[7] = sinfo.TimeStep; meta[8] = CUTOFF_LENGTH; meta[9] = sinfo.BufferLength;
    meta[10] = sinfo.AimedTemperature; meta[11] = sinfo.HeatbathGamma; meta[12] = sinfo.HeatbathTau;
    meta[13] = sinfo.ControlTemperature ? 1 : 0; meta[14] = sinfo.HeatbathType;
    meta[15] = sinfo.CheckListLength;
    put_f64("meta", meta, 16);
  }
  if (state_at.count(0)) dump_state(mdm, 0);
  if (initial_build) mdm->MakePairList();

  std::vector<double> obs((size_t)(steps + 1) * nu * 3, 0.0);
  std::vector<int> builds((size_t)(steps + 1) * nu, 0), pnv((size_t)(steps + 1) * nu, 0),
      tnv((size_t)(steps + 1) * nu, 0);
  KineticEnergyObserver ko;
  PotentialEnergyObserver po(&sinfo);
  VirialObserver vo;
  BuildCounter bc;
  PairDumper pd;
  ForceDumper fd;
  fd.sinfo = &sinfo;
  std::vector<int> last_build(nu, 0);

  for (int s = 0; s <= steps; s++) {
    if (s > 0) mdm->Calculate();
    for (int u = 0; u < nu; u++) {
      MDUnit *mu = mdm->GetMDUnit(u);
      const size_t k = (size_t)s * nu + u;
      const int nb = (int)mu->ObserveDouble(&bc);
      builds[k] = nb;
      pnv[k] = mu->GetParticleNumber();
      tnv[k] = mu->GetTotalParticleNumber();
      obs[3 * k + 0] = mu->ObserveDouble(&ko);
      if (nb > 0) {
        obs[3 * k + 1] = mu->ObserveDouble(&po);
        obs[3 * k + 2] = mu->ObserveDouble(&vo);
      }
      const bool rebuilt = (nb != last_build[u]);
      last_build[u] = nb;
      if (nb > 0 && (pairs_at.count(s) || (pairs_at_build && rebuilt))) {
        pd.step = s; pd.unit = u; pd.csr = pairs_at.count(s) > 0;
        mu->ObserveDouble(&pd);
      }
      if (nb > 0 && force_at.count(s)) {
        fd.step = s; fd.unit = u;
        mu->ObserveDouble(&fd);
      }
    }
    if (s > 0 && state_at.count(s)) dump_state(mdm, s);
    if (s == cs_at) {
      mdm->ChangeScale(cs_alpha);
      double cs[2] = {(double)cs_at, cs_alpha};
      put_f64("change_scale", cs, 2);
    }
  }
  {
    int64_t shape[4] = {steps + 1, nu, 3, 1};
    put_record("obs_KUW", 0, 3, shape, obs.data());
    int64_t sh2[4] = {steps + 1, nu, 1, 1};
    put_record("builds", 1, 2, sh2, builds.data());
    put_record("pn", 1, 2, sh2, pnv.data());
    put_record("tn", 1, 2, sh2, tnv.data());
  }
  if (param->GetBooleanDef("DumpSaveFiles", false)) {
    mdm->SaveAsCdview("final.cd");
    mdm->SaveAsDumpFile("final.dump");
    mdm->SaveConfiguration();
  }
  // the reference's own printed observables at the end, for a human cross-check
  mout << mdm->GetSimulationTime() << " " << mdm->Temperature() << " " << mdm->Pressure() << " "
       << mdm->TotalEnergy() << " #oracle_final" << std::endl;
  fclose(g_fp);
  g_fp = NULL;
}

OracleDump oracle_dump_project;

}  // namespace
