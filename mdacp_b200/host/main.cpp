// mdacp_b200 -- stand-alone C++ driver over the C-ABI (include/mdacp_b200.h).
//
// Mirrors the reference's driver surface for the hot path -- main.cc:8-17, MDManager
// (mdmanager.cc:11-47, 116-148, 262-280, 380-429), ParaInfo (parainfo.cc:18-150, 191-275, one
// process), SimulationInfo (simulationinfo.cc:5-51), Parameter (parameter.cc:35-111) and the four
// project loops Benchmark / Langevin / Collision / Burst (benchmark.cc:13-52, langevin.cc:11-32,
// collision.cc:54-80, burst.cc:104-143), cdview output included (mdmanager.cc:80-105 through the
// asynchronous snapshot of the C-ABI) -- with the same cfg keys and the same printed lines, so a user of the
// reference can run `mdacp_b200 input.cfg` and read the output they are used to.
// Units: MDB200_UNITS (default 1) units in this process, round-robin over the visible GPUs --
// the role OMP_NUM_THREADS plays in the reference.  All numerics happen in libmdacp_b200.so.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include "mdacp_b200.h"

extern "C" {
long mdh_benchmark_config(const double s[3], const double e[3], double density, int unit_id, double *q4, long cap);
long mdh_lattice_config(const double L[3], double density, const double s[3], const double e[3], double *q4, long cap);
long mdh_collision_config(const double L[3], const double s[3], const double e[3], double *q4, double *p4, long cap);
void mdh_initial_velocity(double *p4, long n, double V0, int seed);
long mdh_droplet_config(const double L[3], double density, double radius, const double s[3], const double e[3],
                        double *q4, long cap);
int mdh_save_cdview(const char *filename, int append, const double *q4, const double *p4, const int *type, long pn);
}

static const double CUTOFF_LENGTH = 3.0;  // mdconfig.h:15
static const char *MDACP_VERSION = "MDACP Ver 2.21 (B200 hot path: libmdacp_b200)";

static void check(int rc) {
  if (rc != MDB200_OK) {  // show_error + exit(1), mdconfig.h:21
    std::cout << "# Error in libmdacp_b200 (code " << rc << ")" << std::endl << mdb200_last_error() << std::endl;
    exit(1);
  }
}
static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct Parameter {  // parameter.cc:35-111
  std::map<std::string, std::string> params;
  bool valid = true;
  void LoadFromFile(const char *fn) {
    std::ifstream is(fn);
    if (is.fail()) { std::cout << "Could not open input file " << fn << std::endl; valid = false; }
    std::string line;
    while (getline(is, line)) {
      size_t idx = line.find("=");
      if (idx != std::string::npos) params.insert(std::make_pair(line.substr(0, idx), line.substr(idx + 1)));
    }
  }
  bool Contains(const std::string &k) const { return params.count(k) > 0; }
  std::string GetStringDef(const std::string &k, const std::string &v) const { return Contains(k) ? params.at(k) : v; }
  double GetDoubleDef(const std::string &k, double v) const { return Contains(k) ? atof(params.at(k).c_str()) : v; }
  double GetDouble(const std::string &k) const { return GetDoubleDef(k, 0.0); }
  int GetIntegerDef(const std::string &k, int v) const { return Contains(k) ? atoi(params.at(k).c_str()) : v; }
  bool GetBooleanDef(const std::string &k, bool v) const {
    if (!Contains(k)) return v;
    return params.at(k) == "yes" || params.at(k) == "Yes";
  }
};

static void dims_create(int n, int dims[3]) {  // balanced, non-increasing (MPI_Dims_create)
  dims[0] = dims[1] = dims[2] = 1;
  std::vector<int> f;
  int m = n;
  for (int p = 2; p * p <= m; p++) while (m % p == 0) { f.push_back(p); m /= p; }
  if (m > 1) f.push_back(m);
  for (int k = (int)f.size() - 1; k >= 0; k--) *std::min_element(dims, dims + 3) *= f[k];
  std::sort(dims, dims + 3, [](int a, int b) { return a > b; });
}

struct MDManager {
  Parameter param;
  int num_units = 1;
  int grid[3];
  double L[3];
  double TimeStep, AimedTemperature, HeatbathTau, HeatbathGamma;
  bool ControlTemperature;
  int HeatbathType, CheckListLength;
  std::vector<mdb200_unit *> devs;
  std::vector<std::vector<double>> q, p;  // host Variables per unit (rows of 4)
  std::vector<long> pn;
  double s_time = 0.0;

  int uid_of(int ix, int iy, int iz) const { return ix + iy * grid[0] + iz * grid[0] * grid[1]; }

  explicit MDManager(const char *cfgfile) {
    std::cout << "# " << MDACP_VERSION << std::endl;
    param.LoadFromFile(cfgfile);
    const char *env = getenv("MDB200_UNITS");
    num_units = env ? std::max(1, atoi(env)) : 1;
    std::cout << "# 1 Process(es), " << num_units << " Unit(s) on GPU" << std::endl;
    int l3[3];
    dims_create(num_units, l3);
    grid[0] = l3[2]; grid[1] = l3[1]; grid[2] = l3[0];  // reversed for threads (parainfo.cc:31-34)
    if (param.Contains("LocalGridX")) {
      grid[0] = param.GetIntegerDef("LocalGridX", 1);
      grid[1] = param.GetIntegerDef("LocalGridY", 1);
      grid[2] = param.GetIntegerDef("LocalGridZ", 1);
    }
    std::cout << "# Units = (" << grid[0] << "x" << grid[1] << "x" << grid[2] << ")" << std::endl;
    if (grid[0] * grid[1] * grid[2] != num_units) { std::cout << "Invalid LocalGrid Numbers" << std::endl; exit(1); }
    // SimulationInfo (simulationinfo.cc:5-51)
    TimeStep = param.GetDoubleDef("TimeStep", 0.001);
    AimedTemperature = param.GetDoubleDef("AimedTemperature", 1.0);
    ControlTemperature = param.GetBooleanDef("ControlTemperature", false);
    HeatbathTau = param.GetDoubleDef("HeatbathTau", 0.1);
    HeatbathGamma = param.GetDoubleDef("HeatbathGamma", 0.1);
    std::string hb = param.GetStringDef("HeatbathType", "NoseHoover");
    HeatbathType = (hb == "Langevin") ? MDB200_HT_LANGEVIN : MDB200_HT_NOSEHOOVER;
    CheckListLength = param.GetIntegerDef("CheckListLength", 500);
    L[0] = L[1] = L[2] = 0.0;
    if (param.Contains("SystemSize")) L[0] = L[1] = L[2] = param.GetDouble("SystemSize");
    else if (param.Contains("UnitLength"))
      for (int d = 0; d < 3; d++) L[d] = param.GetDouble("UnitLength") * grid[d];
    if (param.Contains("SystemSizeX")) L[0] = param.GetDouble("SystemSizeX");
    if (param.Contains("SystemSizeY")) L[1] = param.GetDouble("SystemSizeY");
    if (param.Contains("SystemSizeZ")) L[2] = param.GetDouble("SystemSizeZ");
    int ndev = 0;
    check(mdb200_device_count(&ndev));
    devs.resize(num_units);
    q.resize(num_units); p.resize(num_units); pn.assign(num_units, 0);
    for (int id = 0; id < num_units; id++) {
      mdb200_params P;
      memset(&P, 0, sizeof(P));
      int pos[3] = {id % grid[0], (id / grid[0]) % grid[1], id / grid[0] / grid[1]};
      static const int diff[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
      for (int d = 0; d < 3; d++) { P.L[d] = L[d]; P.grid_size[d] = grid[d]; P.grid_pos[d] = pos[d]; }
      P.time_step = TimeStep; P.cutoff = CUTOFF_LENGTH; P.buffer_length = 0.3;
      P.aimed_temperature = AimedTemperature; P.heatbath_tau = HeatbathTau; P.heatbath_gamma = HeatbathGamma;
      P.control_temperature = ControlTemperature; P.heatbath_type = HeatbathType; P.check_list_length = CheckListLength;
      P.unit_id = id;
      for (int dir = 0; dir < 6; dir++) {  // ParaInfo::GetNeighborID, parainfo.cc:191-224
        int n[3];
        for (int d = 0; d < 3; d++) n[d] = (pos[d] + diff[dir][d] + grid[d]) % grid[d];
        P.neighbor_id[dir] = uid_of(n[0], n[1], n[2]);
        P.neighbor_rank[dir] = 0;
      }
      check(mdb200_create(&P, id % ndev, &devs[id]));
    }
    check(mdb200_comm_init(devs.data(), num_units, 1, 0, NULL));
  }
  ~MDManager() { for (auto *d : devs) mdb200_destroy(d); }

  void Rect(int id, double s[3], double e[3]) { check(mdb200_get_rect(devs[id], s, e)); }
  void Upload(void) {
    for (int id = 0; id < num_units; id++)
      check(mdb200_upload(devs[id], (const double(*)[4])q[id].data(), (const double(*)[4])p[id].data(), NULL, pn[id]));
  }
  void SetInitialVelocity(double v0) {  // MDUnit::SetInitialVelocity: seed = unit id (mdunit.h:49)
    for (int id = 0; id < num_units; id++) {
      p[id].assign((size_t)4 * std::max<long>(pn[id], 1), 0.0);
      mdh_initial_velocity(p[id].data(), pn[id], v0, id);
    }
  }
  void SetSeed(int seed) { for (auto *d : devs) check(mdb200_set_seed(d, seed)); }
  void SetControlTemperature(bool b) {
    ControlTemperature = b;
    for (auto *d : devs) check(mdb200_set_control_temperature(d, b ? 1 : 0));
  }
  // MDManager::SaveAsCdviewSequential / SaveAsCdview (mdmanager.cc:80-105): the snapshot of every unit is queued
  // first (device -> pinned host, asynchronous), then the units are appended in order
  int cdview_index = 0;
  void SaveAsCdviewSequential(void) {
    char filename[256];
    snprintf(filename, sizeof(filename), "conf%04d.cd", cdview_index++);
    for (auto *d : devs) check(mdb200_snapshot_begin(d));
    { std::ofstream ofs(filename); }
    for (auto *d : devs) {
      const double(*sq)[4] = NULL, (*sp)[4] = NULL;
      const int *st = NULL;
      long long n = 0;
      check(mdb200_snapshot_wait(d, &sq, &sp, &st, &n));
      if (mdh_save_cdview(filename, 1, (const double *)sq, (const double *)sp, st, (long)n) != 0) {
        std::cout << "# Error: cannot write " << filename << std::endl;
        exit(1);
      }
    }
  }
  void MakePairList(void) { check(mdb200_make_pairlist(devs.data(), num_units)); }
  void Calculate(void) { check(mdb200_step(devs.data(), num_units, 1)); s_time += TimeStep; }
  void Observe(double o[4]) { check(mdb200_observe(devs.data(), num_units, o)); }
  unsigned long GetTotalParticleNumber(void) { double o[4]; Observe(o); return (unsigned long)o[3]; }
  double Temperature(void) { double o[4]; Observe(o); return o[0] / o[3] / 1.5; }
  double TotalEnergy(void) { double o[4]; Observe(o); return o[0] / o[3] + o[1] / o[3]; }
  double Pressure(void) {
    double o[4]; Observe(o);
    const double phi = o[2] / o[3], T = o[0] / o[3] / 1.5, V = L[0] * L[1] * L[2];
    return (T - phi) * o[3] / V;
  }
  void ShowSystemInformation(void) {  // simulationinfo.cc:54-74
    unsigned long n = 0;
    for (long x : pn) n += (unsigned long)x;
    std::cout << "# Number of Particles = " << n << std::endl;
    std::cout << "# System Size = (" << L[0] << "," << L[1] << "," << L[2] << ")" << std::endl;
    std::cout << "# Density = " << (double)n / (L[0] * L[1] * L[2]) << std::endl;
    std::cout << "# Cutoff Length  = " << CUTOFF_LENGTH << std::endl;
    std::cout << "# TimeStep = " << TimeStep << std::endl;
    std::cout << "# ControlTemperature = " << (ControlTemperature ? "yes" : "no") << std::endl;
    if (ControlTemperature) {
      std::cout << "# AimedTemperature =" << AimedTemperature << std::endl;
      if (HeatbathType == MDB200_HT_NOSEHOOVER) {
        std::cout << "# HeatbathType = NoseHoover" << std::endl << "# HeatbathTau = " << HeatbathTau << std::endl;
      } else {
        std::cout << "# HeatbathType = Langevin" << std::endl << "# HeatbathGamma = " << HeatbathGamma << std::endl;
      }
    }
    std::cout << "# IsPeriodic = yes" << std::endl;
  }
};

static void generate(MDManager &m, const std::string &mode) {
  for (int id = 0; id < m.num_units; id++) {
    double s[3], e[3];
    m.Rect(id, s, e);
    const double vol = (e[0] - s[0]) * (e[1] - s[1]) * (e[2] - s[2]);
    long cap = (long)(vol * 1.0) + 4096;
    m.q[id].assign((size_t)4 * cap, 0.0);
    long n;
    if (mode == "Benchmark") {
      n = mdh_benchmark_config(s, e, m.param.GetDoubleDef("Density", 0.7), id, m.q[id].data(), cap);
    } else if (mode == "Langevin") {
      n = mdh_lattice_config(m.L, m.param.GetDoubleDef("Density", 0.5), s, e, m.q[id].data(), cap);
    } else if (mode == "Burst") {
      n = mdh_droplet_config(m.L, m.param.GetDoubleDef("Density", 0.5), m.param.GetDoubleDef("Radius", 10), s, e,
                             m.q[id].data(), cap);
    } else {
      m.p[id].assign((size_t)4 * cap, 0.0);
      n = mdh_collision_config(m.L, s, e, m.q[id].data(), m.p[id].data(), cap);
    }
    if (n < 0) { std::cout << "# Error: too many particles for the host buffer" << std::endl; exit(1); }
    m.pn[id] = n;
  }
}

static void run_benchmark(MDManager &m) {  // benchmark.cc:13-52
  generate(m, "Benchmark");
  m.SetInitialVelocity(m.param.GetDoubleDef("InitialVelocity", 1.0));
  const int T_LOOP = m.param.GetIntegerDef("ThermalizeLoop", 150);
  const int LOOP = m.param.GetIntegerDef("TotalLoop", 1000);
  const int OBSERVE_LOOP = m.param.GetIntegerDef("ObserveLoop", 100);
  m.ShowSystemInformation();
  m.Upload();
  m.MakePairList();
  for (int i = 0; i < T_LOOP; i++) m.Calculate();
  const double start = now();
  for (int i = 0; i < LOOP; i++) {
    m.Calculate();
    if (i % OBSERVE_LOOP == 0)
      std::cout << m.s_time << " " << m.Temperature() << " " << m.Pressure() << " " << m.TotalEnergy() << " #observe"
                << std::endl;
  }
  for (auto *d : m.devs) mdb200_synchronize(d);
  const double sec = now() - start;
  const unsigned long n = m.GetTotalParticleNumber();
  const double mups = (double)LOOP * (double)n / sec / 1.0e6;
  std::cout << "# N = " << n << " " << sec << " [SEC] " << mups << " [MUPS]" << std::endl;
}

static void run_langevin(MDManager &m) {  // langevin.cc:11-32
  generate(m, "Langevin");
  m.SetInitialVelocity(m.param.GetDoubleDef("InitialVelocity", 1.0));
  const int LOOP = m.param.GetIntegerDef("TotalLoop", 1000);
  const int OBSERVE_LOOP = m.param.GetIntegerDef("ObserveLoop", 100);
  m.Upload();
  m.SetSeed(m.param.GetIntegerDef("Seed", 1));
  m.SetControlTemperature(true);
  m.ShowSystemInformation();
  m.MakePairList();
  for (int i = 0; i < LOOP; i++) {
    m.Calculate();
    if (i % OBSERVE_LOOP == 0)
      std::cout << m.s_time << " " << m.Temperature() << " " << m.Pressure() << " " << m.TotalEnergy() << std::endl;
  }
}

static void run_collision(MDManager &m) {  // collision.cc:54-80
  generate(m, "Collision");
  m.Upload();
  m.SaveAsCdviewSequential();
  const int OBSERVE_LOOP = m.param.GetIntegerDef("ObserveLoop", 100);
  const int TOTAL_LOOP = m.param.GetIntegerDef("TotalLoop", 1000);
  const bool save_cdview_file = m.param.GetBooleanDef("SaveCdviewFile", false);
  m.ShowSystemInformation();
  for (int i = 0; i < TOTAL_LOOP; i++) {
    m.Calculate();  // the first call builds the pair list (pairlist.cc:27-29)
    if (i % OBSERVE_LOOP == 0) {
      std::cout << m.s_time << " " << m.Temperature() << "# observe" << std::endl;
      if (save_cdview_file) m.SaveAsCdviewSequential();
    }
  }
}

static void run_burst(MDManager &m) {  // burst.cc:104-143
  generate(m, "Burst");
  m.SetInitialVelocity(m.param.GetDoubleDef("InitialVelocity", 1.0));
  const int T_LOOP = m.param.GetIntegerDef("ThermalizeLoop", 150);
  const int LOOP = m.param.GetIntegerDef("TotalLoop", 1000);
  const int OBSERVE_LOOP = m.param.GetIntegerDef("ObserveLoop", 100);
  m.ShowSystemInformation();
  m.Upload();
  m.MakePairList();
  m.SetControlTemperature(true);
  for (int i = 0; i < T_LOOP; i++) {
    m.Calculate();
    if (i % OBSERVE_LOOP == 0) m.SaveAsCdviewSequential();
  }
  m.SetControlTemperature(false);
  const double start = now();
  for (int i = 0; i < LOOP; i++) {
    m.Calculate();
    if (i % OBSERVE_LOOP == 0) {
      std::cout << m.s_time << " " << m.Temperature() << " " << m.Pressure() << " " << m.TotalEnergy() << "# observe"
                << std::endl;
      m.SaveAsCdviewSequential();
    }
  }
  for (auto *d : m.devs) mdb200_synchronize(d);
  const double sec = now() - start;
  const unsigned long n = m.GetTotalParticleNumber();
  const double mups = (double)LOOP * (double)n / sec / 1.0e6;
  std::cout << "# N = " << n << " " << sec << " [SEC] " << mups << " [MUPS]" << std::endl;
}

int main(int argc, char **argv) {
  setvbuf(stdout, NULL, _IOLBF, 0);
  std::cout.precision(6);
  const char *cfg = "input.cfg";
  if (argc > 1) cfg = argv[1];
  else std::cout << "# Input file is not specified. input.cfg is used." << std::endl;
  MDManager mdm(cfg);
  if (!mdm.param.valid) { std::cout << "Program is aborted." << std::endl; return 0; }
  if (!mdm.param.Contains("Mode")) { std::cout << "Mode is not defined in inputfile." << std::endl; return 0; }
  const std::string mode = mdm.param.GetStringDef("Mode", "");
  std::cout << "# Mode = " << mode << std::endl;
  if (mode == "Benchmark") run_benchmark(mdm);
  else if (mode == "Langevin") run_langevin(mdm);
  else if (mode == "Collision") run_collision(mdm);
  else if (mode == "Burst") run_burst(mdm);
  else std::cout << "Mode " << mode << " is not found." << std::endl;
  return 0;
}
