// libmdacp_host.so -- host-side set-up code the hot path needs around it, in C++ like the
// reference's driver: the initial-state generators and the velocity initialiser.  They use the
// same libstdc++ facilities as the reference (std::mt19937, std::uniform_real_distribution), so
// on the same toolchain the generated states are bit-identical to the reference's
// (tests/test_hostlib.py checks that against reference dumps and golden fixtures).
//
//   mdh_benchmark_config   ExtractConfigurationMaker::Execute   confmaker.cc:174-252
//   mdh_lattice_config     ConfigurationMaker::Execute          confmaker.cc:9-44
//   mdh_collision_config   BallAdder (x2) in Collision::Run     collision.cc:8-61
//   mdh_initial_velocity   Variables::SetInitialVelocity        variables.cc:82-108
//   mdh_droplet_config     DropletMaker::Execute                burst.cc:10-64
// and the file writers fed by mdb200_download / mdb200_snapshot_wait (byte-identical output: the
// same iostream formatting, chunks formatted by several threads and written in order):
//   mdh_save_cdview        MDUnit::SaveAsCdview                 mdunit.cc:52-67 (file handling mdmanager.cc:90-105)
//   mdh_save_dump          FileDumper::Execute                  mdmanager.cc:475-505
//   mdh_save_configuration Variables::SaveConfiguration         variables.cc:45-57
//
// All write rows of 4 doubles (the reference's `double q[N][4]`), lane 3 = 0, and return the
// number of particles written, or -1 if `cap` rows are not enough.
#include <math.h>
#include <algorithm>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

struct Sink {
  double *q, *p;
  long cap, n;
  const double *s, *e;
  bool overflow;
  // MDUnit::AddParticle (mdunit.cc:44-49) + MDRect::IsInside (mdrect.h:33-36) + Variables::AddParticle
  inline void add(const double x[3], const double v[3]) {
    if (!(s[0] <= x[0] && x[0] < e[0] && s[1] <= x[1] && x[1] < e[1] && s[2] <= x[2] && x[2] < e[2])) return;
    if (n >= cap) { overflow = true; return; }
    double *qq = q + 4 * n;
    qq[0] = x[0]; qq[1] = x[1]; qq[2] = x[2]; qq[3] = 0.0;
    if (p) {
      double *pp = p + 4 * n;
      pp[0] = v ? v[0] : 0.0; pp[1] = v ? v[1] : 0.0; pp[2] = v ? v[2] : 0.0; pp[3] = 0.0;
    }
    n++;
  }
};

// Formats rows [0, n) with `row(os, i)` on several threads (one ostringstream per chunk, default iostream
// state = what a fresh std::ofstream has) and writes the chunks in order.
template <typename RowFn>
int write_rows_text(const char *filename, bool append, long n, RowFn row) {
  std::ofstream ofs(filename, append ? (std::ios::out | std::ios::app) : std::ios::out);
  if (!ofs) return -1;
  const long chunk = 1L << 16;
  const long nchunks = (n + chunk - 1) / chunk;
  unsigned hw = std::thread::hardware_concurrency();
  const long nthreads = std::max<long>(1, std::min<long>(hw ? hw : 1, std::min<long>(nchunks, 32)));
  for (long c0 = 0; c0 < nchunks; c0 += nthreads) {
    const long nb = std::min<long>(nthreads, nchunks - c0);
    std::vector<std::string> out((size_t)nb);
    std::vector<std::thread> th;
    for (long t = 0; t < nb; t++)
      th.emplace_back([&, t]() {
        std::ostringstream os;
        const long b = (c0 + t) * chunk, e = std::min<long>(n, b + chunk);
        for (long i = b; i < e; i++) row(os, i);
        out[(size_t)t] = os.str();
      });
    for (auto &x : th) x.join();
    for (long t = 0; t < nb; t++) ofs.write(out[(size_t)t].data(), (std::streamsize)out[(size_t)t].size());
  }
  ofs.close();
  return ofs ? 0 : -1;
}

}  // namespace

extern "C" {

long mdh_benchmark_config(const double s[3], const double e[3], double density, int unit_id, double *q4, long cap) {
  const double w[3] = {e[0] - s[0], e[1] - s[1], e[2] - s[2]};
  double sp = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  const int sx = static_cast<unsigned long int>(w[0] / sp) + 1;
  const int sy = static_cast<unsigned long int>(w[1] / sp) + 1;
  const int sz = static_cast<unsigned long int>(w[2] / sp) + 1;
  const int n_full = sx * sy * sz * 4;
  const int n_exact = static_cast<int>(w[0] * w[1] * w[2] * density);
  const int n_delete = n_full - n_exact;
  // same draw sequence and acceptance rule as confmaker.cc:190-199 (std::find there; a hash set
  // here gives the identical vector without the O(n_delete^2) scan)
  std::vector<int> vec_delete;
  std::unordered_set<int> seen;
  std::mt19937 mt(unit_id);
  std::uniform_real_distribution<double> ud(0.0, 1.0);
  while ((int)vec_delete.size() < n_delete) {
    int v = static_cast<unsigned long int>(n_full * ud(mt));
    if (seen.insert(v).second) vec_delete.push_back(v);
  }
  std::sort(vec_delete.begin(), vec_delete.end());
  sp = w[0] / sx;
  const double hs = sp * 0.5;
  Sink sink{q4, nullptr, cap, 0, s, e, false};
  double x[3];
  int n = 0;
  size_t n_pos = 0;
  auto site = [&](double a, double b, double c) {
    x[0] = a; x[1] = b; x[2] = c;
    if (n_pos < vec_delete.size() && vec_delete[n_pos] == n) n_pos++;
    else sink.add(x, nullptr);
    n++;
  };
  for (int iz = 0; iz < sz; iz++)
    for (int iy = 0; iy < sy; iy++)
      for (int ix = 0; ix < sx; ix++) {
        const double bx = s[0] + static_cast<double>(ix) * sp;
        const double by = s[1] + static_cast<double>(iy) * sp;
        const double bz = s[2] + static_cast<double>(iz) * sp;
        site(bx, by, bz);
        site(bx, by + hs, bz + hs);
        site(bx + hs, by, bz + hs);
        site(bx + hs, by + hs, bz);
      }
  return sink.overflow ? -1 : sink.n;
}

long mdh_lattice_config(const double L[3], double density, const double s[3], const double e[3], double *q4,
                        long cap) {
  const double sp = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  const double hs = sp * 0.5;
  const int sx = static_cast<int>(L[0] / sp);
  const int sy = static_cast<int>(L[1] / sp);
  const int sz = static_cast<int>(L[2] / sp);
  const double eps = 0.0000001;
  Sink sink{q4, nullptr, cap, 0, s, e, false};
  double x[3];
  for (int iz = 0; iz < sz; iz++)
    for (int iy = 0; iy < sy; iy++)
      for (int ix = 0; ix < sx; ix++) {
        x[0] = static_cast<double>(ix) * sp + eps;
        x[1] = static_cast<double>(iy) * sp + eps;
        x[2] = static_cast<double>(iz) * sp + eps;
        sink.add(x, nullptr);
        x[0] = static_cast<double>(ix) * sp + eps;
        x[1] = static_cast<double>(iy) * sp + hs + eps;
        x[2] = static_cast<double>(iz) * sp + hs + eps;
        sink.add(x, nullptr);
        x[0] = static_cast<double>(ix) * sp + hs + eps;
        x[1] = static_cast<double>(iy) * sp + eps;
        x[2] = static_cast<double>(iz) * sp + hs + eps;
        sink.add(x, nullptr);
        x[0] = static_cast<double>(ix) * sp + hs + eps;
        x[1] = static_cast<double>(iy) * sp + hs + eps;
        x[2] = static_cast<double>(iz) * sp + eps;
        sink.add(x, nullptr);
      }
  return sink.overflow ? -1 : sink.n;
}

static void add_ball(Sink &sink, const double c[3], double r, double xv, double density) {
  const double sp = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  const int is = static_cast<int>(2 * r / sp);
  const double hs = sp * 0.5;
  const double v[3] = {xv, 0, 0};
  auto add = [&](const double x[3]) {
    const double dx = x[0] - c[0], dy = x[1] - c[1], dz = x[2] - c[2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 > r * r) return;
    sink.add(x, v);
  };
  for (int iz = 0; iz < is; iz++)
    for (int iy = 0; iy < is; iy++)
      for (int ix = 0; ix < is; ix++) {
        const double x1[3] = {ix * sp - r + c[0], iy * sp - r + c[1], iz * sp - r + c[2]};
        const double x2[3] = {x1[0] + hs, x1[1], x1[2]};
        const double x3[3] = {x1[0], x1[1] + hs, x1[2]};
        const double x4[3] = {x1[0], x1[1], x1[2] + hs};
        add(x1); add(x2); add(x3); add(x4);
      }
}

long mdh_collision_config(const double L[3], const double s[3], const double e[3], double *q4, double *p4,
                          long cap) {
  Sink sink{q4, p4, cap, 0, s, e, false};
  const double c1[3] = {L[0] * 0.25, L[1] * 0.5, L[2] * 0.5};
  const double c2[3] = {L[0] * 0.75, L[1] * 0.5, L[2] * 0.5};
  add_ball(sink, c1, L[1] * 0.25, 1.0, 0.374);   // collision.cc:57-58
  add_ball(sink, c2, L[1] * 0.25, -1.0, 0.374);  // collision.cc:59-60
  return sink.overflow ? -1 : sink.n;
}

void mdh_initial_velocity(double *p4, long n, double V0, int seed) {
  std::mt19937 mt(seed);
  std::uniform_real_distribution<double> ud(0.0, 1.0);
  for (long i = 0; i < n; i++) {
    double z = ud(mt) * 2.0 - 1.0;
    double phi = ud(mt) * 2.0 * M_PI;
    p4[4 * i + 0] = V0 * sqrt(1 - z * z) * cos(phi);
    p4[4 * i + 1] = V0 * sqrt(1 - z * z) * sin(phi);
    p4[4 * i + 2] = V0 * z;
    p4[4 * i + 3] = 0.0;
  }
  double avx = 0, avy = 0, avz = 0;
  for (long i = 0; i < n; i++) {
    avx += p4[4 * i + 0];
    avy += p4[4 * i + 1];
    avz += p4[4 * i + 2];
  }
  avx /= static_cast<double>(n);
  avy /= static_cast<double>(n);
  avz /= static_cast<double>(n);
  for (long i = 0; i < n; i++) {
    p4[4 * i + 0] -= avx;
    p4[4 * i + 1] -= avy;
    p4[4 * i + 2] -= avz;
  }
}

// DropletMaker::Execute (burst.cc:30-64) with AddSub (burst.cc:17-29): global FCC lattice, sites inside the
// sphere of `radius` around the box centre (c = L[X]/2 on all three axes, as the reference has it).
long mdh_droplet_config(const double L[3], double density, double radius, const double s[3], const double e[3],
                        double *q4, long cap) {
  const double sp = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  const double hs = sp * 0.5;
  const int sx = static_cast<int>(L[0] / sp);
  const int sy = static_cast<int>(L[1] / sp);
  const int sz = static_cast<int>(L[2] / sp);
  const double eps = 0.0000001;
  const double c = L[0] * 0.5;
  Sink sink{q4, nullptr, cap, 0, s, e, false};
  auto add = [&](const double x[3]) {
    const double dx = x[0] - c, dy = x[1] - c, dz = x[2] - c;
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < radius * radius) sink.add(x, nullptr);
  };
  double x[3];
  for (int iz = 0; iz < sz; iz++)
    for (int iy = 0; iy < sy; iy++)
      for (int ix = 0; ix < sx; ix++) {
        x[0] = static_cast<double>(ix) * sp + eps;
        x[1] = static_cast<double>(iy) * sp + eps;
        x[2] = static_cast<double>(iz) * sp + eps;
        add(x);
        x[0] = static_cast<double>(ix) * sp + eps;
        x[1] = static_cast<double>(iy) * sp + hs + eps;
        x[2] = static_cast<double>(iz) * sp + hs + eps;
        add(x);
        x[0] = static_cast<double>(ix) * sp + hs + eps;
        x[1] = static_cast<double>(iy) * sp + eps;
        x[2] = static_cast<double>(iz) * sp + hs + eps;
        add(x);
        x[0] = static_cast<double>(ix) * sp + hs + eps;
        x[1] = static_cast<double>(iy) * sp + hs + eps;
        x[2] = static_cast<double>(iz) * sp + eps;
        add(x);
      }
  return sink.overflow ? -1 : sink.n;
}

// MDUnit::SaveAsCdview (mdunit.cc:52-67): "0 type qx qy qz px py pz\n" per local particle.  append = 0 truncates
// first, as rank 0 does in MDManager::SaveAsCdview (mdmanager.cc:90-94); the units are then appended in order.
int mdh_save_cdview(const char *filename, int append, const double *q4, const double *p4, const int *type, long pn) {
  return write_rows_text(filename, append != 0, pn, [=](std::ostream &os, long i) {
    os << "0 ";
    os << (type ? type[i] : 0) << " ";
    os << q4[4 * i + 0] << " ";
    os << q4[4 * i + 1] << " ";
    os << q4[4 * i + 2] << " ";
    os << p4[4 * i + 0] << " ";
    os << p4[4 * i + 1] << " ";
    os << p4[4 * i + 2] << "\n";
  });
}

// Variables::SaveConfiguration (variables.cc:45-57): "qx qy qz px py pz\n" for rows [0, tn) -- the reference
// writes the ghosts too (their momenta are scratch there; the caller passes whatever it wants written).
int mdh_save_configuration(const char *filename, const double *q4, const double *p4, long tn) {
  return write_rows_text(filename, false, tn, [=](std::ostream &os, long i) {
    os << q4[4 * i + 0];
    os << " " << q4[4 * i + 1];
    os << " " << q4[4 * i + 2];
    os << " " << p4[4 * i + 0];
    os << " " << p4[4 * i + 1];
    os << " " << p4[4 * i + 2];
    os << "\n";
  });
}

// FileDumper::Execute (mdmanager.cc:487-500): qx qy qz px py pz as raw doubles per local particle.
int mdh_save_dump(const char *filename, int append, const double *q4, const double *p4, long pn) {
  std::ofstream ofs(filename, append ? (std::ios::out | std::ios::binary | std::ios::app) : (std::ios::out | std::ios::binary));
  if (!ofs) return -1;
  std::vector<double> buf;
  const long chunk = 1L << 16;
  for (long b = 0; b < pn; b += chunk) {
    const long e = std::min(pn, b + chunk);
    buf.resize((size_t)(e - b) * 6);
    for (long i = b; i < e; i++) {
      double *o = &buf[(size_t)(i - b) * 6];
      o[0] = q4[4 * i + 0]; o[1] = q4[4 * i + 1]; o[2] = q4[4 * i + 2];
      o[3] = p4[4 * i + 0]; o[4] = p4[4 * i + 1]; o[5] = p4[4 * i + 2];
    }
    ofs.write((const char *)buf.data(), (std::streamsize)(buf.size() * sizeof(double)));
  }
  ofs.close();
  return ofs ? 0 : -1;
}

}  // extern "C"
