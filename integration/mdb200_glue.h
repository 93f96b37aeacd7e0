// Glue between the reference driver (MDUnit / MDManager, unchanged apart from the forwards that
// integration/apply_binding.py inserts) and libmdacp_b200.so.  Our code; it is copied next to a scratch copy of the
// reference sources at build time (integration/Makefile) and never lives in the reference tree.
//
// Residency.  The reference keeps all particle data in the host `Variables` of each MDUnit; the library keeps it
// on the device.  Mdb200Vars stands where `Variables *vars` stood in MDUnit (mdunit.h:27): every host access
// (`vars->...`, GetVariables(), the configuration makers' AddParticle, the cdview / dump writers) first brings the
// host copy up to date (one download, only if the device has stepped since), and every device call first uploads
// the host copy IF the host code changed it since (checksum of q, p, type and the particle count).  Projects
// (benchmark.cc, langevin.cc, collision.cc, burst.cc), cfg parsing and all output therefore run unmodified.
#ifndef MDB200_GLUE_H
#define MDB200_GLUE_H
#include <vector>
#include "mdacp_b200.h"

class Variables;
class SimulationInfo;
class ParaInfo;
class MDUnit;

class Mdb200UnitGlue {
public:
  mdb200_unit *dev;
  bool dev_newer;               // the device has stepped since the host copy was last brought up to date
  bool host_valid;              // the host copy has been touched since the last upload (checked by checksum)
  unsigned long long host_sum;  // checksum of the host copy at the last upload / download
  int host_pn;
  Mdb200UnitGlue(void) : dev(0), dev_newer(false), host_valid(true), host_sum(0), host_pn(-1) {}
  void Create(int id, SimulationInfo *sinfo, ParaInfo *pinfo);
  void Destroy(void);
  void ToHost(Variables *vars);     // before any host access
  void ToDevice(Variables *vars);   // before any device call
  int ParticleNumber(Variables *vars);
  int TotalParticleNumber(Variables *vars);
  void Stepped(void);               // after a device call that changed the state
};

// stands where `Variables *vars` stood (mdunit.h:27)
class Mdb200Vars {
public:
  Variables *ptr;
  Mdb200UnitGlue *glue;
  Mdb200Vars(void) : ptr(0), glue(0) {}
  Mdb200Vars &operator=(Variables *p) { ptr = p; return *this; }
  Variables *operator->(void) { if (glue) glue->ToHost(ptr); return ptr; }
  operator Variables *(void) { if (glue) glue->ToHost(ptr); return ptr; }
};

namespace mdb200_glue {
void Check(int rc);  // non-zero: the reference's own behaviour, show_error + exit(1) (mdconfig.h:21)
void CommInit(std::vector<MDUnit *> &mdv, int num_procs, int rank);
void MakePairList(std::vector<MDUnit *> &mdv);
void Calculate(std::vector<MDUnit *> &mdv);
void SendBorderParticles(std::vector<MDUnit *> &mdv);
bool IsPairListExpired(std::vector<MDUnit *> &mdv);
void ChangeScale(std::vector<MDUnit *> &mdv, double alpha);
// raw sums {K, U, W, N} over all units and ranks (observer.cc:5-79 + mdmanager.cc:359-366), computed once per state
const double *Observe(std::vector<MDUnit *> &mdv);
void SetControlTemperature(std::vector<MDUnit *> &mdv, bool b);
void SetAimedTemperature(std::vector<MDUnit *> &mdv, double t);
}  // namespace mdb200_glue
#endif
