#!/usr/bin/env python3
"""Binds libmdacp_b200.so into a SCRATCH COPY of the reference driver (sections 1-2).

    apply_binding.py <dir with a copy of /root/reference/*.cc *.h>

Edits four files of the copy in place -- mdunit.h, mdunit.cc, mdmanager.h, mdmanager.cc -- and nothing else:
main.cc, the projects, Parameter, ParaInfo, SimulationInfo, the configuration makers and all printing stay as they
are.  Every edit is anchored on the exact reference text it replaces and fails loudly if that text is not found
exactly once, so a changed reference cannot be patched silently.  No reference source is stored in this repository:
the anchors below are the signatures / single lines being replaced, everything inserted is ours.
"""
import os
import sys


def sub_once(text, old, new, what):
    if text.count(old) != 1:
        raise SystemExit("apply_binding: anchor for %s found %d times: %r" % (what, text.count(old), old))
    return text.replace(old, new)


def replace_body(text, signature, body, what):
    """replaces the brace-matched body that follows `signature` (which ends right before the opening brace)"""
    if text.count(signature) != 1:
        raise SystemExit("apply_binding: signature for %s found %d times" % (what, text.count(signature)))
    start = text.index(signature) + len(signature)
    open_at = text.index("{", start)
    depth, i = 0, open_at
    while True:
        ch = text[i]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return text[:open_at] + "{\n" + body + "}" + text[i + 1:]


def main(d):
    def rd(n):
        return open(os.path.join(d, n)).read()

    def wr(n, t):
        open(os.path.join(d, n), "w").write(t)

    # ---------------------------------------------------------------- mdunit.h (mdunit.h:27, 42, 49-59)
    t = rd("mdunit.h")
    t = sub_once(t, '#include "fcalculator.h"\n', '#include "fcalculator.h"\n#include "mdb200_glue.h"\n', "glue include")
    t = sub_once(t, "  Variables *vars;\n", "  Mdb200Vars vars;      // host Variables behind a residency check\n"
                                           "  Mdb200UnitGlue glue;  // device twin of this unit\n", "vars member")
    t = sub_once(t, "  Variables *GetVariables(void) {return vars;}\n",
                 "  Variables *GetVariables(void) {return vars;}\n"
                 "  Variables *HostVariables(void) {return vars.ptr;}\n"
                 "  Mdb200UnitGlue *GetGlue(void) {return &glue;}\n", "accessors")
    t = sub_once(t, "  int GetParticleNumber(void) {return vars->GetParticleNumber();}\n",
                 "  int GetParticleNumber(void) {return glue.ParticleNumber(vars.ptr);}\n", "GetParticleNumber")
    t = sub_once(t, "  int GetTotalParticleNumber(void) {return vars->GetTotalParticleNumber();}\n",
                 "  int GetTotalParticleNumber(void) {return glue.TotalParticleNumber(vars.ptr);}\n", "GetTotalParticleNumber")
    fw = "{glue.ToDevice(vars.ptr); mdb200_glue::Check(%s); glue.Stepped();}"
    t = sub_once(t, "  void CalculateForce(void) {ForceCalculator::CalculateForce(vars, mesh, sinfo);}\n",
                 "  void CalculateForce(void) " + fw % "mdb200_calculate_force(glue.dev)" + "\n", "CalculateForce")
    t = sub_once(t, "  void UpdatePositionHalf(void) {ForceCalculator::UpdatePositionHalf(vars, sinfo);}\n",
                 "  void UpdatePositionHalf(void) " + fw % "mdb200_update_position_half(glue.dev)" + "\n", "UpdatePositionHalf")
    t = sub_once(t, "  void HeatbathZeta(double t) {ForceCalculator::HeatbathZeta(vars, t, sinfo);}\n",
                 "  void HeatbathZeta(double t) " + fw % "mdb200_heatbath_zeta(glue.dev, t)" + "\n", "HeatbathZeta")
    t = sub_once(t, "  void HeatbathMomenta(void) {ForceCalculator::HeatbathMomenta(vars, sinfo);}\n",
                 "  void HeatbathMomenta(void) " + fw % "mdb200_heatbath_momenta(glue.dev)" + "\n", "HeatbathMomenta")
    t = sub_once(t, "  void SetSeed(const int seed){mt.seed(seed + GetID());}\n",
                 "  void SetSeed(const int seed){mt.seed(seed + GetID()); mdb200_glue::Check(mdb200_set_seed(glue.dev, seed));}\n",
                 "SetSeed")
    t = sub_once(t, "  void Langevin(void) {ForceCalculator::Langevin(vars, sinfo, mt);}\n",
                 "  void Langevin(void) " + fw % "mdb200_langevin(glue.dev)" + "\n", "Langevin")
    wr("mdunit.h", t)

    # ---------------------------------------------------------------- mdunit.cc (mdunit.cc:8-33)
    t = rd("mdunit.cc")
    t = sub_once(t, "  mt.seed(GetID());\n}\n", "  mt.seed(GetID());\n  glue.Create(id, sinfo, pinfo);\n  vars.glue = &glue;\n}\n",
                 "constructor")
    t = sub_once(t, "  delete vars;\n", "  vars.glue = 0;\n  glue.Destroy();\n  delete vars.ptr;\n", "destructor")
    wr("mdunit.cc", t)

    # ---------------------------------------------------------------- mdmanager.h (mdmanager.h:79-80)
    t = rd("mdmanager.h")
    t = sub_once(t, "  void SetControlTemperature(bool b) {sinfo->ControlTemperature = b;}\n",
                 "  void SetControlTemperature(bool b) {sinfo->ControlTemperature = b; mdb200_glue::SetControlTemperature(mdv, b);}\n",
                 "SetControlTemperature")
    t = sub_once(t, "  void SetAimedTemperature(double t) {sinfo->AimedTemperature = t;}\n",
                 "  void SetAimedTemperature(double t) {sinfo->AimedTemperature = t; mdb200_glue::SetAimedTemperature(mdv, t);}\n",
                 "SetAimedTemperature")
    wr("mdmanager.h", t)

    # ---------------------------------------------------------------- mdmanager.cc
    t = rd("mdmanager.cc")
    # constructor (mdmanager.cc:11-47): wire the units of this process (and the other ranks) together
    t = sub_once(t, "  s_time = 0.0;\n}\n", "  s_time = 0.0;\n  mdb200_glue::CommInit(mdv, num_procs, rank);\n}\n", "constructor")
    # the step (mdmanager.cc:116-148)
    t = replace_body(t, "void\nMDManager::Calculate(void) ",
                     "  static StopWatch swAll(GetRank(), \"all\");\n  swAll.Start();\n"
                     "  mdb200_glue::Calculate(mdv);  // expiry test, rebuild, drift, ghosts, kick (+ thermostat), drift\n"
                     "  s_time += sinfo->TimeStep;\n  swAll.Stop();\n", "Calculate")
    t = replace_body(t, "void\nMDManager::MakePairList(void) ", "  mdb200_glue::MakePairList(mdv);\n", "MakePairList")  # :262-280
    t = replace_body(t, "void\nMDManager::SendBorderParticles(void) ", "  mdb200_glue::SendBorderParticles(mdv);\n",
                     "SendBorderParticles")  # :283-292
    t = replace_body(t, "bool\nMDManager::IsPairListExpired(void) ", "  return mdb200_glue::IsPairListExpired(mdv);\n",
                     "IsPairListExpired")  # :369-377
    # observers (mdmanager.cc:390-429): one fused device observation per state, {K, U, W, N}
    t = replace_body(t, "double\nMDManager::ConfigurationTemperature(void) ",
                     "  const double *o = mdb200_glue::Observe(mdv);\n  return o[2] / o[3];\n", "ConfigurationTemperature")
    t = replace_body(t, "double\nMDManager::KineticEnergy(void) ",
                     "  const double *o = mdb200_glue::Observe(mdv);\n  return o[0] / o[3];\n", "KineticEnergy")
    t = replace_body(t, "double\nMDManager::PotentialEnergy(void) ",
                     "  const double *o = mdb200_glue::Observe(mdv);\n  return o[1] / o[3];\n", "PotentialEnergy")
    t = replace_body(t, "double\nMDManager::Pressure(void) ",
                     "  const double *o = mdb200_glue::Observe(mdv);\n  const double pn = o[3];\n  const double phi = o[2] / pn;\n"
                     "  const double T = Temperature();\n  const double V = sinfo->L[X] * sinfo->L[Y] * sinfo->L[Z];\n"
                     "  return (T - phi) * pn / V;\n", "Pressure")
    # MDManager::ChangeScale (mdmanager.cc:432-443)
    t = replace_body(t, "void\nMDManager::ChangeScale(double alpha) ",
                     "  mdb200_glue::ChangeScale(mdv, alpha);  // rebuild, scale box / rects / positions / cells, rebuild\n"
                     "  sinfo->L[X] = sinfo->L[X] * alpha;\n  sinfo->L[Y] = sinfo->L[Y] * alpha;\n  sinfo->L[Z] = sinfo->L[Z] * alpha;\n"
                     "  for (int i = 0; i < num_threads; i++) mdv[i]->GetRect()->ChangeScale(alpha);\n"
                     "  mout << \"# Change Scale \" << alpha << std::endl;\n", "ChangeScale")
    wr("mdmanager.cc", t)
    print("apply_binding: mdunit.h, mdunit.cc, mdmanager.h, mdmanager.cc patched in", d)


if __name__ == "__main__":
    main(sys.argv[1])
