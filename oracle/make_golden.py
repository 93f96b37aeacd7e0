"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/mdacp_scalar).

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists and `make -C oracle ref` has
been done); the fixtures are committed so that the GPU box -- which has no /root/reference --
can check both the oracle and the CUDA path against the reference's own outputs.

Large integer arrays (pair lists) are stored as (count, sha256 of the canonical sorted array,
sha256 of the array in the reference's generation order); everything else at full precision.
"""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refdump  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (cfg, nunits, steps[, reference binary under oracle/_ref; default mdacp_scalar])
CASES = {
    # cutoff 2.5 (BASELINE.json configs[0]; mdconfig.h:15-16 -- a compile-time constant of the reference, so these
    # two come from oracle/_ref/mdacp_scalar_c25, the scalar build of a scratch copy with that one line edited)
    "bench1_c25": ({"DumpInit": "Benchmark", "SystemSize": 14, "TimeStep": 0.004, "InitialVelocity": 2.5}, 1, 100,
                   "mdacp_scalar_c25"),
    "bench8_c25": ({"DumpInit": "Benchmark", "SystemSize": 14, "TimeStep": 0.004, "InitialVelocity": 2.0}, 8, 100,
                   "mdacp_scalar_c25"),
    "bench1": ({"DumpInit": "Benchmark", "SystemSize": 14, "TimeStep": 0.004, "InitialVelocity": 2.5}, 1, 100),
    "bench8": ({"DumpInit": "Benchmark", "SystemSize": 14, "TimeStep": 0.004, "InitialVelocity": 2.0}, 8, 100),
    "slab2": ({"DumpInit": "Benchmark", "UnitLength": 10, "TimeStep": 0.004, "InitialVelocity": 2.0,
               "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 100),
    "langevinT0": ({"DumpInit": "Langevin", "UnitLength": 12, "TimeStep": 0.005, "InitialVelocity": 1.8,
                    "Density": 0.7, "AimedTemperature": 0.0, "HeatbathGamma": 0.1, "HeatbathType": "Langevin",
                    "Seed": 12345}, 1, 100),
    "langevin": ({"DumpInit": "Langevin", "UnitLength": 12, "TimeStep": 0.005, "InitialVelocity": 1.8,
                  "Density": 0.7, "AimedTemperature": 0.9, "HeatbathGamma": 0.1, "HeatbathType": "Langevin",
                  "Seed": 12345}, 1, 100),
    "collision2": ({"DumpInit": "Collision", "SystemSizeX": 40, "SystemSizeY": 30, "SystemSizeZ": 30,
                    "TimeStep": 0.004, "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 300),
    "nosehoover": ({"DumpInit": "Benchmark", "SystemSize": 14, "TimeStep": 0.002, "InitialVelocity": 1.5,
                    "ControlTemperature": "yes", "HeatbathType": "NoseHoover", "AimedTemperature": 1.0,
                    "HeatbathTau": 0.1}, 1, 100),
    # multi-unit versions of the thermostat / droplet cases (global temperature reduction, migration between
    # eight sub-boxes, empty and nearly empty units)
    "langevinT0_8": ({"DumpInit": "Langevin", "UnitLength": 8, "TimeStep": 0.005, "InitialVelocity": 1.8,
                      "Density": 0.7, "AimedTemperature": 0.0, "HeatbathGamma": 0.1, "HeatbathType": "Langevin",
                      "Seed": 12345}, 8, 100),
    "nosehoover8": ({"DumpInit": "Benchmark", "SystemSize": 16, "TimeStep": 0.002, "InitialVelocity": 1.5,
                     "ControlTemperature": "yes", "HeatbathType": "NoseHoover", "AimedTemperature": 1.0,
                     "HeatbathTau": 0.1}, 8, 100),
    "collision8": ({"DumpInit": "Collision", "SystemSizeX": 40, "SystemSizeY": 30, "SystemSizeZ": 30,
                    "TimeStep": 0.004}, 8, 300),
    # MDManager::ChangeScale (no caller in the shipped projects): compress by 2 % after step 40, expand after ... once
    "scale8": ({"DumpInit": "Benchmark", "SystemSize": 16, "TimeStep": 0.004, "InitialVelocity": 2.0,
                "DumpChangeScaleAt": 40, "DumpChangeScaleAlpha": 0.98}, 8, 80),
    "scale1": ({"DumpInit": "Benchmark", "SystemSize": 14, "TimeStep": 0.004, "InitialVelocity": 2.0,
                "DumpChangeScaleAt": 30, "DumpChangeScaleAlpha": 1.03}, 1, 80),
    # the reference's own file writers on the final state (SaveAsCdview, SaveAsDumpFile, SaveConfiguration)
    "files2": ({"DumpInit": "Benchmark", "UnitLength": 8, "TimeStep": 0.004, "InitialVelocity": 2.0,
                "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1, "DumpSaveFiles": "yes"}, 2, 40),
}


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def make(name):
    cfg, nunits, steps = CASES[name][:3]
    exe = os.path.join(HERE, "_ref", CASES[name][3]) if len(CASES[name]) > 3 else refdump.REF_SCALAR
    kv = dict(cfg)
    init_build = kv.get("DumpInit", "Benchmark") != "Collision"
    first = 0 if init_build else 1
    kv.update({"DumpSteps": steps, "DumpStateAt": "0,%d" % steps, "DumpPairsAt": "",
               "DumpForceAt": "%d,%d" % (first, steps), "DumpPairsAtBuild": "yes"})
    workdir = tempfile.mkdtemp(prefix="gold_")
    rec, out = refdump.run_reference(kv, nunits=nunits, exe=exe, workdir=workdir)
    files = {}
    for fn in sorted(os.listdir(workdir)):
        if fn.startswith("final.") or fn.endswith(".cnf"):
            files["file." + fn] = np.frombuffer(open(os.path.join(workdir, fn), "rb").read(), dtype=np.uint8)
    shutil.rmtree(workdir)
    keep = {}
    for k, v in rec.items():
        if k.startswith("pair_key."):
            tag = k[len("pair_key."):]
            key, partner = v, rec["pair_partner." + tag]
            canon = refdump.pair_set(key, partner).astype(np.int32)
            keep["paircount." + tag] = np.array([len(key)], dtype=np.int64)
            keep["pairhash." + tag] = sha(canon)
            keep["pairorder." + tag] = sha(np.stack([key, partner], axis=1).astype(np.int32))
        elif k.startswith("pair_partner.") or k.startswith("csr_"):
            continue
        else:
            keep[k] = v
    keep.update(files)
    keep["cfg_keys"] = np.array(list(cfg.keys()))
    keep["cfg_vals"] = np.array([str(x) for x in cfg.values()])
    keep["steps"] = np.array([steps])
    keep["init_build"] = np.array([1 if init_build else 0])
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **keep)
    print("%-12s units %d steps %d N %d builds %s -> %s (%.0f kB)" % (
        name, nunits, steps, rec["pn"][0].sum(), rec["builds"][-1].tolist(), os.path.relpath(path, ROOT),
        os.path.getsize(path) / 1e3))




# ---------------------------------------------------------------------------------------------
# Whole-project runs of the unmodified reference (its own Mode= loops, its own printed lines and cdview
# files): expected output for the stand-alone driver tests (tests/test_gpu_cxx_driver.py).
# ---------------------------------------------------------------------------------------------
PROJECT_RUNS = {
    "burst": "Mode=Burst\nDensity=0.8\nThermalizeLoop=100\nAimedTemperature=1.5\nSystemSize=30\nRadius=6\n"
             "TimeStep=0.005\nTotalLoop=201\nObserveLoop=100\nInitialVelocity=0.9\n",
    "collision": "Mode=Collision\nSystemSizeX=40\nSystemSizeY=30\nSystemSizeZ=30\nTimeStep=0.004\nTotalLoop=101\n"
                 "ObserveLoop=50\nSaveCdviewFile=yes\n",
}


def make_project_run(name):
    import subprocess
    out = os.path.join(GOLDEN, "projects", name)
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="proj_")
    with open(os.path.join(work, "input.cfg"), "w") as f:
        f.write(PROJECT_RUNS[name])
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([refdump.REF_SCALAR, "input.cfg"], cwd=work, env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=3600, check=True)
    with open(os.path.join(out, "input.cfg"), "w") as f:
        f.write(PROJECT_RUNS[name])
    with open(os.path.join(out, "stdout.txt"), "w") as f:
        f.write(res.stdout)
    files = sorted(fn for fn in os.listdir(work) if fn.endswith(".cd"))
    keep = [files[0], files[-1]] if len(files) > 1 else files  # first and last snapshot
    for fn in keep:
        shutil.copy(os.path.join(work, fn), os.path.join(out, fn))
    shutil.rmtree(work)
    print("%-12s project run -> %s (%s)" % (name, os.path.relpath(out, ROOT), ", ".join(keep)))


if __name__ == "__main__":
    if not refdump.have_ref():
        sys.exit("oracle/_ref/mdacp_scalar missing: run `make -C oracle ref` where /root/reference exists")
    for n in (sys.argv[1:] or sorted(CASES) + sorted(PROJECT_RUNS)):
        if n in CASES:
            make(n)
        else:
            make_project_run(n)
