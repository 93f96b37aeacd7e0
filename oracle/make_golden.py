"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/mdacp_scalar).

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists and `make -C oracle ref` has
been done); the fixtures are committed so that the GPU box -- which has no /root/reference --
can check both the oracle and the CUDA path against the reference's own outputs.

Large integer arrays (pair lists) are stored as (count, sha256 of the canonical sorted array,
sha256 of the array in the reference's generation order); everything else at full precision.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refdump  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (cfg, nunits, steps)
CASES = {
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
}


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def make(name):
    cfg, nunits, steps = CASES[name]
    kv = dict(cfg)
    init_build = kv.get("DumpInit", "Benchmark") != "Collision"
    first = 0 if init_build else 1
    kv.update({"DumpSteps": steps, "DumpStateAt": "0,%d" % steps, "DumpPairsAt": "",
               "DumpForceAt": "%d,%d" % (first, steps), "DumpPairsAtBuild": "yes"})
    rec, out = refdump.run_reference(kv, nunits=nunits)
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


if __name__ == "__main__":
    if not refdump.have_ref():
        sys.exit("oracle/_ref/mdacp_scalar missing: run `make -C oracle ref` where /root/reference exists")
    for n in (sys.argv[1:] or sorted(CASES)):
        make(n)
