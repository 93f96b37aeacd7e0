"""Pins the plain-C oracle (oracle/mdacp_oracle.c) against the UNMODIFIED reference, bit for bit.

Runs where oracle/_ref/mdacp_scalar exists (built by oracle/Makefile from /root/reference); the
same comparisons against the committed fixtures live in tests/test_oracle_golden.py, so the pin
also holds where the reference is absent.
"""
import pytest

from common import check_oracle_against_records, refdump

CASES = {
    "bench1": ({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.004, "InitialVelocity": 2.5}, 1, 120),
    "bench8": ({"DumpInit": "Benchmark", "SystemSize": 28, "TimeStep": 0.004, "InitialVelocity": 2.0}, 8, 80),
    "slab2": ({"DumpInit": "Benchmark", "UnitLength": 16, "TimeStep": 0.004, "InitialVelocity": 2.0,
               "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 60),
    "langevinT0": ({"DumpInit": "Langevin", "UnitLength": 16, "TimeStep": 0.005, "InitialVelocity": 1.8,
                    "Density": 0.7, "AimedTemperature": 0.0, "HeatbathGamma": 0.1, "HeatbathType": "Langevin",
                    "Seed": 12345}, 1, 60),
    "langevin": ({"DumpInit": "Langevin", "UnitLength": 16, "TimeStep": 0.005, "InitialVelocity": 1.8,
                  "Density": 0.7, "AimedTemperature": 0.9, "HeatbathGamma": 0.1, "HeatbathType": "Langevin",
                  "Seed": 12345, "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 60),
    "collision2": ({"DumpInit": "Collision", "SystemSizeX": 40, "SystemSizeY": 30, "SystemSizeZ": 30,
                    "TimeStep": 0.004, "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 300),
    "nosehoover": ({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.002, "InitialVelocity": 1.5,
                    "ControlTemperature": "yes", "HeatbathType": "NoseHoover", "AimedTemperature": 1.0,
                    "HeatbathTau": 0.1}, 1, 60),
    # MDManager::ChangeScale in the middle of the run (compression, 8 units; expansion, 1 unit)
    "scale8": ({"DumpInit": "Benchmark", "SystemSize": 24, "TimeStep": 0.004, "InitialVelocity": 2.0,
                "DumpChangeScaleAt": 30, "DumpChangeScaleAlpha": 0.97}, 8, 70),
    "scale1": ({"DumpInit": "Benchmark", "SystemSize": 18, "TimeStep": 0.004, "InitialVelocity": 2.0,
                "DumpChangeScaleAt": 25, "DumpChangeScaleAlpha": 1.04}, 1, 70),
}


def dump_cfg(cfg, steps):
    kv = dict(cfg)
    init_build = kv.get("DumpInit", "Benchmark") != "Collision"
    first = 0 if init_build else 1
    kv.update({"DumpSteps": steps, "DumpStateAt": "0,%d,%d" % (steps // 2, steps),
               "DumpPairsAt": "%d,%d" % (first, steps), "DumpForceAt": "%d,%d" % (first, steps),
               "DumpPairsAtBuild": "yes"})
    return kv, init_build


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_bitwise(name):
    if not refdump.have_ref():
        pytest.skip("oracle/_ref/mdacp_scalar not built (needs /root/reference)")
    cfg, nunits, steps = CASES[name]
    kv, init_build = dump_cfg(cfg, steps)
    rec, _ = refdump.run_reference(kv, nunits=nunits)
    seed = int(cfg["Seed"]) if cfg.get("DumpInit") == "Langevin" else None
    check_oracle_against_records(rec, steps, init_build, seed)
