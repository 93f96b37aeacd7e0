"""Pins the plain-C oracle (oracle/mdacp_oracle.c) against the UNMODIFIED reference, bit for bit.

Runs where oracle/_ref/mdacp_scalar exists (built by oracle/Makefile from /root/reference); the
same comparisons against the committed fixtures live in tests/test_oracle_golden.py, so the pin
also holds where the reference is absent.
"""
import numpy as np
import pytest

from common import refdump, states_from_dump
from oracle import pyoracle

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
}


def check_oracle_against_records(rec, steps, init_build, seed=None):
    """Replays a reference dump on the C oracle and asserts bit-equality everywhere."""
    meta = refdump.meta_of(rec)
    nu = meta["nunits"]
    osys = pyoracle.System.from_meta(meta, capacity=60000)
    try:
        for u, st in enumerate(states_from_dump(rec, 0, nu)):
            osys.set_unit(u, st["q"], st["p"], st["type"])
        if seed is not None:
            osys.set_seed(seed)
        if init_build:
            osys.make_pair_list()
        last = [0] * nu
        for s in range(steps + 1):
            if s > 0:
                osys.calculate(1)
            for u in range(nu):
                c = osys.counts(u)
                assert c["nbuilds"] == rec["builds"][s, u]
                assert c["pn"] == rec["pn"][s, u] and c["tn"] == rec["tn"][s, u]
                o = osys.observe(u)
                assert np.array_equal(o[:3], rec["obs_KUW"][s, u]), (s, u, o[:3], rec["obs_KUW"][s, u])
                if c["nbuilds"] != last[u]:
                    last[u] = c["nbuilds"]
                    kname = "pair_key.s%d.u%d" % (s, u)
                    if kname in rec:
                        k, p = osys.pairs(u)
                        assert np.array_equal(k, rec[kname])  # same ORDER, not only the same set
                        assert np.array_equal(p, rec["pair_partner.s%d.u%d" % (s, u)])
                if ("csr_list.s%d.u%d" % (s, u)) in rec:
                    ptr, cnt, lst = osys.csr(u)
                    assert np.array_equal(ptr, rec["csr_pointer.s%d.u%d" % (s, u)])
                    assert np.array_equal(cnt, rec["csr_count.s%d.u%d" % (s, u)])
                    assert np.array_equal(lst, rec["csr_list.s%d.u%d" % (s, u)])
                if ("impulse.s%d.u%d" % (s, u)) in rec:
                    assert np.array_equal(osys.impulse(u), rec["impulse.s%d.u%d" % (s, u)])
                if ("q.s%d.u%d" % (s, u)) in rec and s > 0:
                    st = osys.state(u)
                    assert np.array_equal(st["q"], rec["q.s%d.u%d" % (s, u)])
                    assert np.array_equal(st["p"], rec["p.s%d.u%d" % (s, u)])
                    assert np.array_equal(st["type"], rec["type.s%d.u%d" % (s, u)])
                    assert osys.zeta(u) == rec["zeta.s%d.u%d" % (s, u)][0]
    finally:
        osys.close()


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
