"""GPU parity against the UNMODIFIED reference run on the box's own CPU (oracle/_ref/mdacp_scalar,
built here by oracle/Makefile from /root/reference and shipped with the snapshot).

Identical initial states (dumped from the reference driver, uploaded through the C-ABI):
  * pair-index sets at every rebuild: bit-exact (reference indices);
  * rebuild step numbers, pn, tn per step: exact;
  * kick impulse per particle: <= 1e-10 * max(|F_i|, F_rms);
  * K, U, W at every step: <= 1e-9 relative; T, P, E likewise;
  * q, p, type after 100 steps in reference index order: <= 1e-10 / exact.
"""
import numpy as np
import pytest

from common import cfg_for, force_err, refdump, states_from_dump

pytestmark = pytest.mark.gpu

TOL_OBS = 1e-9
TOL_FORCE = 1e-10
TOL_STATE = 1e-10


def _need_ref():
    if not refdump.have_ref():
        pytest.skip("oracle/_ref/mdacp_scalar not present (build it where /root/reference exists)")


def _run_case(cfg, nunits, steps, state_at_end=True, pairs_every_build=True):
    from mdacp_b200 import driver
    _need_ref()
    kv = dict(cfg)
    kv.update({"DumpSteps": steps, "DumpStateAt": "0,%d" % steps, "DumpPairsAt": "", "DumpForceAt": "",
               "DumpPairsAtBuild": "yes"})
    init_build = kv.get("DumpInit", "Benchmark") != "Collision"
    if init_build:
        kv["DumpForceAt"] = "0,%d" % steps
    rec, out = refdump.run_reference(kv, nunits=nunits)
    meta = refdump.meta_of(rec)
    mdm = driver.MDManager(cfg_for(meta), num_procs=1, rank=0, num_threads=nunits, cutoff=meta["cutoff"])
    try:
        assert tuple(mdm.pinfo.GetGridSize()) == meta["grid"]
        mdm.upload(states_from_dump(rec, 0, nunits))
        if kv.get("DumpInit") == "Langevin":
            mdm.SetSeed(int(kv.get("Seed", 1)))
        builds_ref = rec["builds"]
        pn_ref, tn_ref = rec["pn"], rec["tn"]
        obs_ref = rec["obs_KUW"]
        if init_build:
            mdm.MakePairList()
        worst = {"K": 0.0, "U": 0.0, "W": 0.0}
        last_builds = [0] * nunits
        for s in range(steps + 1):
            if s > 0:
                mdm.Calculate()
            for u in range(nunits):
                unit = mdm.GetMDUnit(u)
                c = unit.counts()
                assert c["nbuilds"] == builds_ref[s, u], "rebuild schedule differs at step %d unit %d" % (s, u)
                assert c["pn"] == pn_ref[s, u] and c["tn"] == tn_ref[s, u], "pn/tn differ at step %d unit %d" % (s, u)
                rebuilt = c["nbuilds"] != last_builds[u]
                last_builds[u] = c["nbuilds"]
                if rebuilt and pairs_every_build:
                    k, p = unit.pairs()
                    mine = refdump.pair_set(k, p)
                    ref = refdump.pair_set(rec["pair_key.s%d.u%d" % (s, u)], rec["pair_partner.s%d.u%d" % (s, u)])
                    assert mine.shape == ref.shape, "pair count differs at step %d unit %d: %s vs %s" % (
                        s, u, mine.shape, ref.shape)
                    assert np.array_equal(mine, ref), "pair set differs at step %d unit %d" % (s, u)
                if ("impulse.s%d.u%d" % (s, u)) in rec:
                    f = unit.impulse()
                    fr = rec["impulse.s%d.u%d" % (s, u)]
                    e = force_err(f, fr)
                    assert e <= TOL_FORCE, "impulse error %.3e at step %d unit %d" % (e, s, u)
                o = unit.observe()
                r = obs_ref[s, u]
                for idx, name in enumerate(("K", "U", "W")):
                    if c["nbuilds"] == 0 and idx > 0:
                        continue
                    den = abs(r[idx]) if r[idx] != 0 else 1.0
                    if name == "W":
                        # the virial is a sum with cancellation (near-zero pressure droplets): its error
                        # is measured on the scale it enters the pressure with, P = (2K/3 - W) / V
                        den = max(den, abs(r[0]))
                    worst[name] = max(worst[name], abs(o[idx] - r[idx]) / den)
            for name in worst:
                assert worst[name] <= TOL_OBS, "%s differs by %.3e (relative) at step %d" % (name, worst[name], s)
        # printed observables, combined exactly as mdmanager.cc:390-429
        tot = obs_ref[steps].sum(axis=0)
        N = float(pn_ref[steps].sum())
        ref_o = driver.observables_from_raw([tot[0], tot[1], tot[2], N], meta["L"])
        my_o = mdm.observables()
        for k in ("T", "P", "E"):
            assert abs(my_o[k] - ref_o[k]) <= TOL_OBS * max(abs(ref_o[k]), 1e-300), (k, my_o[k], ref_o[k])
        if state_at_end:
            for u in range(nunits):
                d = mdm.GetMDUnit(u).download()
                qr = rec["q.s%d.u%d" % (steps, u)]
                pr = rec["p.s%d.u%d" % (steps, u)]
                tr = rec["type.s%d.u%d" % (steps, u)]
                assert np.array_equal(d["type"], tr), "reference index order differs (unit %d)" % u
                assert d["q"].shape == qr.shape
                assert np.max(np.abs(d["q"] - qr)) <= TOL_STATE * max(meta["L"]), "q differs (unit %d)" % u
                assert np.max(np.abs(d["p"] - pr)) <= TOL_STATE * max(1.0, float(np.abs(pr).max())), "p differs"
        return worst
    finally:
        mdm.close()


def test_benchmark_single_unit_100_steps():
    """benchmark.cfg-shaped, scaled down to 5,600 particles, 1 unit (periodic self-ghosts)."""
    _run_case({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.001, "InitialVelocity": 1.0}, 1, 100)


def test_benchmark_hot_single_unit_rebuilds():
    """hotter and longer so that several margin-based rebuilds and migrations happen."""
    _run_case({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.004, "InitialVelocity": 2.5}, 1, 150)


def test_benchmark_2x2x2_units_on_one_gpu():
    """8 units (2x2x2 domain decomposition) driven in one process on one device: the same
    exchange logic the multi-GPU path uses, against the reference with 8 OpenMP units."""
    _run_case({"DumpInit": "Benchmark", "SystemSize": 28, "TimeStep": 0.004, "InitialVelocity": 2.0}, 8, 100)


def test_slab_grid_2x1x1():
    _run_case({"DumpInit": "Benchmark", "UnitLength": 16, "TimeStep": 0.004, "InitialVelocity": 2.0,
               "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 60)


def test_langevin_zero_temperature_exact():
    """Langevin path with AimedTemperature = 0: noise amplitude 0, deterministic friction, and the
    reference's missing second half drift (SURVEY.md finding 5)."""
    _run_case({"DumpInit": "Langevin", "UnitLength": 16, "TimeStep": 0.005, "InitialVelocity": 1.8,
               "Density": 0.7, "AimedTemperature": 0.0, "HeatbathGamma": 0.1, "HeatbathType": "Langevin",
               "Seed": 12345}, 1, 100)


def test_collision_first_step_builds():
    """collision.cfg-shaped: vacuum + two droplets, no MakePairList before the loop (Q22)."""
    _run_case({"DumpInit": "Collision", "SystemSizeX": 60, "SystemSizeY": 30, "SystemSizeZ": 30,
               "TimeStep": 0.002}, 1, 100)


def test_collision_two_units_load_imbalance():
    # droplets 5 apart so that they touch, exchange ghosts and migrate across the unit boundary
    _run_case({"DumpInit": "Collision", "SystemSizeX": 40, "SystemSizeY": 30, "SystemSizeZ": 30,
               "TimeStep": 0.004, "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 400)


def test_nose_hoover():
    _run_case({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.002, "InitialVelocity": 1.5,
               "ControlTemperature": "yes", "HeatbathType": "NoseHoover", "AimedTemperature": 1.0,
               "HeatbathTau": 0.1}, 1, 100)
