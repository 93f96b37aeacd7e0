"""Multi-process parity of the CUDA path against the CPU oracle -- TEST INFRASTRUCTURE (uses oracle/ as the checker).

One process per GPU, one unit per process, ghost refresh / migration / reductions through the library's own
transport (peer-to-peer windows or NCCL).  Every rank runs its own copy of the (small) oracle system with the SAME
unit grid (SURVEY.md finding 3) and compares its own unit at every step:
pair set exact at every rebuild, rebuild schedule / pn / tn exact, K, U, W <= 1e-9, final q, p <= 1e-10,
reference index order exact, and the all-reduced global sums against the oracle's.

Used by scripts/multigpu_parity.py (stand-alone, torchrun), tests/test_gpu_multiprocess.py (spawns torchrun) and
bench.py's `parity_n` block (after its timed regions, untimed)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run_parity(world, rank, local_rank, unique_id, unit_length=14.0, steps=100, dt=0.004, v0=2.0, observe_every=1):
    """-> dict(ok, error, grid, n_unit, builds, worst_KUW, worst_q, worst_p); never raises on a parity failure."""
    from common import TOL_OBS, TOL_STATE, refdump
    from mdacp_b200 import driver, hostlib
    from oracle import pyoracle
    res = {"ok": False, "error": None, "steps": steps, "unit_length": unit_length}
    cfg = {"UnitLength": repr(unit_length), "TimeStep": repr(dt), "InitialVelocity": repr(v0)}
    mdm = driver.MDManager(cfg, num_procs=world, rank=rank, num_threads=1, devices=[local_rank], unique_id=unique_id)
    osys = None
    try:
        grid = tuple(mdm.pinfo.GetGridSize())
        L = mdm.sinfo.L
        res["grid"] = list(grid)
        osys = pyoracle.System(L, dt, grid=grid, capacity=300000)
        assert osys.nunits == world
        for u in range(world):
            lo, hi = osys.rect(u)
            q = hostlib.benchmark_config(lo, hi, 0.7, u)
            p = hostlib.initial_velocity(len(q), v0, u)
            t = np.arange(len(q), dtype=np.int32) + 1000000 * u
            osys.set_unit(u, q, p, t)
            if u == rank:
                s, e = mdm.GetRect(0)
                assert np.allclose(s, lo) and np.allclose(e, hi)
                mdm.upload([{"q": q, "p": p, "type": t}])
        unit = mdm.GetMDUnit(0)
        mdm.MakePairList()
        osys.make_pair_list()
        worst, last = 0.0, 0
        c = unit.counts()
        for s in range(steps + 1):
            if s > 0:
                mdm.Calculate()
                osys.calculate(1)
            c, oc = unit.counts(), osys.counts(rank)
            assert c["nbuilds"] == oc["nbuilds"], "step %d: builds %d vs %d" % (s, c["nbuilds"], oc["nbuilds"])
            assert c["pn"] == oc["pn"] and c["tn"] == oc["tn"], "step %d: pn/tn %s vs %s" % (s, c, oc)
            if c["nbuilds"] != last:
                last = c["nbuilds"]
                k, p = unit.pairs()
                ko, po = osys.pairs(rank)
                assert np.array_equal(refdump.pair_set(k, p), refdump.pair_set(ko, po)), "step %d: pair set" % s
            if s % observe_every == 0 or s == steps:
                o, r = unit.observe(), osys.observe(rank)
                for i in range(3):
                    worst = max(worst, abs(o[i] - r[i]) / max(abs(r[i]), abs(r[0])))
                assert worst <= TOL_OBS, "step %d: K/U/W differ by %.3e" % (s, worst)
        tot = mdm.observe_raw()   # all-reduced through the library's transport
        ref = osys.observe(-1)
        for i in range(4):
            assert abs(tot[i] - ref[i]) <= TOL_OBS * max(abs(ref[i]), abs(ref[0])), ("global", i, tot[i], ref[i])
        d, st = unit.download(), osys.state(rank)
        assert np.array_equal(d["type"], st["type"]), "reference index order differs"
        wq = float(np.max(np.abs(d["q"] - st["q"]))) / max(L) if len(d["q"]) else 0.0
        wp = float(np.max(np.abs(d["p"] - st["p"]))) / max(1.0, float(np.abs(st["p"]).max())) if len(d["p"]) else 0.0
        assert wq <= TOL_STATE and wp <= TOL_STATE, "final state differs: q %.3e p %.3e" % (wq, wp)
        res.update({"ok": True, "n_unit": int(c["pn"]), "builds": int(c["nbuilds"]), "worst_KUW": worst,
                    "worst_q": wq, "worst_p": wp})
    except AssertionError as ex:
        res["error"] = "rank %d: %s" % (rank, ex)
    finally:
        if osys is not None:
            osys.close()
        mdm.close()
    return res
