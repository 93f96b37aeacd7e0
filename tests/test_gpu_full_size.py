"""GPU, BASELINE.json full size (configs[1]: benchmark.cfg scaled to 16,034,412 particles on one B200):
the oracle cannot run this in seconds, so the CUDA path is held to size-independent properties of the
reference's algorithm --
  * particle number and the reference index space are conserved through rebuilds and migration;
  * Newton's third law: the impulses of one kick sum to zero (the full list evaluates every pair twice,
    independently, so this also checks that both evaluations agree);
  * total momentum is conserved (from the first rebuild on, see the test) and total energy to the 1e-5 the
    reference's stale-ghost definition of U allows;
  * every local stays inside the unit's box after the periodic wrap (variables.cc:59-79);
  * the pair-list build is idempotent and deterministic: rebuilding on unchanged positions gives the same pair
    count and bit-identical U and W;
  * the pair count equals the analytic value for the lattice state within the vacancy noise.
MDB200_TEST_UNIT_LENGTH shrinks the box for a quick local run."""
import os

import numpy as np
import pytest

from mdacp_b200 import driver, hostlib

pytestmark = pytest.mark.gpu
UNIT_LENGTH = float(os.environ.get("MDB200_TEST_UNIT_LENGTH", "284"))


def test_full_size_properties():
    cfg = {"UnitLength": repr(UNIT_LENGTH), "TimeStep": "0.001", "InitialVelocity": "1.0"}
    mdm = driver.MDManager(cfg)
    try:
        s, e = mdm.GetRect(0)
        q = hostlib.benchmark_config(s, e, 0.7, 0)
        n = len(q)
        if UNIT_LENGTH == 284.0:
            assert n == 16034412  # M2
        p = hostlib.initial_velocity(n, 1.0, 0)
        gid = np.arange(n, dtype=np.int32)
        mdm.upload([{"q": q, "p": p, "type": gid}])
        unit = mdm.GetMDUnit(0)
        mdm.MakePairList()
        c0 = unit.counts()
        o0 = unit.observe()
        assert c0["pn"] == n and c0["tn"] > n
        # listed pairs per particle of the 0.7-density FCC lattice at search length 3.3: 86 directed
        # partners (shells up to 3.095), minus the vacancies
        per = c0["npairs"] / n
        assert 42.0 < per < 44.5, per
        # idempotent, deterministic rebuild
        mdm.MakePairList()
        c1 = unit.counts()
        o1 = unit.observe()
        assert c1["npairs"] == c0["npairs"] and c1["tn"] == c0["tn"]
        assert o1[1] == o0[1] and o1[2] == o0[2], "U, W not bit-identical after an idempotent rebuild"
        # The generated lattice has a plane of particles exactly ON the low faces of the box; the reference's
        # ghost predicate is strict (mdrect.cc:24-47), so those are not ghosted until they have
        # moved and the list has been rebuilt: momentum is only conserved from the first rebuild on.  Run
        # through it first.
        mdm.Calculate(120)
        c2 = unit.counts()
        assert c2["pn"] == n, "particle number changed"
        assert c2["nbuilds"] > c1["nbuilds"], "no rebuild inside the window"
        # Newton's third law on one kick.  Ghosts are refreshed first: after a step they are half a step older
        # than the locals (SURVEY.md finding 2), which would break the symmetry of the boundary pairs.
        mdm.SendBorderParticles()
        f = unit.impulse()
        frms = float(np.sqrt((f ** 2).sum(axis=1).mean()))
        assert np.abs(f.sum(axis=0)).max() <= 1e-9 * frms * n ** 0.5, (f.sum(axis=0), frms)
        del f
        d1 = unit.download()
        o1 = unit.observe()
        e1 = (o1[0] + o1[1]) / n
        p1 = d1["p"].sum(axis=0)
        del d1
        nb = unit.counts()["nbuilds"]
        steps = 0
        while unit.counts()["nbuilds"] == nb and steps < 200:  # a window that ends right after a rebuild
            mdm.Calculate(10)
            steps += 10
        o2 = unit.observe()
        e2 = (o2[0] + o2[1]) / n
        # the reference's potential energy mixes full-step locals with half-step ghosts (SURVEY.md finding 2):
        # E is conserved to ~1e-5 relative, not better
        assert abs(e2 - e1) <= 5e-5 * abs(e1), (e1, e2)
        d = unit.download()
        assert d["pn"] == n
        assert np.array_equal(np.sort(d["type"]), gid), "reference index space not conserved"
        pt = d["p"].sum(axis=0)
        prms = float(np.sqrt((d["p"] ** 2).sum(axis=1).mean()))
        assert np.abs(pt - p1).max() <= 1e-9 * prms * n ** 0.5, (pt, p1)
        # the wrap happens at rebuilds; between rebuilds a particle may sit up to the skin outside
        lo = np.asarray(s) - 0.3
        hi = np.asarray(e) + 0.3
        qq = d["q"][:n, :3]
        assert (qq >= lo).all() and (qq <= hi).all()
    finally:
        mdm.close()
