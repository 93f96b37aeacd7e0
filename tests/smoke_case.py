"""One small invocation of the hot path on cuda:0, checked against the oracle (used by
__graft_entry__.smoke() and by the gpu tests)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import TOL_FORCE, TOL_OBS, force_err, refdump  # noqa: E402
from mdacp_b200 import driver, hostlib  # noqa: E402
from oracle import pyoracle  # noqa: E402


def run(verbose=False, L=14.0, steps=20):
    cfg = {"SystemSize": repr(L), "TimeStep": "0.004", "InitialVelocity": "2.0"}
    mdm = driver.MDManager(cfg)
    osys = pyoracle.System([L] * 3, 0.004, capacity=40000)
    try:
        s, e = mdm.GetRect(0)
        q = hostlib.benchmark_config(s, e, 0.7, 0)
        p = hostlib.initial_velocity(len(q), 2.0, 0)
        mdm.upload([{"q": q, "p": p}])
        osys.set_unit(0, q, p)
        mdm.MakePairList()
        osys.make_pair_list()
        unit = mdm.GetMDUnit(0)
        k, pr = unit.pairs()
        ko, po = osys.pairs(0)
        assert np.array_equal(refdump.pair_set(k, pr), refdump.pair_set(ko, po)), "pair set differs from the oracle"
        ef = force_err(unit.impulse(), osys.impulse(0))
        assert ef <= TOL_FORCE, "impulse differs from the oracle by %.3e" % ef
        mdm.Calculate(steps)
        osys.calculate(steps)
        o, r = unit.observe(), osys.observe(0)
        err = max(abs(o[i] - r[i]) / max(abs(r[i]), abs(r[0])) for i in range(3))
        assert err <= TOL_OBS, "K/U/W differ from the oracle by %.3e" % err
        assert unit.counts()["nbuilds"] == osys.counts(0)["nbuilds"]
        if verbose:
            print("smoke ok: N=%d pairs=%d impulse err %.2e, K/U/W err %.2e after %d steps, %d list builds" % (
                len(q), len(k), ef, err, steps, unit.counts()["nbuilds"]))
        return ef, err
    finally:
        mdm.close()
        osys.close()
