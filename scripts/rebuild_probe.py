"""Rebuild timing probe: runs enough steps for several margin-triggered rebuilds (MDB200_DEBUG=1 prints the stages)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdacp_b200 import driver  # noqa: E402
from perf_probe import fcc  # noqa: E402

L = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 450
q, p = fcc(L)
mdm = driver.MDManager({"SystemSize": repr(L), "TimeStep": "0.001"})
u = mdm.GetMDUnit(0)
mdm.upload([{"q": q, "p": p}])
mdm.MakePairList()
u.synchronize()
t0 = time.time()
mdm.Calculate(nsteps)
u.synchronize()
t1 = time.time()
c = u.counts()
print("L=%g N=%d: %d steps %.3f ms/step wall, %d builds, npairs/N %.2f" % (L, len(q), nsteps, 1e3 * (t1 - t0) / nsteps,
                                                                        c["nbuilds"], c["npairs"] / len(q)))
ls = u.list_stats()
print("  list: entries %d, rows mean %.1f max %d; walked by the force kernel %d (+%.1f%% padding)" % (
    ls["entries"], ls["entries"] / len(q), ls["max_row"], ls["walked_entries"],
    100.0 * (ls["walked_entries"] / max(ls["entries"], 1) - 1)))
mdm.close()
