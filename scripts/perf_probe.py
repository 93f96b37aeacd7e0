"""Quick performance probe (not the bench): FCC LJ at several sizes, per-phase device timers."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdacp_b200 import capi, driver  # noqa: E402


def fcc(L, density=0.7, seed=0, v0=1.0):
    s = 1.0 / (density * 0.25) ** (1.0 / 3.0)
    n = int(L / s)
    s = L / n
    hs = 0.5 * s
    g = np.arange(n) * s
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    base = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    offs = np.array([[0, 0, 0], [0, hs, hs], [hs, 0, hs], [hs, hs, 0]])
    q = (base[:, None, :] + offs[None, :, :]).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    q += rng.uniform(-0.02, 0.02, q.shape)
    q = np.mod(q, L)
    p = rng.normal(size=q.shape)
    p *= v0 / np.sqrt((p ** 2).sum(axis=1, keepdims=True))
    p -= p.mean(axis=0)
    return q, p


def main():
    sizes = [float(x) for x in sys.argv[1:]] or [50.0, 120.0]
    for L in sizes:
        q, p = fcc(L)
        n = q.shape[0]
        mdm = driver.MDManager({"SystemSize": repr(L), "TimeStep": "0.001"})
        u = mdm.GetMDUnit(0)
        t0 = time.time()
        mdm.upload([{"q": q, "p": p}])
        t1 = time.time()
        mdm.MakePairList()
        u.synchronize()
        t2 = time.time()
        c = u.counts()
        print("L=%g N=%d tn=%d npairs=%d (%.2f/particle) upload %.3fs first build %.3fs" % (
            L, n, c["tn"], c["npairs"], c["npairs"] / n, t1 - t0, t2 - t1), flush=True)
        ls = u.list_stats()
        print("  list: %d force-warps, %d lanes per local, entries %d padded %d (+%.1f%%) mean row %.1f max %d" % (
            ls["warps"], ls["lanes_per_local"], ls["entries"], ls["padded_entries"],
            100.0 * (ls["padded_entries"] / max(ls["entries"], 1) - 1), ls["entries"] / max(n, 1), ls["max_row"]), flush=True)
        mdm.Calculate(20)  # warm-up
        u.synchronize()
        o0 = mdm.observables()
        u.timers(reset=True)
        nst = 100
        t3 = time.time()
        mdm.Calculate(nst)
        u.synchronize()
        t4 = time.time()
        tm = u.timers()
        c2 = u.counts()
        o1 = mdm.observables()
        print("  %d steps: %.3f ms/step wall, %.3f G updates/s, builds %d, launches/step %.1f, E %.10f -> %.10f" % (
            nst, 1e3 * (t4 - t3) / nst, n * nst / (t4 - t3) / 1e9, c2["nbuilds"] - c["nbuilds"], tm["launches"] / nst,
            o0["E"], o1["E"]), flush=True)
        u.set_profiling(True)
        u.timers(reset=True)
        mdm.Calculate(nst)
        u.synchronize()
        tm = u.timers()
        nb = u.counts()["nbuilds"] - c2["nbuilds"]
        print("  phases (ms/step): force %.3f comm %.3f integrate %.3f check %.3f | pair %.3f ms per rebuild (%d rebuilds)" % (
            tm["force"] / nst, tm["comm"] / nst, tm["integrate"] / nst, tm["check"] / nst,
            tm["pair"] / max(nb, 1), nb), flush=True)
        P = c["npairs"] / n
        tf = tm["force"] / nst * 1e-3
        if tf > 0:
            print("  force kernel: %.3f G particles/s; algorithmic %.2f TFLOP/s (30 flop/pair), %.2f TB/s (4P+96 B)" % (
                n / tf / 1e9, 30 * P * n / tf / 1e12, (4 * P + 96) * n / tf / 1e12), flush=True)
        mdm.close()


if __name__ == "__main__":
    main()
