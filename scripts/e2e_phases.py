"""Where the end-to-end time goes (not the bench): the phases of bench.py's `e2e` region at the benchmark size, by
host wall clock with a synchronisation after each: upload from pinned host, first Calculate (list build), K - 1 steps,
observation, download.  MDB200_DEBUG=1 adds the laps inside the rebuild on stderr."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdacp_b200 import driver, hostlib  # noqa: E402


BEGIN = os.environ.get("E2E_UPLOAD_BEGIN", "1") != "0"


def main():
    UL = float(sys.argv[1]) if len(sys.argv) > 1 else 284.0
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    therm = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    cfg = {"Mode": "Benchmark", "UnitLength": repr(UL), "TimeStep": "0.001"}
    mdm = driver.MDManager(cfg, devices=[0], cutoff=3.0)
    unit = mdm.GetMDUnit(0)
    s, e = mdm.GetRect(0)
    vol = (e[0] - s[0]) * (e[1] - s[1]) * (e[2] - s[2])
    cap = int((int(vol * 0.7) + 64) * 1.02) + 65536
    q_pin = torch.zeros((cap, 4), dtype=torch.float64).pin_memory()
    p_pin = torch.zeros((cap, 4), dtype=torch.float64).pin_memory()
    qh, ph = q_pin.numpy(), p_pin.numpy()
    n = hostlib.benchmark_config(s, e, 0.7, 0, out=qh).shape[0]
    hostlib.initial_velocity(n, 1.0, 0, out=ph)
    unit.upload4(qh[:n], ph[:n])
    mdm.MakePairList()
    for i in range(therm):
        mdm.Calculate()
    dd = unit.download(want_type=False, out=(qh, ph))
    n = dd["pn"]
    unit.synchronize()
    for rep in range(3):
        t = [time.time()]
        (unit.upload4_begin if BEGIN else unit.upload4)(qh[:n], ph[:n]); t.append(time.time())   # no synchronize: it would wait for the momenta
        mdm.Calculate(); unit.synchronize(); t.append(time.time())
        obs = mdm.observables(); unit.synchronize(); t.append(time.time())
        for i in range(K - 1):
            mdm.Calculate()
        unit.synchronize(); t.append(time.time())
        unit.download(want_type=False, out=(qh, ph)); unit.synchronize(); t.append(time.time())
        d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
        tot = 1e3 * (t[-1] - t[0])
        print("rep %d N=%d K=%d: upload %.1f ms (%.1f GB/s) | first Calculate %.1f | observe %.1f | %d steps %.1f (%.3f/step) | "
              "download %.1f ms (%.1f GB/s) | total %.1f ms -> %.3f G updates/s" % (
                  rep, n, K, d[0], 64e-6 * n / d[0], d[1], d[2], K - 1, d[3], d[3] / max(K - 1, 1), d[4], 64e-6 * n / d[4], tot,
                  n * K / tot / 1e6), flush=True)
    mdm.close()


if __name__ == "__main__":
    main()
