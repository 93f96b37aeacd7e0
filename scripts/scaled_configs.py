"""BASELINE.json configs[3] / [4] at scale on ONE GPU (not the bench): langevin.cfg with UnitLength=284
(16 M particles, Langevin thermostat) and collision.cfg with the box scaled x10 per axis (two droplets of
radius 150 in a 1200 x 600 x 600 box, ~10.6 M particles, the rest vacuum).  Prints ms/step (CUDA events on the
unit's stream), particle-updates/s and the observables, so that the non-NVE step variants and the sparse cell grid are
exercised at production size.

    python scripts/scaled_configs.py [langevin] [collision] [--steps 100] [--warmup 50] [--scale 10]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdacp_b200 import driver, hostlib  # noqa: E402


def timed(mdm, unit, warmup, steps):
    mdm.Calculate(warmup)
    unit.synchronize()
    c0 = unit.counts()
    unit.event_record(0)
    t0 = time.time()
    mdm.Calculate(steps)
    unit.event_record(1)
    ms = unit.event_elapsed_ms(0, 1)
    wall = time.time() - t0
    c1 = unit.counts()
    return ms / steps, wall, c1["nbuilds"] - c0["nbuilds"]


def run_langevin(args):
    cfg = {"UnitLength": repr(args.unit_length), "TimeStep": "0.005", "AimedTemperature": "0.9", "HeatbathGamma": "0.1",
           "HeatbathType": "Langevin", "ControlTemperature": "yes"}
    mdm = driver.MDManager(cfg)
    try:
        s, e = mdm.GetRect(0)
        t0 = time.time()
        q = hostlib.lattice_config(mdm.sinfo.L, s, e, density=0.7)
        p = hostlib.initial_velocity(len(q), 1.8, 0)
        n = len(q)
        print("langevin: N=%d generated in %.1f s" % (n, time.time() - t0), flush=True)
        mdm.upload([{"q": q, "p": p}])
        mdm.SetSeed(12345)
        mdm.MakePairList()
        unit = mdm.GetMDUnit(0)
        ms, wall, nb = timed(mdm, unit, args.warmup, args.steps)
        o = mdm.observables()
        print("langevin: %.3f ms/step (%d rebuilds in %d steps), %.3f G particle-updates/s; T=%.4f (aimed 0.9) P=%.4f E=%.4f" % (
            ms, nb, args.steps, n / ms / 1e6, o["T"], o["P"], o["E"]), flush=True)
    finally:
        mdm.close()


def run_collision(args):
    k = args.scale
    L = [120.0 * k, 60.0 * k, 60.0 * k]
    cfg = {"SystemSizeX": repr(L[0]), "SystemSizeY": repr(L[1]), "SystemSizeZ": repr(L[2]), "TimeStep": "0.002"}
    mdm = driver.MDManager(cfg)
    try:
        s, e = mdm.GetRect(0)
        t0 = time.time()
        q, p = hostlib.collision_config(L, s, e)
        n = len(q)
        print("collision: box %gx%gx%g N=%d generated in %.1f s" % (L[0], L[1], L[2], n, time.time() - t0), flush=True)
        mdm.upload([{"q": q, "p": p}])
        unit = mdm.GetMDUnit(0)
        ms, wall, nb = timed(mdm, unit, args.warmup, args.steps)  # the first step builds the list (collision.cc)
        o = mdm.observables()
        c = unit.counts()
        d = unit.download(want_q=False, want_type=False)
        ptot = d["p"].sum(axis=0)
        print("collision: %.3f ms/step (%d rebuilds in %d steps), %.3f G particle-updates/s; pairs/particle %.1f; "
              "T=%.4f E=%.4f; total momentum %s" % (ms, nb, args.steps, n / ms / 1e6, c["npairs"] / n, o["T"], o["E"],
                                                    np.array2string(ptot, precision=3)), flush=True)
        assert c["pn"] == n
    finally:
        mdm.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["langevin", "collision"])
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--scale", type=float, default=10.0)
    ap.add_argument("--unit-length", type=float, default=284.0)
    a = ap.parse_args()
    if "langevin" in a.which:
        run_langevin(a)
    if "collision" in a.which:
        run_collision(a)
