"""Shared helpers for the parity tests: run the oracle, feed the same state to the C-ABI."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refdump  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def cfg_for(rec_meta, extra=None):
    """cfg dict for mdacp_b200.driver.MDManager that reproduces the dump's SimulationInfo."""
    m = rec_meta
    cfg = {
        "SystemSizeX": repr(m["L"][0]), "SystemSizeY": repr(m["L"][1]), "SystemSizeZ": repr(m["L"][2]),
        "TimeStep": repr(m["dt"]), "AimedTemperature": repr(m["T0"]), "HeatbathGamma": repr(m["gamma"]),
        "HeatbathTau": repr(m["tau"]), "ControlTemperature": "yes" if m["control_temperature"] else "no",
        "HeatbathType": "Langevin" if m["heatbath_type"] == 1 else "NoseHoover",
        "CheckListLength": str(m["check_list_length"]),
        "LocalGridX": str(m["grid"][0]), "LocalGridY": str(m["grid"][1]), "LocalGridZ": str(m["grid"][2]),
    }
    if extra:
        cfg.update(extra)
    return cfg


def states_from_dump(rec, step, nunits):
    out = []
    for u in range(nunits):
        q = rec["q.s%d.u%d" % (step, u)]
        p = rec["p.s%d.u%d" % (step, u)]
        t = rec["type.s%d.u%d" % (step, u)]
        out.append({"q": q[:p.shape[0]], "p": p, "type": t})
    return out


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor) if floor else np.abs(b)
    return float(np.max(np.abs(a - b) / np.where(den == 0, 1.0, den))) if a.size else 0.0


def force_err(f, fref):
    """|dF| <= tol * max(|F_i|, F_rms): per-particle error normalised as SURVEY.md section 8a defines."""
    f = np.asarray(f)
    fref = np.asarray(fref)
    mag = np.sqrt((fref ** 2).sum(axis=1))
    rms = float(np.sqrt((mag ** 2).mean())) if mag.size else 1.0
    den = np.maximum(mag, rms)
    err = np.sqrt(((f - fref) ** 2).sum(axis=1)) / den
    return float(err.max()) if err.size else 0.0
