"""Host-side generators (mdacp_b200/host/hostlib.cpp) against the reference's initial states:
bit-identical q, p and particle order, from the golden fixtures (and live dumps where available)."""
import numpy as np
import pytest

from common import load_golden, refdump
from mdacp_b200 import driver, hostlib


def _generate(cfg, nunits):
    param = driver.Parameter.from_dict(cfg)
    pinfo = driver.ParaInfo(1, nunits, param)
    sinfo = driver.SimulationInfo(param, pinfo.GetGridSize())
    mode = cfg.get("DumpInit", "Benchmark")
    v0 = param.GetDoubleDef("InitialVelocity", 1.0)
    out = []
    for u in range(nunits):
        s, e = driver.unit_rect(sinfo, pinfo, u)
        if mode == "Benchmark":
            q = hostlib.benchmark_config(s, e, param.GetDoubleDef("Density", 0.7), u)
            p = hostlib.initial_velocity(len(q), v0, u)
        elif mode == "Langevin":
            q = hostlib.lattice_config(sinfo.L, s, e, param.GetDoubleDef("Density", 0.5))
            p = hostlib.initial_velocity(len(q), v0, u)
        else:
            q, p = hostlib.collision_config(sinfo.L, s, e)
        out.append((q, p))
    return out


@pytest.mark.parametrize("name", ["bench1", "bench8", "slab2", "langevinT0", "collision2"])
def test_generators_match_reference_initial_state(name):
    rec, cfg, steps, init_build = load_golden(name)
    nunits = refdump.meta_of(rec)["nunits"]
    for u, (q, p) in enumerate(_generate(cfg, nunits)):
        qr, pr = rec["q.s0.u%d" % u], rec["p.s0.u%d" % u]
        assert q.shape[0] == qr.shape[0], (name, u, q.shape, qr.shape)
        assert np.array_equal(q[:, :3], qr)
        assert np.array_equal(p[:, :3], pr)


def test_benchmark_cfg_as_shipped_counts():
    """benchmark.cfg verbatim (SystemSize=50, density 0.7): N = 87,500 with one unit."""
    q = hostlib.benchmark_config([0, 0, 0], [50, 50, 50], 0.7, 0)
    assert q.shape[0] == 87500
