"""GPU: on-device initial states (mdacp_b200/csrc/generate.cu) against the reference's
own initial states (committed golden dumps) and the host generators that are bit-identical to them
(mdacp_b200/host/hostlib.cpp).  Positions, particle order and drift velocities: EXACT.  Random momenta: the same
std::mt19937 draws, cos / sin from CUDA's libm instead of glibc's -- a few ulp, asserted as <= 8 ulp of v0."""
import numpy as np
import pytest

from common import cfg_for, load_golden, refdump
from mdacp_b200 import driver, hostlib

pytestmark = pytest.mark.gpu
ULP = 2.0 ** -52


def _generate_on_device(mdm, cfg, mode):
    param = driver.Parameter.from_dict(cfg)   # the dump's cfg text: Density, InitialVelocity
    v0 = param.GetDoubleDef("InitialVelocity", 1.0)
    L = mdm.sinfo.L
    out = []
    for i in range(mdm.GetTotalThreads()):
        unit = mdm.GetMDUnit(i)
        if mode == "Benchmark":
            n = unit.generate_benchmark(param.GetDoubleDef("Density", 0.7))
            unit.generate_initial_velocity(v0, i)
        elif mode == "Langevin":
            n = unit.generate_lattice(param.GetDoubleDef("Density", 0.5))
            unit.generate_initial_velocity(v0, i)
        else:  # Collision::Run (collision.cc:54-61)
            n = unit.generate_ball([L[0] * 0.25, L[1] * 0.5, L[2] * 0.5], L[1] * 0.25, 1.0, 0.374)
            n += unit.generate_ball([L[0] * 0.75, L[1] * 0.5, L[2] * 0.5], L[1] * 0.25, -1.0, 0.374)
        d = unit.download()
        assert d["pn"] == n and d["tn"] == n
        out.append(d)
    return out


@pytest.mark.parametrize("name", ["bench1", "bench8", "slab2", "langevinT0", "collision2", "collision8", "bench1_c25"])
def test_device_generators_match_reference_initial_state(name):
    rec, cfg, steps, init_build = load_golden(name)
    meta = refdump.meta_of(rec)
    nunits = meta["nunits"]
    mode = cfg.get("DumpInit", "Benchmark")
    mdm = driver.MDManager(cfg_for(meta), num_procs=1, rank=0, num_threads=nunits, cutoff=meta["cutoff"])
    try:
        v0 = driver.Parameter.from_dict(cfg).GetDoubleDef("InitialVelocity", 1.0)
        for u, d in enumerate(_generate_on_device(mdm, cfg, mode)):
            qr, pr = rec["q.s0.u%d" % u], rec["p.s0.u%d" % u]
            assert d["q"].shape[0] == qr.shape[0], (name, u, d["q"].shape, qr.shape)
            assert np.array_equal(d["q"], qr), (name, u)
            assert np.all(d["type"] == 0)
            if mode == "Collision":
                assert np.array_equal(d["p"], pr)
            else:
                assert np.max(np.abs(d["p"] - pr)) <= 8 * ULP * v0, (name, u, np.max(np.abs(d["p"] - pr)))
    finally:
        mdm.close()


def test_device_generated_state_runs_like_the_uploaded_one():
    """The generated store is a normal one: 40 steps from it agree with 40 steps from the uploaded host state to the
    few-ulp difference of the initial momenta (same particle counts, same rebuild schedule, K,U,W to 1e-9)."""
    rec, cfg, steps, init_build = load_golden("bench8")
    meta = refdump.meta_of(rec)
    nunits = meta["nunits"]
    res = []
    for device_init in (False, True):
        mdm = driver.MDManager(cfg_for(meta), num_procs=1, rank=0, num_threads=nunits, cutoff=meta["cutoff"])
        try:
            if device_init:
                _generate_on_device(mdm, cfg, "Benchmark")
            else:
                mdm.upload([{"q": rec["q.s0.u%d" % u], "p": rec["p.s0.u%d" % u]} for u in range(nunits)])
            mdm.MakePairList()
            mdm.Calculate(40)
            o = mdm.observe_raw()
            c = [mdm.GetMDUnit(u).counts() for u in range(nunits)]
            res.append((np.array(o), [(x["pn"], x["tn"], x["nbuilds"], x["npairs"]) for x in c]))
        finally:
            mdm.close()
    assert res[0][1] == res[1][1]
    assert np.allclose(res[0][0], res[1][0], rtol=1e-9, atol=0.0)


def test_generators_refuse_a_stepped_store_and_clear_resets_it():
    from mdacp_b200 import capi
    mdm = driver.MDManager({"Mode": "Benchmark", "SystemSize": "12", "TimeStep": "0.001"}, cutoff=3.0)
    try:
        unit = mdm.GetMDUnit(0)
        n = unit.generate_benchmark(0.7)
        assert n == int(12 ** 3 * 0.7)
        unit.generate_initial_velocity(1.0, 0)
        mdm.Calculate(3)
        with pytest.raises(capi.MDB200Error):
            unit.generate_benchmark(0.7)
        unit.clear()
        assert unit.counts()["pn"] == 0
        assert unit.generate_benchmark(0.7) == n
    finally:
        mdm.close()


@pytest.mark.parametrize("L,nunits", [(50.0, 1), (60.0, 8)])
def test_device_generators_match_host_generators(L, nunits):
    """Larger than the fixtures (benchmark.cfg as shipped: 87,500 particles; 8 units of 27,000): every maker against
    its host twin, incl. the droplet maker of Burst mode."""
    cfg = {"Mode": "Benchmark", "SystemSize": repr(L), "TimeStep": "0.001"}
    mdm = driver.MDManager(cfg, num_procs=1, rank=0, num_threads=nunits, cutoff=3.0)
    try:
        Ls = mdm.sinfo.L
        for i in range(nunits):
            unit = mdm.GetMDUnit(i)
            s, e = mdm.GetRect(i)
            # benchmark maker + velocities
            n = unit.generate_benchmark(0.7)
            unit.generate_initial_velocity(1.3, i)
            d = unit.download()
            qh = hostlib.benchmark_config(s, e, 0.7, i)
            ph = hostlib.initial_velocity(len(qh), 1.3, i)
            assert n == len(qh) and np.array_equal(d["q"], qh[:, :3])
            assert np.max(np.abs(d["p"] - ph[:, :3])) <= 8 * ULP * 1.3
            # the momenta of a unit add up to (almost) nothing (variables.cc:92-107)
            assert np.max(np.abs(d["p"].sum(axis=0))) < 1e-9
            # global lattice
            unit.clear()
            n = unit.generate_lattice(0.5)
            qh = hostlib.lattice_config(Ls, s, e, 0.5)
            d = unit.download()
            assert n == len(qh) and np.array_equal(d["q"], qh[:, :3]) and not d["p"].any()
            # droplet
            unit.clear()
            n = unit.generate_lattice(0.5, radius=0.3 * L)
            qh = hostlib.droplet_config(Ls, s, e, 0.5, 0.3 * L)
            d = unit.download()
            assert n == len(qh) and np.array_equal(d["q"], qh[:, :3])
            # two balls appended one after the other
            unit.clear()
            n = unit.generate_ball([Ls[0] * 0.25, Ls[1] * 0.5, Ls[2] * 0.5], Ls[1] * 0.25, 1.0, 0.374)
            n += unit.generate_ball([Ls[0] * 0.75, Ls[1] * 0.5, Ls[2] * 0.5], Ls[1] * 0.25, -1.0, 0.374)
            qh, ph = hostlib.collision_config(Ls, s, e)
            d = unit.download()
            assert n == len(qh) and np.array_equal(d["q"], qh[:, :3]) and np.array_equal(d["p"], ph[:, :3])
            unit.clear()
    finally:
        mdm.close()


def test_device_generator_at_the_benchmark_size():
    """BASELINE.json configs[1]: 16,034,412 particles (MDB200_TEST_UNIT_LENGTH shrinks it): the on-device maker gives the
    host maker's positions bit for bit and its momenta to an ulp, and a step runs from the generated store."""
    import os
    UL = float(os.environ.get("MDB200_TEST_UNIT_LENGTH", "284"))
    mdm = driver.MDManager({"Mode": "Benchmark", "UnitLength": repr(UL), "TimeStep": "0.001"}, cutoff=3.0)
    try:
        unit = mdm.GetMDUnit(0)
        s, e = mdm.GetRect(0)
        n = unit.generate_benchmark(0.7)
        unit.generate_initial_velocity(1.0, 0)
        if UL == 284.0:
            assert n == 16034412
        d = unit.download(want_type=False)
        qh = hostlib.benchmark_config(s, e, 0.7, 0)
        assert n == len(qh) and np.array_equal(d["q"], qh[:, :3])
        del qh
        ph = hostlib.initial_velocity(n, 1.0, 0)
        assert np.max(np.abs(d["p"] - ph[:, :3])) <= 8 * ULP
        del ph, d
        mdm.Calculate(3)
        o = mdm.observables()
        assert o["N"] == n and abs(o["K"] - 0.5) < 1e-3   # |p| = 1 for every particle
    finally:
        mdm.close()
