"""GPU: mdb200_upload_begin (momenta still travelling while the first pair-list build runs on the positions) must be
indistinguishable from mdb200_upload: bit-identical states after the same steps, for a single unit (the late-momenta
path: upload rows carried through migration and cell sort, gathered afterwards) and for several units (which wait)."""
import numpy as np
import pytest
import torch

from common import cfg_for, load_golden, refdump
from mdacp_b200 import driver

pytestmark = pytest.mark.gpu


def _pinned(a):
    t = torch.zeros((a.shape[0], 4), dtype=torch.float64).pin_memory()
    t.numpy()[:, :a.shape[1]] = a
    return t


def _run(meta, states, nsteps, begin, monkeypatch=None):
    nunits = meta["nunits"]
    mdm = driver.MDManager(cfg_for(meta), num_procs=1, rank=0, num_threads=nunits, cutoff=meta["cutoff"])
    keep = []
    try:
        for u, (q, p) in enumerate(states):
            unit = mdm.GetMDUnit(u)
            tq, tp = _pinned(q), _pinned(p)
            keep.append((tq, tp))
            if begin:
                unit.upload4_begin(tq.numpy(), tp.numpy())
            else:
                unit.upload4(tq.numpy(), tp.numpy())
        mdm.Calculate(nsteps)
        out = [mdm.GetMDUnit(u).download() for u in range(nunits)]
        cnt = [mdm.GetMDUnit(u).counts() for u in range(nunits)]
        return out, cnt
    finally:
        mdm.close()


@pytest.mark.parametrize("name", ["bench1", "collision2", "bench8", "bench1_c25"])
def test_upload_begin_equals_upload(name):
    rec, cfg, steps, init_build = load_golden(name)
    meta = refdump.meta_of(rec)
    nunits = meta["nunits"]
    init = [(rec["q.s0.u%d" % u], rec["p.s0.u%d" % u]) for u in range(nunits)]
    # a state with particles outside their boxes (so that the first build migrates): 45 steps without looking
    warm, _ = _run(meta, init, 45, False)
    states = [(d["q"][:d["pn"]], d["p"]) for d in warm]
    a, ca = _run(meta, states, 30, False)
    b, cb = _run(meta, states, 30, True)
    assert ca == cb
    for x, y in zip(a, b):
        assert x["pn"] == y["pn"]
        assert np.array_equal(x["q"][:x["pn"]], y["q"][:y["pn"]])
        assert np.array_equal(x["p"], y["p"])
        assert np.array_equal(x["type"], y["type"])


def test_momenta_are_there_for_whoever_asks_first():
    """download / observe straight after mdb200_upload_begin (no build in between) see the uploaded momenta."""
    rec, cfg, steps, init_build = load_golden("bench1")
    meta = refdump.meta_of(rec)
    q, p = rec["q.s0.u0"], rec["p.s0.u0"]
    mdm = driver.MDManager(cfg_for(meta), num_procs=1, rank=0, num_threads=1, cutoff=meta["cutoff"])
    try:
        unit = mdm.GetMDUnit(0)
        tq, tp = _pinned(q), _pinned(p)
        unit.upload4_begin(tq.numpy(), tp.numpy())
        d = unit.download()
        assert np.array_equal(d["q"], q) and np.array_equal(d["p"], p)
        unit.upload4_begin(tq.numpy(), tp.numpy())
        k = mdm.observe_raw()[0]
        assert abs(k - 0.5 * (p ** 2).sum()) <= 1e-12 * k
        unit.upload4_begin(tq.numpy(), tp.numpy())
        unit.upload4_begin(tq.numpy(), tp.numpy())   # twice in a row
        mdm.Calculate(5)
        assert mdm.GetMDUnit(0).counts()["pn"] == q.shape[0]
    finally:
        mdm.close()
