"""GPU: edge cases of the hot path through the C-ABI -- empty and nearly empty units, a pair exactly at the
analytic LJ value, particles that sit exactly on a face, and the reference's abort conditions, which the
library reports as error codes instead of exit(1):
"Too fast particles" (pairlist.cc:59-63) and "Invalid position" (variables.cc:70-77)."""
import numpy as np
import pytest

from mdacp_b200 import capi, driver

pytestmark = pytest.mark.gpu
L = 12.0


def _mdm(**extra):
    cfg = {"SystemSize": repr(L), "TimeStep": "0.001"}
    cfg.update(extra)
    return driver.MDManager(cfg)


def test_empty_unit_steps_and_observes():
    mdm = _mdm()
    try:
        mdm.upload([{"q": np.zeros((0, 3)), "p": np.zeros((0, 3))}])
        mdm.MakePairList()
        mdm.Calculate(5)
        u = mdm.GetMDUnit(0)
        c = u.counts()
        assert c["pn"] == 0 and c["tn"] == 0 and c["npairs"] == 0
        o = u.observe()
        assert o[0] == 0.0 and o[1] == 0.0 and o[2] == 0.0 and o[3] == 0.0
        d = u.download()
        assert d["pn"] == 0
    finally:
        mdm.close()


def test_single_particle_moves_freely():
    mdm = _mdm()
    try:
        q = np.array([[6.0, 6.0, 6.0]])
        p = np.array([[0.5, -0.25, 0.125]])
        mdm.upload([{"q": q, "p": p}])
        mdm.MakePairList()
        mdm.Calculate(10)
        u = mdm.GetMDUnit(0)
        assert u.counts()["npairs"] == 0
        d = u.download()
        np.testing.assert_allclose(d["p"][0], p[0], rtol=0, atol=0)
        np.testing.assert_allclose(d["q"][0], q[0] + 10 * 0.001 * p[0], rtol=0, atol=1e-14)
        o = u.observe()
        assert abs(o[0] - 0.5 * (p ** 2).sum()) < 1e-15 and o[1] == 0.0
    finally:
        mdm.close()


def test_two_particles_match_the_analytic_pair_force_and_energy():
    mdm = _mdm()
    try:
        r = 1.3
        q = np.array([[5.0, 6.0, 6.0], [5.0 + r, 6.0, 6.0]])
        mdm.upload([{"q": q, "p": np.zeros((2, 3))}])
        mdm.MakePairList()
        u = mdm.GetMDUnit(0)
        assert u.counts()["npairs"] == 1
        f = u.impulse() / 0.001  # impulse = F dt
        rc = 3.0
        c2 = 6.0 * rc ** -14 - 3.0 * rc ** -8          # variables.cc:21-27
        c0 = -rc ** -12 + rc ** -6 - c2 * rc ** 2
        # fcalculator.cc:213-221: df = (24 r^6 - 48) / r^14 + 8 C2, force on i along d = q_j - q_i
        df = (24.0 * r ** 6 - 48.0) / r ** 14 + 8.0 * c2
        np.testing.assert_allclose(f[0], [df * r, 0.0, 0.0], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(f[1], [-df * r, 0.0, 0.0], rtol=1e-12, atol=1e-13)
        o = u.observe()
        upair = 4.0 * (r ** -12 - r ** -6 + c2 * r * r + c0)  # observer.cc:18-48
        assert abs(o[1] - upair) <= 1e-12 * abs(upair)
        w = df * r * r / 3.0                                   # observer.cc:51-79
        assert abs(o[2] - w) <= 1e-12 * abs(w)
    finally:
        mdm.close()


def test_pair_across_the_periodic_boundary_and_particle_on_a_face():
    mdm = _mdm()
    try:
        # one particle exactly ON the low x face (not ghosted low-side: the predicate is strict, mdrect.cc:24-47),
        # its partner just inside the high x face: the pair exists only through the partner's ghost image
        q = np.array([[0.0, 6.0, 6.0], [L - 1.2, 6.0, 6.0]])
        mdm.upload([{"q": q, "p": np.zeros((2, 3))}])
        mdm.MakePairList()
        u = mdm.GetMDUnit(0)
        c = u.counts()
        assert c["pn"] == 2
        # particle 1 is ghosted to the low side (x - L = -1.2) and pairs with particle 0; particle 0 is not ghosted
        assert c["tn"] == 3 and c["npairs"] == 1
        f = u.impulse()
        assert f[0, 0] != 0.0 and f[1, 0] == 0.0  # the reference's one-sided pair
    finally:
        mdm.close()


def test_too_fast_particle_is_reported():
    mdm = _mdm()
    try:
        q = np.array([[3.0, 3.0, 3.0], [8.0, 8.0, 8.0]])
        p = np.array([[400.0, 0.0, 0.0], [0.0, 0.0, 0.0]])  # 2 |p| dt = 0.8 > skin 0.3
        mdm.upload([{"q": q, "p": p}])
        mdm.MakePairList()
        with pytest.raises(capi.MDB200Error) as ei:
            mdm.Calculate(1)
        assert ei.value.code == capi.ETOOFAST
    finally:
        mdm.close()


def test_invalid_position_is_reported():
    mdm = _mdm()
    try:
        q = np.array([[3.0, 3.0, 3.0], [3.0 * L, 8.0, 8.0]])  # still outside after one periodic wrap
        mdm.upload([{"q": q, "p": np.zeros((2, 3))}])
        with pytest.raises(capi.MDB200Error) as ei:
            mdm.MakePairList()
        assert ei.value.code == capi.EPOSITION
    finally:
        mdm.close()
