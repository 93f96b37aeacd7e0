"""GPU: Langevin thermostat with thermal noise.  The reference draws its noise
from one sequential std::mt19937 per unit, which a per-particle counter RNG cannot reproduce bit-wise,
so parity is statistical here (and exact at AimedTemperature = 0: golden `langevinT0`):
  * one thermostat application on p = 0 gives N(0, 2 gamma T0 dt) per component (mean, variance,
    kurtosis, no repetition between steps);
  * a full run relaxes to the same <T> as the oracle's ensemble (same gamma, dt, half-drift quirk)."""
import numpy as np
import pytest

from mdacp_b200 import driver, hostlib
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def _setup(L, T0, gamma, dt, v0, density=0.7):
    cfg = {"SystemSize": repr(L), "TimeStep": repr(dt), "AimedTemperature": repr(T0), "HeatbathGamma": repr(gamma),
           "HeatbathType": "Langevin", "ControlTemperature": "yes"}
    mdm = driver.MDManager(cfg)
    s, e = mdm.GetRect(0)
    q = hostlib.lattice_config([L] * 3, s, e, density)
    p = hostlib.initial_velocity(len(q), v0, 0)
    return mdm, q, p


def test_noise_moments_single_application():
    T0, gamma, dt = 0.9, 0.1, 0.005
    mdm, q, p = _setup(24.0, T0, gamma, dt, 1.0)
    try:
        unit = mdm.GetMDUnit(0)
        mdm.upload([{"q": q, "p": np.zeros_like(p)}])
        mdm.SetSeed(12345)
        unit.langevin()
        a = unit.download()["p"].ravel().copy()
        var = 2.0 * gamma * T0 * dt  # (xi dt)^2 with <xi^2> = 2 gamma T0 / dt  (fcalculator.cc:771)
        n = a.size
        assert abs(a.mean()) < 5 * np.sqrt(var / n)
        assert abs(a.var() / var - 1.0) < 5 * np.sqrt(2.0 / n)
        kurt = ((a - a.mean()) ** 4).mean() / a.var() ** 2
        assert abs(kurt - 3.0) < 5 * np.sqrt(24.0 / n)
        # the x, y, z draws of a particle are independent
        b = a.reshape(-1, 3)
        c = np.corrcoef(b.T)
        assert abs(c[0, 1]) < 5 / np.sqrt(len(b)) and abs(c[0, 2]) < 5 / np.sqrt(len(b))
        # a different seed gives a different stream
        mdm.upload([{"q": q, "p": np.zeros_like(p)}])
        mdm.SetSeed(54321)
        unit.langevin()
        a2 = unit.download()["p"].ravel()
        assert abs(np.corrcoef(a, a2)[0, 1]) < 5 / np.sqrt(n)
    finally:
        mdm.close()


def test_equilibrium_temperature_matches_oracle_ensemble():
    T0, gamma, dt, L, v0, steps = 0.9, 1.0, 0.005, 12.0, 1.8, 1600
    mdm, q, p = _setup(L, T0, gamma, dt, v0)
    osys = pyoracle.System([L] * 3, dt, T0=T0, gamma=gamma, control_temperature=True, heatbath_type=1,
                           capacity=40000)
    try:
        mdm.upload([{"q": q, "p": p}])
        mdm.SetSeed(12345)
        osys.set_unit(0, q, p)
        osys.set_seed(12345)
        mdm.MakePairList()
        osys.make_pair_list()
        tg, to = [], []
        for s in range(steps):
            mdm.Calculate()
            osys.calculate(1)
            if s >= steps // 2 and s % 4 == 0:
                o = mdm.GetMDUnit(0).observe()
                tg.append(o[0] / o[3] / 1.5)
                r = osys.observe(0)
                to.append(r[0] / r[3] / 1.5)
        tg, to = np.array(tg), np.array(to)
        n = len(q)
        sigma = T0 * np.sqrt(2.0 / (3.0 * n))          # canonical fluctuation of T for n particles
        # 1/gamma = 200 steps of correlation time => ~4 independent blocks in the 800-step window
        tol = 4.0 * sigma / np.sqrt(4.0) + 0.01
        assert abs(tg.mean() - to.mean()) < tol, (tg.mean(), to.mean(), tol)
        assert abs(tg.mean() - T0) < 0.1 * T0, tg.mean()
        assert abs(tg.std() / to.std() - 1.0) < 0.5, (tg.std(), to.std())
    finally:
        mdm.close()
        osys.close()
