"""GPU parity against the UNMODIFIED reference run on the box's own CPU (oracle/_ref/mdacp_scalar,
built here by oracle/Makefile from /root/reference and shipped with the snapshot).

Identical initial states (dumped from the reference driver, uploaded through the C-ABI):
  * pair-index sets at every rebuild: bit-exact (reference indices);
  * rebuild step numbers, pn, tn per step: exact;
  * kick impulse per particle: <= 1e-10 * max(|F_i|, F_rms);
  * K, U, W at every step: <= 1e-9 relative; T, P, E likewise;
  * q, p, type after 100 steps in reference index order: <= 1e-10 / exact.
"""
import pytest

from common import check_cuda_against_records, refdump

pytestmark = pytest.mark.gpu

def _need_ref():
    if not refdump.have_ref():
        pytest.skip("oracle/_ref/mdacp_scalar not present (build it where /root/reference exists)")


def _run_case(cfg, nunits, steps):
    _need_ref()
    kv = dict(cfg)
    init_build = kv.get("DumpInit", "Benchmark") != "Collision"
    kv.update({"DumpSteps": steps, "DumpStateAt": "0,%d" % steps, "DumpPairsAt": "",
               "DumpForceAt": ("0,%d" % steps) if init_build else "", "DumpPairsAtBuild": "yes"})
    rec, _ = refdump.run_reference(kv, nunits=nunits)
    seed = int(kv.get("Seed", 1)) if kv.get("DumpInit") == "Langevin" else None
    return check_cuda_against_records(rec, steps, init_build, seed)


def test_benchmark_single_unit_100_steps():
    """benchmark.cfg-shaped, scaled down to 5,600 particles, 1 unit (periodic self-ghosts)."""
    _run_case({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.001, "InitialVelocity": 1.0}, 1, 100)


def test_benchmark_hot_single_unit_rebuilds():
    """hotter and longer so that several margin-based rebuilds and migrations happen."""
    _run_case({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.004, "InitialVelocity": 2.5}, 1, 150)


def test_benchmark_2x2x2_units_on_one_gpu():
    """8 units (2x2x2 domain decomposition) driven in one process on one device: the same
    exchange logic the multi-GPU path uses, against the reference with 8 OpenMP units."""
    _run_case({"DumpInit": "Benchmark", "SystemSize": 28, "TimeStep": 0.004, "InitialVelocity": 2.0}, 8, 100)


def test_slab_grid_2x1x1():
    _run_case({"DumpInit": "Benchmark", "UnitLength": 16, "TimeStep": 0.004, "InitialVelocity": 2.0,
               "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 60)


def test_langevin_zero_temperature_exact():
    """Langevin path with AimedTemperature = 0: noise amplitude 0, deterministic friction, and the
    reference's missing second half drift (SURVEY.md finding 5)."""
    _run_case({"DumpInit": "Langevin", "UnitLength": 16, "TimeStep": 0.005, "InitialVelocity": 1.8,
               "Density": 0.7, "AimedTemperature": 0.0, "HeatbathGamma": 0.1, "HeatbathType": "Langevin",
               "Seed": 12345}, 1, 100)


def test_collision_first_step_builds():
    """collision.cfg-shaped: vacuum + two droplets, no MakePairList before the loop (Q22)."""
    _run_case({"DumpInit": "Collision", "SystemSizeX": 60, "SystemSizeY": 30, "SystemSizeZ": 30,
               "TimeStep": 0.002}, 1, 100)


def test_collision_two_units_load_imbalance():
    # droplets 5 apart so that they touch, exchange ghosts and migrate across the unit boundary
    _run_case({"DumpInit": "Collision", "SystemSizeX": 40, "SystemSizeY": 30, "SystemSizeZ": 30,
               "TimeStep": 0.004, "LocalGridX": 2, "LocalGridY": 1, "LocalGridZ": 1}, 2, 400)


def test_nose_hoover():
    _run_case({"DumpInit": "Benchmark", "SystemSize": 20, "TimeStep": 0.002, "InitialVelocity": 1.5,
               "ControlTemperature": "yes", "HeatbathType": "NoseHoover", "AimedTemperature": 1.0,
               "HeatbathTau": 0.1}, 1, 100)
