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

import os


def _need_ref(exe=refdump.REF_SCALAR):
    if not refdump.have_ref(exe):
        pytest.skip("%s not present (build it where /root/reference exists)" % os.path.relpath(exe))


def _run_case(cfg, nunits, steps, exe=refdump.REF_SCALAR, pairs_at_build=True, observe_every=1):
    """pairs_at_build=False (the >= 1e6-particle cases): the reference dumps its pair arrays after the initial
    build and at the last step only -- the list of the last rebuild -- instead of at every rebuild."""
    _need_ref(exe)
    kv = dict(cfg)
    init_build = kv.get("DumpInit", "Benchmark") != "Collision"
    kv.update({"DumpSteps": steps, "DumpStateAt": "0,%d" % steps,
               "DumpPairsAt": "" if pairs_at_build else "0,%d" % steps,
               "DumpForceAt": ("0,%d" % steps) if init_build else "",
               "DumpPairsAtBuild": "yes" if pairs_at_build else "no"})
    rec, _ = refdump.run_reference(kv, nunits=nunits, exe=exe)
    seed = int(kv.get("Seed", 1)) if kv.get("DumpInit") == "Langevin" else None
    return check_cuda_against_records(rec, steps, init_build, seed, observe_every=observe_every)


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


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[0] and sizes the committed fixtures cannot hold
# ---------------------------------------------------------------------------------------------
SHIPPED = {"DumpInit": "Benchmark", "SystemSize": 50, "TimeStep": 0.001, "InitialVelocity": 1.0}


def test_benchmark_cfg_as_shipped():
    """benchmark.cfg verbatim (SystemSize 50 -> 87,500 particles, dt 0.001, v0 1.0, cutoff 3.0 as compiled,
    one unit): pair sets, rebuild schedule, impulses, K/U/W at every step and the final state over 100 steps."""
    w = _run_case(SHIPPED, 1, 100)
    print("benchmark.cfg as shipped:", w)


def test_benchmark_cfg_cutoff_2_5():
    """The same at cutoff 2.5 (what BASELINE.json quotes for configs[0]; mdconfig.h:15-16 is a compile-time
    constant, so the reference side is oracle/_ref/mdacp_scalar_c25, see oracle/Makefile)."""
    w = _run_case(SHIPPED, 1, 100, exe=os.path.join(os.path.dirname(refdump.REF_SCALAR), "mdacp_scalar_c25"))
    print("benchmark.cfg at cutoff 2.5:", w)


def test_cutoff_2_5_eight_units_hot():
    _run_case({"DumpInit": "Benchmark", "SystemSize": 24, "TimeStep": 0.004, "InitialVelocity": 2.0}, 8, 100,
              exe=os.path.join(os.path.dirname(refdump.REF_SCALAR), "mdacp_scalar_c25"))


def test_million_particles_eight_units():
    """1.09 M particles on 2x2x2 units (136 k each, 17 cells per axis and unit), hot enough for several
    margin-based rebuilds with migration in 60 steps: 64-bit list offsets, large cell grids and the
    rebuild path at a size no committed fixture reaches."""
    w = _run_case({"DumpInit": "Benchmark", "SystemSize": 116, "TimeStep": 0.004, "InitialVelocity": 2.0}, 8, 60,
                  pairs_at_build=False, observe_every=10)
    print("1.09 M particles, 8 units:", w)


def test_million_particles_one_unit():
    """1.33 M particles in ONE unit (the reference's N = 1e6 cap raised in oracle/_ref/mdacp_scalar_big):
    37 cells per axis, 75 M listed pairs."""
    w = _run_case({"DumpInit": "Benchmark", "SystemSize": 124, "TimeStep": 0.004, "InitialVelocity": 2.0}, 1, 40,
                  exe=os.path.join(os.path.dirname(refdump.REF_SCALAR), "mdacp_scalar_big"),
                  pairs_at_build=False, observe_every=10)
    print("1.33 M particles, 1 unit:", w)
