"""CPU: the pieces of bench.py that decide how the reference's CPU run uses the host."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_unit_grid_uses_the_cores_and_keeps_units_cubic():
    prev = 0
    for cores in range(1, 200):
        a, b, c = bench.unit_grid_for(cores)
        assert a <= b <= c <= 2 * a
        assert a * b * c <= cores
        assert a * b * c >= prev  # never fewer units with more cores
        prev = a * b * c
    assert bench.unit_grid_for(8) == (2, 2, 2)
    assert bench.unit_grid_for(16) == (2, 2, 4)
    assert bench.unit_grid_for(27) == (3, 3, 3)
    assert bench.unit_grid_for(64) == (4, 4, 4)


def test_clock_sampler_without_nvml_is_harmless():
    s = bench.ClockSampler(0)
    s.prepare()
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
