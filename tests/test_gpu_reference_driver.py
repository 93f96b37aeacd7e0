"""GPU: the REFERENCE'S OWN DRIVER bound to libmdacp_b200.so (integration/_build/mdacp_b200_ref).

integration/Makefile compiles main.cc, the projects (benchmark.cc, langevin.cc, collision.cc, burst.cc), Parameter,
ParaInfo, SimulationInfo, the configuration makers and all printing from /root/reference unchanged; only mdunit.h,
mdunit.cc, mdmanager.h and mdmanager.cc of a scratch copy receive the forwards of INTEGRATION.md
(integration/apply_binding.py).  These tests run that binary on the GPU and hold its output to the unmodified
reference's: committed expected output where the reference is absent, the live oracle/_ref/mdacp_scalar where it
travelled with the snapshot."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "integration", "_build", "mdacp_b200_ref")
REF = os.path.join(ROOT, "oracle", "_ref", "mdacp_scalar")
GOLDEN_PROJECTS = os.path.join(ROOT, "tests", "golden", "projects")

BENCHMARK_CFG = """Mode=Benchmark
ThermalizeLoop=150
SystemSize=50
TimeStep=0.001
TotalLoop=201
Restart=yes
AimedTemperature=1.5
ControlTemperature=no
InitialVelocity=1.0
"""
# stdout of the unmodified scalar reference for BENCHMARK_CFG (benchmark.cfg as shipped with TotalLoop cut to 201)
BENCHMARK_EXPECTED = ["0.151 0.219084 -3.4862 -4.43809 #observe", "0.251 0.227322 -3.0562 -4.43804 #observe",
                      "0.351 0.230586 -2.92141 -4.43804 #observe"]


def _run(exe, cfg_text, cwd, threads=1):
    (cwd / "input.cfg").write_text(cfg_text)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    res = subprocess.run([exe, "input.cfg"], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    return res.stdout


def _need_exe():
    if not os.path.exists(EXE):
        pytest.skip("integration/_build/mdacp_b200_ref not built (needs /root/reference at build time)")


def _stable(lines):
    """the lines that do not carry wall-clock times"""
    out = []
    for l in lines:
        s = l.strip()
        if "[SEC]" in s or "[MUPS]" in s or s.startswith("# all") or s.startswith("# force") or s.startswith("# comm") \
                or s.startswith("# pair") or "sec]" in s.lower():
            continue
        out.append(s)
    return out


def _numbers_close(a, b, rtol):
    """same tokens; numeric tokens equal to rtol (the reference prints 6 significant digits)"""
    ta, tb = a.split(), b.split()
    if len(ta) != len(tb):
        return False
    for x, y in zip(ta, tb):
        if x == y:
            continue
        try:
            fx, fy = float(x), float(y)
        except ValueError:
            return False
        if abs(fx - fy) > rtol * max(abs(fx), abs(fy), 1e-300):
            return False
    return True


def _compare_with_live_reference(cfg_text, tmp_path, threads=1, rtol=2e-5):
    if not os.path.exists(REF):
        return None
    (tmp_path / "ref").mkdir()
    (tmp_path / "gpu").mkdir()
    ref = _stable(_run(REF, cfg_text, tmp_path / "ref", threads).splitlines())
    got = _stable(_run(EXE, cfg_text, tmp_path / "gpu", threads).splitlines())
    assert len(got) == len(ref), "\n".join(got[-20:]) + "\n---\n" + "\n".join(ref[-20:])
    for g, r in zip(got, ref):
        assert _numbers_close(g, r, rtol), (g, r)
    return got


def test_benchmark_cfg_through_the_reference_driver(tmp_path):
    """benchmark.cfg as shipped (87,500 particles, cutoff 3.0) through main.cc + benchmark.cc of the reference"""
    _need_exe()
    out = _run(EXE, BENCHMARK_CFG, tmp_path).splitlines()
    lines = [l.strip() for l in out]
    assert "# Number of Particles = 87500" in lines
    assert "# Cutoff Length  = 3" in lines
    assert [l for l in lines if l.endswith("#observe")] == BENCHMARK_EXPECTED, "\n".join(lines)
    mups = [l for l in lines if l.startswith("# N = 87500 ") and l.endswith("[MUPS]")]
    assert mups, "\n".join(lines)
    print(mups[-1])


def test_benchmark_eight_units_matches_the_live_reference(tmp_path):
    """OMP_NUM_THREADS=8: eight MDUnits (2x2x2) on one device, every printed line against the unmodified reference"""
    _need_exe()
    cfg = "Mode=Benchmark\nThermalizeLoop=50\nSystemSize=32\nTimeStep=0.004\nTotalLoop=101\nObserveLoop=20\nInitialVelocity=2.0\n"
    got = _compare_with_live_reference(cfg, tmp_path, threads=8)
    if got is None:
        pytest.skip("oracle/_ref/mdacp_scalar not present")
    assert len([l for l in got if l.endswith("#observe")]) == 6


def test_langevin_zero_temperature_matches_the_live_reference(tmp_path):
    """langevin.cc with AimedTemperature=0 (noise amplitude 0: the thermostat is deterministic friction)"""
    _need_exe()
    cfg = ("Mode=Langevin\nInitialVelocity=1.8\nTimeStep=0.005\nUnitLength=16\nObserveLoop=50\nTotalLoop=201\n"
           "AimedTemperature=0.0\nDensity=0.70\nHeatbathTau=0.1\nHeatbathGamma=0.1\nHeatbathType=Langevin\nSeed=12345\n")
    got = _compare_with_live_reference(cfg, tmp_path)
    if got is None:
        pytest.skip("oracle/_ref/mdacp_scalar not present")


def _cd_rows(path):
    return np.array([l.split() for l in open(path).read().splitlines()], dtype=float)


@pytest.mark.parametrize("name", ["burst", "collision"])
def test_project_run_matches_the_reference(name, tmp_path):
    """Mode=Burst (droplet, Nose-Hoover thermalisation switched off mid-run, burst.cc:104-143) and Mode=Collision
    with SaveCdviewFile=yes (collision.cc:54-80) through the reference's own project code and writers: printed
    lines and cdview files against tests/golden/projects/ (written by the unmodified reference)."""
    _need_exe()
    gold = os.path.join(GOLDEN_PROJECTS, name)
    out = _run(EXE, open(os.path.join(gold, "input.cfg")).read(), tmp_path)
    ref_out = open(os.path.join(gold, "stdout.txt")).read().splitlines()
    want = [l.strip() for l in ref_out if l.endswith("# observe")]
    got = [l.strip() for l in out.splitlines() if l.endswith("# observe")]
    assert len(got) == len(want), out
    for g, w in zip(got, want):
        assert _numbers_close(g, w, 2e-5), (g, w)
    for key in ("# Number of Particles", "# System Size", "# Density", "# TimeStep", "# ControlTemperature"):
        assert [l for l in ref_out if l.startswith(key)] == [l for l in out.splitlines() if l.startswith(key)]
    for fn in sorted(f for f in os.listdir(gold) if f.endswith(".cd")):
        assert (tmp_path / fn).exists(), fn
        if name == "collision" and fn == "conf0000.cd":
            assert (tmp_path / fn).read_bytes() == open(os.path.join(gold, fn), "rb").read()
            continue
        ref, mine = _cd_rows(os.path.join(gold, fn)), _cd_rows(str(tmp_path / fn))
        assert mine.shape == ref.shape, fn
        assert np.array_equal(mine[:, :2], ref[:, :2])
        assert np.abs(mine[:, 2:] - ref[:, 2:]).max() <= 2e-5 * max(1.0, np.abs(ref[:, 2:]).max()), fn
