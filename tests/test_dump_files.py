"""Dump path: the host-side writers produce the reference's files byte for byte, and the
asynchronous device snapshot that feeds them returns exactly what a download would.

CPU part: tests/golden/files2.npz holds the files the UNMODIFIED reference wrote with its own
MDManager::SaveAsCdview / SaveAsDumpFile / SaveConfiguration (oracle/ref_harness, DumpSaveFiles=yes) next to the
state they were written from; our writers are fed that state."""
import os

import numpy as np
import pytest

from common import GOLDEN, cfg_for, load_golden, refdump, states_from_dump
from mdacp_b200 import hostlib


def _final_states(rec, steps):
    nu = refdump.meta_of(rec)["nunits"]
    return states_from_dump(rec, steps, nu), [rec["q.s%d.u%d" % (steps, u)] for u in range(nu)]


def test_writers_are_byte_identical_to_the_reference(tmp_path):
    rec, cfg, steps, _ = load_golden("files2")
    states, q_all = _final_states(rec, steps)
    cd, dump = str(tmp_path / "final.cd"), str(tmp_path / "final.dump")
    for u, st in enumerate(states):  # rank 0 truncates, the units append in order (mdmanager.cc:90-105)
        hostlib.save_cdview(cd, st["q"], st["p"], st["type"], append=u > 0)
        hostlib.save_dump(dump, st["q"], st["p"], append=u > 0)
    assert open(cd, "rb").read() == rec["file.final.cd"].tobytes()
    assert open(dump, "rb").read() == rec["file.final.dump"].tobytes()
    # SaveConfiguration writes locals and ghosts; the ghosts' momenta are scratch in the reference
    # (fcalculator.cc:273-275), so only the locals' lines are comparable
    for u, st in enumerate(states):
        name = str(tmp_path / ("conf%03d.cnf" % u))
        hostlib.save_configuration(name, st["q"], st["p"])
        ref_lines = rec["file.conf%03d.cnf" % u].tobytes().split(b"\n")
        mine = open(name, "rb").read().split(b"\n")
        pn = st["p"].shape[0]
        assert mine[:pn] == ref_lines[:pn]
        assert len(ref_lines) - 1 == q_all[u].shape[0]  # tn lines in the reference's file


def test_many_rows_keep_their_order(tmp_path):
    """the text writer formats chunks on several threads: order and content must not depend on that"""
    n = 200001
    rng = np.random.default_rng(5)
    q = rng.uniform(0, 100, (n, 4))
    p = rng.normal(size=(n, 4)) * 10.0 ** rng.integers(-8, 3, (n, 1))
    t = np.arange(n, dtype=np.int32)
    name = str(tmp_path / "big.cd")
    hostlib.save_cdview(name, q, p, t)
    lines = open(name).read().split("\n")
    assert len(lines) == n + 1 and lines[-1] == ""
    for i in (0, 1, 65535, 65536, 65537, 131072, n - 1):
        f = lines[i].split(" ")
        assert f[0] == "0" and int(f[1]) == i
        assert f[2:] == ["%g" % v for v in (q[i, 0], q[i, 1], q[i, 2], p[i, 0], p[i, 1], p[i, 2])]


def test_droplet_generator_known_answer():
    """burst.cfg-shaped droplet: the unmodified reference prints `# Number of Particles = 724` for
    SystemSize=30, Radius=6, Density=0.8 (tests/golden/projects/burst/stdout.txt)."""
    out = open(os.path.join(GOLDEN, "projects", "burst", "stdout.txt")).read()
    assert "# Number of Particles = 724" in out
    q = hostlib.droplet_config([30.0] * 3, [0.0] * 3, [30.0] * 3, density=0.8, radius=6.0)
    assert len(q) == 724
    r = np.sqrt(((q[:, :3] - 15.0) ** 2).sum(axis=1))
    assert r.max() < 6.0
    # split over two units: same particles, partitioned
    qa = hostlib.droplet_config([30.0] * 3, [0.0] * 3, [15.0, 30.0, 30.0], density=0.8, radius=6.0)
    qb = hostlib.droplet_config([30.0] * 3, [15.0, 0.0, 0.0], [30.0] * 3, density=0.8, radius=6.0)
    assert len(qa) + len(qb) == 724


def test_collision_initial_cdview_is_byte_identical(tmp_path):
    """collision.cc:54-62 writes conf0000.cd before the first step: generator + writer, no GPU involved"""
    L = [40.0, 30.0, 30.0]
    q, p = hostlib.collision_config(L, [0.0] * 3, L)
    name = str(tmp_path / "conf0000.cd")
    hostlib.save_cdview(name, q, p, np.zeros(len(q), dtype=np.int32))
    assert open(name, "rb").read() == open(os.path.join(GOLDEN, "projects", "collision", "conf0000.cd"), "rb").read()


@pytest.mark.gpu
def test_snapshot_equals_download_and_survives_later_steps(tmp_path, monkeypatch):
    from mdacp_b200 import driver
    rec, cfg, steps, _ = load_golden("files2")
    meta = refdump.meta_of(rec)
    nu = meta["nunits"]
    mdm = driver.MDManager(cfg_for(meta), num_threads=nu, cutoff=meta["cutoff"])
    try:
        mdm.upload(states_from_dump(rec, 0, nu))
        mdm.MakePairList()
        mdm.Calculate(steps)
        down = [mdm.GetMDUnit(u).download() for u in range(nu)]
        mdm.BeginSnapshot()
        mdm.Calculate(25)  # the step loop goes on (through a rebuild) while the copy is in flight
        monkeypatch.chdir(tmp_path)
        mdm.SaveAsCdview("final.cd")      # picks the open snapshot up
        snaps = [mdm.GetMDUnit(u).snapshot_wait() for u in range(nu)]
        for d, sn in zip(down, snaps):
            assert sn["pn"] == d["pn"]
            assert np.array_equal(sn["q4"][:, :3], d["q"][:d["pn"]])
            assert np.array_equal(sn["p4"][:, :3], d["p"])
            assert np.array_equal(sn["type"], d["type"])
            assert not sn["q4"][:, 3].any() and not sn["p4"][:, 3].any()
        # the file: same particles in the same order as the reference's, values to the 6 digits it prints
        ref = np.array([l.split() for l in rec["file.final.cd"].tobytes().decode().splitlines()], dtype=float)
        got = np.array([l.split() for l in open("final.cd").read().splitlines()], dtype=float)
        assert got.shape == ref.shape
        assert np.array_equal(got[:, :2], ref[:, :2])
        assert np.abs(got[:, 2:] - ref[:, 2:]).max() <= 2e-5 * max(1.0, np.abs(ref[:, 2:]).max())
        # a second, synchronous save after more steps differs (the snapshot was not a view of live state)
        mdm.SaveAsCdview("later.cd")
        assert open("later.cd").read() != open("final.cd").read()
        assert mdm.SaveAsCdviewSequential() == "conf0000.cd" and mdm.SaveAsDumpFileSequential() == "conf0000.dump"
        assert os.path.getsize("conf0000.dump") == 48 * sum(d["pn"] for d in down)
    finally:
        mdm.close()
