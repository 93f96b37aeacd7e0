"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path -- rank <-> unit-grid mapping,
neighbour tables, sub-boxes and the unique-id broadcast plumbing bench.py uses -- plus the oracle run
as two units, checked rank against rank.  No GPU, no NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ul, ret):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mdacp_b200 import driver, hostlib
        param = driver.Parameter.from_dict({"UnitLength": repr(ul)})
        pinfo = driver.ParaInfo(world, 1, param)
        sinfo = driver.SimulationInfo(param, pinfo.GetGridSize())
        P = driver.make_params(sinfo, pinfo, rank)
        # unique id broadcast exactly as bench.py does it (fake id: no NCCL on CPU)
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf.copy_(torch.arange(128, dtype=torch.uint8))
        dist.broadcast(buf, 0)
        assert bytes(buf.numpy().tobytes()) == bytes(range(128))
        # gather every rank's neighbour table and sub-box
        mine = torch.tensor(list(P.neighbor_id) + list(P.neighbor_rank) + list(P.grid_pos), dtype=torch.int64)
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        tab = np.stack([t.numpy() for t in allt])
        opp = [1, 0, 3, 2, 5, 4]
        for r in range(world):
            for d in range(6):
                nb = tab[r, d]
                assert tab[r, 6 + d] == nb  # one unit per rank: rank == unit id
                assert tab[nb, opp[d]] == r, "neighbour tables are not symmetric"
        s, e = driver.unit_rect(sinfo, pinfo, rank)
        box = torch.tensor(s + e, dtype=torch.float64)
        boxes = [torch.zeros_like(box) for _ in range(world)]
        dist.all_gather(boxes, box)
        vol = sum(float((b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2])) for b in boxes)
        assert abs(vol - sinfo.L[0] * sinfo.L[1] * sinfo.L[2]) < 1e-9 * vol
        # particle counts per unit from the host generator sum to the global count
        q = hostlib.benchmark_config(s, e, 0.7, rank)
        n = torch.tensor([len(q)], dtype=torch.int64)
        dist.all_reduce(n)
        ret[rank] = (int(n.item()), tuple(pinfo.GetGridSize()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_host_logic(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 12.0, ret), nprocs=world, join=True)
    assert ret[0] == ret[1]
    n, grid = ret[0]
    assert grid == (2, 1, 1)
    assert n == 2 * int(12.0 ** 3 * 0.7)


def test_reference_arm_nonzero_ranks_exit_quietly():
    """bench.py --impl reference under torchrun: only rank 0 works, the others exit 0 without output"""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


class _FakeUnit:
    """stands in for capi.Unit: hands out a fixed snapshot (the device side is covered by tests/test_dump_files.py)"""

    def __init__(self, q4, p4, t):
        self.snap = {"pn": len(t), "q4": q4, "p4": p4, "type": t}

    def snapshot_begin(self):
        pass

    def snapshot_wait(self):
        return self.snap


def _state(seed, n):
    rng = np.random.default_rng(seed)
    q4, p4 = np.zeros((n, 4)), np.zeros((n, 4))
    q4[:, :3] = rng.uniform(0, 20, (n, 3))
    p4[:, :3] = rng.normal(size=(n, 3))
    return q4, p4, np.arange(n, dtype=np.int32) + 1000 * seed


def _writer_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mdacp_b200 import driver
        mdm = object.__new__(driver.MDManager)  # host logic only: no device behind it
        mdm.rank, mdm.num_procs, mdm.num_threads = rank, world, 2
        mdm.units = [_FakeUnit(*_state(2 * rank + k, 300 + 7 * rank + k)) for k in range(2)]
        mdm.barrier = dist.barrier
        mdm._snapshot_open, mdm._cdview_index, mdm._dump_index = False, 0, 0
        os.chdir(outdir)
        assert mdm.SaveAsCdviewSequential() == "conf0000.cd"   # mdmanager.cc:80-105
        assert mdm.SaveAsDumpFileSequential() == "conf0000.dump"  # mdmanager.cc:465-505
        assert mdm.SaveAsCdviewSequential() == "conf0001.cd"
    finally:
        dist.destroy_process_group()


def test_rank_serialised_writers(tmp_path):
    """SaveAsCdview / SaveAsDumpFile: rank 0 truncates, then the ranks append in rank order, each its units in
    order -- the file equals what one process holding all four units would have written."""
    sys.path.insert(0, ROOT)
    from mdacp_b200 import hostlib
    world = 2
    mp.spawn(_writer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cd, dump = str(tmp_path / "expect.cd"), str(tmp_path / "expect.dump")
    first = True
    for rank in range(world):
        for k in range(2):
            q4, p4, t = _state(2 * rank + k, 300 + 7 * rank + k)
            hostlib.save_cdview(cd, q4, p4, t, append=not first)
            hostlib.save_dump(dump, q4, p4, append=not first)
            first = False
    assert (tmp_path / "conf0000.cd").read_bytes() == open(cd, "rb").read()
    assert (tmp_path / "conf0001.cd").read_bytes() == open(cd, "rb").read()
    assert (tmp_path / "conf0000.dump").read_bytes() == open(dump, "rb").read()
