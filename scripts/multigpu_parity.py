"""Multi-GPU parity: one process per GPU (torchrun), one unit per process, NCCL halo exchange and
migration inside libmdacp_b200.so, checked against the CPU oracle run with the SAME unit grid
(SURVEY.md finding 3).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multigpu_parity.py [--unit-length 12] [--steps 100]

Every rank runs its own copy of the (cheap, small) oracle system and compares its own unit:
pair set exact, rebuild schedule / pn / tn exact, K,U,W <= 1e-9, final q,p <= 1e-10.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from common import TOL_OBS, TOL_STATE, refdump  # noqa: E402
from mdacp_b200 import capi, driver, hostlib  # noqa: E402
from oracle import pyoracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--unit-length", type=float, default=12.0)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--dt", type=float, default=0.004)
    ap.add_argument("--v0", type=float, default=2.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", lrank))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())

    cfg = {"UnitLength": repr(args.unit_length), "TimeStep": repr(args.dt), "InitialVelocity": repr(args.v0)}
    mdm = driver.MDManager(cfg, num_procs=world, rank=rank, num_threads=1, devices=[lrank], unique_id=uid)
    grid = tuple(mdm.pinfo.GetGridSize())
    L = mdm.sinfo.L
    osys = pyoracle.System(L, args.dt, grid=grid, capacity=200000)
    assert osys.nunits == world
    for u in range(world):
        lo, hi = osys.rect(u)
        q = hostlib.benchmark_config(lo, hi, 0.7, u)
        p = hostlib.initial_velocity(len(q), args.v0, u)
        t = np.arange(len(q), dtype=np.int32) + 1000000 * u
        osys.set_unit(u, q, p, t)
        if u == rank:
            s, e = mdm.GetRect(0)
            assert np.allclose(s, lo) and np.allclose(e, hi)
            mdm.upload([{"q": q, "p": p, "type": t}])
    unit = mdm.GetMDUnit(0)
    mdm.MakePairList()
    osys.make_pair_list()
    worst = 0.0
    last = 0
    for s in range(args.steps + 1):
        if s > 0:
            mdm.Calculate()
            osys.calculate(1)
        c, oc = unit.counts(), osys.counts(rank)
        assert c["nbuilds"] == oc["nbuilds"], "rank %d step %d: builds %d vs %d" % (rank, s, c["nbuilds"], oc["nbuilds"])
        assert c["pn"] == oc["pn"] and c["tn"] == oc["tn"], "rank %d step %d: pn/tn %s vs %s" % (rank, s, c, oc)
        if c["nbuilds"] != last:
            last = c["nbuilds"]
            k, p = unit.pairs()
            ko, po = osys.pairs(rank)
            assert np.array_equal(refdump.pair_set(k, p), refdump.pair_set(ko, po)), "rank %d step %d: pair set" % (rank, s)
        o, r = unit.observe(), osys.observe(rank)
        for i in range(3):
            worst = max(worst, abs(o[i] - r[i]) / max(abs(r[i]), abs(r[0])))
        assert worst <= TOL_OBS, "rank %d step %d: K/U/W differ by %.3e" % (rank, s, worst)
    # all-reduced observables through NCCL vs the oracle's global sums
    tot = mdm.observe_raw()
    ref = osys.observe(-1)
    for i in range(4):
        assert abs(tot[i] - ref[i]) <= TOL_OBS * max(abs(ref[i]), abs(ref[0])), ("global", i, tot[i], ref[i])
    d, st = unit.download(), osys.state(rank)
    assert np.array_equal(d["type"], st["type"]), "rank %d: reference index order differs" % rank
    assert np.max(np.abs(d["q"] - st["q"])) <= TOL_STATE * max(L)
    assert np.max(np.abs(d["p"] - st["p"])) <= TOL_STATE * max(1.0, float(np.abs(st["p"]).max()))
    print("rank %d/%d grid %s: OK  N_unit=%d builds=%d worst K/U/W rel err %.2e" % (
        rank, world, grid, c["pn"], c["nbuilds"], worst), flush=True)
    mdm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
