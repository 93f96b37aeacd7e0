"""Per-phase device timers of a multi-process run (torchrun, one unit per process): where a step's time goes
with the peer-to-peer ghost refresh (default) or NCCL only (MDB200_P2P=0)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mdacp_b200 import capi, driver, hostlib  # noqa: E402

UL = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
NST = int(sys.argv[2]) if len(sys.argv) > 2 else 100
world, rank, lrank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", lrank))
buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(buf, 0)
mdm = driver.MDManager({"UnitLength": repr(UL), "TimeStep": "0.001"}, num_procs=world, rank=rank, num_threads=1,
                       devices=[lrank], unique_id=bytes(buf.cpu().numpy().tobytes()))
s, e = mdm.GetRect(0)
q = hostlib.benchmark_config(s, e, 0.7, rank)
p = hostlib.initial_velocity(len(q), 1.0, rank)
mdm.upload([{"q": q, "p": p}])
mdm.MakePairList()
u = mdm.GetMDUnit(0)
mdm.Calculate(30)
for level in (0, 2):
    u.set_profiling(level)
    u.timers(reset=True)
    u.synchronize()
    dist.barrier()
    t0 = time.time()
    mdm.Calculate(NST)
    u.synchronize()
    wall = time.time() - t0
    tm = u.timers()
    if rank == 0:
        print("P2P=%s profiling %d: %.3f ms/step wall | force %.3f comm %.3f integrate %.3f check %.3f pair(total) %.3f ms" % (
            os.environ.get("MDB200_P2P", "1"), level, 1e3 * wall / NST, tm["force"] / NST, tm["comm"] / NST,
            tm["integrate"] / NST, tm["check"] / NST, tm["pair"]), flush=True)
mdm.close()
dist.destroy_process_group()
