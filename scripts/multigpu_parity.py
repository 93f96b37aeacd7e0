"""Multi-GPU parity: one process per GPU (torchrun), one unit per process, checked against the CPU oracle run with
the SAME unit grid (tests/multigpu_parity_lib.py has the comparison).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multigpu_parity.py [--unit-length 14] [--steps 100]

Environment switches of the transport are honoured (MDB200_P2P=0: NCCL send/recv instead of the peer-to-peer
windows; MDB200_OVERLAP=0/1: in-line / overlapped exchange).  Exit code 1 if any rank fails.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mdacp_b200 import capi  # noqa: E402
import multigpu_parity_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--unit-length", type=float, default=14.0)
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
    res = multigpu_parity_lib.run_parity(world, rank, lrank, uid, args.unit_length, args.steps, args.dt, args.v0)
    sys.stdout.write("rank %d/%d: %s\n" % (rank, world, json.dumps(res)))   # one write: lines of two ranks do not mix
    sys.stdout.flush()
    bad = torch.tensor([0 if res["ok"] else 1], device="cuda")
    dist.all_reduce(bad)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(bad.item()) else 0)


if __name__ == "__main__":
    main()
