"""GPU, several processes: the transport between processes (comm.cu) against the oracle.

One process per GPU (torchrun), one unit per process.  Covers the CUDA-IPC ghost windows with their arrive / ack
mailboxes (default), the NCCL send/recv fallback (MDB200_P2P=0), the overlapped exchange on its own stream
(MDB200_OVERLAP=1) and the in-line one (=0), migration between processes and the re-export of the windows when a
rebuild re-allocates the position buffers (small initial capacity forced through MDB200_DEBUG_SMALL_CAPACITY).
Needs >= 2 GPUs (NCCL refuses two ranks on one device): skipped on a one-GPU box, where bench.py's `parity_n`
block (SCALE runs) is the evidence instead."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from mdacp_b200 import capi
    return capi.device_count()


def _torchrun(n, env_extra, port, args=()):
    env = dict(os.environ)
    env.update(env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "multigpu_parity.py")] + list(args)
    res = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    # (two ranks may print on one line: count the reports, not the lines)
    assert res.stdout.count('"ok": true') == n and '"ok": false' not in res.stdout, res.stdout[-4000:]
    return res.stdout


CASES = [
    ("p2p_inline", {"MDB200_P2P": "1", "MDB200_OVERLAP": "0"}),
    ("p2p_overlap", {"MDB200_P2P": "1", "MDB200_OVERLAP": "1"}),
    ("nccl_inline", {"MDB200_P2P": "0", "MDB200_OVERLAP": "0"}),
    ("nccl_overlap", {"MDB200_P2P": "0", "MDB200_OVERLAP": "1"}),
]


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("name,env", CASES)
def test_units_in_separate_processes_match_the_oracle(n, name, env):
    if _gpus() < n:
        pytest.skip("needs %d GPUs" % n)
    _torchrun(n, env, 29600 + n * 10 + [c[0] for c in CASES].index(name), ["--unit-length", "14", "--steps", "100"])


def test_two_processes_long_hot_run_with_reallocation():
    """300 hot steps on 2 GPUs from a deliberately small initial capacity: rebuilds re-allocate q / q_alt, the
    receivers re-export their windows, replaced allocations wait in the graveyard until the senders unmapped them."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _torchrun(2, {"MDB200_DEBUG_SMALL_CAPACITY": "1"}, 29690, ["--unit-length", "12", "--steps", "300", "--v0", "2.5"])
