"""Reader / runner for the oracle dump harness (oracle/ref_harness/dump_project.cc).

TEST INFRASTRUCTURE ONLY: imported by tests/, tools under oracle/ and bench.py's CPU legs.
The harness drives the UNMODIFIED reference (oracle/_ref/mdacp_scalar, built by oracle/Makefile
from /root/reference) through its public API and writes named binary records.
"""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SCALAR = os.path.join(HERE, "_ref", "mdacp_scalar")
REF_AVX2 = os.path.join(HERE, "_ref", "mdacp_avx2")
REF_AVX2_BIG = os.path.join(HERE, "_ref", "mdacp_avx2_big")

_DTYPES = {0: np.float64, 1: np.int32, 2: np.int64}


def have_ref(path=REF_SCALAR):
    return os.path.isfile(path) and os.access(path, os.X_OK)


def read_dump(path):
    """-> dict name -> ndarray"""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        name = data[off:off + 48].split(b"\0", 1)[0].decode()
        dtype, ndim = struct.unpack_from("<ii", data, off + 48)
        shape = struct.unpack_from("<4q", data, off + 56)
        off += 88
        shp = tuple(shape[:ndim])
        n = int(np.prod(shp)) if ndim else 1
        dt = np.dtype(_DTYPES[dtype])
        arr = np.frombuffer(data, dtype=dt, count=n, offset=off).reshape(shp).copy()
        off += n * dt.itemsize
        out[name] = arr
    return out


def run_reference(cfg, nunits=1, exe=REF_SCALAR, timeout=3600, workdir=None):
    """Run the reference with Mode=OracleDump.  cfg: dict of cfg keys (DumpFile is set here).
    Returns (records dict, stdout text)."""
    if not have_ref(exe):
        raise FileNotFoundError(exe)
    own = workdir is None
    tmp = tempfile.mkdtemp(prefix="orc_") if own else workdir
    dump = os.path.join(tmp, "dump.bin")
    cfgp = os.path.join(tmp, "dump.cfg")
    kv = dict(cfg)
    kv["Mode"] = "OracleDump"
    kv["DumpFile"] = dump
    with open(cfgp, "w") as f:
        for k, v in kv.items():
            f.write("%s=%s\n" % (k, v))
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(nunits)
    res = subprocess.run([exe, cfgp], cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=timeout, text=True)
    if res.returncode != 0 or not os.path.exists(dump):
        raise RuntimeError("reference run failed (%d):\n%s" % (res.returncode, res.stdout[-2000:]))
    rec = read_dump(dump)
    if own:
        for fn in os.listdir(tmp):
            os.remove(os.path.join(tmp, fn))
        os.rmdir(tmp)
    return rec, res.stdout


def meta_of(rec):
    m = rec["meta"]
    return {
        "nunits": int(m[0]), "grid": (int(m[1]), int(m[2]), int(m[3])),
        "L": (float(m[4]), float(m[5]), float(m[6])), "dt": float(m[7]), "cutoff": float(m[8]),
        "buffer": float(m[9]), "T0": float(m[10]), "gamma": float(m[11]), "tau": float(m[12]),
        "control_temperature": int(m[13]), "heatbath_type": int(m[14]), "check_list_length": int(m[15]),
    }


def pair_set(key, partner):
    """canonical sorted (key, partner) array, shape (n, 2)"""
    # one 64-bit sort key per pair (indices are non-negative 32-bit): much faster than a lexsort on 1e8 pairs
    k = (np.asarray(key, dtype=np.int64) << 32) | np.asarray(partner, dtype=np.int64)
    k.sort()
    return np.stack([k >> 32, k & 0xFFFFFFFF], axis=1)
