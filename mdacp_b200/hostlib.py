"""ctypes binding of libmdacp_host.so: the host-side initial-state generators (C++, same libstdc++
facilities as the reference driver; see mdacp_b200/host/hostlib.cpp)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmdacp_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libmdacp_host.so is not built (%s); run __graft_entry__.build()" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        L.mdh_benchmark_config.restype = C.c_long
        L.mdh_benchmark_config.argtypes = [dp, dp, C.c_double, C.c_int, vp, C.c_long]
        L.mdh_lattice_config.restype = C.c_long
        L.mdh_lattice_config.argtypes = [dp, C.c_double, dp, dp, vp, C.c_long]
        L.mdh_collision_config.restype = C.c_long
        L.mdh_collision_config.argtypes = [dp, dp, dp, vp, vp, C.c_long]
        L.mdh_initial_velocity.restype = None
        L.mdh_initial_velocity.argtypes = [vp, C.c_long, C.c_double, C.c_int]
        L.mdh_droplet_config.restype = C.c_long
        L.mdh_droplet_config.argtypes = [dp, C.c_double, C.c_double, dp, dp, vp, C.c_long]
        L.mdh_save_cdview.restype = C.c_int
        L.mdh_save_cdview.argtypes = [C.c_char_p, C.c_int, vp, vp, vp, C.c_long]
        L.mdh_save_dump.restype = C.c_int
        L.mdh_save_dump.argtypes = [C.c_char_p, C.c_int, vp, vp, C.c_long]
        L.mdh_save_configuration.restype = C.c_int
        L.mdh_save_configuration.argtypes = [C.c_char_p, vp, vp, C.c_long]
        _lib = L
    return _lib


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def benchmark_config(s, e, density=0.7, unit_id=0, out=None):
    """ExtractConfigurationMaker (confmaker.cc:174-252) for the unit box [s,e): -> q (n,4)."""
    w = [e[d] - s[d] for d in range(3)]
    cap = int(w[0] * w[1] * w[2] * density) + 64
    q = np.zeros((cap, 4)) if out is None else out
    n = lib().mdh_benchmark_config(_d3(s), _d3(e), float(density), int(unit_id), _vp(q), q.shape[0])
    if n < 0:
        raise MemoryError("benchmark_config: output too small")
    return q[:n]


def lattice_config(L, s, e, density=0.5):
    """ConfigurationMaker (confmaker.cc:9-44): global FCC lattice filtered to the unit box."""
    cap = int((e[0] - s[0]) * (e[1] - s[1]) * (e[2] - s[2]) * density * 1.3) + 1024
    q = np.zeros((cap, 4))
    n = lib().mdh_lattice_config(_d3(L), float(density), _d3(s), _d3(e), _vp(q), cap)
    if n < 0:
        raise MemoryError("lattice_config: output too small")
    return q[:n]


def collision_config(L, s, e):
    """The two droplets of Collision::Run (collision.cc:54-61) filtered to the unit box."""
    cap = int((e[0] - s[0]) * (e[1] - s[1]) * (e[2] - s[2]) * 0.374 * 1.3) + 1024
    q = np.zeros((cap, 4))
    p = np.zeros((cap, 4))
    n = lib().mdh_collision_config(_d3(L), _d3(s), _d3(e), _vp(q), _vp(p), cap)
    if n < 0:
        raise MemoryError("collision_config: output too small")
    return q[:n], p[:n]


def initial_velocity(n, V0, seed, out=None):
    """Variables::SetInitialVelocity(V0, seed = unit id) (variables.cc:82-108): -> p (n,4)."""
    p = np.zeros((n, 4)) if out is None else out
    lib().mdh_initial_velocity(_vp(p), int(n), float(V0), int(seed))
    return p[:n]


def droplet_config(L, s, e, density=0.5, radius=10.0):
    """DropletMaker (burst.cc:10-64): global FCC lattice sites inside the sphere around the box centre."""
    vol = (e[0] - s[0]) * (e[1] - s[1]) * (e[2] - s[2])
    cap = int(min(vol, 4.2 * radius ** 3 + 64) * density * 1.3) + 1024
    q = np.zeros((cap, 4))
    n = lib().mdh_droplet_config(_d3(L), float(density), float(radius), _d3(s), _d3(e), _vp(q), cap)
    if n < 0:
        raise MemoryError("droplet_config: output too small")
    return q[:n]


def _rows4(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 4 or not a.flags["C_CONTIGUOUS"]:
        b = np.zeros((a.shape[0], 4))
        b[:, :a.shape[1]] = a
        a = b
    return a


def save_cdview(filename, q, p, type_=None, append=False):
    """MDUnit::SaveAsCdview (mdunit.cc:52-67) for one unit's locals; append=False truncates first
    (mdmanager.cc:90-94).  q, p: (n,3) or (n,4) rows in reference index order."""
    q4, p4 = _rows4(q), _rows4(p)
    n = min(q4.shape[0], p4.shape[0])
    t = None if type_ is None else np.ascontiguousarray(type_, dtype=np.int32)
    rc = lib().mdh_save_cdview(os.fsencode(filename), 1 if append else 0, _vp(q4), _vp(p4),
                               None if t is None else _vp(t), n)
    if rc != 0:
        raise IOError("cannot write " + str(filename))


def save_dump(filename, q, p, append=False):
    """FileDumper::Execute (mdmanager.cc:487-500): raw doubles qx qy qz px py pz per local."""
    q4, p4 = _rows4(q), _rows4(p)
    n = min(q4.shape[0], p4.shape[0])
    if lib().mdh_save_dump(os.fsencode(filename), 1 if append else 0, _vp(q4), _vp(p4), n) != 0:
        raise IOError("cannot write " + str(filename))


def save_configuration(filename, q, p):
    """Variables::SaveConfiguration (variables.cc:45-57): text rows "qx qy qz px py pz"."""
    q4, p4 = _rows4(q), _rows4(p)
    n = min(q4.shape[0], p4.shape[0])
    if lib().mdh_save_configuration(os.fsencode(filename), _vp(q4), _vp(p4), n) != 0:
        raise IOError("cannot write " + str(filename))
