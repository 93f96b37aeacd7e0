"""ctypes binding of oracle/liboracle.so (the plain-C restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class Params(C.Structure):
    _fields_ = [
        ("L", C.c_double * 3), ("time_step", C.c_double), ("cutoff", C.c_double), ("buffer_length", C.c_double),
        ("aimed_temperature", C.c_double), ("heatbath_tau", C.c_double), ("heatbath_gamma", C.c_double),
        ("control_temperature", C.c_int), ("heatbath_type", C.c_int), ("check_list_length", C.c_int),
        ("grid", C.c_int * 3), ("capacity", C.c_int), ("pairs_per_particle", C.c_int),
    ]


_lib = None


def build():
    subprocess.run(["make", "oracle"], cwd=HERE, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(Params)]
        L.orc_destroy.argtypes = [vp]
        L.orc_num_units.argtypes = [vp]
        L.orc_get_rect.argtypes = [vp, C.c_int, dp, dp]
        L.orc_set_unit.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int]
        L.orc_add_particle.argtypes = [vp, dp, dp, C.c_int]
        L.orc_set_seed.argtypes = [vp, C.c_int]
        L.orc_set_control_temperature.argtypes = [vp, C.c_int]
        L.orc_make_pair_list.argtypes = [vp]
        L.orc_calculate.argtypes = [vp, C.c_int]
        L.orc_change_scale.argtypes = [vp, C.c_double]
        L.orc_update_position_half.argtypes = [vp]
        L.orc_send_border_particles.argtypes = [vp]
        L.orc_calculate_force.argtypes = [vp]
        L.orc_is_pairlist_expired.argtypes = [vp]
        L.orc_observe.argtypes = [vp, C.c_int, dp]
        L.orc_get_counts.argtypes = [vp, C.c_int, ip, ip, C.POINTER(C.c_long), ip]
        L.orc_get_state.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_get_pairs.argtypes = [vp, C.c_int, vp, vp]
        L.orc_get_csr.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_get_impulse.argtypes = [vp, C.c_int, vp]
        L.orc_get_zeta.restype = C.c_double
        L.orc_get_zeta.argtypes = [vp, C.c_int]
        L.orc_simulation_time.restype = C.c_double
        L.orc_simulation_time.argtypes = [vp]
        L.orc_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class System:
    """The reference's MDManager surface for the hot path, on the CPU oracle."""

    def __init__(self, L, dt, grid=(1, 1, 1), cutoff=3.0, buffer_length=0.3, T0=1.0, tau=0.1, gamma=0.1,
                 control_temperature=False, heatbath_type=0, check_list_length=500, capacity=200000,
                 pairs_per_particle=80):
        P = Params()
        for d in range(3):
            P.L[d] = L[d]
            P.grid[d] = grid[d]
        P.time_step, P.cutoff, P.buffer_length = dt, cutoff, buffer_length
        P.aimed_temperature, P.heatbath_tau, P.heatbath_gamma = T0, tau, gamma
        P.control_temperature = 1 if control_temperature else 0
        P.heatbath_type, P.check_list_length = heatbath_type, check_list_length
        P.capacity, P.pairs_per_particle = capacity, pairs_per_particle
        self.P = P
        self.L = tuple(L)
        self.h = lib().orc_create(C.byref(P))
        self.nunits = lib().orc_num_units(self.h)

    @classmethod
    def from_meta(cls, meta, **kw):
        return cls(meta["L"], meta["dt"], meta["grid"], meta["cutoff"], meta["buffer"], meta["T0"], meta["tau"],
                   meta["gamma"], bool(meta["control_temperature"]), meta["heatbath_type"],
                   meta["check_list_length"], **kw)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rect(self, unit):
        lo = (C.c_double * 3)()
        hi = (C.c_double * 3)()
        lib().orc_get_rect(self.h, unit, lo, hi)
        return np.array(lo[:]), np.array(hi[:])

    def set_unit(self, unit, q, p, type_=None):
        q = np.asarray(q, dtype=np.float64)
        p = np.asarray(p, dtype=np.float64)
        n = q.shape[0]
        q4 = np.zeros((n, 4))
        p4 = np.zeros((n, 4))
        q4[:, :q.shape[1]] = q
        p4[:, :p.shape[1]] = p
        t = None if type_ is None else np.ascontiguousarray(type_, dtype=np.int32)
        rc = lib().orc_set_unit(self.h, unit, _p(q4), _p(p4), _p(t), n)
        if rc:
            raise RuntimeError("oracle capacity exceeded")

    def set_seed(self, seed):
        lib().orc_set_seed(self.h, int(seed))

    def make_pair_list(self):
        lib().orc_make_pair_list(self.h)

    def calculate(self, nsteps=1):
        rc = lib().orc_calculate(self.h, int(nsteps))
        if rc:
            raise RuntimeError("oracle error %d: %s" % (rc, lib().orc_last_error().decode()))

    def change_scale(self, alpha):
        lib().orc_change_scale(self.h, float(alpha))

    def update_position_half(self):
        lib().orc_update_position_half(self.h)

    def send_border_particles(self):
        lib().orc_send_border_particles(self.h)

    def calculate_force(self):
        lib().orc_calculate_force(self.h)

    def is_pairlist_expired(self):
        return bool(lib().orc_is_pairlist_expired(self.h))

    def observe(self, unit=-1):
        o = (C.c_double * 4)()
        lib().orc_observe(self.h, unit, o)
        return np.array(o[:])

    def counts(self, unit):
        pn, tn, nb = C.c_int(), C.c_int(), C.c_int()
        npairs = C.c_long()
        lib().orc_get_counts(self.h, unit, C.byref(pn), C.byref(tn), C.byref(npairs), C.byref(nb))
        return {"pn": pn.value, "tn": tn.value, "npairs": npairs.value, "nbuilds": nb.value}

    def state(self, unit):
        c = self.counts(unit)
        q = np.zeros((max(c["tn"], 1), 4))
        p = np.zeros((max(c["pn"], 1), 4))
        t = np.zeros(max(c["pn"], 1), dtype=np.int32)
        lib().orc_get_state(self.h, unit, _p(q), _p(p), _p(t))
        return {"pn": c["pn"], "tn": c["tn"], "q": q[:c["tn"], :3], "p": p[:c["pn"], :3], "type": t[:c["pn"]]}

    def pairs(self, unit):
        n = self.counts(unit)["npairs"]
        k = np.zeros(max(n, 1), dtype=np.int32)
        p = np.zeros(max(n, 1), dtype=np.int32)
        lib().orc_get_pairs(self.h, unit, _p(k), _p(p))
        return k[:n], p[:n]

    def csr(self, unit):
        c = self.counts(unit)
        ptr = np.zeros(max(c["pn"], 1), dtype=np.int32)
        cnt = np.zeros(max(c["pn"], 1), dtype=np.int32)
        lst = np.zeros(max(c["npairs"], 1), dtype=np.int32)
        lib().orc_get_csr(self.h, unit, _p(ptr), _p(cnt), _p(lst))
        return ptr[:c["pn"]], cnt[:c["pn"]], lst[:c["npairs"]]

    def impulse(self, unit):
        pn = self.counts(unit)["pn"]
        f = np.zeros((max(pn, 1), 3))
        lib().orc_get_impulse(self.h, unit, _p(f))
        return f[:pn]

    def zeta(self, unit):
        return lib().orc_get_zeta(self.h, unit)
