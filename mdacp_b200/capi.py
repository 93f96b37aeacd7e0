"""ctypes binding of libmdacp_b200.so (include/mdacp_b200.h).

This is the only way Python reaches the hot path: there is no CPU fallback.  Importing this
module without the built library raises; calling a compute entry without a CUDA device returns
MDB200_ECUDA, which is raised as MDB200Error.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDB200_LIB_PATH") or os.path.join(_HERE, "lib", "libmdacp_b200.so")

OK, ECUDA, ECAPACITY, ETOOFAST, EPOSITION, ECELL, EINVAL, ECOMM = range(8)
HT_NOSEHOOVER, HT_LANGEVIN = 0, 1

# every symbol include/mdacp_b200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "mdb200_abi_version", "mdb200_last_error", "mdb200_device_count", "mdb200_create", "mdb200_destroy",
    "mdb200_get_rect", "mdb200_upload", "mdb200_download", "mdb200_get_counts", "mdb200_comm_unique_id",
    "mdb200_comm_init", "mdb200_make_pairlist", "mdb200_step", "mdb200_is_pairlist_expired",
    "mdb200_update_position_half", "mdb200_send_border_particles", "mdb200_calculate_force",
    "mdb200_heatbath_zeta", "mdb200_heatbath_momenta", "mdb200_langevin", "mdb200_set_seed", "mdb200_get_zeta",
    "mdb200_set_zeta", "mdb200_set_control_temperature", "mdb200_set_aimed_temperature", "mdb200_observe",
    "mdb200_observe_unit", "mdb200_get_pairs", "mdb200_get_impulse", "mdb200_timers", "mdb200_set_profiling",
    "mdb200_synchronize", "mdb200_event_record", "mdb200_event_elapsed_ms", "mdb200_list_stats",
    "mdb200_snapshot_begin", "mdb200_snapshot_wait", "mdb200_change_scale", "mdb200_measure_fp64_peak",
    "mdb200_clear", "mdb200_generate_benchmark", "mdb200_generate_lattice", "mdb200_generate_ball",
    "mdb200_generate_initial_velocity", "mdb200_upload_begin",
]


class MDB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mdb200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("L", C.c_double * 3), ("time_step", C.c_double), ("cutoff", C.c_double), ("buffer_length", C.c_double),
        ("aimed_temperature", C.c_double), ("heatbath_tau", C.c_double), ("heatbath_gamma", C.c_double),
        ("control_temperature", C.c_int), ("heatbath_type", C.c_int), ("check_list_length", C.c_int),
        ("grid_size", C.c_int * 3), ("grid_pos", C.c_int * 3), ("unit_id", C.c_int), ("neighbor_id", C.c_int * 6),
        ("neighbor_rank", C.c_int * 6), ("pn_capacity", C.c_longlong), ("tn_capacity", C.c_longlong),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libmdacp_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` -- there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.mdb200_last_error.restype = C.c_char_p
        vp, ip, dp, llp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_longlong)
        L.mdb200_device_count.argtypes = [ip]
        L.mdb200_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
        L.mdb200_destroy.argtypes = [vp]
        L.mdb200_get_rect.argtypes = [vp, dp, dp]
        L.mdb200_upload.argtypes = [vp, vp, vp, vp, C.c_longlong]
        L.mdb200_upload_begin.argtypes = [vp, vp, vp, vp, C.c_longlong]
        L.mdb200_clear.argtypes = [vp]
        L.mdb200_generate_benchmark.argtypes = [vp, C.c_double, llp]
        L.mdb200_generate_lattice.argtypes = [vp, C.c_double, C.c_double, llp]
        L.mdb200_generate_ball.argtypes = [vp, dp, C.c_double, C.c_double, C.c_double, llp]
        L.mdb200_generate_initial_velocity.argtypes = [vp, C.c_double, C.c_int]
        L.mdb200_download.argtypes = [vp, vp, vp, vp, C.c_longlong, llp, llp]
        L.mdb200_get_counts.argtypes = [vp, llp, llp, llp, llp]
        L.mdb200_comm_unique_id.argtypes = [vp]
        L.mdb200_comm_init.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]
        L.mdb200_make_pairlist.argtypes = [C.POINTER(vp), C.c_int]
        L.mdb200_step.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
        L.mdb200_is_pairlist_expired.argtypes = [C.POINTER(vp), C.c_int, ip]
        L.mdb200_update_position_half.argtypes = [vp]
        L.mdb200_send_border_particles.argtypes = [C.POINTER(vp), C.c_int]
        L.mdb200_calculate_force.argtypes = [vp]
        L.mdb200_heatbath_zeta.argtypes = [vp, C.c_double]
        L.mdb200_heatbath_momenta.argtypes = [vp]
        L.mdb200_langevin.argtypes = [vp]
        L.mdb200_set_seed.argtypes = [vp, C.c_int]
        L.mdb200_get_zeta.argtypes = [vp, dp]
        L.mdb200_set_zeta.argtypes = [vp, C.c_double]
        L.mdb200_set_control_temperature.argtypes = [vp, C.c_int]
        L.mdb200_set_aimed_temperature.argtypes = [vp, C.c_double]
        L.mdb200_observe.argtypes = [C.POINTER(vp), C.c_int, dp]
        L.mdb200_observe_unit.argtypes = [vp, dp]
        L.mdb200_get_pairs.argtypes = [vp, vp, vp, C.c_longlong, llp]
        L.mdb200_get_impulse.argtypes = [vp, vp, C.c_longlong]
        L.mdb200_timers.argtypes = [vp, dp, llp, C.c_int]
        L.mdb200_set_profiling.argtypes = [vp, C.c_int]
        L.mdb200_synchronize.argtypes = [vp]
        L.mdb200_event_record.argtypes = [vp, C.c_int]
        L.mdb200_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, dp]
        L.mdb200_list_stats.argtypes = [vp, dp]
        L.mdb200_measure_fp64_peak.argtypes = [C.c_int, dp]
        L.mdb200_snapshot_begin.argtypes = [vp]
        L.mdb200_change_scale.argtypes = [C.POINTER(vp), C.c_int, C.c_double]
        L.mdb200_snapshot_wait.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), llp]
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise MDB200Error(rc, lib().mdb200_last_error().decode(errors="replace"))


def measure_fp64_peak(device=0):
    """measured FP64 TFLOP/s of a device (DFMA loop)"""
    t = C.c_double()
    check(lib().mdb200_measure_fp64_peak(int(device), C.byref(t)))
    return t.value


def device_count():
    n = C.c_int(0)
    rc = lib().mdb200_device_count(C.byref(n))
    return n.value if rc == OK else 0


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Unit:
    """One reference MDUnit on one device (opaque handle)."""

    def __init__(self, params, device=0):
        self.params = params
        self.handle = C.c_void_p()
        check(lib().mdb200_create(C.byref(params), device, C.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().mdb200_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rect(self):
        s = (C.c_double * 3)()
        e = (C.c_double * 3)()
        check(lib().mdb200_get_rect(self.handle, s, e))
        return np.array(s[:]), np.array(e[:])

    def upload(self, q, p, type_=None):
        """q, p: (n,3) or (n,4) float64 in the reference's index order."""
        q = np.asarray(q, dtype=np.float64)
        p = np.asarray(p, dtype=np.float64)
        n = q.shape[0]
        q4 = np.zeros((n, 4))
        p4 = np.zeros((n, 4))
        q4[:, :q.shape[1]] = q
        p4[:, :p.shape[1]] = p
        t = None if type_ is None else np.ascontiguousarray(type_, dtype=np.int32)
        check(lib().mdb200_upload(self.handle, _ptr(q4), _ptr(p4), _ptr(t), n))

    def upload4(self, q4, p4, type_=None):
        """zero-copy variant: C-contiguous (n,4) float64 arrays (may be pinned)."""
        assert q4.dtype == np.float64 and p4.dtype == np.float64 and q4.shape[1] == 4 and p4.shape[1] == 4
        check(lib().mdb200_upload(self.handle, _ptr(q4), _ptr(p4), _ptr(type_), q4.shape[0]))

    # on-device initial states (include/mdacp_b200.h; generate.cu): each returns the number of particles appended
    def clear(self):
        check(lib().mdb200_clear(self.handle))

    def generate_benchmark(self, density=0.7):
        n = C.c_longlong()
        check(lib().mdb200_generate_benchmark(self.handle, density, C.byref(n)))
        return n.value

    def generate_lattice(self, density, radius=0.0):
        n = C.c_longlong()
        check(lib().mdb200_generate_lattice(self.handle, density, radius, C.byref(n)))
        return n.value

    def generate_ball(self, c, r, xv, density):
        n = C.c_longlong()
        c3 = (C.c_double * 3)(*[float(x) for x in c])
        check(lib().mdb200_generate_ball(self.handle, c3, r, xv, density, C.byref(n)))
        return n.value

    def generate_initial_velocity(self, v0, seed):
        check(lib().mdb200_generate_initial_velocity(self.handle, v0, int(seed)))

    def upload4_begin(self, q4, p4, type_=None):
        """mdb200_upload_begin: returns once the positions are on the device; p4 (page-locked) must stay untouched
        until the next call on this unit has returned."""
        assert q4.dtype == np.float64 and p4.dtype == np.float64 and q4.shape[1] == 4 and p4.shape[1] == 4
        self._pending_upload = (q4, p4, type_)   # keep the host arrays alive
        check(lib().mdb200_upload_begin(self.handle, _ptr(q4), _ptr(p4), _ptr(type_), q4.shape[0]))

    def counts(self):
        a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_longlong()
        check(lib().mdb200_get_counts(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"pn": a.value, "tn": b.value, "npairs": c.value, "nbuilds": d.value}

    def download(self, want_q=True, want_p=True, want_type=True, out=None):
        c = self.counts()
        pn, tn = c["pn"], c["tn"]
        rows = max(tn, 1)
        if out is None:
            q = np.zeros((rows, 4)) if want_q else None
            p = np.zeros((rows, 4)) if want_p else None
        else:
            q, p = out  # caller's (possibly pinned) (n,4) buffers; ghosts are returned as far as they fit
            rows = min(x.shape[0] for x in (q, p) if x is not None)
            tn = min(tn, rows)
        t = np.zeros(rows, dtype=np.int32) if want_type else None
        a, b = C.c_longlong(), C.c_longlong()
        check(lib().mdb200_download(self.handle, _ptr(q), _ptr(p), _ptr(t), rows, C.byref(a), C.byref(b)))
        return {"pn": pn, "tn": tn, "q": None if q is None else q[:tn, :3], "p": None if p is None else p[:pn, :3],
                "type": None if t is None else t[:pn]}

    def snapshot_begin(self):
        """Queue an asynchronous copy of q, p, type (locals, reference order) to pinned host memory."""
        check(lib().mdb200_snapshot_begin(self.handle))

    def snapshot_wait(self):
        """-> {"pn", "q4", "p4", "type"}: read-only views of the library's pinned buffers (rows of 4 doubles),
        valid until the next snapshot_begin / close on this unit."""
        q, p, t, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_longlong()
        check(lib().mdb200_snapshot_wait(self.handle, C.byref(q), C.byref(p), C.byref(t), C.byref(n)))
        pn = n.value
        if pn == 0:
            return {"pn": 0, "q4": np.zeros((0, 4)), "p4": np.zeros((0, 4)), "type": np.zeros(0, dtype=np.int32)}
        q4 = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_double)), shape=(pn, 4))
        p4 = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(pn, 4))
        ty = np.ctypeslib.as_array(C.cast(t, C.POINTER(C.c_int)), shape=(pn,))
        return {"pn": pn, "q4": q4, "p4": p4, "type": ty}

    def pairs(self):
        n = C.c_longlong()
        check(lib().mdb200_get_pairs(self.handle, None, None, 0, C.byref(n)))
        k = np.zeros(max(n.value, 1), dtype=np.int32)
        p = np.zeros(max(n.value, 1), dtype=np.int32)
        check(lib().mdb200_get_pairs(self.handle, _ptr(k), _ptr(p), n.value, C.byref(n)))
        return k[:n.value], p[:n.value]

    def impulse(self):
        pn = self.counts()["pn"]
        f = np.zeros((max(pn, 1), 3))
        check(lib().mdb200_get_impulse(self.handle, _ptr(f), max(pn, 1)))
        return f[:pn]

    def observe(self):
        o = (C.c_double * 4)()
        check(lib().mdb200_observe_unit(self.handle, o))
        return np.array(o[:])

    def update_position_half(self):
        check(lib().mdb200_update_position_half(self.handle))

    def calculate_force(self):
        check(lib().mdb200_calculate_force(self.handle))

    def langevin(self):
        check(lib().mdb200_langevin(self.handle))

    def heatbath_momenta(self):
        check(lib().mdb200_heatbath_momenta(self.handle))

    def set_seed(self, seed):
        check(lib().mdb200_set_seed(self.handle, int(seed)))

    @property
    def zeta(self):
        z = C.c_double()
        check(lib().mdb200_get_zeta(self.handle, C.byref(z)))
        return z.value

    @zeta.setter
    def zeta(self, v):
        check(lib().mdb200_set_zeta(self.handle, float(v)))

    def set_profiling(self, level):
        """0 off; 1 force kernel only (no extra syncs); 2 / True every phase"""
        level = 2 if level is True else int(level)
        check(lib().mdb200_set_profiling(self.handle, level))

    def list_stats(self):
        o = (C.c_double * 8)()
        check(lib().mdb200_list_stats(self.handle, o))
        return {"warps": int(o[0]), "padded_entries": int(o[1]), "max_row": int(o[2]), "entries": int(o[3]),
                "lanes_per_local": int(o[4]), "walked_entries": int(o[5]), "derive_ms": float(o[6]),
                "derive_launches": int(o[7])}

    def event_record(self, slot):
        check(lib().mdb200_event_record(self.handle, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        check(lib().mdb200_event_elapsed_ms(self.handle, a, b, C.byref(ms)))
        return ms.value

    def timers(self, reset=False):
        ms = (C.c_double * 6)()
        n = (C.c_longlong * 2)()
        check(lib().mdb200_timers(self.handle, ms, n, 1 if reset else 0))
        names = ["all", "force", "comm", "pair", "integrate", "check"]
        d = dict(zip(names, ms[:]))
        d["launches"] = n[0]
        d["force_launches"] = n[1]
        return d

    def synchronize(self):
        check(lib().mdb200_synchronize(self.handle))


class Group:
    """The units of this process driven together (what MDManager holds in `mdv`)."""

    def __init__(self, units, nranks=1, rank=0, unique_id=None):
        self.units = list(units)
        self.n = len(self.units)
        self._arr = (C.c_void_p * self.n)(*[u.handle for u in self.units])
        uid = None
        if unique_id is not None:
            self._uid = C.create_string_buffer(bytes(unique_id), 128)
            uid = C.cast(self._uid, C.c_void_p)
        check(lib().mdb200_comm_init(self._arr, self.n, nranks, rank, uid))

    def make_pairlist(self):
        check(lib().mdb200_make_pairlist(self._arr, self.n))

    def step(self, nsteps=1):
        check(lib().mdb200_step(self._arr, self.n, int(nsteps)))

    def change_scale(self, alpha):
        check(lib().mdb200_change_scale(self._arr, self.n, float(alpha)))

    def is_pairlist_expired(self):
        e = C.c_int()
        check(lib().mdb200_is_pairlist_expired(self._arr, self.n, C.byref(e)))
        return bool(e.value)

    def send_border_particles(self):
        check(lib().mdb200_send_border_particles(self._arr, self.n))

    def observe(self):
        o = (C.c_double * 4)()
        check(lib().mdb200_observe(self._arr, self.n, o))
        return np.array(o[:])

    def synchronize(self):
        for u in self.units:
            u.synchronize()


def comm_unique_id():
    buf = C.create_string_buffer(128)
    check(lib().mdb200_comm_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw
