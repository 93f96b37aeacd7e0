"""Host-side mirror of the reference driver for the hot path (MDManager / ParaInfo /
SimulationInfo / Parameter), in Python over the C-ABI.

Same names, argument meaning and error behaviour as the reference for everything the per-step
path needs, so that parity tests read like the reference's project loops
(benchmark.cc:13-52, langevin.cc:11-32, collision.cc:54-80):

    mdm = MDManager(cfg_text)                 # mdmanager.cc:11-47
    mdm.upload(states)                        # host-generated initial state -> device
    mdm.MakePairList()                        # mdmanager.cc:262-280
    for i in range(LOOP):
        mdm.Calculate()                       # mdmanager.cc:116-148
        if i % OBSERVE_LOOP == 0:
            print(mdm.GetSimulationTime(), mdm.Temperature(), mdm.Pressure(), mdm.TotalEnergy())

All numerics happen in libmdacp_b200.so on the GPU; this file only builds the POD parameter
blocks (unit grid, neighbours, sub-boxes) and sequences calls.
"""
import math

import numpy as np

from . import capi

CUTOFF_LENGTH = 3.0   # mdconfig.h:15 (as shipped)
BUFFER_LENGTH = 0.3   # simulationinfo.cc:6
D_LEFT, D_RIGHT, D_BACK, D_FORWARD, D_DOWN, D_UP = range(6)
_DIFF = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]  # parainfo.cc:8-15


class Parameter:
    """key=value lines, split at the first '=', no trimming (parameter.cc:35-46)."""

    def __init__(self, text=""):
        self.params = {}
        for line in text.splitlines():
            idx = line.find("=")
            if idx >= 0:
                key, value = line[:idx], line[idx + 1:]
                self.params.setdefault(key, value)  # std::map::insert keeps the first

    @classmethod
    def from_dict(cls, d):
        p = cls()
        p.params = {str(k): str(v) for k, v in d.items()}
        return p

    def Contains(self, key):
        return key in self.params

    def GetStringDef(self, key, value):
        return self.params.get(key, value)

    def GetDoubleDef(self, key, value):
        return _atof(self.params[key]) if key in self.params else value

    def GetDouble(self, key):
        return self.GetDoubleDef(key, 0.0)

    def GetIntegerDef(self, key, value):
        return _atoi(self.params[key]) if key in self.params else value

    def GetInteger(self, key):
        return self.GetIntegerDef(key, 0)

    def GetBooleanDef(self, key, value):
        if key not in self.params:
            return value
        return self.params[key] in ("yes", "Yes")  # parameter.cc:101-111


def _atof(s):
    import re
    m = re.match(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", s)
    return float(m.group(0)) if m else 0.0


def _atoi(s):
    import re
    m = re.match(r"\s*[+-]?\d+", s)
    return int(m.group(0)) if m else 0


def dims_create(n, ndims=3):
    """Balanced non-increasing factorisation, as MPI_Dims_create returns it (8 -> 2,2,2)."""
    dims = [1] * ndims
    f, m, p = [], n, 2
    while p * p <= m:
        while m % p == 0:
            f.append(p)
            m //= p
        p += 1
    if m > 1:
        f.append(m)
    for k in reversed(f):
        i = dims.index(min(dims))
        dims[i] *= k
    return sorted(dims, reverse=True)


class ParaInfo:
    """Unit grid (parainfo.cc:18-121): procs x threads, both 3-D; unit id <-> grid position."""

    def __init__(self, num_procs, num_threads, param):
        self.num_procs, self.num_threads = num_procs, num_threads
        g3 = dims_create(num_procs)
        l3 = dims_create(num_threads)
        self.mpi_grid_size = [g3[0], g3[1], g3[2]]
        self.openmp_grid_size = [l3[2], l3[1], l3[0]]  # reversed (parainfo.cc:31-34)
        self.valid = True
        if param.Contains("GridX"):
            if not (param.Contains("GridY") and param.Contains("GridZ")):
                self.valid = False
            else:
                self.mpi_grid_size = [param.GetInteger("GridX"), param.GetInteger("GridY"), param.GetInteger("GridZ")]
        if param.Contains("LocalGridX"):
            if not (param.Contains("LocalGridY") and param.Contains("LocalGridZ")):
                self.valid = False
            else:
                self.openmp_grid_size = [param.GetInteger("LocalGridX"), param.GetInteger("LocalGridY"),
                                         param.GetInteger("LocalGridZ")]
        gx, gy, gz = self.mpi_grid_size
        lx, ly, lz = self.openmp_grid_size
        if num_procs != gx * gy * gz or num_threads != lx * ly * lz:
            self.valid = False
            return
        n = num_procs * num_threads
        self.grid = [0] * n
        self.pos = [None] * n
        glx, gly = gx * lx, gy * ly
        for i in range(num_procs):  # parainfo.cc:124-150
            gix, giy, giz = i % gx, (i // gx) % gy, i // gx // gy
            for j in range(num_threads):
                lix, liy, liz = j % lx, (j // lx) % ly, j // lx // ly
                ix, iy, iz = gix * lx + lix, giy * ly + liy, giz * lz + liz
                uid = j + i * num_threads
                self.grid[ix + iy * glx + iz * glx * gly] = uid
                self.pos[uid] = (ix, iy, iz)

    def IsValid(self):
        return self.valid

    def GetGridSize(self):
        return [self.mpi_grid_size[d] * self.openmp_grid_size[d] for d in range(3)]

    def GetGridPosition(self, uid):
        return list(self.pos[uid])

    def GetNeighborID(self, uid, direction):
        g = self.GetGridSize()
        pos = [(self.pos[uid][d] + _DIFF[direction][d]) % g[d] for d in range(3)]
        return self.grid[pos[0] + pos[1] * g[0] + pos[2] * g[0] * g[1]]

    def IsOverBoundary(self, uid, direction):
        g = self.GetGridSize()
        return any(not (0 <= self.pos[uid][d] + _DIFF[direction][d] < g[d]) for d in range(3))

    def RankOf(self, uid):
        return uid // self.num_threads


class SimulationInfo:
    """simulationinfo.cc:5-51"""

    def __init__(self, param, grid_size, cutoff=CUTOFF_LENGTH):
        self.BufferLength = BUFFER_LENGTH
        self.Cutoff = cutoff
        self.SearchLength = cutoff + self.BufferLength
        self.TimeStep = param.GetDoubleDef("TimeStep", 0.001)
        self.IsPeriodic = param.GetBooleanDef("IsPeriodic", True)
        self.AimedTemperature = param.GetDoubleDef("AimedTemperature", 1.0)
        self.ControlTemperature = param.GetBooleanDef("ControlTemperature", False)
        self.HeatbathTau = param.GetDoubleDef("HeatbathTau", 0.1)
        self.HeatbathGamma = param.GetDoubleDef("HeatbathGamma", 0.1)
        self.SortParticle = param.GetBooleanDef("SortParticle", False)
        hb = param.GetStringDef("HeatbathType", "NoseHoover")
        if hb == "NoseHoover":
            self.HeatbathType = capi.HT_NOSEHOOVER
        elif hb == "Langevin":
            self.HeatbathType = capi.HT_LANGEVIN
        else:
            raise ValueError("Error: Unknown Heatbath type " + hb)
        self.CheckListLength = param.GetIntegerDef("CheckListLength", 500)
        self.L = [0.0, 0.0, 0.0]
        if param.Contains("SystemSize"):
            self.L = [param.GetDouble("SystemSize")] * 3
        elif param.Contains("UnitLength"):
            ul = param.GetDouble("UnitLength")
            self.L = [ul * grid_size[d] for d in range(3)]
        for d, k in enumerate(("SystemSizeX", "SystemSizeY", "SystemSizeZ")):
            if param.Contains(k):
                self.L[d] = param.GetDouble(k)


def unit_rect(sinfo, pinfo, uid):
    """mdunit.cc:14-24"""
    g = pinfo.GetGridSize()
    pos = pinfo.GetGridPosition(uid)
    s, e = [], []
    for d in range(3):
        ul = sinfo.L[d] / float(g[d])
        s.append(ul * float(pos[d]))
        e.append(s[d] + ul)
    return s, e


def make_params(sinfo, pinfo, uid, pn_capacity=0, tn_capacity=0):
    P = capi.Params()
    for d in range(3):
        P.L[d] = sinfo.L[d]
        P.grid_size[d] = pinfo.GetGridSize()[d]
        P.grid_pos[d] = pinfo.GetGridPosition(uid)[d]
    P.time_step = sinfo.TimeStep
    P.cutoff = sinfo.Cutoff
    P.buffer_length = sinfo.BufferLength
    P.aimed_temperature = sinfo.AimedTemperature
    P.heatbath_tau = sinfo.HeatbathTau
    P.heatbath_gamma = sinfo.HeatbathGamma
    P.control_temperature = 1 if sinfo.ControlTemperature else 0
    P.heatbath_type = sinfo.HeatbathType
    P.check_list_length = sinfo.CheckListLength
    P.unit_id = uid
    for direction in range(6):
        nid = pinfo.GetNeighborID(uid, direction)
        P.neighbor_id[direction] = nid
        P.neighbor_rank[direction] = pinfo.RankOf(nid)
    P.pn_capacity = pn_capacity
    P.tn_capacity = tn_capacity
    return P


class MDManager:
    """The per-step surface of the reference MDManager (mdmanager.h:37-84) over the C-ABI.

    num_procs x num_threads units as in the reference; this process owns the num_threads units
    of `rank` (one GPU each, round-robin over `devices`).  num_procs > 1 needs `unique_id`
    (bytes from capi.comm_unique_id() on rank 0, broadcast by the caller).
    """

    def __init__(self, cfg, num_procs=1, rank=0, num_threads=1, devices=None, unique_id=None,
                 cutoff=CUTOFF_LENGTH, pn_capacity=0, tn_capacity=0, barrier=None):
        self.param = cfg if isinstance(cfg, Parameter) else (
            Parameter.from_dict(cfg) if isinstance(cfg, dict) else Parameter(cfg))
        self.num_procs, self.rank, self.num_threads = num_procs, rank, num_threads
        self.pinfo = ParaInfo(num_procs, num_threads, self.param)
        if not self.pinfo.IsValid():
            raise ValueError("Invalid Grid Numbers")
        self.sinfo = SimulationInfo(self.param, self.pinfo.GetGridSize(), cutoff)
        ndev = capi.device_count()
        if ndev < 1:
            raise capi.MDB200Error(capi.ECUDA, "no CUDA device: the hot path has no CPU fallback")
        devices = list(devices) if devices is not None else [i % ndev for i in range(num_threads)]
        self.units = []
        for t in range(num_threads):
            uid = t + rank * num_threads
            P = make_params(self.sinfo, self.pinfo, uid, pn_capacity, tn_capacity)
            self.units.append(capi.Unit(P, devices[t % len(devices)]))
        self.group = capi.Group(self.units, num_procs, rank, unique_id)
        self.s_time = 0.0
        self.barrier = barrier            # callable; Communicator::Barrier for the rank-serialised writers
        self._cdview_index = 0            # static index of SaveAsCdviewSequential (mdmanager.cc:82)
        self._dump_index = 0              # static index of SaveAsDumpFileSequential (mdmanager.cc:467)
        self._snapshot_open = False

    # --- set-up -------------------------------------------------------------------------
    def GetTotalThreads(self):
        return self.num_threads

    def GetMDUnit(self, i):
        return self.units[i]

    def GetParameter(self):
        return self.param

    def GetSystemSize(self):
        return self.sinfo.L

    def GetTimeStep(self):
        return self.sinfo.TimeStep

    def GetRect(self, i):
        return unit_rect(self.sinfo, self.pinfo, i + self.rank * self.num_threads)

    def upload(self, states):
        """states: list (one per local unit) of dicts/tuples with q, p[, type]."""
        for u, st in zip(self.units, states):
            if isinstance(st, dict):
                u.upload(st["q"], st["p"], st.get("type"))
            else:
                u.upload(*st)

    def SetSeed(self, seed):
        for u in self.units:
            u.set_seed(seed)

    def SetControlTemperature(self, b):
        self.sinfo.ControlTemperature = bool(b)
        for u in self.units:
            capi.check(capi.lib().mdb200_set_control_temperature(u.handle, 1 if b else 0))

    def SetAimedTemperature(self, t):
        self.sinfo.AimedTemperature = float(t)
        for u in self.units:
            capi.check(capi.lib().mdb200_set_aimed_temperature(u.handle, float(t)))

    # --- the hot path -------------------------------------------------------------------
    def MakePairList(self):
        self.group.make_pairlist()

    def Calculate(self, nsteps=1):
        self.group.step(nsteps)
        for _ in range(nsteps):
            self.s_time += self.sinfo.TimeStep  # mdmanager.cc:146

    def SendBorderParticles(self):
        self.group.send_border_particles()

    def ChangeScale(self, alpha):
        """mdmanager.cc:432-443"""
        self.group.change_scale(alpha)
        self.sinfo.L = [x * alpha for x in self.sinfo.L]

    # --- observation (mdmanager.cc:359-429) -----------------------------------------------
    def GetSimulationTime(self):
        return self.s_time

    def observe_raw(self):
        """{K, U, W, N} raw sums over all units and ranks."""
        return self.group.observe()

    def GetTotalParticleNumber(self):
        return int(self.observe_raw()[3])

    def KineticEnergy(self):
        o = self.observe_raw()
        return o[0] / o[3]

    def PotentialEnergy(self):
        o = self.observe_raw()
        return o[1] / o[3]

    def Temperature(self):
        return self.KineticEnergy() / 1.5

    def TotalEnergy(self):
        o = self.observe_raw()
        return o[0] / o[3] + o[1] / o[3]

    def Pressure(self):
        o = self.observe_raw()
        return observables_from_raw(o, self.sinfo.L)["P"]

    def observables(self):
        return observables_from_raw(self.observe_raw(), self.sinfo.L)

    # --- dump path (mdmanager.cc:80-113, 465-505) --------------------------------------------
    def BeginSnapshot(self):
        """Queue the device -> pinned-host copy of every unit's locals (reference order); the step loop
        may go on, the Save* calls below pick the snapshot up (capi.Unit.snapshot_begin)."""
        for u in self.units:
            u.snapshot_begin()
        self._snapshot_open = True

    def _snapshots(self):
        if not self._snapshot_open:
            self.BeginSnapshot()
        self._snapshot_open = False
        return [u.snapshot_wait() for u in self.units]

    def _serial_append(self, filename, write_unit):
        """Rank 0 truncates, then the ranks append one after the other (mdmanager.cc:90-105 / ExecuteAllSerial
        :454-463), each its units in order: the file is byte-identical to the reference's."""
        from . import hostlib  # noqa: F401
        snaps = self._snapshots()
        if self.rank == 0:
            open(filename, "w").close()
        for r in range(self.num_procs):
            if self.barrier is not None:
                self.barrier()
            if r == self.rank:
                for sn in snaps:
                    write_unit(filename, sn)
        if self.barrier is not None:
            self.barrier()

    def SaveAsCdview(self, filename):
        from . import hostlib
        self._serial_append(filename, lambda fn, sn: hostlib.save_cdview(fn, sn["q4"], sn["p4"], sn["type"], append=True))

    def SaveAsCdviewSequential(self):
        filename = "conf%04d.cd" % self._cdview_index
        self._cdview_index += 1
        self.SaveAsCdview(filename)
        return filename

    def SaveAsDumpFile(self, filename):
        from . import hostlib
        self._serial_append(filename, lambda fn, sn: hostlib.save_dump(fn, sn["q4"], sn["p4"], append=True))

    def SaveAsDumpFileSequential(self):
        filename = "conf%04d.dump" % self._dump_index
        self._dump_index += 1
        self.SaveAsDumpFile(filename)
        return filename

    def SaveConfiguration(self):
        """MDUnit::SaveConfiguration (mdunit.cc:69-75): conf%03d.cnf per unit, locals and ghosts; ghost momenta
        (scratch in the reference) are written as zeros."""
        from . import hostlib
        names = []
        for i, u in enumerate(self.units):
            d = u.download()
            p4 = np.zeros((d["tn"], 4))
            p4[:d["pn"], :3] = d["p"]
            name = "conf%03d.cnf" % (i + self.rank * self.num_threads)
            hostlib.save_configuration(name, d["q"], p4)
            names.append(name)
        return names

    def close(self):
        for u in self.units:
            u.close()


def observables_from_raw(o, L):
    """T, P, E, K, U from raw {K,U,W,N} exactly as mdmanager.cc:390-429 combines them."""
    K, U, W, N = (float(x) for x in o)
    k = K / N
    T = k / 1.5
    phi = W / N
    V = L[0] * L[1] * L[2]
    return {"T": T, "P": (T - phi) * N / V, "E": k + U / N, "K": k, "U": U / N, "N": int(N)}
