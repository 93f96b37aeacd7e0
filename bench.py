#!/usr/bin/env python
"""bench.py -- LJ particle-updates/sec of the hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W              # our arm (libmdacp_b200.so)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU path

A "step" is one MDManager::Calculate() (mdmanager.cc:116-148) on the benchmark.cfg-shaped workload:
FCC LJ, density 0.7, cutoff 3.0 (as shipped), dt 0.001, v0 1.0, one cubic unit of edge UNIT_LENGTH per
GPU (284 -> 16,034,412 particles/GPU; BASELINE.json configs[1]/[2]).  Like Benchmark::Run (benchmark.cc:24-45) the
state is thermalised first (ThermalizeLoop = 150 steps, untimed set-up, --thermalize), then W warm-up steps, then the
K timed steps with whatever margin-based pair-list rebuilds, ghost refreshes and observations (every 100 steps) fall
into them.  value = N_total * K / seconds.  Because K may be too short to hold a rebuild (one every ~85 steps), the line
also carries `steady_state`: the same metric over a long window (>= 300 steps, several rebuilds) measured right after,
plus the cost and cadence of the rebuild and of the inner-list derivation.

Rank 0 prints ONE JSON line (contract in the task statement): value (state resident in HBM), e2e
(same metric through the C-ABI from/to pinned HOST buffers: upload + first list build + K steps +
observation + download inside the timed region), roofline of the force kernel, cpu_baseline.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "LJ particle-updates/sec"
UNIT = "particle-updates/s"
OBSERVE_LOOP = 100
DENSITY = 0.7
V0 = 1.0
DT = 0.001
CUTOFF = 3.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML (nvidia-ml-py) every 50 ms from a
    thread while the timed region runs -- the recipe's nvidia-smi clocks line, without the
    process start-up latency that would miss a sub-second region."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.h = None
        self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.05)

    def prepare(self):
        """NVML initialisation (can take ~100 ms on a fresh box): done BEFORE the barrier that opens the timed
        region, so that the other ranks do not spend it waiting for rank 0 inside their timed region."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            # the first query of each kind can take > 100 ms on an 8-GPU box and holds a driver lock the rank's own
            # launches then wait for (seen as one sample in a 180 ms window and a 9 ms/step outlier): ask once here
            for _ in range(2):
                nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                (nv.nvmlDeviceGetCurrentClocksEventReasons if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                 else nv.nvmlDeviceGetCurrentClocksThrottleReasons)(self.h)
                nv.nvmlDeviceGetPowerUsage(self.h)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None and self.h is None:
            self.prepare()
        if self.nv is None:
            self.thread = None
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is None:
            return out
        self.stop_flag.set()
        self.thread.join(timeout=2)
        nv = self.nv
        if not self.samples:
            return out
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for s in self.samples:
            bits |= s[1]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, v in names.items() if bits & v)
        out.update({"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "power_w_max": max(s[2] for s in self.samples)})
        return out


# ------------------------------------------------------------------------------------------
# reference / CPU baseline
# ------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cubic_units(cores):
    """largest n^3 <= cores: Benchmark mode asserts cubic sub-boxes (confmaker.cc:181-182)"""
    n = 1
    while (n + 1) ** 3 <= cores:
        n += 1
    return n ** 3


def unit_grid_for(cores):
    """Unit grid a x b x c (a <= b <= c <= 2a) with the most units <= cores: the sub-boxes stay cubes when the box is
    a x b x c unit edges, so every host core gets a unit (16 cores -> 2x2x4, not only 2x2x2)."""
    best = (1, 1, 1)
    for a in range(1, 9):
        for b in range(a, 2 * a + 1):
            for c in range(b, 2 * a + 1):
                if a * b * c <= cores and (a * b * c, -c) > (best[0] * best[1] * best[2], -best[2]):
                    best = (a, b, c)
    return best


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return float(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def run_reference_cpu(unit_length, n_gpus_equiv, therm, loop, timeout=1500):
    """Times the reference's own CPU implementation (oracle/_ref, unmodified sources) on the host
    cores: Mode=Benchmark on the SAME system as the GPU arm (global box = the GPU run's box), with
    as many OpenMP units as fit a cubic grid.  Returns dict(value, cores, kind, sample, ...)."""
    from oracle import refdump
    from mdacp_b200 import driver
    cores = host_cores()
    # bounded sample: the CPU cannot hold the N-GPU weak-scaled system (N x 16 M particles), and its
    # throughput does not grow with the box, so every N is compared with a box of one-GPU size or a small multiple:
    # CPU units of edge unit_length / 2 (what 2x2x2 units make of the one-GPU box), one per host core on the largest
    # grid that keeps them cubes -- 8 cores: the one-GPU box itself; 16 cores: 2x2x4 units, i.e. twice that box.
    lgrid = list(unit_grid_for(cores)) if cores >= 8 else [2, 2, 2]
    units = lgrid[0] * lgrid[1] * lgrid[2]
    # at most ~2.1 one-GPU systems in total (many-core hosts get smaller units rather than a bigger box)
    n_cap = 2.1 * DENSITY * unit_length ** 3
    edge = min(unit_length / 2.0, (n_cap / DENSITY / units) ** (1.0 / 3.0))
    need_gb = 0.8e-6 * DENSITY * edge ** 3 * units  # measured: 12.2 GB for 16 M particles
    if need_gb * 1.5 > mem_available_gb() > 0.0:
        lgrid, units, edge = [2, 2, 2], 8, unit_length / 2.0
    L = [edge * lgrid[0], edge * lgrid[1], edge * lgrid[2]]
    n_total = int(L[0] * L[1] * L[2] * DENSITY)
    per_unit = n_total / units * 1.35 + 50000  # locals + ghosts
    exe = refdump.REF_AVX2 if per_unit < 950000 else refdump.REF_AVX2_BIG
    kind = "reference"
    if not refdump.have_ref(exe):
        return run_oracle_port_cpu(therm, loop)
    cfg = ("Mode=Benchmark\nThermalizeLoop=%d\nTotalLoop=%d\nObserveLoop=%d\nTimeStep=%r\nInitialVelocity=%r\n"
           "SystemSizeX=%r\nSystemSizeY=%r\nSystemSizeZ=%r\nLocalGridX=%d\nLocalGridY=%d\nLocalGridZ=%d\n" % (
               therm, loop, OBSERVE_LOOP, DT, V0, L[0], L[1], L[2], lgrid[0], lgrid[1], lgrid[2]))
    tmp = tempfile.mkdtemp(prefix="refbench_")
    with open(os.path.join(tmp, "bench.cfg"), "w") as f:
        f.write(cfg)
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(units)
    env["OMP_PROC_BIND"] = "true"
    t0 = time.time()
    res = subprocess.run([exe, "bench.cfg"], cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=timeout)
    wall = time.time() - t0
    m = re.search(r"# N = (\d+) ([0-9.eE+-]+) \[SEC\] ([0-9.eE+-]+) \[MUPS\]", res.stdout)
    if not m:
        log("reference run failed:\n" + res.stdout[-2000:])
        return {"value": None, "unit": UNIT, "cores": units, "kind": kind, "sample": "failed: " + res.stdout[-200:]}
    n, sec, mups = int(m.group(1)), float(m.group(2)), float(m.group(3))
    return {"value": mups * 1e6, "unit": UNIT, "cores": units, "kind": kind,
            "sample": "unmodified reference (%s, AVX2 kernel, %d OpenMP units %dx%dx%d) Mode=Benchmark box %gx%gx%g "
                      "N=%d, %d thermalise + %d timed steps; %.1f s timed, %.0f s wall incl. set-up; host %s, %d cores" % (
                          os.path.basename(exe), units, lgrid[0], lgrid[1], lgrid[2], L[0], L[1], L[2], n, therm, loop,
                          sec, wall, cpu_model(), cores),
            "n": n, "sec": sec, "steps": loop}


def run_reference_cpu_project(workload, unit_length, collision_scale, loop, timeout=1500):
    """The reference's own CPU path on the other BASELINE.json workloads (Mode=Langevin / Mode=Collision print no
    timing line, benchmark.cfg as shipped does): the unmodified binary is run twice, with TotalLoop = a and a + loop
    steps, and the difference of the wall-clock times is the time of `loop` steps -- set-up (lattice generation, first
    pair list) cancels.  Bounded samples: one cubic unit per host core as in run_reference_cpu."""
    from oracle import refdump
    cores = host_cores()
    lgrid = list(unit_grid_for(cores)) if cores >= 8 else [2, 2, 2]
    units = lgrid[0] * lgrid[1] * lgrid[2]
    if workload == "shipped":
        n3 = cubic_units(cores)
        g = round(n3 ** (1.0 / 3.0))
        lgrid, units = [g, g, g], n3
        base = "Mode=Benchmark\nThermalizeLoop=150\nSystemSize=50\nTimeStep=0.001\nInitialVelocity=1.0\n"
        what = "benchmark.cfg as shipped (SystemSize 50), %d OpenMP units" % units
    elif workload == "langevin":
        edge = min(unit_length / 2.0, (2.1 * DENSITY * unit_length ** 3 / DENSITY / units) ** (1.0 / 3.0))
        L = [edge * lgrid[0], edge * lgrid[1], edge * lgrid[2]]
        base = ("Mode=Langevin\nInitialVelocity=1.8\nTimeStep=0.005\nSystemSizeX=%r\nSystemSizeY=%r\nSystemSizeZ=%r\n"
                "ObserveLoop=100000\nAimedTemperature=0.9\nDensity=0.70\nHeatbathTau=0.1\nHeatbathGamma=0.1\n"
                "HeatbathType=Langevin\nSeed=12345\n" % (L[0], L[1], L[2]))
        what = "langevin.cfg scaled to a %gx%gx%g box, %d OpenMP units" % (L[0], L[1], L[2], units)
    else:
        k = collision_scale
        # the droplets sit in the x halves: units along x first keeps every core busy
        lgrid = [4, 2, 2] if cores >= 16 else [2, 2, 2]
        units = lgrid[0] * lgrid[1] * lgrid[2]
        base = ("Mode=Collision\nSystemSizeX=%r\nSystemSizeY=%r\nSystemSizeZ=%r\nTimeStep=0.002\nObserveLoop=100000\n"
                "DropletSpeed=1.0\nSaveCdviewFile=no\n" % (120.0 * k, 60.0 * k, 60.0 * k))
        what = "collision.cfg scaled x%g per axis, %d OpenMP units" % (k, units)
    base += "LocalGridX=%d\nLocalGridY=%d\nLocalGridZ=%d\n" % (lgrid[0], lgrid[1], lgrid[2])
    exe = refdump.REF_AVX2_BIG if refdump.have_ref(refdump.REF_AVX2_BIG) else refdump.REF_AVX2
    if not refdump.have_ref(exe):
        return {"value": None, "unit": UNIT, "cores": units, "kind": "reference", "sample": "oracle/_ref not built"}
    env = dict(os.environ, OMP_NUM_THREADS=str(units), OMP_PROC_BIND="true")
    times, n = [], None
    first = 30 if workload != "shipped" else 0   # past the first builds / the start-up transient
    for total in (first, first + loop):
        tmp = tempfile.mkdtemp(prefix="refproj_")
        with open(os.path.join(tmp, "input.cfg"), "w") as f:
            f.write(base + "TotalLoop=%d\n" % total)
        t0 = time.time()
        res = subprocess.run([exe, "input.cfg"], cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                             text=True, timeout=timeout)
        times.append(time.time() - t0)
        m = re.search(r"# Number of Particles = (\d+)", res.stdout)
        if m:
            n = int(m.group(1))
        if res.returncode != 0 or n is None:
            return {"value": None, "unit": UNIT, "cores": units, "kind": "reference", "sample": "failed: " + res.stdout[-300:]}
    sec = max(times[1] - times[0], 1e-9)
    return {"value": n * loop / sec, "unit": UNIT, "cores": units, "kind": "reference",
            "sample": "unmodified reference (%s, AVX2 kernel): %s, N=%d; wall time of TotalLoop=%d minus TotalLoop=%d: "
                      "%.1f s for %d steps; host %s, %d cores" % (os.path.basename(exe), what, n, first + loop, first, sec,
                                                                  loop, cpu_model(), cores),
            "n": n, "sec": sec, "steps": loop}


def run_oracle_port_cpu(therm, loop):
    """Fallback when oracle/_ref is absent: the C restatement (oracle/liboracle.so), 8 OpenMP units,
    on a bounded 2M-particle sample of the same lattice."""
    import numpy as np
    from oracle import pyoracle
    from mdacp_b200 import hostlib
    Ls = 142.0
    osys = pyoracle.System([Ls] * 3, DT, grid=(2, 2, 2), cutoff=CUTOFF, capacity=600000)
    n = 0
    for u in range(8):
        lo, hi = osys.rect(u)
        q = hostlib.benchmark_config(lo, hi, DENSITY, u)
        p = hostlib.initial_velocity(len(q), V0, u)
        osys.set_unit(u, q, p)
        n += len(q)
    osys.make_pair_list()
    osys.calculate(therm)
    t0 = time.time()
    osys.calculate(loop)
    sec = time.time() - t0
    osys.close()
    return {"value": n * loop / sec, "unit": UNIT, "cores": min(8, host_cores()), "kind": "port",
            "sample": "C oracle port, 2x2x2 units, box %g^3 N=%d, %d+%d steps, %.1f s; host %s" % (
                Ls, n, therm, loop, sec, cpu_model()), "n": n, "sec": sec, "steps": loop}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mdacp_b200 import capi, driver, hostlib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log("warning: WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    UL = float(args.unit_length)
    wl = args.workload
    # --- the workload (BASELINE.json configs): cfg of the reference project, host-side generator, step variant ---
    if wl == "benchmark":      # configs[1] / [2]: benchmark.cfg scaled, one cubic unit of edge UL per GPU (weak scaling)
        cfg = {"Mode": "Benchmark", "UnitLength": repr(UL), "TimeStep": repr(DT), "InitialVelocity": repr(V0)}
        wl_text = ("benchmark.cfg-shaped LJ FCC, UnitLength=%g per GPU, density 0.7, cutoff 3.0 (as shipped), skin 0.3, "
                   "dt 0.001, v0 1.0" % UL)
        scaling = "weak"
    elif wl == "shipped":      # configs[0]: benchmark.cfg verbatim (SystemSize 50 -> 87,500 particles), one unit
        cfg = {"Mode": "Benchmark", "SystemSize": "50", "TimeStep": repr(DT), "InitialVelocity": repr(V0)}
        wl_text = "benchmark.cfg as shipped: SystemSize=50 (87,500 particles), cutoff 3.0, dt 0.001, v0 1.0, one unit"
        scaling = "strong"
    elif wl == "langevin":     # configs[3]: langevin.cfg with UnitLength = UL per GPU (langevin.cc:11-32)
        cfg = {"Mode": "Langevin", "UnitLength": repr(UL), "TimeStep": "0.005", "AimedTemperature": "0.9",
               "HeatbathGamma": "0.1", "HeatbathTau": "0.1", "HeatbathType": "Langevin", "ControlTemperature": "yes",
               "Density": "0.70", "InitialVelocity": "1.8", "Seed": "12345"}
        wl_text = ("langevin.cfg scaled: Langevin thermostat (gamma 0.1, T0 0.9, per-particle counter RNG), global FCC lattice "
                   "density 0.70, UnitLength=%g per GPU, dt 0.005, v0 1.8, cutoff 3.0" % UL)
        scaling = "weak"
    else:                      # configs[4]: collision.cfg with the box scaled x10 per axis (collision.cc:54-80)
        k = float(args.collision_scale)
        cfg = {"Mode": "Collision", "SystemSizeX": repr(120.0 * k), "SystemSizeY": repr(60.0 * k),
               "SystemSizeZ": repr(60.0 * k), "TimeStep": "0.002"}
        wl_text = ("collision.cfg scaled x%g per axis: two FCC droplets (density 0.374, radius %g, +-1 along x) in a "
                   "%gx%gx%g box, the rest vacuum, dt 0.002, cutoff 3.0; the SAME box on every GPU count" % (
                       k, 15.0 * k, 120.0 * k, 60.0 * k, 60.0 * k))
        scaling = "strong"
    mdm = driver.MDManager(cfg, num_procs=world, rank=rank, num_threads=1, devices=[local_rank], unique_id=uid,
                           cutoff=CUTOFF)
    unit = mdm.GetMDUnit(0)
    s, e = mdm.GetRect(0)
    grid = mdm.pinfo.GetGridSize()

    # --- synthetic input: the reference's own generator semantics (host side, C++: mdacp_b200/host/hostlib.cpp) ---
    t0 = time.time()
    vol = (e[0] - s[0]) * (e[1] - s[1]) * (e[2] - s[2])
    if wl in ("benchmark", "shipped"):
        cap = int((int(vol * DENSITY) + 64) * 1.02) + 65536  # head-room: migration changes the per-unit count
    elif wl == "langevin":
        cap = int(vol * 0.70 * 1.05) + 65536
    else:
        # two balls of radius L_y / 4 at density 0.374 (collision.cc:54-61), wherever they fall among the units: a unit holds
        # at most both of them (the box is mostly vacuum: sizing by its volume would pin 13 GB of host memory)
        r_ball = 0.25 * 60.0 * float(args.collision_scale)
        cap = min(int(vol * 0.374 * 1.3), int(2 * 4.19 * r_ball ** 3 * 0.374 * 1.15)) + 65536
    q_pin = torch.zeros((cap, 4), dtype=torch.float64).pin_memory()
    p_pin = torch.zeros((cap, 4), dtype=torch.float64).pin_memory()
    qh, ph = q_pin.numpy(), p_pin.numpy()
    if args.device_init:
        # the same makers on the device (generate.cu): positions bit-identical to the host makers', momenta
        # to an ulp (CUDA's cos / sin) -- an alternative provenance of the initial state, hence a flag
        if wl in ("benchmark", "shipped"):
            n_local = unit.generate_benchmark(DENSITY)
            unit.generate_initial_velocity(V0, rank)
        elif wl == "langevin":
            n_local = unit.generate_lattice(0.70)
            unit.generate_initial_velocity(1.8, rank)
        else:
            L = mdm.sinfo.L
            n_local = unit.generate_ball([L[0] * 0.25, L[1] * 0.5, L[2] * 0.5], L[1] * 0.25, 1.0, 0.374)
            n_local += unit.generate_ball([L[0] * 0.75, L[1] * 0.5, L[2] * 0.5], L[1] * 0.25, -1.0, 0.374)
        unit.synchronize()
        if n_local > cap:
            raise MemoryError("pinned host buffers too small for %d particles" % n_local)
    elif wl in ("benchmark", "shipped"):
        n_local = hostlib.benchmark_config(s, e, DENSITY, rank, out=qh).shape[0]
        hostlib.initial_velocity(n_local, V0, rank, out=ph)
    elif wl == "langevin":
        q0 = hostlib.lattice_config(mdm.sinfo.L, s, e, density=0.70)
        n_local = q0.shape[0]
        qh[:n_local] = q0
        hostlib.initial_velocity(n_local, 1.8, rank, out=ph)
        del q0
    else:
        q0, p0 = hostlib.collision_config(mdm.sinfo.L, s, e)
        n_local = q0.shape[0]
        qh[:n_local] = q0
        ph[:n_local] = p0
        del q0, p0
    t_gen = time.time() - t0
    n_total = int(allsum(float(n_local)))
    n_max = int(allmax(float(n_local)))
    log("[rank %d] %s unit grid %s box %s..%s  N_local=%d N_total=%d  (generated %sin %.2f s)" % (
        rank, wl, grid, [round(x, 3) for x in s], [round(x, 3) for x in e], n_local, n_total,
        "on the device " if args.device_init else "", t_gen))
    if wl == "langevin":
        mdm.SetSeed(12345)

    # --- state resident in HBM: upload, first build, warm-up (untimed) ---
    if not args.device_init:
        unit.upload4_begin(qh[:n_local], ph[:n_local])   # the same entry the timed e2e region uses below (warms it up)
    mdm.MakePairList()
    W, K = args.warmup, args.steps
    obs = None
    for i in range(args.thermalize):   # ThermalizeLoop of benchmark.cfg (benchmark.cc:24-27): untimed set-up
        mdm.Calculate()
    for i in range(W):
        mdm.Calculate()
    mdm.observables()   # untimed: the first observation allocates its partial sums and, across ranks, opens the sum all-reduce
    unit.synchronize()
    c0 = unit.counts()
    npairs = c0["npairs"]

    # --- timed region 1: `value` (K steps, rebuilds + ghost refresh + observations inside) ---
    sampler = ClockSampler(local_rank)
    unit.set_profiling(1)  # force kernel timed with CUDA events on the unit's stream, no extra syncs
    unit.timers(reset=True)
    if rank == 0:
        sampler.prepare()
    barrier()
    if rank == 0:
        sampler.start()
    unit.event_record(0)
    tw0 = time.time()
    for i in range(K):
        mdm.Calculate()
        if i % OBSERVE_LOOP == 0:
            obs = mdm.observables()  # T, P, E of benchmark.cc:34-40 (4 doubles back to the host)
    unit.event_record(1)
    unit.synchronize()
    barrier()
    tw1 = time.time()
    ms_dev = unit.event_elapsed_ms(0, 1)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = allmax(ms_dev)
    tm = unit.timers()
    c1 = unit.counts()
    unit.set_profiling(0)
    launches = int(allsum(float(tm["launches"])))
    value = n_total * K / (ms_total * 1e-3)
    rebuilds = c1["nbuilds"] - c0["nbuilds"]
    n_obs = (K + OBSERVE_LOOP - 1) // OBSERVE_LOOP

    # --- steady state: the same metric over a window long enough for several rebuilds (untimed for `value`) ---
    S = max(int(args.steady_steps), 0)
    steady = None
    if S > 0:
        unit.set_profiling(1)
        unit.timers(reset=True)
        cs0 = unit.counts()
        barrier()
        unit.event_record(2)
        for i in range(S):
            mdm.Calculate()
            if i % OBSERVE_LOOP == 0:
                mdm.observables()
        unit.event_record(3)
        unit.synchronize()
        barrier()
        ms_s = allmax(unit.event_elapsed_ms(2, 3))
        tms, lss, cs1 = unit.timers(), unit.list_stats(), unit.counts()
        unit.set_profiling(0)
        nreb = cs1["nbuilds"] - cs0["nbuilds"]
        # one forced rebuild, timed on the device (MDManager::MakePairList on the current state)
        unit.event_record(4)
        mdm.MakePairList()
        unit.event_record(5)
        unit.synchronize()
        reb_ms = allmax(unit.event_elapsed_ms(4, 5))
        walk_ms = tms["force"] / max(tms["force_launches"], 1)
        der_n, der_ms = lss["derive_launches"], lss["derive_ms"] / max(lss["derive_launches"], 1)
        steady = {"steps": S, "thermalize_steps": args.thermalize, "ms_per_step": ms_s / S,
                  "value": n_total * S / (ms_s * 1e-3), "unit": UNIT,
                  "rebuilds": nreb, "rebuild_every_steps": (S / nreb) if nreb else None, "rebuild_ms": reb_ms,
                  "force_walk_ms": walk_ms, "inner_list_derivations": der_n, "derive_ms": der_ms,
                  "derive_every_steps": (S / der_n) if der_n else None,
                  "observations": (S + OBSERVE_LOOP - 1) // OBSERVE_LOOP,
                  "pairs_per_particle": cs1["npairs"] / max(cs1["pn"], 1),
                  "note": "device time over S consecutive steps with rebuilds, derivations and observations at their "
                          "natural cadence; rebuild_ms = one forced MakePairList timed on its own"}

    # --- roofline of the dominant kernel (force kick), measured live in the timed region above ---
    peak, peak_src = measured_peaks()
    P = npairs / max(n_local, 1)  # listed unordered pairs per particle (reference count)
    f_ms = tm["force"] / max(tm["force_launches"], 1)
    alg_bytes = (4.0 * P + 96.0) * n_local   # half-list indices + q read + p read/write
    alg_flops = 30.0 * P * n_local
    achieved = alg_bytes / (f_ms * 1e-3) / 1e9 if f_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "force_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # dram bytes per particle of the force kernel from one ncu --set full capture (static: a number taken
            # under the profiler cannot be re-measured inside this run), scaled to this launch
            traffic = tj["dram_bytes_per_particle"] * n_local
            traffic_src = "static: " + tj.get("source", "ncu --set full capture under profiles/")
        except Exception:
            traffic = None
    try:
        fp64_peak = capi.measure_fp64_peak(local_rank)
        fp64_src = "measured here: DFMA loop, 16 chains per thread (mdb200_measure_fp64_peak)"
    except Exception:
        fp64_peak, fp64_src = 37.0, "nominal B200 FP64 (measurement failed)"
    roofline = {"bound": "hbm", "kernel": "k_tile_force_p (persistent LJ force kick + half drifts)", "achieved": achieved, "peak": peak,
                "limiter": "the kernel is not HBM-bound: ncu has the SM's shared-memory data pipe at 86 % of its peak (the "
                           "partner gathers, 11 wavefronts per warp-step of which 5 are bank conflicts) and FP64 at 50 % "
                           "(profiles/r02_final_force_ncu.txt, DESIGN.md section 5)",
                "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "ms_per_launch": f_ms, "launches_timed": tm["force_launches"],
                "algorithmic_bytes_per_particle": 4.0 * P + 96.0, "pairs_per_particle": P,
                "force_share_of_step": (tm["force"] / ms_total) if ms_total else None,
                "fp64": {"achieved_tflops": alg_flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else 0.0,
                         "peak_tflops": fp64_peak, "peak_source": fp64_src,
                         "frac": (alg_flops / (f_ms * 1e-3) / 1e12 / fp64_peak) if (f_ms > 0 and fp64_peak) else None,
                         "flops_per_pair": 30}}

    # --- timed region 2: `e2e` through the C-ABI with HOST buffers ---
    dd = unit.download(want_type=False, out=(qh, ph))  # current state -> pinned host (untimed)
    n_now = dd["pn"]
    unit.synchronize()
    barrier()
    te0 = time.time()
    unit.upload4_begin(qh[:n_now], ph[:n_now])     # H2D inside the timed region (momenta travel during the first build)
    for i in range(K):
        mdm.Calculate()                            # first call rebuilds the list (pairlist.cc:27-29)
        if i % OBSERVE_LOOP == 0:
            obs2 = mdm.observables()               # D2H of the observation
    unit.download(want_type=False, out=(qh, ph))   # D2H of the final state
    unit.synchronize()
    barrier()
    te1 = time.time()
    e2e_s = allmax(te1 - te0)
    e2e_val = n_total * K / e2e_s
    h2d = (2 * n_now * 32) / K
    d2h = (2 * n_now * 32 + n_obs * 32) / K

    # --- multi-GPU parity (untimed): a small system on the same rank grid against the oracle, through the same
    #     transport the timed regions used (tests/multigpu_parity_lib.py; the oracle is the checker here) ---
    parity = None
    if world > 1 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import multigpu_parity_lib
        pr = multigpu_parity_lib.run_parity(world, rank, local_rank, uid, unit_length=14.0, steps=120, dt=0.004, v0=2.0)
        bad = allsum(0.0 if pr["ok"] else 1.0)
        parity = {"ok": bad == 0.0, "failed_ranks": int(bad), "worst_KUW": allmax(pr.get("worst_KUW") or 0.0),
                  "worst_q": allmax(pr.get("worst_q") or 0.0), "worst_p": allmax(pr.get("worst_p") or 0.0),
                  "builds": pr.get("builds"), "n_unit": pr.get("n_unit"), "grid": pr.get("grid"), "steps": pr["steps"],
                  "error": pr.get("error"),
                  "transport": "NCCL send/recv" if os.environ.get("MDB200_P2P", "1")[:1] == "0" else "peer-to-peer windows",
                  "what": "UnitLength 14 per rank, 120 hot steps (dt 0.004, v0 2.0): pair sets / rebuild schedule / pn / tn "
                          "exact, K,U,W <= 1e-9, final q,p <= 1e-10 against the CPU oracle on the same unit grid"}

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic" + (" (initial state made on the device)" if args.device_init else ""),
            "config": {"workload": "%s; %d particles on rank 0, unit grid %dx%dx%d; BASELINE.json configs[%d]" % (
                           wl_text, n_local, grid[0], grid[1], grid[2],
                           {"benchmark": 1 if world == 1 else 2, "shipped": 0, "langevin": 3, "collision": 4}[wl]),
                       "workload_key": wl,
                       "load_balance": {"particles_max_over_ranks": n_max, "particles_mean": n_total / world,
                                        "efficiency_bound": (n_total / world) / max(n_max, 1)},
                       "n_particles": n_total, "thermalize_steps": args.thermalize, "rebuilds_in_window": rebuilds,
                       "observations_in_window": n_obs,
                       "window_note": ("no rebuild fell into the K timed steps: steady_state.value is the rebuild-inclusive "
                                       "figure" if rebuilds == 0 else "the timed steps contain %d rebuild(s)" % rebuilds),
                       "l2": "inputs (%.1f GB of positions, momenta and pair list per GPU) far exceed the 126 MB L2; "
                             "no flush needed" % ((n_local * 64 + (2 * npairs * 4.4)) / 1e9),
                       "parallelism": ("domain decomposition, one unit per GPU; per-step ghost refresh by %s, migration and "
                                       "reductions over NCCL" % ("NCCL send/recv" if os.environ.get("MDB200_P2P", "1")[:1] == "0"
                                                                 else "peer-to-peer stores over NVLink (CUDA IPC windows)"))
                       if world > 1 else "single unit",
                       "timing": "CUDA events on the unit's stream, max over ranks; host wall %.3f s" % (tw1 - tw0)},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "seconds": e2e_s, "includes": "upload from pinned host (mdb200_upload_begin: the momenta travel during "
                                                  "the first pair-list build), first pair-list build, K steps, "
                                                  "observations, download of q and p"},
            "gpu_launches": launches,
            "roofline": roofline,
            "steady_state": steady,
            "parity_n": parity,
            "observables_last": obs,
        }
    mdm.close()
    return out, rank, world


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--thermalize", type=int, default=150)   # ThermalizeLoop of benchmark.cfg (benchmark.cc:24-27)
    ap.add_argument("--steady-steps", type=int, default=300)  # length of the steady_state window (0 = skip)
    ap.add_argument("--no-parity", action="store_true")      # skip the multi-GPU parity block
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--unit-length", type=float, default=float(os.environ.get("MDB200_BENCH_UNIT_LENGTH", "284")))
    ap.add_argument("--workload", default="benchmark", choices=["benchmark", "shipped", "langevin", "collision"])
    ap.add_argument("--collision-scale", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-init", action="store_true")    # initial state by the on-device makers (generate.cu)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        # same cadence as the GPU arm: ThermalizeLoop steps (benchmark.cc:24-27) + the warm-up, then the timed steps --
        # the reference's own [MUPS] line covers exactly its TotalLoop steps after ThermalizeLoop; K is bounded so that
        # the run ends within a few minutes (about 0.25 s per step for two one-GPU systems on 16 cores)
        therm = args.thermalize + max(3, min(args.warmup, 5))
        loop = max(5, min(args.steps, 20))
        if args.workload == "benchmark":
            r = run_reference_cpu(args.unit_length, args.gpus, therm, loop)
        else:
            r = run_reference_cpu_project(args.workload, args.unit_length, args.collision_scale, loop)
        n_units = r.get("cores", 1)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (r["sec"] / r["steps"] * 1e3) if r.get("sec") else None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload_key": args.workload,
                       "workload": ("benchmark.cfg-shaped LJ FCC on the host CPU (reference arm of the %d-GPU run): one "
                                    "cubic unit per host core, units of the 2x2x2 split of the one-GPU box (UnitLength=%g "
                                    "per GPU), cutoff 3.0; each run a bounded sample: %d thermalise + %d timed steps, one "
                                    "observation per %d steps" % (args.gpus, args.unit_length, therm, loop, OBSERVE_LOOP))
                       if args.workload == "benchmark" else
                       "reference arm of the %d-GPU %s run on the host CPU, bounded sample: %s" % (
                           args.gpus, args.workload, r["sample"])},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": n_units, "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

    out, rank, world = run_ours(args)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and args.workload == "benchmark":
            try:
                # bounded sample of the same workload in the same (thermalised) state: ThermalizeLoop + 12 timed steps
                r = run_reference_cpu(args.unit_length, 1, args.thermalize, 12)
                out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                       "sample": r["sample"]}
            except Exception as ex:  # the baseline is reported, never required for the GPU number
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference",
                                       "sample": "failed: %r" % (ex,)}
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return 0


if __name__ == "__main__":
    sys.exit(main())
