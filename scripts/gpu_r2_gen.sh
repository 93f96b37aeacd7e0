mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generate.py -x -q -m gpu > gpurun_out/gen.log 2>&1; tail -30 gpurun_out/gen.log
timeout 300 python - <<'PY' 2>&1 | tail -8
import time, numpy as np, torch
from mdacp_b200 import driver, hostlib
mdm = driver.MDManager({"Mode": "Benchmark", "UnitLength": "284", "TimeStep": "0.001"}, devices=[0], cutoff=3.0)
u = mdm.GetMDUnit(0)
for rep in range(3):
    u.clear(); u.synchronize()
    t0 = time.time(); n = u.generate_benchmark(0.7); u.synchronize(); t1 = time.time()
    u.generate_initial_velocity(1.0, 0); u.synchronize(); t2 = time.time()
    print("device: %d particles, lattice %.1f ms, velocities %.1f ms" % (n, 1e3*(t1-t0), 1e3*(t2-t1)), flush=True)
d = u.download()
s, e = mdm.GetRect(0)
t0 = time.time(); qh = hostlib.benchmark_config(s, e, 0.7, 0); ph = hostlib.initial_velocity(len(qh), 1.0, 0); t1 = time.time()
print("host: %.1f ms; q equal %s; max |dp| %.3g; |sum p| %.3g" % (1e3*(t1-t0), np.array_equal(d["q"], qh[:, :3]), np.max(np.abs(d["p"]-ph[:, :3])), np.max(np.abs(d["p"].sum(axis=0)))))
mdm.MakePairList(); mdm.Calculate(10); print(mdm.observables())
PY
