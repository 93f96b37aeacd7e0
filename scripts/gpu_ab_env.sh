# usage: gpu_ab_env.sh VAR L  : perf_probe with VAR=0 and VAR=1, then the bench with each
VAR=$1; L=${2:-200}
mkdir -p gpurun_out
for v in 0 1; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 300 python scripts/perf_probe.py $L 2>&1 | grep -E "list:|steps:|phases|force kernel"
done
for v in 0 1; do
  env $VAR=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${VAR}_$v.json 2> gpurun_out/bench_${VAR}_$v.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_${VAR}_$v.json") if l.startswith("{")][-1])
print("$VAR=$v", "value %.3f G/s" % (d["value"] / 1e9), "ms/step %.3f" % d["ms_per_step"], "e2e %.3f G/s" % (d["e2e"]["value"] / 1e9),
      "force ms %.3f" % d["roofline"]["ms_per_launch"], "frac %.3f" % d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"])
PY
done
