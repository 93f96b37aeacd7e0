for v in 10 16 24 32; do
  MDB200_INNER_STEPS=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_inner_$v.json 2> gpurun_out/bench_inner_$v.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_inner_$v.json") if l.startswith("{")][-1])
print("INNER_STEPS=$v", "value %.3f G/s" % (d["value"] / 1e9), "ms/step %.3f" % d["ms_per_step"], "force ms %.3f" % d["roofline"]["ms_per_launch"], "clk", d["clocks"]["sm_mhz"])
PY
done
