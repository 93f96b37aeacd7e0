# usage: gpu_probe_variants.sh L1 [L2 ...]  : perf_probe for every library in build/variants (two rounds, to expose drift)
for round in 1 2; do
for f in build/variants/*.so; do
  echo "== $f (round $round)"
  MDB200_LIB_PATH=$f timeout 300 python scripts/perf_probe.py "$@" 2>&1 | grep -E "phases|force kernel|clocks" &
  pid=$!
  sleep 4; nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader
  wait $pid
done
done
