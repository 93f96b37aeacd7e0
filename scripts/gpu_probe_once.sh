# usage: gpu_probe_once.sh L  : perf_probe once for every library in build/variants
for f in build/variants/*.so; do
  echo "== $f"
  MDB200_LIB_PATH=$f timeout 300 python scripts/perf_probe.py "$@" 2>&1 | grep -E "steps:|phases|force kernel"
done
