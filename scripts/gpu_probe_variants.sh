# usage: gpu_probe_variants.sh L1 [L2 ...]  : perf_probe for every library in build/variants
for f in build/variants/*.so; do
  echo "== $f"
  MDB200_LIB_PATH=$f timeout 300 python scripts/perf_probe.py "$@" 2>&1 | grep -E "^L=|list:|phases|force kernel"
done
