set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for f in build/variants/*.so; do echo "== $f"; MDB200_LIB_PATH=$f timeout 300 python scripts/perf_probe.py 120 200 2>&1 | grep -v "^$"; done
