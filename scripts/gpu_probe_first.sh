for ul in 100 284; do
MDB200_SAMPLE_TILES=1 MDB200_DEBUG=1 timeout 300 python - $ul <<'PY' 2>&1 | grep -E "first build|one-pass|max_row|tiles" | head
import sys
from mdacp_b200 import driver
mdm = driver.MDManager({"Mode": "Benchmark", "UnitLength": sys.argv[1], "TimeStep": "0.001"}, devices=[0], cutoff=3.0)
u = mdm.GetMDUnit(0)
u.generate_benchmark(0.7); u.generate_initial_velocity(1.0, 0)
mdm.MakePairList()
print("max_row", u.list_stats()["max_row"], u.counts())
PY
done
