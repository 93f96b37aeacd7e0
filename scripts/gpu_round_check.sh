# one comprehensive GPU call: parity tests, bench (N=1), launch list, per-launch force timings
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json | cut -c1-1500
MDB200_NO_PRUNE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_noprune.json 2> gpurun_out/bench_n1_noprune.err
cut -c1-200 gpurun_out/bench_n1_noprune.json
MDB200_DEBUG_FORCE=1 timeout 300 python scripts/perf_probe.py 200 > gpurun_out/probe200.log 2>&1
grep -v "force launch" gpurun_out/probe200.log | tail -8
grep "force launch" gpurun_out/probe200.log | sed -n '100,140p'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-300
