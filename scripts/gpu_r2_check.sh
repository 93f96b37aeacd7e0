# full GPU test suite + bench runs (driver-like short window, default thermalised window)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
MDB200_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err
grep -E "rebuild|tiles|realloc|grows" gpurun_out/bench_s20.err | head -30
cut -c1-1800 gpurun_out/bench_s20.json
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cut -c1-1800 gpurun_out/bench_default.json
