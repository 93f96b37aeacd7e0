mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py tests/test_gpu_split_launch.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -4 gpurun_out/quick.log
MDB200_DEBUG=1 timeout 300 python scripts/e2e_phases.py 284 20 150 > gpurun_out/e2e_phases.log 2>&1; grep -E "^rep|first build|one-pass|list grows|realloc" gpurun_out/e2e_phases.log | tail -20
