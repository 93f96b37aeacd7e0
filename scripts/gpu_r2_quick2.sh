mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py tests/test_gpu_reference_driver.py -x -q -m gpu > gpurun_out/quick.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/quick.log | cut -c1-300
timeout 200 python scripts/perf_probe.py 50 52 100 > gpurun_out/probe_small.log 2>&1; echo "probe exit $?"; grep -v "list:" gpurun_out/probe_small.log | cut -c1-250
bash scripts/gpu_r2_var.sh "X=1"
