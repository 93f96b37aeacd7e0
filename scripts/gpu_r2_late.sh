mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_upload_begin.py tests/test_gpu_smoke.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -12 gpurun_out/quick.log
for b in 1 0; do
E2E_UPLOAD_BEGIN=$b timeout 300 python scripts/e2e_phases.py 284 20 150 2>&1 | grep -E "^rep" | tail -2
done
