# quick parity check on the GPU: smoke + committed goldens (+ optional extra pytest arguments)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py -x -q -m gpu "$@" > gpurun_out/quick.log 2>&1
echo "exit $?" >> gpurun_out/quick.log
tail -40 gpurun_out/quick.log
