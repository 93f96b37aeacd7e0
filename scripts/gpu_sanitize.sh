# compute-sanitizer (memcheck, then racecheck + initcheck on the smoke case) over small GPU tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_smoke.py tests/test_gpu_edge_cases.py tests/test_dump_files.py "tests/test_gpu_golden.py::test_cuda_matches_reference_golden[collision2]" "tests/test_gpu_golden.py::test_cuda_matches_reference_golden[scale1]" -m gpu -x -q > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/memcheck_pytest.log; tail -4 gpurun_out/memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_gpu_smoke.py -m gpu -x -q > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -2 gpurun_out/racecheck_pytest.log; tail -3 gpurun_out/racecheck.log
