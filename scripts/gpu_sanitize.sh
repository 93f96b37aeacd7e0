# compute-sanitizer (memcheck, then racecheck on the smoke case and on a persistent / deriving force case) over small GPU tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_smoke.py tests/test_gpu_edge_cases.py "tests/test_gpu_golden.py::test_cuda_matches_reference_golden[collision2]" \
  "tests/test_gpu_generate.py::test_device_generators_match_reference_initial_state" "tests/test_gpu_generate.py::test_generators_refuse_a_stepped_store_and_clear_resets_it" \
  "tests/test_gpu_upload_begin.py" "tests/test_gpu_split_launch.py::test_fused_inner_list_derivation" "tests/test_gpu_split_launch.py::test_first_build_sized_from_a_tile_sample" \
  -m gpu -x -q > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/memcheck_pytest.log; tail -4 gpurun_out/memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_gpu_smoke.py "tests/test_gpu_split_launch.py::test_fused_inner_list_derivation[bench1-1-False]" -m gpu -x -q > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -2 gpurun_out/racecheck_pytest.log; tail -3 gpurun_out/racecheck.log
