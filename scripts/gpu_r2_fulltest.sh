mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_tile_force_p<\(int\)[0-9], \(bool\)[01], \(bool\)1>' -s 1 -c 1 -o gpurun_out/force_derive_r2g python bench.py --steps 10 --warmup 25 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force_dg.log 2>&1; echo "deriving force exit $?"
[ -f gpurun_out/force_derive_r2g.ncu-rep ] && python scripts/ncu_top.py gpurun_out/force_derive_r2g.ncu-rep -1 30 > gpurun_out/force_derive_r2g_top.txt 2>&1
find gpurun_out -name "*.ncu-rep" -size +15M -delete
head -12 gpurun_out/force_derive_r2g_top.txt | cut -c1-140
