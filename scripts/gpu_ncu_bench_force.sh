# ncu --set full of the two force kernels in the thermalised 16 M-particle bench state
# -> gpurun_out/force_inner.ncu-rep (k_force walking the inner list), gpurun_out/force_prune.ncu-rep (deriving launch)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:^k_force$' -s 175 -c 1 \
  -o gpurun_out/force_inner -f python bench.py --steps 40 --warmup 150 --no-cpu-baseline > gpurun_out/ncu_force_inner.log 2>&1
tail -2 gpurun_out/ncu_force_inner.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:^k_force_prune$' -s 9 -c 1 \
  -o gpurun_out/force_prune -f python bench.py --steps 40 --warmup 150 --no-cpu-baseline > gpurun_out/ncu_force_prune.log 2>&1
tail -2 gpurun_out/ncu_force_prune.log | cut -c1-200
ls -la gpurun_out
