# ncu of the force, derive and search kernels in one short bench run (3 separate captures)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_force -s 12 -c 1 -o gpurun_out/force_r2 python bench.py --steps 10 --warmup 20 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_derive -s 2 -c 1 -o gpurun_out/derive_r2 python bench.py --steps 10 --warmup 20 --thermalize 40 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_derive.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 2 -c 2 -o gpurun_out/search_r2 python bench.py --steps 10 --warmup 20 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_search.log 2>&1
ls -la gpurun_out/*.ncu-rep
