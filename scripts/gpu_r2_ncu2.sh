# ncu of the two force kernels (one block per tile / persistent) in short bench runs
mkdir -p gpurun_out
MDB200_FORCE_PERSIST=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_force -s 12 -c 1 -o gpurun_out/force_r2b python bench.py --steps 10 --warmup 20 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force_b.log 2>&1
MDB200_FORCE_PERSIST=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_force -s 12 -c 1 -o gpurun_out/force_r2p python bench.py --steps 10 --warmup 20 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force_p.log 2>&1
ls -la gpurun_out/*.ncu-rep
