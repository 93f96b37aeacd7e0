# final evidence of the round: small-system probe, launch list of a bench run, ncu --set full of the main kernels
mkdir -p gpurun_out
timeout 200 python scripts/perf_probe.py 50 100 > gpurun_out/probe_small.log 2>&1; echo "probe exit $?"; cat gpurun_out/probe_small.log | cut -c1-250
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --thermalize 60 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "launch list exit $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_force -s 12 -c 1 -o gpurun_out/force_r2f python bench.py --steps 10 --warmup 20 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force_f.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_tile_compact|k_redeal" -s 1 -c 2 -o gpurun_out/derive_r2f python bench.py --steps 10 --warmup 20 --thermalize 60 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_derive_f.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 2 -c 1 -o gpurun_out/search_r2f python bench.py --steps 10 --warmup 20 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_search_f.log 2>&1
ls -la gpurun_out/*_r2f.ncu-rep gpurun_out/launches_r2.csv
