# final ncu evidence of round 2: launch list of a bench run, --set full of the force launches (plain and deriving), the
# one-pass search and the generator kernels; summaries are made on the box (scripts/ncu_top.py), reports kept if small
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 20 --warmup 5 --thermalize 60 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "launch list exit $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_force_p -s 14 -c 1 -o gpurun_out/force_r2g python bench.py --steps 10 --warmup 25 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force_g.log 2>&1; echo "force exit $?"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_tile_force_p<[0-9], *[a-z0-9]*, *(1|true)>" -s 1 -c 1 -o gpurun_out/force_derive_r2g python bench.py --steps 10 --warmup 25 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force_dg.log 2>&1; echo "deriving force exit $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 3 -c 1 -o gpurun_out/search_r2g python bench.py --steps 10 --warmup 20 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_search_g.log 2>&1; echo "search exit $?"
cat > gen16.py <<'PY'
from mdacp_b200 import driver
mdm = driver.MDManager({"Mode": "Benchmark", "UnitLength": "284", "TimeStep": "0.001"}, devices=[0], cutoff=3.0)
u = mdm.GetMDUnit(0)
print(u.generate_benchmark(0.7)); u.generate_initial_velocity(1.0, 0); u.synchronize()
PY
timeout 300 ncu --set full --clock-control none -k regex:"k_mt19937|k_gen_" -c 6 -o gpurun_out/gen_r2g python gen16.py > gpurun_out/ncu_gen_g.log 2>&1; echo "gen exit $?"
rm -f gen16.py
for r in force_r2g force_derive_r2g search_r2g; do
  [ -f gpurun_out/$r.ncu-rep ] && python scripts/ncu_top.py gpurun_out/$r.ncu-rep -1 40 > gpurun_out/${r}_top.txt 2>&1
done
[ -f gpurun_out/gen_r2g.ncu-rep ] && ncu -i gpurun_out/gen_r2g.ncu-rep --page raw --csv > gpurun_out/gen_r2g_raw.csv 2>/dev/null
ls -la gpurun_out/
find gpurun_out -name "*.ncu-rep" -size +15M -delete
du -sh gpurun_out
