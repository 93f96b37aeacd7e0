# parity (goldens) + short bench + ncu of the force kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -3 gpurun_out/quick.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s20.json'))
print('s20: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G' % (d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9))
PY
timeout 600 python bench.py --steps 100 --warmup 150 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('default: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G rebuilds %d' % (d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, d['config']['rebuilds_in_window']))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_force -s 12 -c 1 -o gpurun_out/force_r2 python bench.py --steps 10 --warmup 20 --no-cpu-baseline > gpurun_out/ncu_force.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_reference_driver.py -x -q -m gpu > gpurun_out/refdriver.log 2>&1; tail -25 gpurun_out/refdriver.log
