# iteration check: goldens, short bench, launch list of a short bench, ncu of one force launch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -3 gpurun_out/quick.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s20.json'))
s=d.get('steady_state') or {}
print('s20: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step rebuild %.1f ms every %s derive %.2f ms every %s walk %.3f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('rebuild_ms',0), s.get('rebuild_every_steps'), s.get('derive_ms',0), s.get('derive_every_steps'), s.get('force_walk_ms',0)))
print('fp64 peak', d['roofline']['fp64'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --thermalize 100 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches.csv 2>/dev/null | head -40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_force -s 12 -c 1 -o gpurun_out/force_r2 python bench.py --steps 10 --warmup 20 --steady-steps 0 --no-cpu-baseline > gpurun_out/ncu_force.log 2>&1
ls -la gpurun_out/*.ncu-rep
