mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py tests/test_gpu_split_launch.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -4 gpurun_out/quick.log
for wl in collision langevin; do
  MDB200_DEBUG=1 timeout 600 python bench.py --workload $wl --steps 100 --warmup 20 --steady-steps 200 --no-cpu-baseline > gpurun_out/bench_wl_$wl.json 2> gpurun_out/bench_wl_$wl.err; echo "bench $wl exit $?"
  python - $wl <<'PY'
import json,sys
wl=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/bench_wl_%s.json'%wl) if l.startswith('{')][-1])
    s=d.get('steady_state') or {}
    print('%s: value %.3f G  ms/step %.3f  force ms %.3f  e2e %.3f G  launches %s | steady %.3f ms/step rebuild %.1f ms every %.1f derive %.2f walk %.3f' % (wl, d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value']/1e9, d.get('gpu_launches'), s.get('ms_per_step',0), s.get('rebuild_ms',0), s.get('rebuild_every_steps') or 0, s.get('derive_ms',0), s.get('force_walk_ms',0)))
except Exception as e:
    print(wl,'no bench json', e)
PY
  grep -E "tiles hold|tiles [0-9]" gpurun_out/bench_wl_$wl.err | tail -2
done
timeout 600 python bench.py --steps 20 --warmup 5 --steady-steps 200 --no-cpu-baseline > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_s20.json') if l.startswith('{')][-1])
s=d.get('steady_state') or {}
print('s20: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step rebuild %.1f ms every %s derive %.2f ms every %s walk %.3f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('rebuild_ms',0), s.get('rebuild_every_steps'), s.get('derive_ms',0), s.get('derive_every_steps'), s.get('force_walk_ms',0)))
PY
