# cautious check after a batch of kernel changes: every stage under a short timeout
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -5 gpurun_out/quick.log
MDB200_DEBUG=1 timeout 120 python scripts/perf_probe.py 100 > gpurun_out/probe100.log 2>&1; echo "probe100 exit $?"; grep -v "^\[mdb200\] rebuild" gpurun_out/probe100.log | tail -8
MDB200_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 5 --steady-steps 100 --no-cpu-baseline > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err; echo "bench exit $?"; grep -c rebuild gpurun_out/bench_s20.err; tail -4 gpurun_out/bench_s20.err | cut -c1-200
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_s20.json'))
    s=d.get('steady_state') or {}
    print('s20: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step rebuild %.1f ms every %s derive %.2f ms every %s walk %.3f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('rebuild_ms',0), s.get('rebuild_every_steps'), s.get('derive_ms',0), s.get('derive_every_steps'), s.get('force_walk_ms',0)))
    print('fp64', d['roofline']['fp64'])
except Exception as e:
    print('no bench json', e)
PY
