# persistent force kernel: goldens with it forced onto the small tiles too, then the bench with each variant
mkdir -p gpurun_out
MDB200_FORCE_PERSIST=1 timeout 500 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/quick_persist.log 2>&1; echo "goldens (persist forced) exit $?"; tail -4 gpurun_out/quick_persist.log | cut -c1-300
for v in "MDB200_FORCE_PERSIST=1" "MDB200_FORCE_PERSIST=0"; do
  env $v MDB200_DEBUG_FORCE=1 timeout 300 python bench.py --steps 20 --warmup 5 --steady-steps 100 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v exit $?"
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_%s.json'%v))
    s=d.get('steady_state') or {}
    print('%s: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step walk %.3f' % (v, d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('force_walk_ms',0)))
except Exception as e:
    print('no bench json', e)
PY
  grep "persistent:" gpurun_out/bench_$v.err | tail -2 | cut -c1-250
done
