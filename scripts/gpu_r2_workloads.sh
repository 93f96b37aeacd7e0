# the other BASELINE.json workloads through bench.py (GPU arm), plus the full default bench line with its CPU baseline
mkdir -p gpurun_out
for wl in shipped langevin collision; do
  timeout 600 python bench.py --workload $wl --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench_wl_$wl.json 2> gpurun_out/bench_wl_$wl.err; echo "bench $wl exit $?"
  python - $wl <<'PY'
import json,sys
wl=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/bench_wl_%s.json'%wl) if l.startswith('{')][-1])
    s=d.get('steady_state') or {}
    print('%s: %s | value %.3f G  ms/step %.3f  force ms %.3f  e2e %.3f G  launches %s' % (wl, d['config']['workload'][:110], d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value']/1e9, d.get('gpu_launches')))
except Exception as e:
    print(wl,'no bench json', e)
PY
  tail -2 gpurun_out/bench_wl_$wl.err | cut -c1-200
done
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench exit $?"; cut -c1-600 gpurun_out/bench_default.json
