# round-2 final evidence on one GPU: driver-like bench line (with the CPU baseline), default line, reference arm, workloads
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_s20.json 2> gpurun_out/bench_final_s20.err; echo "s20 exit $?"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err; echo "default exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref exit $?"
for wl in shipped langevin collision; do
  timeout 600 python bench.py --workload $wl --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench_wl_$wl.json 2> gpurun_out/bench_wl_$wl.err; echo "bench $wl exit $?"
done
python - <<'PY'
import json
for f in ['final_s20','final_default','wl_shipped','wl_langevin','wl_collision']:
    try:
        d=json.loads([l for l in open('gpurun_out/bench_%s.json'%f) if l.startswith('{')][-1])
        s=d.get('steady_state') or {}
        print('%s: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G launches %s | steady: %.3f G %.3f ms/step rebuild %.1f ms every %s derive %.2f every %s walk %.3f | cpu %s' % (f, d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, d.get('gpu_launches'), s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('rebuild_ms',0), s.get('rebuild_every_steps'), s.get('derive_ms',0), s.get('derive_every_steps'), s.get('force_walk_ms',0), (d.get('cpu_baseline') or {}).get('value')))
    except Exception as e:
        print(f, 'failed', e)
try:
    d=json.loads([l for l in open('gpurun_out/bench_final_ref.json') if l.startswith('{')][-1]); print('reference arm:', d['value']/1e9, 'G', d.get('cpu_baseline',{}).get('sample','')[:160])
except Exception as e:
    print('ref failed', e)
PY
