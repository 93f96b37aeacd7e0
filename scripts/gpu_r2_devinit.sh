# bench.py --device-init (on-device makers) and the collision workload's buffers: quick check
mkdir -p gpurun_out
timeout 300 python bench.py --workload collision --steps 20 --warmup 5 --steady-steps 0 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "collision exit $?"; tail -1 gpurun_out/bench_c.err | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --steady-steps 0 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "default exit $?"
python - <<'PY'
import json
for f in ['bench_c','bench_q']:
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
    print(f, 'value %.3f G ms/step %.3f e2e %.3f G N %d clocks %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['config']['n_particles'], d['clocks']))
PY
