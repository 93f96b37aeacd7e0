mkdir -p gpurun_out
for wl in benchmark collision; do
timeout 300 python bench.py --workload $wl --device-init --steps 20 --warmup 5 --steady-steps 0 --no-cpu-baseline > gpurun_out/bench_devinit_$wl.json 2> gpurun_out/bench_devinit_$wl.err; echo "exit $?"
python - $wl <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_devinit_%s.json'%sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], 'value %.3f G ms/step %.3f e2e %.3f G N %d obs %s | %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['config']['n_particles'], d['observables_last'], d['data']))
PY
tail -1 gpurun_out/bench_devinit_$wl.err | cut -c1-200
done
