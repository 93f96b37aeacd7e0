mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 --steady-steps 100 --no-parity > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2b.json') if l.startswith('{')][-1])
s=d.get('steady_state') or {}
print('N=%d: value %.3f G  ms/step %.3f  force ms %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step | clocks %s' % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), d['clocks']))
PY
