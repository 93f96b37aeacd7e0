mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_split_launch.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -4 gpurun_out/quick.log
timeout 900 python -m pytest tests/test_gpu_multiprocess.py -x -q -m gpu > gpurun_out/mgpu_pytest_n2.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/mgpu_pytest_n2.log | cut -c1-300
for ov in 1 0; do
MDB200_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 --steady-steps 100 > gpurun_out/bench_n2_ov$ov.json 2> gpurun_out/bench_n2_ov$ov.err; echo "bench exit $?"
python - $ov <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/bench_n2_ov%s.json'%sys.argv[1]) if l.startswith('{')][-1])
    s=d.get('steady_state') or {}
    print('overlap=%s N=%d: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step derive %.2f walk %.3f rebuild %.1f' % (sys.argv[1], d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('derive_ms',0), s.get('force_walk_ms',0), s.get('rebuild_ms',0)))
    print('parity_n', (d.get('parity_n') or {}).get('ok'))
except Exception as e:
    print('no bench json', e)
PY
done
