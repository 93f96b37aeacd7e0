# 8 GPUs: the default transport of an 8-rank run (peer-to-peer windows, overlapped exchange) and the NCCL in-line variant
# against the oracle, the bench with its parity_n block, and BASELINE.json configs[3] / [4] on 8 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -x -q -m gpu -k "p2p_overlap-env1-8 or nccl_inline-env2-8" > gpurun_out/mgpu_pytest_n8.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/mgpu_pytest_n8.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --steady-steps 100 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench exit $?"
timeout 400 $TR bench.py --gpus 8 --workload langevin --steps 40 --warmup 10 --steady-steps 0 --no-parity > gpurun_out/bench_n8_langevin.json 2> gpurun_out/bench_n8_langevin.err; echo "langevin exit $?"
timeout 400 $TR bench.py --gpus 8 --workload collision --steps 60 --warmup 10 --steady-steps 0 --no-parity > gpurun_out/bench_n8_collision.json 2> gpurun_out/bench_n8_collision.err; echo "collision exit $?"
python - <<'PY'
import json
for f in ['n8','n8_langevin','n8_collision']:
    try:
        d=json.loads([l for l in open('gpurun_out/bench_%s.json'%f) if l.startswith('{')][-1])
        s=d.get('steady_state') or {}
        print('%s N=%d: value %.3f G  ms/step %.3f  force ms %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step derive %.2f rebuild %.1f | %d particles, rebuilds %s' % (f, d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('derive_ms',0), s.get('rebuild_ms',0), d['config']['n_particles'], d['config']['rebuilds_in_window']))
        if d.get('parity_n'): print('  parity_n ok=%s worst KUW %.2e q %.2e p %.2e transport %s' % (d['parity_n']['ok'], d['parity_n']['worst_KUW'], d['parity_n']['worst_q'], d['parity_n']['worst_p'], d['parity_n']['transport']))
    except Exception as e:
        print(f, 'no bench json', e)
PY
tail -3 gpurun_out/bench_n8_collision.err | cut -c1-300
