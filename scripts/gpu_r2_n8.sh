# 8 GPUs: the default transport of an 8-rank run (peer-to-peer windows, overlapped exchange) and the in-line variant
# against the oracle, then the bench with its parity_n block
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -x -q -m gpu -k "p2p_overlap-env1-8 or nccl_inline-env2-8" > gpurun_out/mgpu_pytest_n8.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/mgpu_pytest_n8.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5 --steady-steps 100 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n8.json') if l.startswith('{')][-1])
    s=d.get('steady_state') or {}
    print('N=%d: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step' % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0)))
    print('parity_n', d.get('parity_n'))
except Exception as e:
    print('no bench json', e)
PY
tail -5 gpurun_out/bench_n8.err | cut -c1-300
