# fused derivation inside the persistent force launch: parity (goldens through the persistent kernel), then A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_golden.py tests/test_gpu_split_launch.py -x -q -m gpu > gpurun_out/quick.log 2>&1; tail -4 gpurun_out/quick.log
for fd in 1 0; do
  MDB200_FUSED_DERIVE=$fd timeout 600 python bench.py --steps 20 --warmup 5 --steady-steps 200 --no-cpu-baseline > gpurun_out/bench_fd$fd.json 2> gpurun_out/bench_fd$fd.err
  python - $fd <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_fd%s.json'%sys.argv[1]) if l.startswith('{')][-1])
s=d.get('steady_state') or {}
print('fused_derive=%s: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step rebuild %.1f ms every %s derive %.2f ms every %s walk %.3f' % (sys.argv[1], d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('rebuild_ms',0), s.get('rebuild_every_steps'), s.get('derive_ms',0), s.get('derive_every_steps'), s.get('force_walk_ms',0)))
PY
done
MDB200_DEBUG_FORCE=1 timeout 300 python scripts/e2e_phases.py 284 5 40 2>&1 | grep -E "deriving|rep " | head -12
MDB200_DEBUG=1 timeout 300 python scripts/e2e_phases.py 284 20 150 > gpurun_out/e2e_phases.log 2>&1; grep -E "^rep|lap|rebuild|search|migration|border|cell" gpurun_out/e2e_phases.log | tail -40
