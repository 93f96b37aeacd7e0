# bench under a list of environment variants: bash scripts/gpu_r2_var.sh "A=1 B=2" "A=0" ...
mkdir -p gpurun_out
for v in "$@"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --steady-steps 100 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench [$v] exit $?"
  python - "$tag" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_%s.json'%v))
    s=d.get('steady_state') or {}
    print('%s: value %.3f G  ms/step %.3f  force ms %.3f frac %.3f  e2e %.3f G | steady: %.3f G %.3f ms/step walk %.3f rebuild %.1f ms derive %.2f' % (v, d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']/1e9, s.get('value',0)/1e9, s.get('ms_per_step',0), s.get('force_walk_ms',0), s.get('rebuild_ms',0), s.get('derive_ms',0)))
except Exception as e:
    print('no bench json', e)
PY
done
