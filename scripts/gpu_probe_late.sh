for env in "X=1" "MDB200_UPLOAD_ASYNC=0"; do
  env $env MDB200_DEBUG=1 timeout 300 python bench.py --workload shipped --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$env" <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/b.json') if l.startswith('{')][-1])
print(sys.argv[1], 'e2e seconds', d['e2e']['seconds'], 'ms/step', d['ms_per_step'])
PY
  grep -c realloc gpurun_out/b.err; grep realloc gpurun_out/b.err | tail -5
done
