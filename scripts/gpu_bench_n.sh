# usage: gpu_bench_n.sh N TAG [extra bench args]  : bench.py on N GPUs (torchrun for N > 1) -> gpurun_out/bench_TAG.json
N=$1; TAG=$2; shift 2
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
fi
tail -2 gpurun_out/bench_$TAG.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/bench_$TAG.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print("$TAG", "value %.3f G/s" % (d["value"] / 1e9), "ms/step %.3f" % d["ms_per_step"], "e2e %.3f G/s" % (d["e2e"]["value"] / 1e9),
              "force ms %.3f" % d["roofline"]["ms_per_launch"], "frac %.3f" % d["roofline"]["frac"])
PY
