# usage: gpu_ncu_kernel.sh <kernel-regex> <tag> [L] [skip] [count]  -> gpurun_out/<tag>.ncu-rep (--set full)
K=$1; TAG=$2; L=${3:-120}; SKIP=${4:-0}; CNT=${5:-1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT \
  -o gpurun_out/$TAG -f python scripts/perf_probe.py $L > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
