# usage: gpu_ncu_force.sh <lib.so> <tag> [L]   -> gpurun_out/force_<tag>.ncu-rep (one k_force launch, --set full)
LIBF=$1; TAG=$2; L=${3:-120}
MDB200_LIB_PATH=$LIBF timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force -s 30 -c 1 \
  -o gpurun_out/force_$TAG -f python scripts/perf_probe.py $L > gpurun_out/ncu_force_$TAG.log 2>&1
tail -3 gpurun_out/ncu_force_$TAG.log
