# usage: gpu_mgpu.sh N  : multi-GPU parity (N processes, NCCL) with and without the overlapped exchange
N=${1:-2}
for ov in 0 1; do
  export MDB200_OVERLAP=$ov
  echo "== overlap=$ov"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/multigpu_parity.py 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -6
done
