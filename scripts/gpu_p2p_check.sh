# usage: gpu_p2p_check.sh N : multi-process parity with the peer-to-peer ghost refresh (and with NCCL only), then the bench both ways
N=${1:-2}
for p2p in 1 0; do for ov in 0 1; do
  echo "== MDB200_P2P=$p2p MDB200_OVERLAP=$ov"
  MDB200_DEBUG=1 MDB200_P2P=$p2p MDB200_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/multigpu_parity.py 2>&1 | grep -E "OK|peer-to-peer|rror|FAIL|differ" | sort | uniq -c | head -8
done; done
for p2p in 1 0; do MDB200_P2P=$p2p bash scripts/gpu_bench_n.sh $N p2p${p2p}_n$N --no-cpu-baseline; done
