# ncu --set full of the one-pass pair search (tiled and plain) on a 1.9 M-particle Langevin run
mkdir -p gpurun_out
for v in 1 0; do
MDB200_SEARCH_TILE=$v timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:k_pair_search_onepass' -s 2 -c 1 \
  -o gpurun_out/search_tile$v -f python scripts/scaled_configs.py langevin --unit-length 140 --steps 30 --warmup 40 > gpurun_out/ncu_search_tile$v.log 2>&1
tail -2 gpurun_out/ncu_search_tile$v.log | cut -c1-200
done
