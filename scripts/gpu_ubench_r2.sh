# round-2 design micro-benchmarks: tile force walk in shared memory, red.global.add.f64
set -x
mkdir -p gpurun_out
cd scripts/ubench
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > ../../gpurun_out/ubench_clocks.csv &
SMI=$!
( ./tile_force 3 3.15 2368 512; ./tile_force 3 3.3 2368 512; ./tile_force 3 3.05 2368 512; ./tile_force 3 3.15 2368 1024; ./tile_force 2 3.15 4736 256; ./tile_force 2 3.15 4736 512; ./tile_force 3 3.15 2368 768 ) > ../../gpurun_out/ubench_tile_force.txt 2>&1
./red_f64 > ../../gpurun_out/ubench_red_f64.txt 2>&1
kill $SMI
cat ../../gpurun_out/ubench_tile_force.txt ../../gpurun_out/ubench_red_f64.txt
sort ../../gpurun_out/ubench_clocks.csv | uniq -c | sort -rn | head -5
