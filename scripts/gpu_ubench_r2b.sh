set -x
mkdir -p gpurun_out
cd scripts/ubench
( ./tile_force 3 3.15 2368 512; ./tile_force 3 3.15 296 512; ./tile_force 3 3.15 148 512; ./tile_force 2 3.15 4736 256 ) > ../../gpurun_out/ubench_tile_force2.txt 2>&1
cut -c1-300 ../../gpurun_out/ubench_tile_force2.txt
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 3 -c 1 -o ../../gpurun_out/ubench_tile_m0 ./tile_force 3 3.15 2368 512 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 19 -c 1 -o ../../gpurun_out/ubench_tile_m1 ./tile_force 3 3.15 2368 512 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 35 -c 1 -o ../../gpurun_out/ubench_tile_m2 ./tile_force 3 3.15 2368 512 > /dev/null 2>&1
ls -la ../../gpurun_out
