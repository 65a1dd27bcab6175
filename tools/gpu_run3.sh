set -x
ncu --set full --clock-control none --import-source on -k regex:'flat_kernel' -c 3 -o gpurun_out/prof_flat python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
