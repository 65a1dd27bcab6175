set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mme_kernel|nn_tile_kernel' -c 3 -o gpurun_out/prof_v2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
cat gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c3.json
