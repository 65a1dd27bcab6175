python tools/ab_kernels.py C3 "" 2>&1 | tee gpurun_out/ab19.log
