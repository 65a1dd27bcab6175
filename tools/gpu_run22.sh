python tools/ab_kernels.py C3 "" "CELL=0.04" "CELL=0.0667" "CELL=0.1" 2>&1 | tee gpurun_out/ab22.log
