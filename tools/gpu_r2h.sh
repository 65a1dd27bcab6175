# round 2, run H (2 GPUs): the slab-layout tests, then C3 on 2 GPUs with the slab and the replicated layout
set -x
ME_DEBUG_SLAB=1 timeout 600 python -m pytest tests/test_gpu_slab.py -q --timeout=300 2>&1 | grep -v "^\[mapeval\] slab plan: rank [1-9]" | tail -60 | tee gpurun_out/pytest_r2h_slab.log
bash tools/gpu_multi.sh 2:C3:20:slab 2:C3:20:replicated
