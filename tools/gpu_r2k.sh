# round 2, run K (ONE 8-GPU box): the slab-layout tests, C5 (200 M vs 200 M) on 8 GPUs and C3 on 8 GPUs with the slab layout
set -x
timeout 120 python -m pytest tests/test_gpu_slab.py -q --timeout=100 2>&1 | tail -25 | cut -c1-700 | tee gpurun_out/pytest_r2k_slab.log
bash tools/gpu_multi.sh 8:C5:3:slab 8:C3:20:slab
