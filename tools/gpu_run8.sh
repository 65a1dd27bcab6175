python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python tools/ab_kernels.py C3 "ME_NN_VARIANT=81" "ME_NN_VARIANT=82" "ME_NN_VARIANT=161" "ME_NN_VARIANT=162" "ME_NN_VARIANT=81,ME_NN_BLOCKS=64,ME_MME_BLOCKS=256" "ME_NN_VARIANT=81,ME_NN_BLOCKS=256,ME_MME_BLOCKS=1024" "ME_NN_TILE=1" 2>&1 | tee gpurun_out/ab8.log
