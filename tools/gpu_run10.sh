python tools/ab_kernels.py C3 "ME_NN_VARIANT=161" "ME_NN_VARIANT=163" "ME_NN_VARIANT=83" "ME_MME_VARIANT=83" "ME_MME_VARIANT=163" 2>&1 | tee gpurun_out/ab10.log
