for bf in 0 1; do
  touch cloud_map_evaluation_b200/csrc/mme.cu
  if [ $bf = 1 ]; then X="-DME_MME_BRANCHFREE"; else X=""; fi
  make -s -C cloud_map_evaluation_b200/csrc EXTRA="$X" 2>&1 | grep -E " error"
  echo "branch-free $bf"; python tools/ab_kernels.py C3 "" 2>&1 | tail -1
done | tee gpurun_out/ab23.log
