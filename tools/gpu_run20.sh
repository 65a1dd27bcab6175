for mb in 8 10 12; do
  touch cloud_map_evaluation_b200/csrc/nn.cu; make -s -C cloud_map_evaluation_b200/csrc EXTRA="-DME_NN_MIN_BLOCKS=$mb" 2>&1 | grep -E " error"
  echo "NN min blocks $mb"; python tools/ab_kernels.py C3 "" 2>&1 | tail -1
done | tee gpurun_out/ab20.log
