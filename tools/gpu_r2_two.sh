#!/bin/bash
mkdir -p gpurun_out
{
for cfg in c3 c5; do
  GSP_CHOL_ALGO=panel python tools/gpu_dist.py 1 $cfg 3
  echo "two CTAs per SM:"; GSP_GEMM_SUB_TWO=1 GSP_CHOL_ALGO=panel python tools/gpu_dist.py 1 $cfg 3
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_gemm_two_1gpu.log
GSP_GEMM_SUB_TWO=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "potrf or lusim_c3_size or lusim_conditional or c2_size_well or w1_without" 2>&1 | tail -2
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2_size_well or w1_without" 2>&1 | tail -2
