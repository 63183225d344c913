#!/bin/bash
# 2-GPU call: real multi-GPU tests, the panel algorithm on one GPU under the LUSIM parity tests, factorization timing sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
( time python -m pytest tests/test_multi_gpu.py -m gpu -x -q --durations=8 ) > gpurun_out/r2_pytest_multigpu_2.log 2>&1
tail -15 gpurun_out/r2_pytest_multigpu_2.log
( time GSP_CHOL_ALGO=panel python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "potrf or lusim and not c5_full" ) > gpurun_out/r2_pytest_panel_1gpu.log 2>&1
tail -5 gpurun_out/r2_pytest_panel_1gpu.log
{
for cfg in c3 c5; do
  GSP_CHOL_ALGO=recursive python tools/gpu_dist.py 1 $cfg
  for pb in 2 4 8; do GSP_CHOL_ALGO=panel GSP_CHOL_PB=$pb python tools/gpu_dist.py 1 $cfg; done
  for pb in 2 4 8; do GSP_CHOL_PB=$pb python tools/gpu_dist.py 2 $cfg; done
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_dist_sweep_2gpu.log
